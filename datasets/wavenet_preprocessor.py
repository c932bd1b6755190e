"""WaveNet-only preprocessing (reference datasets/wavenet_preprocessor.py:11-154): every wav of a folder -> `audio-<name>.npy` (mu-law
indices int16 / float32 samples, padded and cut to frames * hop_size) + `mel-<name>.npy` ([frames, num_mels] float32) and one `map.txt`
row `audio|mel|mel|<no_g>|time_steps|mel_frames` per utterance - the layout `train.py --model WaveNet --wavenet_input .../map.txt` reads
(wavenet_vocoder/feeder.py: columns 0, 1, 2, 3). Same step order as the Tacotron preprocessor; the mel runs on the fused GPU kernel."""
import os

import numpy as np

from datasets import audio
from wavenet_vocoder.util import is_mulaw, is_mulaw_quantize, mulaw, mulaw_quantize


def build_from_path(hparams, input_dir, mel_dir, wav_dir, n_jobs=1, tqdm=lambda x: x):
    """n_jobs is accepted for signature compatibility: the heavy part is one GPU kernel per utterance, so files are processed in order"""
    rows = []
    for name in tqdm(sorted(os.listdir(input_dir))):
        if not name.lower().endswith(".wav"):
            continue
        r = _process_utterance(mel_dir, wav_dir, name[:-4], os.path.join(input_dir, name), hparams)
        if r is not None:
            rows.append(r)
    return rows


def _process_utterance(mel_dir, wav_dir, index, wav_path, hparams):
    try:
        wav = audio.load_wav(wav_path, sr=hparams.sample_rate)
    except FileNotFoundError:
        print("file %s is not present in the wav folder. skipping!" % wav_path)
        return None
    if hparams.trim_silence:
        wav = audio.trim_silence(wav, hparams)
    emph = audio.preemphasis(wav, hparams.preemphasis, hparams.preemphasize)
    if hparams.rescale:
        wav = wav / np.abs(wav).max() * hparams.rescaling_max
        emph = emph / np.abs(emph).max() * hparams.rescaling_max
        for w in (wav, emph):
            if (w > 1.0).any() or (w < -1.0).any():
                raise RuntimeError("wav has invalid value: %s" % wav_path)
    wav, emph = np.asarray(wav, dtype=np.float32), np.asarray(emph, dtype=np.float32)
    if is_mulaw_quantize(hparams.input_type):
        out = mulaw_quantize(wav, hparams.quantize_channels)
        start, end = audio.start_and_end_indices(out, hparams.silence_threshold)       # cut leading / trailing digital silence
        wav, emph, out = wav[start:end], emph[start:end], out[start:end]
        pad_value, out_dtype = int(mulaw_quantize(np.zeros(1, dtype=np.float32), hparams.quantize_channels)[0]), np.int16
    elif is_mulaw(hparams.input_type):
        out, pad_value, out_dtype = mulaw(wav, hparams.quantize_channels), 0.0, np.float32
    else:
        out, pad_value, out_dtype = wav, 0.0, np.float32
    mel = audio.melspectrogram(emph, hparams).astype(np.float32)                        # [num_mels, frames]
    frames = mel.shape[1]
    if frames > hparams.max_mel_frames and hparams.clip_mels_length:
        return None
    hop = audio.get_hop_size(hparams)
    l_pad, r_pad = audio.librosa_pad_lr(wav, hparams.n_fft, hop, hparams.wavenet_pad_sides)
    out = np.pad(out, (l_pad, r_pad), mode="constant", constant_values=pad_value)
    assert len(out) >= frames * hop
    out = out[:frames * hop]                    # a whole number of hops, so that the conditioning upsamples onto it exactly
    if getattr(hparams, "gin_channels", 0) and hparams.gin_channels > 0:
        raise RuntimeError("global conditioning needs a speaker-id rule here (datasets/wavenet_preprocessor.py:146-149 of the reference)")
    audio_path, mel_path = os.path.join(wav_dir, "audio-%s.npy" % index), os.path.join(mel_dir, "mel-%s.npy" % index)
    np.save(audio_path, out.astype(out_dtype), allow_pickle=False)
    np.save(mel_path, mel.T, allow_pickle=False)
    return (audio_path, mel_path, mel_path, "<no_g>", len(out), frames)
