"""Tacotron Synthesizer (reference tacotron/synthesizer.py): load a checkpoint, then `synthesize(texts, basenames, out_dir, log_dir,
mel_filenames)` writes `mel-<basename>.npy` ([frames, num_mels] float32). GTA mode teacher-forces on the ground-truth mels
(mel_filenames given); natural mode runs the free-running decoder until every row's stop token fires. With a log_dir (eval mode) the
Griffin-Lim previews of the reference are written too (`wavs/wav-<b>-mel.wav`, and `wav-<b>-linear.wav` + `linear-<b>.npy` when the
post-processing net is on; GPU Griffin-Lim of datasets/audio.py) and so are the plots (`plots/alignment-<b>.png`, `plots/mel-<b>.png`,
`plots/linear-<b>.png`; tacotron/utils/plot.py writes the PNGs itself)."""
import os

import numpy as np
import torch

import t2_checkpoint
from datasets import audio
from tacotron.feeder import pad_input, pad_target
from tacotron.models import create_model
from tacotron.utils import plot
from tacotron.utils.text import text_to_sequence


def get_output_lengths(stop_tokens):
    """synthesizer.py:254-257 _get_output_lengths: a row's length is the INDEX of its first rounded stop prediction of 1 (the frame the
    stop fires on is not kept), the whole row when it never fires"""
    out = []
    for row in np.round(np.asarray(stop_tokens)):
        hit = np.nonzero(row == 1)[0]
        out.append(int(hit[0]) if hit.size else len(row))
    return out


class Synthesizer(object):
    def load(self, checkpoint_path, hparams, gta=False, model_name="Tacotron"):
        self._hparams, self.gta = hparams, gta
        self.model = create_model(model_name, hparams)
        variables, _ = t2_checkpoint.load(checkpoint_path)
        self.model.load_variables(variables)
        self._pad = 0
        self._target_pad = -hparams.max_abs_value if hparams.symmetric_mels else 0.0

    def synthesize(self, texts, basenames, out_dir, log_dir, mel_filenames):
        hp = self._hparams
        cleaners = [x.strip() for x in hp.cleaners.split(",")]
        seqs = [np.asarray(text_to_sequence(t, cleaners), dtype=np.int32) for t in texts]
        lens = torch.tensor([len(s) for s in seqs], dtype=torch.int32).cuda()
        T_in = max(len(s) for s in seqs)
        inputs = torch.from_numpy(np.stack([pad_input(s, T_in, self._pad) for s in seqs])).cuda()
        if self.gta:
            targets = [np.load(f) for f in mel_filenames]
            target_lengths = [len(t) for t in targets]
            T_out = max(target_lengths)
            mel = torch.from_numpy(np.stack([pad_target(t, T_out, self._target_pad) for t in targets]).astype(np.float32)).cuda()
            self.model.initialize(inputs, lens, mel, gta=True)
            mels = self.model.tower_mel_outputs[0].cpu().numpy()
            mels = [m[:n] for m, n in zip(mels, target_lengths)]          # take off the batch-wise padding (synthesizer.py:167)
        else:
            self.model.initialize(inputs, lens)
            mels = self.model.tower_mel_outputs[0].cpu().numpy()
            stop = self.model.tower_stop_token_prediction[0].cpu().numpy()
            cut = get_output_lengths(stop)
            mels = [m[:n] for m, n in zip(mels, cut)]
            if hp.predict_linear:       # post-processing net (tacotron.py:203-219): linear spectrograms of the same frames
                linears = [l[:n] for l, n in zip(self.model.tower_linear_outputs[0].cpu().numpy(), cut)]
        lo = -hp.max_abs_value if hp.symmetric_mels else 0.0          # T2_output_range (synthesizer.py:78,157,160): no decay margin here
        mels = [np.clip(m, lo, hp.max_abs_value) for m in mels]
        if hp.predict_linear and not self.gta:
            linears = [np.clip(l, lo, hp.max_abs_value) for l in linears]
        if basenames is None:
            # live mode (synthesizer.py:162-182): Griffin-Lim of the first utterance into temp.wav, then the platform's player - `aplay`
            # on Linux, skipped when the machine has no player / audio device
            import shutil
            if len(mels[0]) < 2:
                return None
            audio.save_wav(audio.inv_mel_spectrogram(mels[0].T, hp), "temp.wav", hp.sample_rate)
            if shutil.which("aplay"):
                os.system("aplay temp.wav")
            return None
        names = []
        for m, b in zip(mels, basenames):
            path = os.path.join(out_dir, "mel-%s.npy" % b)
            np.save(path, m.astype(np.float32), allow_pickle=False)
            names.append(path)
        if log_dir is not None and not self.gta:
            # evaluation artefacts of the reference (synthesizer.py:199-224): Griffin-Lim inversions of the mel and, with the
            # post-processing net, of the linear spectrogram (GPU Griffin-Lim, datasets/audio.py)
            os.makedirs(os.path.join(log_dir, "wavs"), exist_ok=True)
            os.makedirs(os.path.join(log_dir, "plots"), exist_ok=True)
            alignments = self.model.tower_alignments[0].cpu().numpy()                 # [B, T_in, T_out]
            for i, (m, b) in enumerate(zip(mels, basenames)):
                if hp.predict_linear:
                    np.save(os.path.join(out_dir, "linear-%s.npy" % b), linears[i].astype(np.float32), allow_pickle=False)
                if len(m) < 2:          # nothing to invert or draw (an untrained model can fire its stop token on the first frames)
                    continue
                plot.plot_alignment(alignments[i], os.path.join(log_dir, "plots", "alignment-%s.png" % b), title=texts[i], split_title=True,
                                    max_len=len(m))
                plot.plot_spectrogram(m, os.path.join(log_dir, "plots", "mel-%s.png" % b), title=texts[i], split_title=True)
                if hp.predict_linear:
                    plot.plot_spectrogram(linears[i], os.path.join(log_dir, "plots", "linear-%s.png" % b), title=texts[i], split_title=True,
                                          auto_aspect=True)
                audio.save_wav(audio.inv_mel_spectrogram(m.T, hp), os.path.join(log_dir, "wavs", "wav-%s-mel.wav" % b), hp.sample_rate)
                if hp.predict_linear:
                    audio.save_wav(audio.inv_linear_spectrogram(linears[i].T, hp), os.path.join(log_dir, "wavs", "wav-%s-linear.wav" % b), hp.sample_rate)
        return names, ["<no_g>"] * len(names)
