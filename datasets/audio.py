"""Drop-in for the reference's datasets/audio.py: same free functions taking (array, hparams), numpy in / numpy out,
computed by the sm_100a kernels of libt2b200.so (host<->device copies inside each call). No CPU fallback: without
a CUDA device or the built library these functions raise.

Covered (reference datasets/audio.py line numbers): preemphasis :22-25, get_hop_size :54-59, linearspectrogram
:61-68, melspectrogram :70-77, librosa_pad_lr :210-219; plus the mu-law family of wavenet_vocoder/util.py.
Griffin-Lim / LWS inversion, wav IO and trim_silence are outside the hot path.
"""
import numpy as np
import torch

from t2_import import t2

_front_ends = {}


def get_hop_size(hparams):
    hop_size = hparams.hop_size
    if hop_size is None:
        assert hparams.frame_shift_ms is not None
        hop_size = int(hparams.frame_shift_ms / 1000 * hparams.sample_rate)
    return hop_size


def _fe(hparams):
    key = (hparams.sample_rate, hparams.n_fft, get_hop_size(hparams), hparams.win_size, hparams.num_mels, hparams.fmin,
           hparams.fmax, hparams.magnitude_power, hparams.min_level_db, hparams.ref_level_db, hparams.max_abs_value,
           hparams.symmetric_mels, hparams.allow_clipping_in_normalization, hparams.signal_normalization)
    if key not in _front_ends:
        _front_ends[key] = t2.audio.MelFrontEnd(hparams)
    return _front_ends[key]


def _to_dev(wav):
    w = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
    return w.reshape(1, -1).cuda() if w.dim() == 1 else w.cuda()


def preemphasis(wav, k, preemphasize=True):
    if not preemphasize:
        return wav
    return t2.audio.preemphasis(_to_dev(wav), k)[0].cpu().numpy()


def melspectrogram(wav, hparams):
    """wav: 1-D float array -> [num_mels, frames] float32 (audio.py:70-77)."""
    return _fe(hparams)(_to_dev(wav), time_major=False)[0].cpu().numpy()


def linearspectrogram(wav, hparams):
    """wav: 1-D float array -> [n_fft/2+1, frames] float32 (audio.py:61-68)."""
    _, lin = _fe(hparams)(_to_dev(wav), time_major=False, linear=True)
    return lin[0].cpu().numpy()


def melspectrogram_batch(wavs, hparams, preemphasis_coef=0.0, gain=1.0):
    """Batched variant used by the preprocessor: wavs [B, n] -> [B, frames, num_mels] (the on-disk layout)."""
    return _fe(hparams)(_to_dev(wavs), preemphasis=preemphasis_coef, gain=gain, time_major=True).cpu().numpy()


def librosa_pad_lr(x, fsize, fshift, pad_sides=1):
    assert pad_sides in (1, 2)
    pad = (x.shape[0] // fshift + 1) * fshift - x.shape[0]
    if pad_sides == 1:
        return 0, pad
    return pad // 2, pad // 2 + pad % 2


def mulaw_quantize(x, mu=256):
    return t2.audio.mulaw_quantize(_to_dev(x).reshape(-1)).cpu().numpy().reshape(np.shape(x)).astype(np.int64)


def inv_mulaw_quantize(y, mu=256):
    q = torch.from_numpy(np.ascontiguousarray(y, dtype=np.int32)).reshape(-1).cuda()
    return t2.audio.inv_mulaw_quantize(q).cpu().numpy().reshape(np.shape(y))


def mulaw(x, mu=256):
    return t2.audio.mulaw(_to_dev(x).reshape(-1)).cpu().numpy().reshape(np.shape(x))


def inv_mulaw(y, mu=256):
    return t2.audio.inv_mulaw(_to_dev(y).reshape(-1)).cpu().numpy().reshape(np.shape(y))
