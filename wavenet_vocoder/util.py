"""Drop-in for wavenet_vocoder/util.py of the reference: same function names / argument order, the arithmetic runs in
the CUDA kernels of libt2b200 (`t2_mulaw_*`), numpy in / numpy out like the reference's numpy branch (util.py:30-129).
mu is fixed by the kernels to 255 (= quantize_channels - 1 = 256 - 1), the only value the reference's hparams use."""
import numpy as np

from datasets import audio as _audio


def _assert_valid_input_type(s):
    assert s in ("mulaw-quantize", "mulaw", "raw")


def is_mulaw_quantize(s):
    _assert_valid_input_type(s)
    return s == "mulaw-quantize"


def is_mulaw(s):
    _assert_valid_input_type(s)
    return s == "mulaw"


def is_raw(s):
    _assert_valid_input_type(s)
    return s == "raw"


def is_scalar_input(s):
    return is_raw(s) or is_mulaw(s)


mulaw = _audio.mulaw
inv_mulaw = _audio.inv_mulaw
mulaw_quantize = _audio.mulaw_quantize
inv_mulaw_quantize = _audio.inv_mulaw_quantize


def sequence_mask(input_lengths, max_len=None, expand=True):
    """util.py:165-171: float mask [B, T] (or [B, T, 1] when expand)."""
    lengths = np.asarray(input_lengths)
    max_len = int(lengths.max()) if max_len is None else int(max_len)
    m = (np.arange(max_len)[None, :] < lengths[:, None]).astype(np.float32)
    return m[:, :, None] if expand else m


def waveplot(path, y_hat, y_target, hparams, title=None):
    """util.py:174-196"""
    from tacotron.utils import plot
    plot.waveplot(path, y_hat, y_target, hparams, title=title)


def plot_spectrogram(pred_spectrogram, path, title=None, split_title=False, target_spectrogram=None, max_len=None, auto_aspect=False):
    """util.py:198-237 (the same figure as tacotron/utils/plot.py's)"""
    from tacotron.utils import plot
    plot.plot_spectrogram(pred_spectrogram, path, title=title, split_title=split_title, target_spectrogram=target_spectrogram, max_len=max_len,
                          auto_aspect=auto_aspect)
