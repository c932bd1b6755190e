"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's audio front-end.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Restates datasets/audio.py of the reference with numpy/scipy only (librosa is not installable here):
  preemphasis            datasets/audio.py:22-25
  _stft                  datasets/audio.py:178-182  (librosa.stft, center=True, pad_mode='constant')
  _build_mel_basis       datasets/audio.py:243-246  (librosa.filters.mel: Slaney scale, area normalised)
  _amp_to_db/_normalize  datasets/audio.py:248-270
  melspectrogram         datasets/audio.py:70-77
  linearspectrogram      datasets/audio.py:61-68
  librosa_pad_lr         datasets/audio.py:210-219
  inv_linear_spectrogram / inv_mel_spectrogram / _griffin_lim / _istft   datasets/audio.py:97-133,151-161,184-186
and wavenet_vocoder/util.py:30-129 (mu-law family, mu hard-wired to 255).

PINNING. The reference ships no tests or golden vectors and librosa cannot be imported in the build container. Pinned by EXECUTING
the reference's own source (tests/golden/make_reference_vectors.py -> reference_exec.npz, checked by tests/test_reference_pinned.py):
the mu-law family (numpy and tensor code paths, every quantiser bin edge), preemphasis / inv_preemphasis, _amp_to_db / _db_to_amp,
_normalize / _denormalize (4 flag combinations), the padding helpers, start_and_end_indices, and the COMPOSITIONS melspectrogram /
linearspectrogram / inv_linear_spectrogram / inv_mel_spectrogram / _griffin_lim. PARITY UNPINNED for the three librosa primitives
those compositions call (librosa.stft, librosa.istft, librosa.filters.mel): they are restatements of librosa's documented behaviour,
cross-checked against independent implementations (torch.stft, torchaudio's Slaney filterbank: tests/test_oracle_audio.py).
"""
import numpy as np
from scipy import signal


def get_hop_size(hparams):  # datasets/audio.py:54-59
    hop_size = hparams.hop_size
    if hop_size is None:
        assert hparams.frame_shift_ms is not None
        hop_size = int(hparams.frame_shift_ms / 1000 * hparams.sample_rate)
    return hop_size


def preemphasis(wav, k, preemphasize=True):
    if preemphasize:
        return signal.lfilter([1, -k], [1], wav)
    return wav


def inv_preemphasis(wav, k, inv_preemphasize=True):
    if inv_preemphasize:
        return signal.lfilter([1], [1, -k], wav)
    return wav


def hann_window_padded(win_length, n_fft):
    """scipy periodic Hann of win_length, zero-padded (centred) to n_fft — librosa.util.pad_center."""
    w = signal.get_window("hann", win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[lpad:lpad + win_length] = w
    return out


def stft(y, hparams):
    """librosa.stft(y, n_fft, hop_length, win_length, pad_mode='constant') -> complex64 [1 + n_fft/2, frames]."""
    n_fft, hop, win = hparams.n_fft, get_hop_size(hparams), hparams.win_size
    y = np.asarray(y)
    fft_window = hann_window_padded(win, n_fft)
    ypad = np.pad(y, n_fft // 2, mode="constant")
    n_frames = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = ypad[idx] * fft_window[None, :]
    return np.fft.rfft(frames, axis=1).T.astype(np.complex64)


def istft(D, hparams):
    """librosa.istft(D, hop_length, win_length) as called by datasets/audio.py:184-186: per frame irfft -> synthesis window (the
    analysis window) -> overlap-add -> division by the window sum of squares where it is not tiny -> centre trim of n_fft / 2."""
    n_fft, hop = hparams.n_fft, get_hop_size(hparams)
    w = hann_window_padded(hparams.win_size, n_fft)
    n_frames = D.shape[1]
    y = np.zeros(n_fft + hop * (n_frames - 1), dtype=np.float64)
    wss = np.zeros_like(y)
    for k in range(n_frames):
        y[k * hop:k * hop + n_fft] += w * np.fft.irfft(D[:, k], n_fft)
        wss[k * hop:k * hop + n_fft] += w * w
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    return y[n_fft // 2:len(y) - n_fft // 2].astype(np.float32)


def griffin_lim(S, hparams, angles0, iters=None):
    """datasets/audio.py:151-161 with the initial phases given (the reference draws them with np.random.rand)"""
    S_complex = np.abs(S).astype(np.complex128)
    y = istft(S_complex * angles0, hparams)
    for _ in range(hparams.griffin_lim_iters if iters is None else iters):
        angles = np.exp(1j * np.angle(stft(y, hparams)))
        y = istft(S_complex * angles, hparams)
    return y


def _db_to_amp(x):  # datasets/audio.py:252-253
    return np.power(10.0, np.asarray(x) * 0.05)


def _spectrogram_to_wav(S, hparams, angles0, iters):
    """tail of inv_*_spectrogram (audio.py:126-133 without LWS): Griffin-Lim on S ** power, then the inverse pre-emphasis"""
    return inv_preemphasis(griffin_lim(S ** hparams.power, hparams, angles0, iters), hparams.preemphasis, hparams.preemphasize)


def inv_linear_spectrogram(linear_spectrogram, hparams, angles0, iters=None):
    """datasets/audio.py:118-133 with the initial Griffin-Lim phases given"""
    D = _denormalize(linear_spectrogram, hparams) if hparams.signal_normalization else linear_spectrogram
    return _spectrogram_to_wav(_db_to_amp(D + hparams.ref_level_db) ** (1 / hparams.magnitude_power), hparams, angles0, iters)


def inv_mel_spectrogram(mel_spectrogram, hparams, angles0, iters=None):
    """datasets/audio.py:97-112: the mel spectrogram goes back to linear through the pseudo-inverse of the filterbank, floored at
    1e-10 (audio.py:231-241)"""
    D = _denormalize(mel_spectrogram, hparams) if hparams.signal_normalization else mel_spectrogram
    S = np.maximum(1e-10, np.dot(np.linalg.pinv(build_mel_basis(hparams)), _db_to_amp(D + hparams.ref_level_db) ** (1 / hparams.magnitude_power)))
    return _spectrogram_to_wav(S, hparams, angles0, iters)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def build_mel_basis(hparams):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax): Slaney mel scale, triangular, area normalised."""
    assert hparams.fmax <= hparams.sample_rate // 2
    sr, n_fft, n_mels = hparams.sample_rate, hparams.n_fft, hparams.num_mels
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(hparams.fmin), _hz_to_mel(hparams.fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights


_mel_basis_cache = {}


def _linear_to_mel(S, hparams):
    key = (hparams.sample_rate, hparams.n_fft, hparams.num_mels, hparams.fmin, hparams.fmax)
    if key not in _mel_basis_cache:
        _mel_basis_cache[key] = build_mel_basis(hparams)
    return np.dot(_mel_basis_cache[key], S)


def _amp_to_db(x, hparams):
    min_level = np.exp(hparams.min_level_db / 20 * np.log(10))
    return 20 * np.log10(np.maximum(min_level, x))


def _normalize(S, hparams):
    if hparams.allow_clipping_in_normalization:
        if hparams.symmetric_mels:
            return np.clip((2 * hparams.max_abs_value) * ((S - hparams.min_level_db) / (-hparams.min_level_db))
                           - hparams.max_abs_value, -hparams.max_abs_value, hparams.max_abs_value)
        return np.clip(hparams.max_abs_value * ((S - hparams.min_level_db) / (-hparams.min_level_db)), 0,
                       hparams.max_abs_value)
    assert S.max() <= 0 and S.min() - hparams.min_level_db >= 0
    if hparams.symmetric_mels:
        return (2 * hparams.max_abs_value) * ((S - hparams.min_level_db) / (-hparams.min_level_db)) - hparams.max_abs_value
    return hparams.max_abs_value * ((S - hparams.min_level_db) / (-hparams.min_level_db))


def _denormalize(D, hparams):
    if hparams.allow_clipping_in_normalization:
        if hparams.symmetric_mels:
            return (((np.clip(D, -hparams.max_abs_value, hparams.max_abs_value) + hparams.max_abs_value)
                     * -hparams.min_level_db / (2 * hparams.max_abs_value)) + hparams.min_level_db)
        return (np.clip(D, 0, hparams.max_abs_value) * -hparams.min_level_db / hparams.max_abs_value) + hparams.min_level_db
    if hparams.symmetric_mels:
        return ((D + hparams.max_abs_value) * -hparams.min_level_db / (2 * hparams.max_abs_value)) + hparams.min_level_db
    return (D * -hparams.min_level_db / hparams.max_abs_value) + hparams.min_level_db


def linearspectrogram(wav, hparams):
    D = stft(wav, hparams)
    S = _amp_to_db(np.abs(D) ** hparams.magnitude_power, hparams) - hparams.ref_level_db
    return _normalize(S, hparams) if hparams.signal_normalization else S


def melspectrogram(wav, hparams):
    D = stft(wav, hparams)
    S = _amp_to_db(_linear_to_mel(np.abs(D) ** hparams.magnitude_power, hparams), hparams) - hparams.ref_level_db
    return _normalize(S, hparams) if hparams.signal_normalization else S


def librosa_pad_lr(x, fsize, fshift, pad_sides=1):
    assert pad_sides in (1, 2)
    pad = (x.shape[0] // fshift + 1) * fshift - x.shape[0]
    if pad_sides == 1:
        return 0, pad
    return pad // 2, pad // 2 + pad % 2


# ---- wavenet_vocoder/util.py:30-129 (mu is forced to 255 whatever the argument) ---------------------------
# dtype semantics: the reference feeds float32 wavs (librosa.load) through numpy-1.14 value-based casting, i.e.
# every step runs in float32; float64 input stays float64. numpy's float32 log1p is platform dependent (glibc /
# SVML, <= a few ulp), so the float32 path here is DEFINED as "log1p evaluated in float64 then rounded to float32"
# (= a correctly rounded log1pf); every other step is a plain IEEE float32 op. The CUDA kernel follows the same
# definition, which makes mu-law indices bit-reproducible across machines.
def _log1p_like(x):
    if x.dtype == np.float32:
        return np.log1p(x.astype(np.float64)).astype(np.float32)
    return np.log1p(x)


def mulaw(x, mu=256):
    mu = 255
    x = np.asarray(x)
    if x.dtype != np.float32:
        x = x.astype(np.float64)
    dt = x.dtype.type
    return np.sign(x) * _log1p_like(dt(mu) * np.abs(x)) / _log1p_like(np.array(mu, dtype=dt))


def inv_mulaw(y, mu=256):
    mu = 255
    y = np.asarray(y)
    if y.dtype != np.float32:
        y = y.astype(np.float64)
    dt = y.dtype.type
    if y.dtype == np.float32:  # pow evaluated in float64 and rounded (platform-independent definition, see above)
        p = np.power(np.float64(1.0 + mu), np.abs(y).astype(np.float64)).astype(np.float32)
    else:
        p = np.power(dt(1.0 + mu), np.abs(y))
    return np.sign(y) * dt(1.0 / mu) * (p - dt(1.0))


def mulaw_quantize(x, mu=256):
    mu = 255
    y = mulaw(x, mu)
    dt = y.dtype.type
    return ((y + dt(1)) / dt(2) * dt(mu)).astype(np.int64)  # astype(int) truncates toward zero (util.py:99-102,156)


def inv_mulaw_quantize(y, mu=256):
    mu = 255
    y = np.float32(2) * np.asarray(y).astype(np.float32) / np.float32(mu) - np.float32(1)
    return inv_mulaw(y, mu)


def start_and_end_indices(quantized, silence_threshold=2):  # datasets/audio.py:33-44
    for start in range(quantized.size):
        if abs(quantized[start] - 127) > silence_threshold:
            break
    for end in range(quantized.size - 1, 1, -1):
        if abs(quantized[end] - 127) > silence_threshold:
            break
    assert abs(quantized[start] - 127) > silence_threshold
    assert abs(quantized[end] - 127) > silence_threshold
    return start, end
