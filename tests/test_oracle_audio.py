"""Pins the audio oracle against independent implementations (torch.stft, torchaudio Slaney fbanks) and the
invariants the reference implies."""
import numpy as np
import pytest
import torch

from hparams import hparams
from oracle import audio


def _wav(seed=1, n=22050):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 22050.0
    w = 0.5 * np.sin(2 * np.pi * (200 + 2000 * t) * t) + rng.normal(0, 0.05, n)
    return (w / np.abs(w).max() * 0.999).astype(np.float32)


def test_mulaw_known_answers():
    assert audio.mulaw_quantize(0.0) == 127                      # util.py:99-102, audio.py:35,38
    assert audio.mulaw_quantize(np.float32(1.0)) == 255
    assert audio.mulaw_quantize(np.float32(-1.0)) == 0
    x = np.linspace(-1, 1, 4001)
    assert np.abs(audio.inv_mulaw(audio.mulaw(x)) - x).max() < 1e-12
    q = audio.mulaw_quantize(x.astype(np.float32))
    assert q.min() == 0 and q.max() == 255 and np.all(np.diff(q) >= 0)
    r = audio.inv_mulaw_quantize(q)
    assert np.abs(audio.mulaw_quantize(r.astype(np.float32)) - q).max() <= 1


def test_stft_matches_torch_stft():
    w = _wav()
    D = audio.stft(w, hparams)
    assert D.shape == (1025, len(w) // 275 + 1)                  # frames = len//hop + 1
    win = torch.from_numpy(audio.hann_window_padded(1100, 2048))
    T = torch.stft(torch.from_numpy(w).double(), 2048, 275, 2048, window=win, center=True, pad_mode="constant",
                   return_complex=True)
    assert np.abs(T.numpy() - D).max() < 1e-3 * np.abs(D).max()


def test_mel_basis_matches_torchaudio_slaney():
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(1025, 55., 7600., 80, 22050, norm="slaney", mel_scale="slaney")
    ours = audio.build_mel_basis(hparams)
    assert ours.shape == (80, 1025)
    assert np.abs(fb.numpy().T - ours).max() < 1e-6


def test_melspectrogram_shape_range_and_denormalize():
    w = audio.preemphasis(_wav(), hparams.preemphasis)
    m = audio.melspectrogram(w, hparams)
    assert m.shape == (80, 81)
    assert m.min() >= -4.0 and m.max() <= 4.0
    S = audio._amp_to_db(audio._linear_to_mel(np.abs(audio.stft(w, hparams)) ** 2, hparams), hparams) - 20
    assert np.allclose(audio._denormalize(audio._normalize(S, hparams), hparams), np.clip(S, -100, 0))


def test_pad_lr_alignment():
    x = np.zeros(22050)
    l, r = audio.librosa_pad_lr(x, 2048, 275)
    assert (len(x) + l + r) == (len(x) // 275 + 1) * 275


def test_mulaw_properties():
    """size-independent properties of the companding family (wavenet_vocoder/util.py:30-129): odd symmetry, monotone,
    indices cover [0, 255], expand(compress(x)) == x, quantise(expand(q)) within one bin (truncation)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from hypothesis.extra import numpy as hnp

    @settings(max_examples=60, deadline=None)
    @given(hnp.arrays(np.float32, hnp.array_shapes(min_dims=1, max_dims=1, min_side=1, max_side=257),
                      elements=st.floats(-1, 1, width=32)))
    def check(x):
        y = audio.mulaw(x)
        assert np.all(np.abs(y) <= 1.0 + 1e-6)
        assert np.allclose(audio.mulaw(-x), -y, atol=1e-7)
        assert np.allclose(audio.inv_mulaw(y), x, atol=2e-6)
        q = audio.mulaw_quantize(x)
        assert q.min() >= 0 and q.max() <= 255
        xs = np.sort(x)
        assert np.all(np.diff(audio.mulaw_quantize(xs)) >= 0)                  # monotone
        back = audio.mulaw_quantize(audio.inv_mulaw_quantize(q).astype(np.float32))
        assert np.abs(back - q).max() <= 1
    check()
    assert audio.mulaw_quantize(np.zeros(0, np.float32)).shape == (0,)          # empty input


def test_istft_inverts_stft_and_griffin_lim_reduces_the_spectral_error():
    from oracle import audio as oa
    """librosa.istft semantics of the oracle (datasets/audio.py:184-186): exact inverse of the STFT where the window sum is non-zero,
    output length hop * (frames - 1); Griffin-Lim (audio.py:151-161) lowers | |STFT(y)| - S | from its random-phase start."""
    from hparams import hparams
    rng = np.random.default_rng(3)
    hop = oa.get_hop_size(hparams)
    y = (0.4 * np.sin(np.arange(hop * 30) * 0.07) + 0.05 * rng.standard_normal(hop * 30)).astype(np.float32)
    D = oa.stft(y, hparams)
    back = oa.istft(D, hparams)
    assert back.shape == (hop * (D.shape[1] - 1),) == y.shape
    assert np.abs(back - y).max() < 1e-5
    S = np.abs(D).astype(np.float64)
    ang = np.exp(2j * np.pi * rng.random(S.shape))
    err = [np.abs(np.abs(oa.stft(oa.griffin_lim(S, hparams, ang, iters=n), hparams)) - S).mean() for n in (0, 8)]
    assert err[1] < 0.6 * err[0]
