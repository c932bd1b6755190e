"""GPU audio front-end (through the C-ABI) vs the numpy oracle.

Tolerances: mel / linear spectrograms within 1e-3 absolute in the normalised [-4, 4] domain (north-star);
mu-law indices bit-exact."""
import numpy as np
import pytest
import torch

from hparams import hparams
from oracle import audio as oa
from t2_import import t2

pytestmark = pytest.mark.gpu


def _wav(seed, n=22050):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 22050.0
    w = 0.5 * np.sin(2 * np.pi * (200 + 2000 * t) * t) + rng.normal(0, 0.05, n)
    return (w / np.abs(w).max() * 0.999).astype(np.float32)


def test_mulaw_quantize_bit_exact():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 4_000_000), np.linspace(-1, 1, 1_000_001), [0.0, 1.0, -1.0, 1e-8, -1e-8],
                        rng.laplace(0, 0.05, 2_000_000).clip(-1, 1)]).astype(np.float32)
    q_ref = oa.mulaw_quantize(x)
    q = t2.audio.mulaw_quantize(torch.from_numpy(x).cuda()).cpu().numpy()
    assert q.dtype == np.int32
    assert np.array_equal(q, q_ref), "mismatches: %d" % int((q != q_ref).sum())
    assert q[4_000_000 + 1_000_001] == 127          # mulaw_quantize(0) == 127 (audio.py:35)
    y_ref = oa.mulaw(x)
    y = t2.audio.mulaw(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(y, y_ref)


def test_inv_mulaw_roundtrip_bit_exact():
    q = np.arange(256, dtype=np.int32)
    x_ref = oa.inv_mulaw_quantize(q)
    x = t2.audio.inv_mulaw_quantize(torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(x, x_ref.astype(np.float32))
    back = t2.audio.mulaw_quantize(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(back - q).max() <= 1               # quantise(expand(q)) may land one bin lower (truncation)


@pytest.mark.parametrize("n", [22050, 5000, 275 * 40, 1100])
def test_melspectrogram_matches_oracle(n):
    fe = t2.audio.MelFrontEnd(hparams)
    wavs = np.stack([oa.preemphasis(_wav(s, n), 0.97).astype(np.float32) for s in range(3)])
    mel, lin = fe(torch.from_numpy(wavs).cuda(), linear=True)
    mel_t = fe(torch.from_numpy(wavs).cuda(), time_major=False)
    torch.cuda.synchronize()
    for i in range(3):
        ref = oa.melspectrogram(wavs[i], hparams)       # [80, frames]
        assert mel.shape[1:] == (n // 275 + 1, 80)
        err = np.abs(mel[i].cpu().numpy().T - ref).max()
        assert err < 1e-3, "mel max abs err %g" % err
        assert np.abs(mel_t[i].cpu().numpy() - ref).max() < 1e-3
        lref = oa.linearspectrogram(wavs[i], hparams)
        lerr = np.abs(lin[i].cpu().numpy().T - lref).max()
        assert lerr < 1e-3, "linear max abs err %g" % lerr


def test_silence_and_full_scale_edges():
    fe = t2.audio.MelFrontEnd(hparams)
    z = torch.zeros(1, 4000, device="cuda")
    assert torch.all(fe(z) == -4.0)                      # floor -> clipped to -max_abs_value
    one = torch.full((1, 4000), 0.999, device="cuda")
    ref = oa.melspectrogram(np.full(4000, 0.999, dtype=np.float32), hparams)
    assert np.abs(fe(one)[0].cpu().numpy().T - ref).max() < 1e-3


def test_fused_preemphasis_and_gain():
    fe = t2.audio.MelFrontEnd(hparams)
    w = _wav(7)
    pre = oa.preemphasis(w, 0.97)
    gain = 0.999 / np.abs(pre).max()
    ref = oa.melspectrogram(pre * gain, hparams)
    out = fe(torch.from_numpy(w[None]).cuda(), preemphasis=0.97, gain=float(gain))[0].cpu().numpy().T
    assert np.abs(out - ref).max() < 1e-3
    p = t2.audio.preemphasis(torch.from_numpy(w[None]).cuda(), 0.97)[0].cpu().numpy()
    assert np.abs(p - pre).max() < 1e-6


def test_dropin_module_surface():
    from datasets import audio
    w = oa.preemphasis(_wav(3), 0.97)
    assert np.abs(audio.melspectrogram(w, hparams) - oa.melspectrogram(w, hparams)).max() < 1e-3
    assert audio.mulaw_quantize(np.float32(0.0)) == 127
    assert audio.librosa_pad_lr(np.zeros(22050), 2048, 275) == oa.librosa_pad_lr(np.zeros(22050), 2048, 275)


def test_griffin_lim_matches_oracle_with_injected_phases():
    """GPU Griffin-Lim (t2_griffin_lim_f32) vs the librosa-semantics oracle from the SAME initial phases: waveform after 0 and 3 rounds;
    reference datasets/audio.py:151-161,184-186. Tolerance: the oracle's STFT matrix is complex64 (librosa), phases of near-silent bins are
    ill-conditioned, so the comparison is relative to the signal level."""
    import numpy as np
    import torch
    from hparams import hparams
    from oracle import audio as oa
    from t2_import import t2
    rng = np.random.default_rng(4)
    hop = oa.get_hop_size(hparams)
    frames = 24
    y0 = (0.4 * np.sin(np.arange(hop * (frames - 1)) * 0.05) + 0.05 * rng.standard_normal(hop * (frames - 1))).astype(np.float32)
    S = np.abs(oa.stft(y0, hparams)).astype(np.float64)              # [bins, frames]
    u = rng.random(S.shape)
    ang = np.exp(2j * np.pi * u)
    fe = t2.audio.MelFrontEnd(hparams)
    mag = torch.from_numpy(np.ascontiguousarray(S.T, dtype=np.float32))[None].cuda()
    for iters, tol in ((0, 2e-4), (3, 5e-3)):
        ref = oa.griffin_lim(S, hparams, ang, iters=iters)
        ph = torch.from_numpy(np.stack([ang.real.T, ang.imag.T], axis=-1).astype(np.float32))[None].contiguous().cuda()
        wav = fe.griffin_lim(mag, iters, phase=ph)[0].cpu().numpy()
        assert wav.shape == ref.shape == (hop * (frames - 1),)
        rel = np.abs(wav - ref).max() / np.abs(ref).max()
        print("griffin-lim iters %d: max err / max |y| = %.3g" % (iters, rel))
        assert rel < tol
    # spectral convergence of the self-seeded path, and the drop-in inversion of a real linear spectrogram
    e = []
    for iters in (0, 30):
        w = fe.griffin_lim(mag, iters, seed=7)[0].cpu().numpy()
        e.append(np.abs(np.abs(oa.stft(w, hparams)) - S).mean())
    assert e[1] < 0.5 * e[0]
    from datasets import audio
    lin = audio.linearspectrogram(y0, hparams)
    wav = audio.inv_linear_spectrogram(lin, hparams)
    assert wav.shape == (hop * (lin.shape[1] - 1),) and np.isfinite(wav).all()
    # (no faithful round trip is expected: the inversion sharpens with S ** hparams.power and undoes a pre-emphasis the input never had)
    mel_wav = audio.inv_mel_spectrogram(audio.melspectrogram(y0, hparams), hparams)
    assert mel_wav.shape == wav.shape and np.isfinite(mel_wav).all()
