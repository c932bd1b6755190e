"""Fast-WaveNet AR synthesis kernel vs the oracle's incremental forward.

Teacher forcing (the reference's wavenet_synth_debug path, wavenet.py:876-878) makes the network outputs a
deterministic function of the inputs, so the raw outputs are compared directly (bf16 weights vs fp32 oracle:
max abs err <= 4e-2); the sampling functions are checked by re-sampling the CUDA raw outputs with the oracle's
restatement of mixture.py:76-107 using the SAME injected uniforms."""
import math

import numpy as np
import pytest
import torch

from hparams import hparams
from oracle import wavenet as ow
from t2_import import t2

pytestmark = pytest.mark.gpu


def _hp(**kw):
    hp = hparams.copy()
    hp.parse("layers=6,stacks=2,residual_channels=128,gate_channels=256,skip_out_channels=128,"
             "upsample_scales=[4,4],hop_size=16,wavenet_dropout=0.0")
    for k, v in kw.items():
        hp.set_hparam(k, v)
    return hp


@pytest.mark.parametrize("cs", [1, 8, 16])
def test_ar_teacher_forced_mulaw(cs):
    hp = _hp(input_type="mulaw-quantize", quantize_channels=256, out_channels=256)
    B, T = 3, 48
    g = torch.Generator().manual_seed(21)
    params = ow.init_params(hp, seed=21, random_bias=True)
    idx = torch.randint(90, 166, (B, T), generator=g)
    c = torch.rand(B, 80, T // 16, generator=g)
    x = torch.nn.functional.one_hot(idx, 256).float().transpose(1, 2)
    y_par = ow.step(x, c, params, hp).transpose(1, 2)            # [B, T, 256]; == incremental (tests/test_oracle_wavenet.py)
    syn = t2.wavenet.WaveNetSynthesizer(hp, B, T, cluster_size=cs)
    syn.load_params(params)
    # the kernel feeds `initial` at t = 0 and test_inputs[t] as the input of step t + 1
    initial = idx[:, 0].int().cuda()
    ti = torch.cat([idx[:, 1:], idx[:, -1:]], dim=1).int().cuda()
    u = torch.rand(B, T, generator=g)
    out, raw = syn.generate(c.cuda(), initial, test_inputs=ti, u_a=u.cuda(), return_raw=True)
    torch.cuda.synchronize()
    err = (raw.cpu() - y_par).abs()
    assert err.max().item() < 4e-2 and err.mean().item() < 6e-3, (err.max().item(), err.mean().item())
    # categorical sampling by inverse CDF on the CUDA logits
    p = torch.softmax(raw.cpu().double(), -1)
    cdf = p.cumsum(-1)
    ref = (cdf < u.double().unsqueeze(-1)).sum(-1).clamp(max=255)
    agree = (out.cpu().long() == ref).float().mean().item()
    assert agree > 0.97, agree          # disagreements only where u falls within fp32 rounding of a CDF step


@pytest.mark.parametrize("cs", [1, 8])
def test_ar_teacher_forced_mol_raw(cs):
    hp = _hp(input_type="raw", out_channels=30, legacy=False, residual_legacy=False, upsample_type="2D")
    B, T = 2, 40
    g = torch.Generator().manual_seed(22)
    params = ow.init_params(hp, seed=22, random_bias=True)
    w = (torch.rand(B, T, generator=g) * 2 - 1) * 0.8
    c = torch.rand(B, 80, -(-T // 16), generator=g)[:, :, :T // 16 + (1 if T % 16 else 0)]
    Tc = c.shape[-1]
    T = Tc * 16
    w = torch.cat([w, torch.zeros(B, T - w.shape[1])], dim=1)
    y_par = ow.step(w.unsqueeze(1), c, params, hp).transpose(1, 2)   # [B, T, 30]
    syn = t2.wavenet.WaveNetSynthesizer(hp, B, T, cluster_size=cs)
    syn.load_params(params)
    ti = torch.cat([w[:, 1:], w[:, -1:]], dim=1).cuda()
    u_mix = torch.rand(B, T, 10, generator=g).clamp(1e-5, 1 - 1e-5)
    u_log = torch.rand(B, T, generator=g).clamp(1e-5, 1 - 1e-5)
    out, raw = syn.generate(c.cuda(), w[:, 0].contiguous().cuda(), test_inputs=ti, u_a=u_mix.cuda(), u_b=u_log.cuda(),
                            return_raw=True)
    torch.cuda.synchronize()
    err = (raw.cpu() - y_par).abs()
    assert err.max().item() < 4e-2 and err.mean().item() < 6e-3, (err.max().item(), err.mean().item())
    ref = ow.sample_from_discretized_mix_logistic(raw.cpu().transpose(1, 2), hp.log_scale_min, u_mix, u_log)
    assert (out.cpu() - ref).abs().max().item() < 1e-4


def test_ar_teacher_forced_gaussian_raw():
    """out_channels = 2: x = mean + exp(max(log_scale, min)) * n clipped to [-1, 1] (gaussian.py:39-52), normal draws injected"""
    hp = _hp(input_type="raw", out_channels=2)
    B, T = 2, 48
    g = torch.Generator().manual_seed(24)
    params = ow.init_params(hp, seed=24, random_bias=True)
    w = (torch.rand(B, T, generator=g) * 2 - 1) * 0.8
    c = torch.rand(B, 80, T // 16, generator=g)
    y_par = ow.step(w.unsqueeze(1), c, params, hp).transpose(1, 2)   # [B, T, 2]
    syn = t2.wavenet.WaveNetSynthesizer(hp, B, T, cluster_size=8)
    syn.load_params(params)
    ti = torch.cat([w[:, 1:], w[:, -1:]], dim=1).cuda()
    n = torch.randn(B, T, generator=g)
    out, raw = syn.generate(c.cuda(), w[:, 0].contiguous().cuda(), test_inputs=ti, u_b=n.cuda(), return_raw=True)
    torch.cuda.synchronize()
    err = (raw.cpu() - y_par).abs()
    assert err.max().item() < 4e-2 and err.mean().item() < 6e-3, (err.max().item(), err.mean().item())
    ref = ow.sample_from_gaussian(raw.cpu().transpose(1, 2), hp.log_scale_min_gauss, n)
    assert (out.cpu() - ref).abs().max().item() < 1e-4
    free = syn.generate(c.cuda(), w[:, 0].contiguous().cuda(), seed=3).cpu()      # Box-Muller draws from the counter hash
    assert free.min() >= -1 and free.max() <= 1 and free.std() > 0


def test_ar_free_running_is_deterministic_and_in_range():
    hp = _hp(input_type="mulaw-quantize", quantize_channels=256, out_channels=256)
    B, T = 5, 64
    syn = t2.wavenet.WaveNetSynthesizer(hp, B, T, cluster_size=8)
    syn.load_params(ow.init_params(hp, seed=23))
    c = torch.rand(B, 80, T // 16).cuda()
    init = torch.full((B,), 127, dtype=torch.int32, device="cuda")       # mulaw_quantize(0) start token
    a = syn.generate(c, init, seed=5).cpu()
    b = syn.generate(c, init, seed=5).cpu()
    d = syn.generate(c, init, seed=6).cpu()
    assert torch.equal(a, b) and not torch.equal(a, d)
    assert a.min() >= 0 and a.max() <= 255
