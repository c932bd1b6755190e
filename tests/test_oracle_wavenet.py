"""Pins the WaveNet oracle against the only known answers / invariants the reference offers."""
import numpy as np
import torch

from hparams import hparams, paper_hparams
from oracle import wavenet as ow


def small_hp(**kw):
    hp = hparams.copy()
    hp.parse("layers=4,stacks=2,residual_channels=16,gate_channels=32,skip_out_channels=16,"
             "upsample_scales=[2,3],hop_size=6,cin_channels=8,num_mels=8")
    for k, v in kw.items():
        hp.set_hparam(k, v)
    return hp


def test_receptive_field_known_answers():
    assert ow.receptive_field_size(24, 4, 3) == 505      # wavenet.py:54-71 with paper_hparams
    assert ow.receptive_field_size(20, 2, 3) == 4093     # hparams.py defaults


def test_param_counts_match_survey():
    hp = paper_hparams()
    hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256")
    n = sum(int(np.prod(s)) for k, s in ow.param_shapes(hp).items() if not k.startswith("local_cond"))
    assert abs(n / 1e6 - 13.80) < 0.02
    hp2 = paper_hparams()
    n2 = sum(int(np.prod(s)) for k, s in ow.param_shapes(hp2).items() if not k.startswith("local_cond"))
    assert abs(n2 / 1e6 - 13.68) < 0.02


def test_nn_init_is_nearest_neighbour_repeat():
    for ut in ("SubPixel", "2D"):
        hp = small_hp(upsample_type=ut)
        p = ow.init_params(hp)
        c = torch.rand(2, 8, 5)
        up = ow.upsample(c, p, hp)
        assert up.shape == (2, 8, 30)
        ref = c.repeat_interleave(6, dim=-1) * hp.NN_scaler
        assert torch.allclose(up, ref, atol=1e-6)


def test_incremental_matches_parallel_mulaw_quantize():
    hp = small_hp(input_type="mulaw-quantize", quantize_channels=16, out_channels=16)
    torch.manual_seed(0)
    p = ow.init_params(hp, seed=1, random_bias=True)
    B, Tc = 2, 4
    T = Tc * 6
    idx = torch.randint(0, 16, (B, T))
    x = torch.nn.functional.one_hot(idx, 16).float().transpose(1, 2)
    c = torch.rand(B, 8, Tc)
    y_par = ow.step(x, c, p, hp)                       # [B, 16, T]
    # teacher-forced incremental: the input at step t is x[:, :, t]
    init = x[:, :, 0].unsqueeze(1)
    test_inputs = x.transpose(1, 2)[:, 1:, :]
    test_inputs = torch.cat([test_inputs, test_inputs[:, -1:, :]], dim=1)
    _, raw = ow.incremental(init, c, p, hp, T, test_inputs=test_inputs, u_cat=torch.rand(B, T))
    assert torch.allclose(raw.transpose(1, 2), y_par, atol=2e-5)


def test_incremental_matches_parallel_raw_mol():
    hp = small_hp(input_type="raw", out_channels=30, legacy=False, residual_legacy=False)
    p = ow.init_params(hp, seed=2, random_bias=True)
    B, Tc = 2, 3
    T = Tc * 6
    x = torch.rand(B, 1, T) * 2 - 1
    c = torch.rand(B, 8, Tc)
    y_par = ow.step(x, c, p, hp)
    init = x[:, :, 0].unsqueeze(1)
    ti = torch.cat([x.transpose(1, 2)[:, 1:, :], x.transpose(1, 2)[:, -1:, :]], dim=1)
    _, raw = ow.incremental(init, c, p, hp, T, test_inputs=ti)
    assert torch.allclose(raw.transpose(1, 2), y_par, atol=2e-5)


def test_losses_finite_and_sane():
    hp = small_hp(input_type="mulaw-quantize", quantize_channels=16, out_channels=16)
    p = ow.init_params(hp, seed=3)
    idx = torch.randint(0, 16, (2, 24))
    x = torch.nn.functional.one_hot(idx, 16).float().transpose(1, 2)
    c = torch.rand(2, 8, 4)
    loss, grads, _ = ow.train_step(p, x, c, idx, torch.tensor([24, 17]), hp)
    assert abs(loss.item() - np.log(16)) < 0.5
    assert all(torch.isfinite(g).all() for g in grads.values())
    hp2 = small_hp(input_type="raw", out_channels=30)
    p2 = ow.init_params(hp2, seed=4)
    y = torch.rand(2, 24) * 2 - 1
    loss2, grads2, _ = ow.train_step(p2, y.unsqueeze(1), c, y, torch.tensor([24, 20]), hp2)
    assert np.isfinite(loss2.item())
    assert all(torch.isfinite(g).all() for g in grads2.values())
