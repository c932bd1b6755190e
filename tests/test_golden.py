"""Golden vectors (tests/golden/*.npz, frozen by tests/golden/make_golden.py).

CPU: the oracle still reproduces them (guards the checker against drift; the seeded parameter draw is pinned by a checksum).
GPU: the CUDA path reproduces them through the C-ABI with NO oracle arithmetic in the comparison — inputs, expected
outputs and tolerances come from the fixture. Tolerances: mu-law indices bit-exact; mel within 1e-3 (normalised dB domain);
WaveNet / Tacotron losses within 1e-3 / 2e-3 absolute (bf16 GEMM operands vs the fp32 fixture)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

from hparams import hparams  # noqa: E402


def _load(name):
    return np.load(os.path.join(HERE, "golden", name))


# ------------------------------------------------------------------ CPU: oracle vs fixtures
def test_oracle_reproduces_audio_fixture():
    g, now = _load("audio_cfg1.npz"), mg.audio_case()
    assert np.array_equal(g["wav"], now["wav"])
    assert np.array_equal(g["mulaw_q"], now["mulaw_q"])
    assert g["mel"].shape == (80, 81) and np.abs(g["mel"] - now["mel"]).max() < 1e-5
    assert np.abs(g["lin_rows"] - now["lin_rows"]).max() < 1e-5
    assert g["mulaw_q"].min() >= 0 and g["mulaw_q"].max() <= 255


@pytest.mark.parametrize("kind", ["ce", "mol"])
def test_oracle_reproduces_wavenet_fixture(kind):
    g, now = _load("wavenet_%s_tiny.npz" % kind), mg.wavenet_case(kind)
    assert abs(g["param_abs_sum"] - now["param_abs_sum"]) < 1e-6 * g["param_abs_sum"]
    assert abs(g["loss"] - now["loss"]) < 1e-5
    assert np.abs(g["yhat_slice"] - now["yhat_slice"]).max() < 1e-4
    assert np.allclose(g["grad_norms"], now["grad_norms"], rtol=1e-3, atol=1e-7)


def test_oracle_reproduces_tacotron_fixture():
    g, now = _load("tacotron_tiny.npz"), mg.tacotron_case()
    assert abs(g["param_abs_sum"] - now["param_abs_sum"]) < 1e-6 * g["param_abs_sum"]
    assert np.abs(g["parts"] - now["parts"]).max() < 1e-5
    assert np.abs(g["alignments"] - now["alignments"]).max() < 1e-5
    assert np.abs(g["mel_outputs"] - now["mel_outputs"]).max() < 1e-3


# ------------------------------------------------------------------ GPU: CUDA path vs fixtures
@pytest.mark.gpu
def test_cuda_audio_matches_fixture():
    from t2_import import t2
    g = _load("audio_cfg1.npz")
    q = t2.audio.mulaw_quantize(torch.from_numpy(g["wav"]).cuda()).cpu().numpy()
    assert np.array_equal(q, g["mulaw_q"].astype(np.int32))                      # bit-exact
    assert np.array_equal(t2.audio.mulaw(torch.from_numpy(g["wav"]).cuda()).cpu().numpy(), g["mulaw_f"])
    fe = t2.audio.MelFrontEnd(hparams)
    mel, lin = fe(torch.from_numpy(g["pre"][None]).cuda(), linear=True)
    assert np.abs(mel[0].cpu().numpy().T - g["mel"]).max() < 1e-3
    assert np.abs(lin[0].cpu().numpy().T[::41] - g["lin_rows"]).max() < 1e-3
    pre = t2.audio.preemphasis(torch.from_numpy(g["wav"][None]).cuda(), hparams.preemphasis)[0].cpu().numpy()
    assert np.abs(pre - g["pre"]).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ce", "mol"])
def test_cuda_wavenet_matches_fixture(kind):
    from oracle import wavenet as ow        # only for the seeded parameter draw (pinned by the checksum)
    from t2_import import t2
    g = _load("wavenet_%s_tiny.npz" % kind)
    hp = mg.wn_hp(kind)
    params = ow.init_params(hp, seed=int(g["seed"]), random_bias=True)
    assert abs(float(sum(v.double().abs().sum() for v in params.values())) - g["param_abs_sum"]) < 1e-6 * g["param_abs_sum"]
    w = torch.from_numpy(g["w"])
    B, T = w.shape
    model = t2.wavenet.WaveNet(hp, B, T)
    model.load_params(params)
    if kind == "ce":
        xd = t2.audio.mulaw_quantize(w.cuda())
        yd, ldo = xd, 256
    else:
        xd = yd = w.cuda()
        ldo = 32
    logits = torch.zeros(B, T, ldo, device="cuda")
    model.forward(xd, torch.from_numpy(g["c"]).cuda(), yd, torch.from_numpy(g["lengths"]).int().cuda(), logits=logits)
    model.backward()
    torch.cuda.synchronize()
    assert abs(model.loss_value() - float(g["loss"])) < (1e-3 if kind == "ce" else 2e-3)
    got = logits[:, :, :hp.out_channels].cpu().transpose(1, 2)[:, :, ::37].numpy()
    err = np.abs(got - g["yhat_slice"])
    assert err.max() < 4e-2 and err.mean() < 6e-3
    grads = model.export_grads()
    for name, ref in zip(g["grad_names"], g["grad_norms"]):
        if ref > 1e-6:
            assert abs(grads[str(name)].norm().item() - ref) < 0.1 * ref, name


@pytest.mark.gpu
def test_cuda_tacotron_matches_fixture():
    from oracle import tacotron as ot       # only for the seeded parameter draw (pinned by the checksum)
    from t2_import import t2
    g = _load("tacotron_tiny.npz")
    hp = mg.taco_hp()
    params = ot.init_params(hp, seed=int(g["seed"]), random_bias=True)
    assert abs(float(sum(v.double().abs().sum() for v in params.values())) - g["param_abs_sum"]) < 1e-6 * g["param_abs_sum"]
    B, T_in = g["inputs"].shape
    T_out = g["mel"].shape[1]
    model = t2.tacotron.Tacotron(hp, B, T_in, T_out)
    model.load_params(params)
    model.forward(torch.from_numpy(g["inputs"]).int().cuda(), torch.from_numpy(g["lens"]).int().cuda(),
                  torch.from_numpy(g["mel"]).cuda(), torch.from_numpy(g["stop"]).cuda())
    torch.cuda.synchronize()
    los = model.losses()
    for i, k in enumerate(("before", "after", "stop", "reg")):
        assert abs(los[k] - g["parts"][i]) < 2e-3, k
    al = model.workspace_tensor("alignments", (T_out, B, T_in)).float().cpu().transpose(0, 1).numpy()
    assert np.abs(al - g["alignments"]).max() < 2e-2
    dec = model.workspace_tensor("decoder_output", (B, T_out, hp.num_mels)).cpu().numpy()
    assert np.abs(dec - g["decoder_output"]).mean() < 1e-2
    melo = model.workspace_tensor("mel_outputs", (B, T_out, hp.num_mels)).cpu().numpy()
    assert np.abs(melo - g["mel_outputs"]).mean() < 4e-2
