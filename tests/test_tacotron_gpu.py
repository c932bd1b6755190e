"""CUDA Tacotron (through the C-ABI) vs the fp32 CPU oracle, dropout / zoneout off (rates are hparams), same seeded
inputs. Tolerances (bf16 GEMM operands, fp32 accumulate / state): losses <= 2e-3 absolute, mel outputs mean abs err
<= 4e-2 (five batch-normalised postnet layers re-normalise bf16 noise to unit scale), alignments max abs err <= 2e-2.
Gradients vs the fp32 oracle: per tensor cosine >= 0.97 and relative error <= 0.25 (measured: 2-4 % for the large
tensors; 10-18 % for the small encoder-conv / location-attention tensors of the tiny B=3 problem, shrinking as the batch
grows — the bf16 sign-flip noise floor discussed in tests/test_wavenet_gpu.py, amplified by batch-norm over ~100 rows)."""
import pytest
import torch

from hparams import hparams
from oracle import tacotron as ot
from t2_import import t2

pytestmark = pytest.mark.gpu


def _hp(**kw):
    hp = hparams.copy()
    hp.parse("predict_linear=False,tacotron_dropout_rate=0.0,tacotron_zoneout_rate=0.0,enc_conv_channels=256,embedding_dim=256,"
             "encoder_lstm_units=128,decoder_lstm_units=256,postnet_channels=256,prenet_layers=[128,128],attention_dim=128")
    for k, v in kw.items():
        hp.set_hparam(k, v)
    return hp


def _batch(hp, B, T_in, T_out, seed):
    g = torch.Generator().manual_seed(seed)
    inputs = torch.randint(2, 66, (B, T_in), generator=g)
    lens = torch.tensor([T_in] + [max(T_in - 7 * (i + 1), 3) for i in range(B - 1)])
    for b in range(B):
        inputs[b, lens[b]:] = 0
    mel = (torch.randn(B, T_out, hp.num_mels, generator=g) * 1.5 - 1).clamp(-4, 4)
    stop = torch.zeros(B, T_out)
    stop[:, -3:] = 1
    return inputs, lens, mel, stop


def _run_forward(hp, B, T_in, T_out, seed):
    params = ot.init_params(hp, seed=seed, random_bias=True)
    inputs, lens, mel, stop = _batch(hp, B, T_in, T_out, seed)
    ref = ot.forward(params, inputs, lens, mel, hp, training=True)
    _, parts = ot.loss_fn(ref, mel, stop, params, hp)
    model = t2.tacotron.Tacotron(hp, B, T_in, T_out)
    model.load_params(params)
    model.forward(inputs.int().cuda(), lens.int().cuda(), mel.cuda(), stop.cuda())
    torch.cuda.synchronize()
    return model, params, ref, parts, (inputs, lens, mel, stop)


@pytest.mark.parametrize("B,T_in,T_out", [(3, 40, 24), (2, 150, 33)])
def test_forward_matches_oracle(B, T_in, T_out):
    hp = _hp()
    model, params, ref, parts, _ = _run_forward(hp, B, T_in, T_out, 31)
    mem_ref = None
    al = model.workspace_tensor("alignments", (T_out, B, T_in)).float().cpu().transpose(0, 1)
    err_al = (al - ref["alignments"]).abs().max().item()
    dec = model.workspace_tensor("decoder_output", (B, T_out, hp.num_mels)).cpu()
    melo = model.workspace_tensor("mel_outputs", (B, T_out, hp.num_mels)).cpu()
    stop = model.workspace_tensor("stop_logits", (B, T_out)).cpu()
    e_dec = (dec - ref["decoder_output"]).abs()
    e_mel = (melo - ref["mel_outputs"]).abs()
    e_stop = (stop - ref["stop_logits"]).abs()
    los = model.losses()
    print("align err %.3g | dec max %.3g mean %.3g | mel max %.3g mean %.3g | stop max %.3g | losses cuda %s oracle %s" % (
        err_al, e_dec.max(), e_dec.mean(), e_mel.max(), e_mel.mean(), e_stop.max(), los, {k: round(v.item(), 6) for k, v in parts.items()}))
    assert err_al < 2e-2
    assert e_dec.mean().item() < 1e-2 and e_mel.mean().item() < 4e-2 and e_stop.max().item() < 5e-2
    for k in ("before", "after", "stop", "reg"):
        assert abs(los[k] - parts[k].item()) < 2e-3, k


@pytest.mark.parametrize("B,T_in,T_out", [(3, 40, 24), (8, 60, 64)])
def test_backward_matches_oracle(B, T_in, T_out):
    hp = _hp()
    model, params, ref, parts, (inputs, lens, mel, stop) = _run_forward(hp, B, T_in, T_out, 32)
    model.backward()
    torch.cuda.synchronize()
    _, grads_ref, _, _ = ot.train_step(params, inputs, lens, mel, stop, hp)
    grads = model.export_grads()
    report, bad = [], []
    for name, g_ref in grads_ref.items():
        g = grads[name]
        den = g_ref.norm().item()
        rel = (g - g_ref).norm().item() / max(den, 1e-12)
        cos = (g * g_ref).sum().item() / max(den * g.norm().item(), 1e-20)
        report.append("%-60s rel %.4g cos %.4f |ref| %.3g |cuda| %.3g" % (name, rel, cos, den, g.norm().item()))
        if den >= 1e-6 and (rel >= 0.25 or cos < 0.97):
            bad.append(report[-1])
    print("\n".join(report))
    assert not bad, "gradient mismatch:\n" + "\n".join(bad)


def test_adam_global_norm_clip_matches_oracle():
    hp = _hp()
    B, T_in, T_out = 2, 24, 12
    model, params, ref, parts, (inputs, lens, mel, stop) = _run_forward(hp, B, T_in, T_out, 33)
    model.backward()
    grads = model.export_grads()
    state = {}
    p_ref = {k: v.clone() for k, v in params.items() if ot.is_trainable(k)}
    ot.adam_step(p_ref, {k: grads[k] for k in p_ref}, state, hp, 0)
    model.optimizer_step()
    torch.cuda.synchronize()
    p_new = model.export_params()
    for k in p_ref:
        assert (p_new[k] - p_ref[k]).abs().max().item() < 2e-6, k
