"""The two arithmetic modes of the CUDA path against the fp32 oracle (SURVEY §7 step 6 "parity in fp32 first, then
bf16 mode"). The reference computes in fp32 throughout; the product's training / benchmark mode uses bf16 tensor-core operands and
bf16-stored activations (fp32 accumulate). The 'fp32-class' mode carries every activation and weight as a bf16 hi + lo pair through the
SAME tcgen05 kernels (three products per contraction), forward + loss only:

  WaveNet fp32-class   logits max abs err <= 1e-4 (measured 6e-6), loss (CE / MoL NLL) abs err <= 1e-4   -> north-star 1e-3 met
  WaveNet bf16         logits max abs err <= 5e-3, loss abs err <= 1e-3 (measured: 1.7e-3 / 3e-5 at the 24-layer Cfg-2 shape)
  Tacotron             the convolution stacks have the fp32-class mode, the recurrences do not: see the Tacotron test below

so the bf16-mode deviation is operand / storage rounding, not a difference in the algorithm."""
import math

import pytest
import torch

from hparams import hparams
from oracle import wavenet as ow
from t2_import import t2
from parity_util import record
from test_parity_full_gpu import _wn_hp, _wn_inputs, taco_compare

pytestmark = pytest.mark.gpu


def _forward(hp, B, T, seed, precision):
    params = ow.init_params(hp, seed=seed, random_bias=True)
    x, c, y, lengths, xd, yd = _wn_inputs(hp, B, T, seed)
    with torch.no_grad():
        yhat_ref = ow.step(x, c, params, hp)
        loss_ref = ow.loss_fn(yhat_ref, y, lengths, hp).item()
    model = t2.wavenet.WaveNet(hp, B, T, precision=precision)
    model.load_params(params)
    no = 256 if ow.is_mulaw_quantize(hp.input_type) else 32
    logits = torch.zeros(B, T, no, device="cuda")
    model.forward(xd.cuda(), c.cuda(), yd.cuda(), lengths.int().cuda(), logits=logits, save_for_backward=(precision == "bf16"))
    torch.cuda.synchronize()
    err = (logits[:, :, :hp.out_channels].cpu() - yhat_ref.transpose(1, 2)).abs()
    return err.max().item(), err.mean().item(), abs(model.loss_value() - loss_ref), model


@pytest.mark.parametrize("shape", ["small_ce", "cfg2_24L_ce", "cfg4_24L_mol", "small_gauss"])
def test_wavenet_fp32_class_vs_bf16(shape):
    if shape == "small_ce":
        hp = hparams.copy()
        hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256,layers=6,stacks=2,residual_channels=128,gate_channels=256,"
                 "skip_out_channels=128,upsample_scales=[4,4],hop_size=16,wavenet_dropout=0.0")
        B, T = 2, 400
    elif shape == "cfg2_24L_ce":
        hp, B, T = _wn_hp("input_type=mulaw-quantize,quantize_channels=256,out_channels=256"), 2, 7680
    elif shape == "cfg4_24L_mol":
        hp, B, T = _wn_hp("input_type=raw,quantize_channels=65536,out_channels=30"), 2, 4096
    else:
        hp = hparams.copy()
        hp.parse("input_type=raw,out_channels=2,layers=6,stacks=2,residual_channels=256,gate_channels=512,skip_out_channels=256,"
                 "upsample_scales=[4,4],hop_size=16,wavenet_dropout=0.0,legacy=False,residual_legacy=False,upsample_type=2D")
        B, T = 2, 256
    out = {}
    for precision in ("fp32-class", "bf16"):
        mx, mean, dl, model = _forward(hp, B, T, 31, precision)
        out[precision] = (mx, mean, dl)
        del model
        torch.cuda.empty_cache()
    record("wavenet_precision_modes_" + shape, fp32_class_logits_max=out["fp32-class"][0], fp32_class_logits_mean=out["fp32-class"][1],
           fp32_class_loss_err=out["fp32-class"][2], bf16_logits_max=out["bf16"][0], bf16_logits_mean=out["bf16"][1], bf16_loss_err=out["bf16"][2])
    assert out["fp32-class"][0] <= 1e-4 and out["fp32-class"][2] <= 1e-4, out
    # bf16 mode: CE / MoL NLL within 1e-3 (measured 8e-6 .. 6e-5); the Gaussian log-density loss divides by the predicted variance and
    # amplifies the bf16 logit error (~3.8e-3) to ~1e-3, hence 3e-3 there
    assert out["bf16"][0] <= 5e-3 and out["bf16"][2] <= (3e-3 if shape == "small_gauss" else 1e-3), out
    assert out["fp32-class"][0] < 0.1 * out["bf16"][0]


@pytest.mark.parametrize("stochastic", [False, True])
def test_tacotron_fp32_class_conv_stacks_vs_bf16(stochastic):
    """mel-L1 on `mel_outputs` (the north-star parity metric) at the Cfg-3 widths, B = 32, T_in 160, T_out 200.
    bf16 mode: ~2.5e-2 at random init. Two contributions (tools/taco_layer_diag.py): (i) every batch-normalised postnet layer adds
    ~0.2 % of a unit-variance activation through bf16 storage of operands / activations; (ii) at random init the decoder outputs are
    nearly constant over (batch, time) (per-channel std 0.06), so the FIRST postnet batch norm divides by a pre-norm std of 0.03 and
    amplifies the decoder-output deviation (4e-4, bf16 recurrence GEMMs) ~30x. The 'fp32-class' mode runs the convolution stacks on
    bf16 hi + lo operand pairs with fp32 pre-norm activations and removes (i) - mel-L1 halves to ~1.2e-2; (ii) remains because the
    recurrences keep bf16 operands (a split-operand decoder is not implemented). Asserted: the measured values with
    2x margin, alignments / decoder output / losses unchanged or better."""
    from hparams import hparams as hp0
    hp = hp0.copy()
    hp.parse("predict_linear=False" + ("" if stochastic else ",tacotron_dropout_rate=0.0,tacotron_zoneout_rate=0.0"))
    tol = dict(align=6e-4, dec_l1=1.6e-3, stop=5e-3, loss=2e-3, grad_rel=1.0, grad_cos=0.0)
    tag = "tacotron_precision_modes_%s_" % ("stochastic" if stochastic else "deterministic")
    a = taco_compare(tag + "fp32_class", hp, 32, 160, 200, 54, dict(tol, mel_l1=2.5e-2), backward=False, precision="fp32-class").measured
    b = taco_compare(tag + "bf16", hp, 32, 160, 200, 54, dict(tol, mel_l1=5e-2), backward=False, precision="bf16").measured
    assert a["mel_l1"] <= 0.7 * b["mel_l1"], (a["mel_l1"], b["mel_l1"])
    assert a["dec_l1"] <= b["dec_l1"] * 1.05 and a["align_max_err"] <= b["align_max_err"] * 1.5


def test_fp32_class_mode_is_forward_only():
    hp = hparams.copy()
    hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256,layers=4,stacks=2,residual_channels=128,gate_channels=256,"
             "skip_out_channels=128,upsample_scales=[4,4],hop_size=16")
    m = t2.wavenet.WaveNet(hp, 1, 128, precision="fp32-class")
    assert m.cfg.dropout == 0.0
    m.init_variables(seed=1)
    x = torch.zeros(1, 128, dtype=torch.int32, device="cuda")
    c = torch.rand(1, 80, 8, device="cuda")
    ln = torch.tensor([128], dtype=torch.int32, device="cuda")
    with pytest.raises(t2.lib.T2Error):
        m.forward(x, c, x, ln, save_for_backward=True)
    m.forward(x, c, x, ln, save_for_backward=False)
    with pytest.raises(t2.lib.T2Error):
        m.backward()
