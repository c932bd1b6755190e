"""Debug tool (GPU box): where does the Tacotron forward deviate from the fp32 oracle? Compares every conv block of the encoder
and the postnet layer by layer: each CUDA layer output against (a) the oracle's own chain and (b) the oracle layer applied to
the CUDA path's previous-layer output (isolates the error ADDED by that layer from the error it inherits)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hparams import hparams
from oracle import tacotron as ot
from t2_import import t2
from test_parity_full_gpu import taco_batch

hp = hparams.copy()
hp.parse("predict_linear=False,tacotron_dropout_rate=0.0,tacotron_zoneout_rate=0.0")
B, T_in, T_out = 16, 120, 160
params = ot.init_params(hp, seed=53, random_bias=True)
inputs, lens, mel, stop = taco_batch(hp, B, T_in, T_out, 53)
model = t2.tacotron.Tacotron(hp, B, T_in, T_out)
model.load_params(params)
model.forward(inputs.int().cuda(), lens.int().cuda(), mel.cuda(), stop.cuda(), training=True, seed=99)
torch.cuda.synchronize()
with torch.no_grad():
    ref = ot.forward(params, inputs, lens, mel, hp, training=True)
dec = model.workspace_tensor("decoder_output", (B, T_out, hp.num_mels)).cpu()
print("decoder_output: L1 %.3g  (ref std over (b,t) per channel, mean: %.3g ; ref overall std %.3g)" % (
    (dec - ref["decoder_output"]).abs().mean(), ref["decoder_output"].std(dim=(0, 1)).mean(), ref["decoder_output"].std()))


def chain(prefix, n, x_ref, x_cuda, acts, T, tag):
    for i in range(n):
        p = "%s/conv_layer_%d/" % (prefix, i + 1)
        with torch.no_grad():
            y_ref = ot.conv_block(x_ref, params, p, acts[i], True, 0.0)
            y_from_cuda = ot.conv_block(x_cuda, params, p, acts[i], True, 0.0)
        x_c = model.workspace_tensor("%s_x%d" % (tag, i), (B, T, y_ref.shape[-1])).float().cpu()
        # pre-BN statistics of the oracle layer: |mean| / std per channel tells how much batch norm amplifies bf16 rounding of y
        k = params[p + "kernel"]
        z = torch.nn.functional.conv1d(x_ref.transpose(1, 2), k.permute(2, 1, 0).contiguous(), params[p + "bias"], padding=(k.shape[0] - 1) // 2).transpose(1, 2)
        if acts[i] == "relu":
            z = torch.relu(z)
        elif acts[i] == "tanh":
            z = torch.tanh(z)
        ratio = (z.mean(dim=(0, 1)).abs() / z.std(dim=(0, 1))).mean().item()
        print("%s layer %d: vs oracle chain L1 %.4g | error added by this layer (oracle layer on CUDA input) L1 %.4g | pre-BN |mean|/std %.3g | pre-BN std %.3g" % (
            tag, i, (x_c - y_ref).abs().mean(), (x_c - y_from_cuda).abs().mean(), ratio, z.std(dim=(0, 1)).mean()))
        x_ref, x_cuda = y_ref, x_c
    return x_ref, x_cuda


emb = params["inputs_embedding"][inputs]
chain("encoder_convolutions", hp.enc_conv_num_layers, emb, emb, ["relu"] * hp.enc_conv_num_layers, T_in, "enc_conv")
acts = ["tanh"] * (hp.postnet_num_layers - 1) + [None]
xr, xc = chain("postnet_convolutions", hp.postnet_num_layers, ref["decoder_output"], dec, acts, T_out, "post_conv")
melo = model.workspace_tensor("mel_outputs", (B, T_out, hp.num_mels)).cpu()
print("mel_outputs L1 %.4g max %.4g" % ((melo - ref["mel_outputs"]).abs().mean(), (melo - ref["mel_outputs"]).abs().max()))
