"""Regenerates the golden vectors in this directory:  python tests/golden/make_golden.py

The reference (TF1 + librosa 0.5.1) cannot be imported in the build container and ships no golden
vectors, so these fixtures are frozen outputs of the CPU oracle (`oracle/`) on the seeded synthetic inputs of SURVEY.md
§8d — "parity unpinned" against the reference itself; the oracle's independent pins (torch.stft, torchaudio's Slaney
filterbank, analytic mu-law known answers, incremental == parallel) live in tests/test_oracle_*.py. The fixtures guard two
things: (1) the oracle does not drift (tests/test_golden.py, CPU), (2) the CUDA path reproduces them through the C-ABI
(tests/test_golden.py, -m gpu) without the oracle in the loop.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from hparams import hparams  # noqa: E402
from oracle import audio as oa  # noqa: E402
from oracle import tacotron as ot  # noqa: E402
from oracle import wavenet as ow  # noqa: E402


def cfg1_wav(seed=1, n=22050):
    """sine sweep + noise, peak 0.999"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 22050.0
    w = 0.5 * np.sin(2 * np.pi * (200 + 2000 * t) * t) + rng.normal(0, 0.05, n)
    return (w / np.abs(w).max() * 0.999).astype(np.float32)


def audio_case():
    wav = cfg1_wav()
    pre = oa.preemphasis(wav, hparams.preemphasis).astype(np.float32)
    mel = oa.melspectrogram(pre, hparams).astype(np.float32)            # [80, 81]
    lin = oa.linearspectrogram(pre, hparams).astype(np.float32)         # [1025, 81]
    q = oa.mulaw_quantize(wav).astype(np.int16)
    return dict(wav=wav, pre=pre, mel=mel, lin_rows=lin[::41].copy(), mulaw_q=q, mulaw_f=oa.mulaw(wav).astype(np.float32))


def wn_hp(kind):
    hp = hparams.copy()
    hp.parse("layers=4,stacks=2,residual_channels=128,gate_channels=256,skip_out_channels=128,"
             "upsample_scales=[4,4],hop_size=16,wavenet_dropout=0.0")
    if kind == "mol":
        hp.parse("input_type=raw,quantize_channels=65536,out_channels=30,legacy=False,residual_legacy=False,"
                 "upsample_type=2D,residual_channels=256,gate_channels=512,skip_out_channels=256")
    else:
        hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256")
    return hp


def wn_inputs(hp, B, T, seed):
    from scipy.signal import lfilter
    g = torch.Generator().manual_seed(seed)
    hop = math.prod(hp.upsample_scales)
    c = torch.rand(B, hp.cin_channels, T // hop, generator=g)
    e = torch.randn(B, T + 64, generator=g).numpy()
    w = torch.from_numpy(lfilter([1.0], [1.0, -1.6, 0.8], e, axis=1)[:, 64:].copy()).float()
    w = w / w.abs().max() * 0.6
    lengths = torch.tensor([T] + [max(T - 37 * (i + 1), 2) for i in range(B - 1)])
    return c, w, lengths


def wavenet_case(kind, seed=7):
    hp = wn_hp(kind)
    B, T = (2, 512) if kind == "ce" else (3, 256)
    params = ow.init_params(hp, seed=seed, random_bias=True)
    c, w, lengths = wn_inputs(hp, B, T, seed)
    if kind == "ce":
        idx = torch.from_numpy(oa.mulaw_quantize(w.numpy()))
        x = torch.nn.functional.one_hot(idx, hp.quantize_channels).float().transpose(1, 2)
        y = idx
    else:
        x, y = w.unsqueeze(1), w
    loss, grads, yhat = ow.train_step(params, x, c, y, lengths, hp)
    psum = float(sum(v.double().abs().sum() for v in params.values()))
    gn = {k: float(v.norm()) for k, v in grads.items()}
    names = sorted(gn)
    return dict(c=c.numpy(), w=w.numpy(), lengths=lengths.numpy(), loss=np.float64(loss.item()),
                yhat_slice=yhat[:, :, ::37].numpy().astype(np.float32), param_abs_sum=np.float64(psum),
                grad_names=np.array(names), grad_norms=np.array([gn[k] for k in names], np.float64), seed=np.int64(seed))


def taco_hp():
    hp = hparams.copy()
    hp.parse("predict_linear=False,tacotron_dropout_rate=0.0,tacotron_zoneout_rate=0.0,enc_conv_channels=256,embedding_dim=256,"
             "encoder_lstm_units=128,decoder_lstm_units=256,postnet_channels=256,prenet_layers=[128,128],attention_dim=128")
    return hp


def taco_batch(hp, B, T_in, T_out, seed):
    g = torch.Generator().manual_seed(seed)
    inputs = torch.randint(2, 66, (B, T_in), generator=g)
    lens = torch.tensor([T_in] + [max(T_in - 7 * (i + 1), 3) for i in range(B - 1)])
    for b in range(B):
        inputs[b, lens[b]:] = 0
    mel = (torch.randn(B, T_out, hp.num_mels, generator=g) * 1.5 - 1).clamp(-4, 4)
    stop = torch.zeros(B, T_out)
    stop[:, -3:] = 1
    return inputs, lens, mel, stop


def tacotron_case(B=3, T_in=40, T_out=24, seed=11):
    hp = taco_hp()
    params = ot.init_params(hp, seed=seed, random_bias=True)
    inputs, lens, mel, stop = taco_batch(hp, B, T_in, T_out, seed)
    ref = ot.forward(params, inputs, lens, mel, hp, training=True)
    total, parts = ot.loss_fn(ref, mel, stop, params, hp)
    psum = float(sum(v.double().abs().sum() for v in params.values()))
    return dict(inputs=inputs.numpy(), lens=lens.numpy(), mel=mel.numpy(), stop=stop.numpy(), seed=np.int64(seed),
                param_abs_sum=np.float64(psum), loss=np.float64(total.item()),
                parts=np.array([parts[k].item() for k in ("before", "after", "stop", "reg")], np.float64),
                alignments=ref["alignments"].numpy().astype(np.float32),
                decoder_output=ref["decoder_output"].numpy().astype(np.float32),
                mel_outputs=ref["mel_outputs"].numpy().astype(np.float32))


def main():
    np.savez_compressed(os.path.join(HERE, "audio_cfg1.npz"), **audio_case())
    np.savez_compressed(os.path.join(HERE, "wavenet_ce_tiny.npz"), **wavenet_case("ce"))
    np.savez_compressed(os.path.join(HERE, "wavenet_mol_tiny.npz"), **wavenet_case("mol"))
    np.savez_compressed(os.path.join(HERE, "tacotron_tiny.npz"), **tacotron_case())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
