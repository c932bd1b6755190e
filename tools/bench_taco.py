"""Config 3 (BASELINE.json): Tacotron encoder+decoder+postnet TRAINING step, r=1, batch 32, T_in 160, T_out 800, bf16
GEMM operands, synthetic LJSpeech-shaped batch. Prints mel-frames/sec (B*T_out frames per step) for the CUDA path and
a bounded CPU-oracle baseline (B=4 sample)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hparams import hparams
from oracle import tacotron as ot
from t2_import import t2


def batch(hp, B, T_in, T_out, seed=3):
    g = torch.Generator().manual_seed(seed)
    inputs = torch.randint(2, 66, (B, T_in), generator=g)
    lens = torch.randint(60, T_in + 1, (B,), generator=g).sort(descending=True).values
    lens[0] = T_in
    for b in range(B):
        inputs[b, lens[b] - 1] = 1
        inputs[b, lens[b]:] = 0
    tl = torch.randint(T_out // 2, T_out + 1, (B,), generator=g)
    mel = (torch.randn(B, T_out, hp.num_mels, generator=g) * 1.5 - 1).clamp(-4, 4)
    stop = torch.zeros(B, T_out)
    for b in range(B):
        mel[b, tl[b]:] = -4.0
        stop[b, tl[b] - 1:] = 1.0
    return inputs, lens, mel, stop


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 5
    hp = hparams.copy()
    linear = "--linear" in sys.argv               # with the CBHG post-processing net + linear head (the reference's default)
    hp.parse("predict_linear=%s" % linear)
    B, T_in, T_out = 32, 160, 800
    inputs, lens, mel, stop = batch(hp, B, T_in, T_out)
    model = t2.tacotron.Tacotron(hp, B, T_in, T_out)
    model.init_variables(seed=5339)
    args = (inputs.int().cuda(), lens.int().cuda(), mel.cuda(), stop.cuda())
    kw = {}
    if linear:
        g = torch.Generator().manual_seed(9)
        kw["linear_targets"] = (torch.randn(B, T_out, hp.num_freq, generator=g) * 1.5 - 1).clamp(-4, 4).cuda()
    lib = t2.lib.load()

    use_graph = "--graph" in sys.argv
    if use_graph:
        model.capture(*args, **kw)

    def step():
        if use_graph:
            model.train_step()
        else:
            model.train_step(*args, **kw)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    n0 = lib.t2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (lib.t2_launch_count() - n0) // steps
    out = {"metric": "tacotron_train_mel_frames_per_sec", "value": B * T_out / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
           "config": {"workload": "tacotron_train r=1 B=32 T_in=160 T_out=800 predict_linear=False, bf16 GEMM operands / fp32 state, " + ("CUDA graph" if use_graph else "eager launches")},
           "kernel_launches_per_step": int(launches), "loss": model.losses()}
    # bounded CPU baseline: the oracle (fp32, autograd) on B=4 of the same shapes
    Bc = 4
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    p = ot.init_params(hp, seed=5339)
    t0 = time.perf_counter()
    ot.train_step(p, inputs[:Bc], lens[:Bc], mel[:Bc], stop[:Bc], hp)
    dt = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": Bc * T_out / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                           "sample": "1 oracle train step (fwd + autograd bwd), B=%d x T_out=%d, fp32" % (Bc, T_out)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
