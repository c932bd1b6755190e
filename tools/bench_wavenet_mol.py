"""Config 4 (BASELINE.json): WaveNet MoL-10 training step conditioned on GTA-like mels, raw float input, paper widths,
batch 8 x 16128 samples per GPU (16000 rounded up to the 256-sample hop), fwd + MoL NLL + bwd + clip + Adam + EMA, one CUDA
graph per step. Run under torchrun for N > 1 (one NCCL all-reduce of the flat gradient buffer per step). Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    import torch.distributed as dist
    from hparams import hparams
    from t2_import import t2
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hp = hparams.copy()
    hp.parse("input_type=raw,quantize_channels=65536,out_channels=30,layers=24,stacks=4,residual_channels=256,gate_channels=512,"
             "skip_out_channels=256,upsample_scales=[16,16],hop_size=256,wavenet_dropout=0.05")
    B, T = 8, 16128
    from scipy.signal import lfilter
    rng = np.random.default_rng(4 + rank)
    w = lfilter([1.0], [1.0, -1.6, 0.8], rng.standard_normal((B, T + 64)), axis=1)[:, 64:]
    w = (w / np.abs(w).max() * 0.6).astype(np.float32)
    c = np.clip(rng.normal(0.0, 1.5, (B, 80, T // 256)), -4, 4).astype(np.float32) / 8 + 0.5     # GTA-like mels -> [0, 1]
    lengths = np.full((B,), T, dtype=np.int32)
    x = torch.from_numpy(w).to(dev)
    static = (x, torch.from_numpy(c).to(dev), x.clone(), torch.from_numpy(lengths).to(dev))
    model = t2.wavenet.WaveNet(hp, B, T, device=dev)
    model.init_variables(seed=5339)
    model.capture(*static)
    for _ in range(5):
        model.train_step(world_size=world)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.train_step(world_size=world)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    gate_ms = sum(model.time_kernel(0, l, reps=20) for l in (3, 9, 15, 21)) / 4
    flops = 2.0 * B * T * hp.gate_channels * (3 * hp.residual_channels + hp.cin_channels)
    if rank == 0:
        print(json.dumps({"metric": "wavenet_train_audio_samples_per_sec", "value": world * B * T / (ms * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": steps, "ms_per_step": ms, "dtype": "bf16", "data": "synthetic", "scaling": "weak",
                          "config": {"workload": "wavenet_mol10_24L: raw input, MoL-10 NLL, R256/G512/S256, batch 8 x 16128 samples per GPU, dropout 0.05, dp%d" % world},
                          "loss": model.loss_value(),
                          "roofline": {"bound": "tensor", "kernel": "act_gemm_kernel<EPI_GATE,256,NT=2>", "ms_per_launch": gate_ms,
                                       "achieved": flops / (gate_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "flops_per_launch": flops},
                          "workspace_gb": model.workspace.numel() / 1e9}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
