#!/usr/bin/env python
"""bench.py — headline benchmark: WaveNet vocoder TRAIN throughput (audio samples / second), BASELINE.json configs[1].

Workload "wavenet_ce_24L" (SURVEY.md §8d Cfg-2): 24 layers / 4 stacks, residual 256 / gate 512 / skip 256, mu-law-256
one-hot input and softmax-CE loss, local conditioning on 80-band mels through the learnable upsampling net, dropout
0.05, per-GPU batch 2 x 7680 samples (hop 256 = upsample_scales [16,16] so 7680 is hop-aligned), synthetic
LJSpeech-shaped data, random-init weights. One step = forward + loss + backward + gradient all-reduce (N > 1) +
per-tensor clip + Adam + EMA + re-pack of the bf16 operand copies.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (CUDA, one process per GPU)
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  # CPU arm: the oracle restatement of the
                                                                         # reference graph on the host cores (TF1 cannot
                                                                         # be installed here)
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def workload_hparams():
    from hparams import hparams
    hp = hparams.copy()
    hp.parse("input_type=mulaw-quantize,quantize_channels=256,out_channels=256,layers=24,stacks=4,"
             "residual_channels=256,gate_channels=512,skip_out_channels=256,upsample_scales=[16,16],hop_size=256,"
             "wavenet_dropout=0.05")
    return hp


B_PER_GPU, T_STEP = 2, 7680


def synth_batch(hp, B, T, seed, quantize):
    """Synthetic LJSpeech-shaped batch: AR(2) 'speech-like' noise -> mu-law indices; mels ~ U[0,1].
    quantize: float32 [B,T] -> int indices (the CUDA mu-law kernel on the GPU arm, the oracle on the CPU arm)."""
    import numpy as np
    from scipy.signal import lfilter
    rng = np.random.default_rng(seed)
    e = rng.standard_normal((B, T + 64))
    w = lfilter([1.0], [1.0, -1.6, 0.8], e, axis=1)[:, 64:]
    w = (w / np.abs(w).max() * 0.6).astype(np.float32)
    idx = quantize(w).astype(np.int32)
    c = rng.random((B, hp.cin_channels, T // 256), dtype=np.float32)
    lengths = np.full((B,), T, dtype=np.int32)
    return idx, c, lengths


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, reasons, mx = [], set(), 0
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_run(hp, steps, warmup, B, T):
    """Times the oracle (fp32 PyTorch-CPU restatement of the reference graph) on the host cores: one step =
    forward + loss + autograd backward + clip + Adam + EMA on a bounded sample (B x T samples)."""
    import torch
    from oracle import wavenet as ow
    ncores = os.cpu_count() or 1
    params = ow.init_params(hp, seed=5339)
    from oracle import audio as oa
    idx, c, lengths = synth_batch(hp, B, T, 2, oa.mulaw_quantize)
    # pick the intra-op thread count that runs this graph fastest on this host (oversubscribing a 128-core box
    # with 128 threads on these small convolutions is ~5x slower than 32): probe a short forward at each setting
    best = (None, 1e30)
    probe_x = torch.nn.functional.one_hot(torch.from_numpy(idx[:1, :2048]).long(), hp.quantize_channels).float().transpose(1, 2).contiguous()
    probe_c = torch.from_numpy(c[:1, :, :8])
    for nt in sorted({ncores, min(ncores, 64), min(ncores, 32), min(ncores, 16)}):
        torch.set_num_threads(nt)
        with torch.no_grad():
            ow.step(probe_x, probe_c, params, hp)
            t0 = time.perf_counter()
            ow.step(probe_x, probe_c, params, hp)
            dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (nt, dt)
    nthreads = best[0]
    torch.set_num_threads(nthreads)
    idx_t = torch.from_numpy(idx).long()
    x = torch.nn.functional.one_hot(idx_t, hp.quantize_channels).float().transpose(1, 2).contiguous()
    c_t, len_t = torch.from_numpy(c), torch.from_numpy(lengths).long()
    state = {}
    hp_nodrop = hp  # the oracle has no RNG-matched dropout; cost of the mask multiply is negligible on CPU
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss, grads, _ = ow.train_step(params, x, c_t, idx_t, len_t, hp_nodrop)
        ow.adam_step(params, grads, state, hp, i)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": B * T * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": nthreads,
            "sample": "B=%d x T=%d samples per step, %d timed steps, fp32, torch.set_num_threads(%d) of %d host cores (fastest probed setting)" % (
                B, T, len(times), nthreads, ncores),
            "loss": float(loss)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    hp = workload_hparams()
    config = {"workload": "wavenet_ce_24L: WaveNet 24-layer dilated stack fwd+bwd+Adam, mu-law 256 softmax-CE, "
                          "R256/G512/S256, batch %d x %d samples per GPU, dropout 0.05, dp%d" % (B_PER_GPU, T_STEP, max(world, args.gpus)),
              "per_gpu_batch": B_PER_GPU, "samples_per_item": T_STEP, "parallelism": "dp%d" % max(world, 1),
              "l2": "per-step working set 1.6 GB (activations stashed for backward) > 126 MB L2, no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: the SAME per-GPU workload (2 x 7680) at 1 + 2 steps keeps the run within minutes
        steps = min(args.steps, 3)
        r = cpu_reference_run(hp, steps, min(args.warmup, 1), B_PER_GPU, T_STEP)
        line = {"impl": "reference", "metric": "wavenet_train_audio_samples_per_sec", "value": r["value"], "unit": "samples/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "oracle/wavenet.py (fp32 PyTorch-CPU restatement of the reference TF1 graph; TF1 is not installable here)"}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from t2_import import t2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = t2.lib.load()
    model = t2.wavenet.WaveNet(hp, B_PER_GPU, T_STEP, device=dev)
    model.init_variables(seed=5339)
    idx, c, lengths = synth_batch(hp, B_PER_GPU, T_STEP, 2 + rank,
                                  lambda w: t2.audio.mulaw_quantize(torch.from_numpy(w).to(dev)).cpu().numpy())
    pin = [torch.from_numpy(a).pin_memory() for a in (idx, c, idx, lengths)]
    static = [p.to(dev) for p in pin]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    if not args.no_graph:
        model.capture(*static)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, e2e):
        for _ in range(n):
            if e2e:
                model.train_step(*pin, world_size=world)
                loss_host.copy_(model.loss_buf, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            else:
                model.train_step(*(static if args.no_graph else (None, None, None, None)), world_size=world)

    results = {}
    sampler = None
    clocks = None
    for mode in ("resident", "e2e"):
        run_steps(args.warmup, mode == "e2e")
        barrier()
        if mode == "resident":
            sampler = ClockSampler(local_rank)
            if rank == 0:
                sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        run_steps(args.steps, mode == "e2e")
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if mode == "resident" and rank == 0:
            clocks = sampler.stop()
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[mode] = t.item()
    loss = model.loss_value()

    # roofline leg: the dominant kernel is the per-layer gate GEMM (24 launches / step, ~2/3 of the forward FLOPs)
    gate_ms = [model.time_kernel(0, l, reps=20) for l in (3, 9, 15, 21)]
    gate_ms_avg = sum(gate_ms) / len(gate_ms)
    R, G, C = hp.residual_channels, hp.gate_channels, hp.cin_channels
    flops_per_launch = 2.0 * B_PER_GPU * T_STEP * G * (3 * R + C)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("gate_gemm_dram_bytes_per_launch")
    except Exception:
        pass
    achieved = flops_per_launch / (gate_ms_avg * 1e-3) / 1e12

    if rank == 0:
        total_samples = world * B_PER_GPU * T_STEP * args.steps
        value = total_samples / (results["resident"] * 1e-3)
        e2e_value = total_samples / (results["e2e"] * 1e-3)
        # bounded CPU baseline (rank 0, N = 1 only): one warm-up + one timed oracle step on the same workload
        cpu = None
        if world == 1:
            r = cpu_reference_run(hp, 1, 1, B_PER_GPU, T_STEP)
            cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {
            "metric": "wavenet_train_audio_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": results["resident"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config, "loss": loss, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s",
                    "h2d_bytes_per_step": int(sum(p.numel() * p.element_size() for p in pin)), "d2h_bytes_per_step": 8,
                    "ms_per_step": results["e2e"] / args.steps},
            "gpu_launches": int(model.launches_per_step * args.steps),
            "roofline": {"bound": "tensor", "kernel": "act_gemm_kernel<EPI_GATE,256,NT=2> (per-layer dilated-conv + cin gate GEMM)",
                         "timing": "CUDA events around 20 back-to-back launches replayed from one CUDA graph, averaged over layers 3/9/15/21",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "flops_per_launch": flops_per_launch, "ms_per_launch": gate_ms_avg,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
