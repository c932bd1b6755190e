"""Debug tool (GPU box): where does the time between the per-layer GEMM launches of the captured WaveNet step go?
Every act_gemm CTA stamps %globaltimer at entry / after griddepcontrol.wait / at exit (t2_dbg_set_timing_buffer); the captured
graph gives each kernel node its own slice, so one replay yields a device-side timeline of the ~100 dependent GEMM launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2_import import t2
from bench import workload_hparams, synth_batch, B_PER_GPU, T_STEP

L = t2.lib
lib = L.load()
hp = workload_hparams()
m = t2.wavenet.WaveNet(hp, B_PER_GPU, T_STEP)
m.init_variables(seed=1)
idx, c, lengths = synth_batch(hp, B_PER_GPU, T_STEP, 2, lambda w: t2.audio.mulaw_quantize(torch.from_numpy(w).cuda()).cpu().numpy())
x = torch.from_numpy(idx).cuda(); cc = torch.from_numpy(c).cuda(); ln = torch.from_numpy(lengths).cuda()
for _ in range(2):
    m.forward(x, cc, x, ln); m.backward()
torch.cuda.synchronize()
NL = hp.layers
names = []
for l in range(NL):
    names.append("gate%02d" % l)
    if l + 1 < NL:
        names.append("out%02d" % l)
names += ["skip", "final1", "ce_head", "dh2", "dskip"]
for l in range(NL - 1, -1, -1):
    names += ["dz%02d" % l, "dx%02d" % l]
names += ["dcond"]
n_launch, ncta, slots = len(names), 120, 16
buf = torch.zeros(3 * n_launch * ncta * slots, dtype=torch.int64, device="cuda")
lib.t2_dbg_set_timing_buffer(L.ptr(buf))
m.capture(x, cc, x, ln)            # warm-up pass = slices [0, n), captured pass = slices [n, 2n)
lib.t2_dbg_set_timing_buffer(None)
for _ in range(3):
    m._graph.replay()
torch.cuda.synchronize()
t = buf.view(-1, ncta, slots)[n_launch:2 * n_launch].cpu().double()
t0 = t[0, :, 8].min().item()
rows = []
prev_exit = None
for i, name in enumerate(names):
    e, w, x_ = t[i, :, 8] - t0, t[i, :, 9] - t0, t[i, :, 10] - t0
    cyc = (t[i, :, 6] - t[i, :, 0])
    ph = [(t[i, :, k + 1] - t[i, :, k]).mean().item() for k in range(6)]     # clock64 phases: setup, first stage, MMA issue, acc ready, epilogue, teardown
    rows.append({"launch": name, "entry_min_us": e.min().item() / 1e3, "entry_max_us": e.max().item() / 1e3,
                 "wait_done_min_us": w.min().item() / 1e3, "wait_done_max_us": w.max().item() / 1e3,
                 "exit_min_us": x_.min().item() / 1e3, "exit_max_us": x_.max().item() / 1e3,
                 "cta_life_us_mean": (x_ - e).mean().item() / 1e3, "cta_busy_us_mean": (x_ - w).mean().item() / 1e3,
                 "cta_cycles_mean": cyc.mean().item(),
                 "gap_prev_exit_to_wait_done_us": None if prev_exit is None else (w.min().item() - prev_exit) / 1e3,
                 "n_sms": int(t[i, :, 11].unique().numel()), "phase_cycles": ph,
                 "epi_start_after_entry_cycles": (t[i, :, 4] - t[i, :, 0]).mean().item()})
    prev_exit = x_.max().item()
print("%-9s %9s %9s %9s %9s %9s %9s %8s %8s %8s" % ("launch", "entry_min", "entry_max", "wait_min", "wait_max", "exit_min", "exit_max", "busy", "gap", "period"))
for i, r in enumerate(rows):
    period = rows[i + 1]["wait_done_min_us"] - r["wait_done_min_us"] if i + 1 < len(rows) else float("nan")
    print("%-9s %9.2f %9.2f %9.2f %9.2f %9.2f %9.2f %8.2f %8s %8.2f" % (
        r["launch"], r["entry_min_us"], r["entry_max_us"], r["wait_done_min_us"], r["wait_done_max_us"], r["exit_min_us"], r["exit_max_us"],
        r["cta_busy_us_mean"], "-" if r["gap_prev_exit_to_wait_done_us"] is None else "%.2f" % r["gap_prev_exit_to_wait_done_us"], period))
print("\nclock64 phases (mean cycles per CTA): setup | wait first stage | MMA issue (all k-blocks) | last MMA -> accumulator ready | epilogue (first half start -> last half end; overlaps MMA issue when NT=2) | teardown")
for r in rows:
    if r["launch"] in ("gate05", "out05", "gate17", "out17", "skip", "final1", "ce_head", "dz17", "dx17", "dz05", "dx05"):
        print("%-8s" % r["launch"], " ".join("%8.0f" % v for v in r["phase_cycles"]), " | epilogue starts %.0f cycles after entry, CTA %.0f cycles" % (r["epi_start_after_entry_cycles"], r["cta_cycles_mean"]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gap_probe.json", "w"), indent=1)
