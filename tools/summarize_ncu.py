"""Condenses the ncu outputs of tools/profile_r2.sh (gpurun_out/) into the tracked summaries under profiles/:
  r02_launch_summary_<workload>.txt   per-kernel launch counts, time share and DRAM bytes of ONE training step (cold-cache, serialised)
  r02_dram_traffic.json               measured DRAM read + write bytes per step / per gate-GEMM launch (bench.py `roofline.traffic`)
  r02_ncu_full_<kernel>.txt           the counters the north star asks for (tensor-pipe %, DRAM / L2 throughput, achieved GB/s, stalls)
Run here (no GPU needed): python tools/summarize_ncu.py"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"\(t2::GemmArgs\)|\(t2::WgradArgs\)|t2::\(anonymous namespace\)::|t2::|void ", "", name)
    name = re.sub(r"\(int\)", "", name)
    return re.sub(r"\(.*\)$", "", name)[:70]


def launches(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    ui = h.index("Metric Unit")
    out = collections.OrderedDict()
    for r in rows[1:]:
        d = out.setdefault(r[ii], {"name": r[ki]})
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        if "byte" in unit:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if r[mi].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)          # -> microseconds
        d[r[mi]] = v
    return list(out.values())


def one_step(ls, marker):
    idx = [i for i, l in enumerate(ls) if marker in l["name"]]
    if len(idx) >= 2:
        return ls[idx[0]:idx[1]]
    return ls


def summarize_list(workload, marker):
    path = os.path.join(GO, "r2_launches_%s.csv" % workload)
    if not os.path.exists(path):
        return None
    ls = one_step(launches(path), marker)
    agg = collections.OrderedDict()
    for l in ls:
        a = agg.setdefault(short(l["name"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += l.get("gpu__time_duration.sum", 0.0)
        a[2] += l.get("dram__bytes_read.sum", 0.0)
        a[3] += l.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values()) or 1.0
    lines = ["# one training step of `bench.py --workload %s --no-graph` under ncu (--cache-control none, --clock-control none):" % workload,
             "# per-launch times are serialised (no overlap between streams / dependent launches): compare SHARES, not absolutes",
             "%-70s %6s %10s %7s %12s %12s" % ("kernel", "count", "time_us", "share", "dram_rd_MB", "dram_wr_MB")]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-70s %6d %10.1f %6.1f%% %12.2f %12.2f" % (k, a[0], a[1], 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6))
    rd, wr = sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
    lines.append("%-70s %6d %10.1f %6.1f%% %12.2f %12.2f" % ("TOTAL", sum(a[0] for a in agg.values()), tot, 100.0, rd / 1e6, wr / 1e6))
    open(os.path.join(PR, "r02_launch_summary_%s.txt" % workload), "w").write("\n".join(lines) + "\n")
    gate = [l for l in ls if re.search(r"act_gemm2?_kernel<(\(int\))?0,", l["name"])]
    res = {workload + "_step_dram_bytes": rd + wr, workload + "_step_dram_read_bytes": rd, workload + "_step_dram_write_bytes": wr,
           workload + "_step_kernel_time_us_serialised": tot}
    if gate:
        res[workload + "_gate_dram_bytes_per_launch"] = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in gate) / len(gate)
        res[workload + "_gate_time_share"] = sum(l.get("gpu__time_duration.sum", 0) for l in gate) / tot
    return res


WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "launch__cluster_size"]


def summarize_full(tag):
    rep = os.path.join(GO, "r2_full_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    h, units = rows[0], rows[1]
    lines = ["# ncu --set full --clock-control none, %s (gpurun_out/r2_full_%s.ncu-rep; one row per captured launch)" % (tag, tag)]
    for r in rows[2:]:
        lines.append("kernel: " + short(r[h.index("Kernel Name")]))
        vals = {}
        for m in WANT:
            if m in h:
                lines.append("  %-75s %s %s" % (m, r[h.index(m)], units[h.index(m)]))
                vals[m] = (r[h.index(m)], units[h.index(m)])
        try:
            t = float(vals["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}[vals["gpu__time_duration.sum"][1]]
            by = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                by += float(vals[m][0].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[vals[m][1]]
            lines.append("  %-75s %.1f GB/s" % ("achieved DRAM bandwidth (read + write bytes / duration)", by / t / 1e9))
        except Exception:
            pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    hdr = next((r for r in srows if r and r[0] == "Address"), None)
    if hdr:
        si = hdr.index("# Samples")
        stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        body = [r for r in srows if len(r) == len(hdr) and r[0].startswith("0x")]
        half = body[:len(body) // 2] if len(body) > 10 else body            # two launches captured: the first one's instructions
        tot = sum(int(r[si] or 0) for r in half) or 1
        agg = sorted(((hdr[i], sum(int(r[i] or 0) for r in half)) for i in stall), key=lambda kv: -kv[1])[:6]
        lines.append("  warp-stall samples (first captured launch, %d samples): " % tot + ", ".join("%s %.0f%%" % (k, 100.0 * v / tot) for k, v in agg))
    open(os.path.join(PR, "r02_ncu_full_%s.txt" % tag), "w").write("\n".join(lines) + "\n")


def main():
    traffic = {"source": "tools/profile_r2.sh: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none over one eager "
                         "training step (profiles/r02_launch_summary_*.txt)"}
    for wl, marker in (("wavenet_ce", "pack_kernel"), ("wavenet_default", "pack_kernel"), ("tacotron", "tpack_kernel")):
        r = summarize_list(wl, marker)
        if r:
            traffic.update(r)
    if len(traffic) > 1:
        path = os.path.join(PR, "r02_dram_traffic.json")
        try:                       # keep the figures of workloads whose launch list is not in gpurun_out/ this time
            old = json.load(open(path))
            traffic = dict({k: v for k, v in old.items() if k not in traffic}, **traffic)
        except Exception:
            pass
        json.dump(traffic, open(path, "w"), indent=1)
    for tag in ("gate", "out", "dz", "dx", "wgrad", "lstm", "tout", "attf", "attb", "ar", "stft", "gru"):
        summarize_full(tag)
    print("\n".join(sorted(f for f in os.listdir(PR) if f.startswith("r02_"))))


if __name__ == "__main__":
    main()
