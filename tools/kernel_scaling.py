"""Debug tool (GPU box): per-launch time of the four per-layer GEMMs of the paper-width WaveNet at different batch sizes
(60 / 120 / 240 / 480 CTAs): separates per-SM latency (time independent of the CTA count) from shared-resource limits."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2_import import t2
from bench import workload_hparams, synth_batch

wl = sys.argv[1] if len(sys.argv) > 1 else "wavenet_ce"
hp = workload_hparams(wl)
T = 7680
rows = []
for B in (1, 2, 4, 8):
    m = t2.wavenet.WaveNet(hp, B, T)
    m.init_variables(seed=1)
    idx, c, lengths = synth_batch(hp, B, T, 2, lambda w: t2.audio.mulaw_quantize(torch.from_numpy(w).cuda()).cpu().numpy())
    x = torch.from_numpy(idx).cuda(); cc = torch.from_numpy(c).cuda(); ln = torch.from_numpy(lengths).cuda()
    for _ in range(2):
        m.forward(x, cc, x, ln); m.backward()
    torch.cuda.synchronize()
    r = {"B": B, "ctas": B * T // 128}
    for which, name in enumerate(("gate", "out", "dz", "dx")):
        r[name + "_us"] = 1e3 * sum(m.time_kernel(which, l, 20) for l in (3, 9, 15)) / 3
    rows.append(r)
    print(json.dumps(r), flush=True)
    del m
    torch.cuda.empty_cache()
