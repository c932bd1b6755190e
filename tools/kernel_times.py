"""Debug tool (GPU box): isolated device time of the four per-layer GEMMs of the WaveNet step (t2_wn_time_kernel)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t2_import import t2
from bench import workload_hparams, synth_batch, WN_SHAPES
name = sys.argv[1] if len(sys.argv) > 1 else "wavenet_ce"
hp = workload_hparams(name)
B, T = WN_SHAPES[name]
m = t2.wavenet.WaveNet(hp, B, T)
m.init_variables(seed=1)
q = None if hp.input_type != "mulaw-quantize" else (lambda w: t2.audio.mulaw_quantize(torch.from_numpy(w).cuda()).cpu().numpy())
x, c, ln = (torch.from_numpy(a).cuda() for a in synth_batch(hp, B, T, 2, q))
for _ in range(2):
    m.forward(x, c, x, ln); m.backward()
torch.cuda.synchronize()
out = {"lib": os.environ.get("T2_LIB", "default"), "cluster": os.environ.get("T2_CLUSTER", "1")}
for which, tag in ((0, "gate"), (4, "gate_nostash"), (1, "out"), (2, "dz"), (3, "dx")):
    out[tag + "_us"] = round(1e3 * sum(m.time_kernel(which, l, reps=20) for l in (3, 9, 15)) / 3, 2)
print(json.dumps(out))
