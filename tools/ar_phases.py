"""clock64 phase stamps of one layer pass (t = 64, l = 7, CTA 0) of the AR synthesis kernel, paper widths, B = 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hparams import hparams
from t2_import import t2

hp = hparams.copy()
hp.parse("layers=24,stacks=4,residual_channels=256,gate_channels=512,skip_out_channels=256,upsample_scales=[11,25],"
         "input_type=mulaw-quantize,quantize_channels=256,out_channels=256")
for cs in (16, 8):
    T = 275 * 4
    syn = t2.wavenet.WaveNetSynthesizer(hp, 1, T, cluster_size=cs)
    syn.init_variables(seed=5)
    c = torch.rand(1, 80, T // 275, device="cuda")
    init = torch.full((1,), 127, dtype=torch.int32).cuda()
    syn.generate(c, init, seed=1)
    buf = torch.zeros(16, dtype=torch.int64, device="cuda")
    lib = t2.lib.load()
    t2.lib.check(lib.t2_dbg_ar_stamps(t2.lib.ptr(buf)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); syn.generate(c, init, seed=2); e1.record()
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    t2.lib.check(lib.t2_dbg_ar_stamps(None))
    names = ["gather", "publish+wait weights", "matvec1", "gate+bcast", "cluster.sync", "matvec2", "out+bcast", "cluster.sync", "copy"]
    print("CS=%d  %.1f us/sample | " % (cs, 1e3 * e0.elapsed_time(e1) / T) + " | ".join("%s %d" % (n, t[i + 1] - t[i]) for i, n in enumerate(names)) + " | layer total %d" % (t[9] - t[0]))
