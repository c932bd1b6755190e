"""Phase timing (clock64 stamps, CTA 0) of the Tacotron attention kernels at Cfg-3 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hparams import hparams
from t2_import import t2
from tools.bench_taco import batch

hp = hparams.copy()
hp.parse("predict_linear=False")
B, T_in, T_out = 32, 160, 40
inputs, lens, mel, stop = batch(hp, B, T_in, T_out)
model = t2.tacotron.Tacotron(hp, B, T_in, T_out)
model.init_variables(seed=5339)
args = (inputs.int().cuda(), lens.int().cuda(), mel.cuda(), stop.cuda())
model.train_step(*args)
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
lib = t2.lib.load()
t2.lib.check(lib.t2_dbg_att_stamps(t2.lib.ptr(buf)))
model.train_step(*args)
torch.cuda.synchronize()
t = buf.cpu().tolist()
t2.lib.check(lib.t2_dbg_att_stamps(None))
f = [t[i + 1] - t[i] for i in range(6)]
b = [t[i + 1] - t[i] for i in range(16, 23)]
print("att_fwd cycles: load %d | query %d | energies %d | softmax %d | context %d | write %d | total %d" % (*f, t[6] - t[0]))
print("att_bwd cycles: load %d | query %d | dalpha %d | softmax-bwd+energies %d | dq/dU %d | dcum %d | dh2ext %d | total %d" % (*b, t[23] - t[16]))
