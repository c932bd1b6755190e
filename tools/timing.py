"""Debug tool (GPU box): per-CTA phase timing of the GEMM kernels via clock64 stamps."""
import ctypes, sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
buf = torch.zeros(512 * 16, dtype=torch.int64, device='cuda')
lib.t2_dbg_set_timing_buffer(L.ptr(buf))
names = ['entry->setup', 'setup->first stage', 'first stage->MMAs issued', 'MMAs issued->acc ready', 'acc ready->epi done', 'epi done->teardown']
def show(tag, ncta):
    torch.cuda.synchronize()
    t = buf.view(-1, 16)[:ncta].double().cpu()
    d = [(t[:, 1] - t[:, 0]), (t[:, 2] - t[:, 1]), (t[:, 3] - t[:, 2]), (t[:, 4] - t[:, 3]), (t[:, 5] - t[:, 4]), (t[:, 6] - t[:, 5])]
    print(tag, 'total %.0f cyc |' % (t[:, 6] - t[:, 0]).mean().item(), ' | '.join('%s %.0f' % (n, v.mean().item()) for n, v in zip(names, d)))
    buf.zero_()
for which, tag, n in ((0, 'gate GEMM  K=896  N=512', 240), (1, 'out GEMM   K=256  N=256', 120), (2, 'dz GEMM    K=512  N=256', 120), (3, 'dx GEMM    K=1536 N=256', 120)):
    lib.t2_dbg_set_timing_buffer(None)
    ms = m.time_kernel(which, 9, reps=50)          # device time per launch: 50 launches in one CUDA graph
    lib.t2_dbg_set_timing_buffer(L.ptr(buf))
    lib.t2_dbg_set_timing_buffer(L.ptr(buf))
    m.time_kernel(which, 9, reps=1)                # warm-up launch + one stamped graph launch: read the LAST slice
    show('%s  (%.2f us/launch back-to-back in a graph)' % (tag, ms * 1e3), n)
# out GEMM etc. via a full forward: the buffer keeps the LAST kernel that ran with <= 512 CTAs (the CE head)
lib.t2_dbg_set_timing_buffer(None)
