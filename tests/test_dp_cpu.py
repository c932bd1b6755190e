"""Data-parallel semantics on CPU (gloo, world_size 2): sum-all-reduce of per-rank gradients scaled by 1/N equals the
gradient of the mean of the tower losses, and clip-then-Adam on the averaged gradient keeps replicas bit-identical —
the order the reference uses (wavenet.py:561-613: average towers, clip, Adam)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hparams import hparams
    from oracle import wavenet as ow
    hp = hparams.copy()
    hp.parse("input_type=mulaw-quantize,quantize_channels=16,out_channels=16,layers=2,stacks=1,residual_channels=8,"
             "gate_channels=16,skip_out_channels=8,upsample_scales=[2,2],hop_size=4,cin_channels=4,num_mels=4")
    torch.manual_seed(0)
    params = ow.init_params(hp, seed=3)
    g = torch.Generator().manual_seed(100)            # the SAME global batch on every rank, sharded by rank
    idx = torch.randint(0, 16, (2 * world, 16), generator=g)
    c = torch.rand(2 * world, 4, 4, generator=g)
    lengths = torch.full((2 * world,), 16)
    sl = slice(2 * rank, 2 * rank + 2)
    x = torch.nn.functional.one_hot(idx[sl], 16).float().transpose(1, 2)
    loss, grads, _ = ow.train_step(params, x, c[sl], idx[sl], lengths[sl], hp)
    names = sorted(grads)
    flat = torch.cat([grads[k].reshape(-1) for k in names])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world
    off = 0
    avg = {}
    for k in names:
        n = grads[k].numel()
        avg[k] = flat[off:off + n].view_as(grads[k]).clone()
        off += n
    state = {}
    ow.adam_step(params, avg, state, hp, 0)
    pflat = torch.cat([params[k].reshape(-1) for k in names])
    gathered = [torch.zeros_like(pflat) for _ in range(world)]
    dist.all_gather(gathered, pflat)
    if rank == 0:
        # single-process reference: mean of the two tower losses
        p0 = ow.init_params(hp, seed=3)
        tot = {k: torch.zeros_like(v) for k, v in p0.items()}
        for r in range(world):
            s2 = slice(2 * r, 2 * r + 2)
            xr = torch.nn.functional.one_hot(idx[s2], 16).float().transpose(1, 2)
            _, gr, _ = ow.train_step(p0, xr, c[s2], idx[s2], lengths[s2], hp)
            for k in tot:
                tot[k] += gr[k] / world
        err = max((tot[k] - avg[k]).abs().max().item() for k in names)
        same = all(torch.equal(gathered[0], gathered[r]) for r in range(world))
        torch.save({"err": err, "same": same}, out)
    dist.destroy_process_group()


def test_allreduce_mean_then_clip_then_adam(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["err"] < 1e-6
    assert res["same"]


def _bucket_worker(rank, world, port, out):
    """PRODUCT host logic of the overlapped data-parallel step (tacotron-2_b200/wavenet.py train_step): the gradient buffer is reduced in
    the bucket ranges of grad_buckets(), in the order the captured graphs finish them; here on CPU tensors over gloo"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from t2_import import t2
    from bench import workload_hparams
    hp = workload_hparams("wavenet_ce")
    cfg = t2.wavenet.make_config(hp, 2, 7680)
    tensors, n_params = t2.wavenet.param_table(cfg)
    g = torch.Generator().manual_seed(10 + rank)
    grads = torch.randn(n_params, generator=g)
    mono = grads.clone()
    dist.all_reduce(mono, op=dist.ReduceOp.SUM)
    res = {}
    for G in (1, 3, 6):
        groups, rest = t2.wavenet.grad_buckets(tensors, hp.layers, n_params, G)
        cover = sorted(groups + rest)
        assert cover[0][0] == 0 and cover[-1][1] == n_params and all(a[1] == b[0] for a, b in zip(cover, cover[1:])), cover
        buf = grads.clone()
        works = [dist.all_reduce(buf[a:b], op=dist.ReduceOp.SUM, async_op=True) for a, b in groups + rest]
        for w in works:
            w.wait()
        res[G] = torch.equal(buf, mono)
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_product_gradient_buckets_over_gloo(tmp_path):
    out = str(tmp_path / "buckets.pt")
    mp.spawn(_bucket_worker, args=(2, 31500 + os.getpid() % 2000, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res == {1: True, 3: True, 6: True}
