"""World-size-2 data-parallel plumbing on CPU (gloo): gradient averaging over the flat buffer, parameter broadcast,
shard ranges.  (The NCCL path is the same code with backend='nccl'; bench.py exercises it on GPUs.)"""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, ws, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    from text2img_ekl_b200 import parallel
    from text2img_ekl_b200.engine import FlatGrads
    assert parallel.init_from_env("gloo") == (rank, ws)
    torch.manual_seed(rank)                       # different init per rank on purpose
    net = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Linear(3, 2))
    net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last)
    parallel.broadcast_params([net])
    w0 = net[0].weight.detach().clone()
    fg = FlatGrads(net.parameters())
    for p in net.parameters():
        p.grad.fill_(float(rank + 1))             # rank 0 -> 1, rank 1 -> 2 : mean 1.5
    ar = parallel.make_allreduce()
    ar(fg.flat)
    ok = all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5)) for p in net.parameters())
    gathered = [torch.zeros_like(w0) for _ in range(ws)]
    dist.all_gather(gathered, w0)
    same = torch.equal(gathered[0], gathered[1])
    lo, hi = parallel.shard_range(10)
    if rank == 0:
        torch.save(dict(ok=ok, same=same, shard=(lo, hi), layout=net[0].weight.grad.stride() == net[0].weight.stride()), out)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world_size_2(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, 29611, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"], "all-reduce(mean) over the flat gradient buffer"
    assert r["same"], "broadcast makes replicas identical"
    assert r["shard"] == (0, 5)
    assert r["layout"], "gradient views keep the parameter's (channels_last) layout"


def test_single_process_has_no_collective():
    sys.path.insert(0, ROOT)
    from text2img_ekl_b200 import parallel
    assert parallel.make_allreduce() is None
    assert parallel.shard_range(7, 0, 2) == (0, 4) and parallel.shard_range(7, 1, 2) == (4, 7)
