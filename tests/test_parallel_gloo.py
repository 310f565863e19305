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


def _tail_worker(rank, ws, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    from text2img_ekl_b200 import ops, parallel
    from text2img_ekl_b200.engine import FlatGrads
    assert parallel.init_from_env("gloo") == (rank, ws)

    class Net(torch.nn.Module):                     # three blocks, marks behind the first two (like _DBase._trunk)
        def __init__(self):
            super().__init__()
            self.b1, self.b2, self.b3 = torch.nn.Linear(5, 7), torch.nn.Linear(7, 6), torch.nn.Linear(6, 3)

        def forward(self, x):
            x = ops.grad_mark(torch.tanh(self.b1(x)), self, self.b1)
            x = ops.grad_mark(torch.tanh(self.b2(x)), self, self.b2)
            return self.b3(x)

    torch.manual_seed(0)
    net = Net()
    fg = FlatGrads(net.parameters())
    offs, off = [], 0
    for p in fg.params:
        offs.append(off)
        off += p.numel()
    tail = parallel.TailAllreduce(net, fg.params, offs, off, fg.flat, min_bytes=4)
    ops.GRAD_MARKS[id(net)] = tail.on_mark
    x = torch.randn(4, 5, generator=torch.Generator().manual_seed(100 + rank))      # per-rank batch
    # marks fired while the bucketer is idle (another network's backward) must do nothing
    net(x).square().sum().backward()
    idle_ok = tail.works == [] and tail.lo == off
    local = fg.flat.clone()
    fg.zero()
    tail.begin()
    net(x).square().sum().backward()
    in_flight = len(tail.works)                     # slices that went out DURING backward
    n_slices = tail.finish()
    gathered = [torch.zeros_like(local) for _ in range(ws)]
    dist.all_gather(gathered, local)
    want = sum(gathered) / ws
    if rank == 0:
        torch.save(dict(idle_ok=idle_ok, in_flight=in_flight, n_slices=n_slices, err=float((fg.flat - want).abs().max()),
                        ends=sorted(set(tail.end_of.values())), total=off), out)
    ops.GRAD_MARKS.clear()
    dist.barrier()
    dist.destroy_process_group()


def test_tail_first_allreduce_equals_flat_allreduce(tmp_path):
    """parallel.TailAllreduce (opt-in overlap of a network's gradient all-reduce with its own backward): slices leave
    from the tail of the flat buffer as marks fire, every element is reduced exactly once, and the result equals the
    mean of the ranks' gradients."""
    out = str(tmp_path / "t.pt")
    mp.spawn(_tail_worker, args=(2, 29613, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["idle_ok"]
    assert r["in_flight"] == 2 and r["n_slices"] == 3, r          # b3 tail, then b2, then the head (b1) in finish()
    assert r["err"] < 1e-6, r


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
