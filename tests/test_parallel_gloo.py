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
    red = parallel.make_reducer(fg.flat)            # whole-buffer reduction (no marks): the generator's case
    assert red is not None and red.mode == "bf16"
    red.begin()
    g, is_bf16 = red.finish()
    ok = is_bf16 and g.dtype == torch.bfloat16 and red.slices == [(0, fg.flat.numel())] and all(
        bool((g[o:o + p.numel()].float() == 1.5).all()) for p, o in zip(fg.params, fg.offsets))
    red32 = parallel.GradReducer(fg.flat, mode="fp32")
    red32.begin()
    g32, is16 = red32.finish()
    ok = ok and (not is16) and g32 is fg.flat and all(torch.allclose(p.grad, torch.full_like(p.grad, 1.5)) for p in net.parameters())
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
    res = {}
    for mode in ("fp32", "bf16"):
        red = parallel.GradReducer(fg.flat, net, fg.params, fg.offsets, mode=mode, min_bytes=4)
        ops.GRAD_MARKS[id(net)] = red.on_mark
        x = torch.randn(4, 5, generator=torch.Generator().manual_seed(100 + rank))      # per-rank batch
        # marks fired while the reducer is idle (another network's backward) must do nothing
        fg.zero()
        net(x).square().sum().backward()
        idle_ok = red.slices == [] and red.lo == red.total
        local = fg.flat.clone()
        fg.zero()
        red.begin()
        net(x).square().sum().backward()
        in_flight = len(red.slices)                     # slices that went out DURING backward
        g, is_bf16 = red.finish()
        gathered = [torch.zeros_like(local) for _ in range(ws)]
        dist.all_gather(gathered, local)
        want = sum(gathered) / ws
        if mode == "bf16":
            want = sum(t.bfloat16().float() for t in gathered) / ws          # each rank's payload is rounded once
        cover = sorted(red.slices)
        res[mode] = dict(idle_ok=idle_ok, in_flight=in_flight, n_slices=len(red.slices), is_bf16=is_bf16,
                         err=float((g.float() - want).abs().max()), scale=float(want.abs().max()),
                         covers=cover[0][0] == 0 and cover[-1][1] == red.total and all(a[1] == b[0] for a, b in zip(cover, cover[1:])))
        ops.GRAD_MARKS.clear()
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_tail_first_reduction_equals_flat_allreduce(tmp_path):
    """parallel.GradReducer (overlap of a network's gradient exchange with its own backward): slices leave from the
    tail of the flat buffer as marks fire, every element is reduced exactly once, and the result equals the mean of the
    ranks' gradients -- exactly in fp32 mode, to bf16 rounding of the payload in bf16 mode."""
    out = str(tmp_path / "t.pt")
    mp.spawn(_tail_worker, args=(2, 29613, out), nprocs=2, join=True)
    res = torch.load(out)
    for mode, r in res.items():
        assert r["idle_ok"] and r["covers"], (mode, r)
        assert r["in_flight"] == 2 and r["n_slices"] == 3, (mode, r)     # b3 tail, then b2, then the head (b1) in finish()
        assert r["is_bf16"] == (mode == "bf16")
        assert r["err"] <= (1e-6 if mode == "fp32" else 2 ** -8 * r["scale"]), (mode, r)


def test_gloo_world_size_2(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, 29611, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"], "mean over the flat gradient buffer (bf16 payload and fp32 in place)"
    assert r["same"], "broadcast makes replicas identical"
    assert r["shard"] == (0, 5)
    assert r["layout"], "gradient views keep the parameter's (channels_last) layout"


def test_single_process_has_no_collective():
    sys.path.insert(0, ROOT)
    from text2img_ekl_b200 import parallel
    assert parallel.make_reducer(torch.zeros(8)) is None
    assert parallel.shard_range(7, 0, 2) == (0, 4) and parallel.shard_range(7, 1, 2) == (4, 7)


def _replicate_worker(rank, ws, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    from text2img_ekl_b200 import cub_trainer_splitz_cap_ca as cub
    tr = cub.condGANTrainer.__new__(cub.condGANTrainer)        # the constructor selects a CUDA device; the method under test does not
    tr.device = torch.device("cpu")
    torch.manual_seed(10 + rank)                               # weights_init draws differ per process, as in a real launch
    tr.netG = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5))
    tr.netsD = [torch.nn.Linear(5, 3), torch.nn.Linear(5, 2)]
    for n in [tr.netG] + tr.netsD:
        n.apply(cub.weights_init)
    tr.netG[1].running_mean.fill_(float(rank))
    torch.manual_seed(0)                                       # the user's global seed is the same on every rank
    tr._replicate()
    flat = torch.cat([t.detach().flatten() for n in [tr.netG] + tr.netsD for t in list(n.parameters()) + list(n.buffers())]).float()
    gathered = [torch.zeros_like(flat) for _ in range(ws)]
    dist.all_gather(gathered, flat)
    draw = torch.randn(4)                                      # in-step noise / eps / seed draws must differ per rank
    draws = [torch.zeros(4) for _ in range(ws)]
    dist.all_gather(draws, draw)
    if rank == 0:
        torch.save(dict(same=torch.equal(gathered[0], gathered[1]), differ=not torch.equal(draws[0], draws[1]),
                        world=(tr.rank, tr.world_size)), out)
    dist.barrier()
    dist.destroy_process_group()


def test_trainer_setup_replicates_rank0_and_offsets_rng(tmp_path):
    """condGANTrainer.setup() -> _replicate(): the public multi-GPU entry joins the process group by itself, every
    replica starts from rank 0's weights AND buffers (what nn.DataParallel's replicate does, cub:139,163), and the
    per-rank RNG streams differ so that the replicas draw different noise."""
    out = str(tmp_path / "s.pt")
    mp.spawn(_replicate_worker, args=(2, 29621, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["same"] and r["differ"] and r["world"] == (0, 2), r


# ---------------------------------------------------------------- bench.py control flow under two ranks
def _bench_worker(rank, ws, port, out, asymmetric):
    """bench.drive (the part of bench.py that runs steps) with a stub step: every step all-reduces, like a real
    data-parallel step.  `asymmetric` re-creates round 1's bug (a step executed by rank 0 only)."""
    import argparse
    import datetime
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import bench
    dist.init_process_group("gloo", init_method="env://", timeout=datetime.timedelta(seconds=8))
    calls = [0]

    def step(i=0):
        t = torch.ones(4) * (rank + 1)
        dist.all_reduce(t)
        calls[0] += 1
        assert float(t[0]) == 3.0

    def roofline():
        step()
        return {"frac": 0.5}
    a = argparse.Namespace(steps=3, warmup=2, no_profile=False)
    ok, err = True, ""
    try:
        if asymmetric:
            ms, ms2, roof = bench.drive(step, step, roofline if rank == 0 else None, a, ws, on_gpu=False)
        else:
            ms, ms2, roof = bench.drive(step, step, roofline, a, ws, on_gpu=False)
        dist.barrier()
    except Exception as ex:  # noqa: BLE001
        ok, err = False, str(ex)[:200]
    torch.save(dict(ok=ok, err=err, calls=calls[0]), "%s.%d" % (out, rank))
    if ok:
        dist.destroy_process_group()
    else:
        os._exit(0)


def test_bench_control_flow_is_rank_symmetric(tmp_path):
    """bench.py at N > 1: both timed regions AND the roofline pass run on every rank (each issues collectives).  The
    stub step all-reduces; the symmetric flow finishes with the same number of steps on both ranks, and the flow with a
    rank-0-only pass (round 1's deadlock) is detected by the collective time-out instead of passing silently."""
    out = str(tmp_path / "b.pt")
    mp.spawn(_bench_worker, args=(2, 29617, out, False), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert r0["ok"] and r1["ok"], (r0, r1)
    assert r0["calls"] == r1["calls"] == (2 + 3) + (2 + 3) + 1, (r0, r1)
    mp.spawn(_bench_worker, args=(2, 29619, out, True), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert not (r0["ok"] and r1["ok"]), "a rank-asymmetric step must not go unnoticed"


def test_bench_roofline_runs_on_every_rank():
    """Static guard for the same property in bench.run_ours / measure: nothing that runs a step sits under a
    `rank == 0` condition (only the sampler, the JSON print and the single-GPU baselines do)."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and "rank == 0" in ast.unparse(node.test) and "ws == 1" not in ast.unparse(node.test):
            body = ast.unparse(node)
            for name in ("kernel_roofline", "train_step", "replay(", "drive(", "timed_region(", "d_step", "g_step"):
                if name in body:
                    bad.append((node.lineno, name))
    assert not bad, bad
