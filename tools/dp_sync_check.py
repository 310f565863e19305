#!/usr/bin/env python
"""Run under torchrun with N ranks on N GPUs: a few training steps of a config on DIFFERENT per-rank batches, then check
that every replica holds bit-identical parameters, Adam moments and bf16 shadows (identical averaged gradients + identical
element-wise updates), i.e. that the tail-first reducer / overlapped optimiser keep the replicas in sync.  Prints one JSON
line on rank 0; exit code 1 on divergence.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_sync_check.py [--config 3stages]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="3stages")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--graph", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from text2img_ekl_b200 import configs, parallel
    from text2img_ekl_b200.synthetic import SyntheticLoader
    rank, ws = parallel.init_from_env()
    Trainer = configs.setup(a.config, batch=a.batch)
    torch.manual_seed(100 + rank)              # different initial weights per rank: _replicate() must override them
    tr = Trainer(None, None, 64)
    tr.setup()
    loader = SyntheticLoader(a.batch, getattr(tr, "CLS_KIND", "index"), rank=rank, pool=a.steps)
    gs = None
    if a.graph:
        from text2img_ekl_b200.engine import GraphedStep
        gs = GraphedStep(tr, loader.pool[0])
        for i in range(a.steps):
            gs.step(loader.pool[i])
    else:
        for i in range(a.steps):
            tr.train_step(loader.pool[i])
    torch.cuda.synchronize()
    worst, bad = 0.0, []
    opts = [("G", tr.optimizerG)] + [("D%d" % i, o) for i, o in enumerate(tr.optimizersD)]
    for name, o in opts:
        for buf in ("flat_p", "exp_avg", "exp_avg_sq", "shadow"):
            t = getattr(o, buf).float()
            ref = t.clone()
            dist.broadcast(ref, 0)
            d = float((t - ref).abs().max())
            worst = max(worst, d)
            if d != 0.0:
                bad.append("%s.%s" % (name, buf))
    flag = torch.tensor([len(bad)], device="cuda")
    dist.all_reduce(flag)
    moved = float((tr.optimizerG.exp_avg.abs().sum() > 0)) * float((tr.optimizersD[-1].exp_avg.abs().sum() > 0))
    if rank == 0:
        print(json.dumps({"world": ws, "config": a.config, "steps": a.steps, "graph": a.graph, "diverged_buffers": int(flag.item()),
                          "max_abs_diff_rank": worst, "updated": bool(moved), "grad_comm": parallel.grad_comm_mode(),
                          "tail_adam": os.environ.get("EKL_TAIL_ADAM", "1")}), flush=True)
    rc = 1 if flag.item() or not moved else 0
    sys.stdout.flush()
    from bench import teardown            # captured graphs hold NCCL kernels: destroy them before the process group
    teardown(ws, gs)
    sys.exit(rc)


if __name__ == "__main__":
    main()
