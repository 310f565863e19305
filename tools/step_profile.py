#!/usr/bin/env python
"""Per-kernel device-time table of the graph-replayed training step (CUPTI via torch.profiler; development tool —
bench.py numbers are never taken under a profiler).

    python tools/step_profile.py --config 3stages --steps 3 [--top 60] [--json out.json]
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
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--top", type=int, default=70)
    ap.add_argument("--json", default="")
    ap.add_argument("--eager", action="store_true")
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    from bench import DEFAULT_BATCH
    from text2img_ekl_b200 import configs
    from text2img_ekl_b200.engine import GraphedStep
    from text2img_ekl_b200.synthetic import SyntheticLoader
    B = a.batch or DEFAULT_BATCH[a.config]
    Trainer = configs.setup(a.config, batch=B)
    torch.manual_seed(0)
    tr = Trainer(None, None, 64)
    tr.setup()
    loader = SyntheticLoader(B, getattr(tr, "CLS_KIND", "index"), pool=1)
    if a.eager:
        for _ in range(3):
            tr.train_step(loader.pool[0])
        run = lambda: tr.train_step(loader.pool[0])
    else:
        gs = GraphedStep(tr, loader.pool[0])
        for _ in range(3):
            gs.replay()
        run = gs.replay
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / a.steps
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            run()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        if t <= 0:
            continue
        rows.append((ev.key, ev.count / a.steps, t / a.steps))
    rows.sort(key=lambda r: -r[2])
    tot = sum(r[2] for r in rows)
    print("step %.3f ms (unprofiled, %s); kernel time sum %.3f ms/step; %d kernels/step" %
          (step_ms, "eager" if a.eager else "graph", tot / 1e3, int(sum(r[1] for r in rows))))
    print("%-100s %8s %10s %7s %8s" % ("kernel", "n/step", "us/step", "share", "avg us"))
    for k, n, t in rows[: a.top]:
        print("%-100s %8.1f %10.1f %6.1f%% %8.1f" % (k[:100], n, t, 100 * t / tot, t / n))
    ours = sum(t for k, n, t in rows if "ekl" in k or "<unnamed>" in k or "anonymous" in k)
    print("library kernels (anonymous namespace): %.1f%% of kernel time" % (100 * ours / tot))
    if a.json:
        json.dump({"step_ms": step_ms, "kernel_ms": tot / 1e3, "rows": rows}, open(a.json, "w"))


if __name__ == "__main__":
    main()
