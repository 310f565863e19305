#!/usr/bin/env python
"""Timeline of ONE graph-replayed training step (CUPTI kernel records with their streams): which stream is busy when,
how long each phase of the step takes and how much of it is a single stream running alone.  Development tool.

    python tools/step_trace.py --config 3stages --json gpurun_out/trace.json      # on the GPU box
    python tools/step_trace.py --read gpurun_out/trace.json                       # here: summary of a saved trace
"""
import argparse
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)[:60]


def capture(a):
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
    gs = GraphedStep(tr, loader.pool[0])
    for _ in range(3):
        gs.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        gs.replay()
        torch.cuda.synchronize()
    tmp = a.json + ".chrome"
    prof.export_chrome_trace(tmp)
    ev = []
    for e in json.load(open(tmp))["traceEvents"]:
        if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
            ar = e.get("args", {})
            ev.append((short(e["name"]), int(ar.get("stream", 0)), float(e["ts"]), float(e["dur"]), list(ar.get("grid", []))))
    os.remove(tmp)
    ev.sort(key=lambda x: x[2])
    t0 = ev[0][2]
    ev = [(n, s, round(ts - t0, 3), round(d, 3), g) for n, s, ts, d, g in ev]
    json.dump({"config": a.config, "batch": B, "events": ev}, open(a.json, "w"))
    return ev


def summarize(ev, top=12):
    ev = [tuple(e[:4]) for e in ev]          # (name, stream, start us, duration us[, grid])
    end = max(ts + d for _, _, ts, d in ev)
    streams = collections.OrderedDict()
    for n, s, ts, d in ev:
        streams.setdefault(s, []).append((ts, d, n))
    print("step span %.1f us, %d kernels, %d streams" % (end, len(ev), len(streams)))
    for s, lst in streams.items():
        busy = sum(d for _, d, _ in lst)
        print("  stream %-4d first %8.1f last %8.1f busy %8.1f us (%3d kernels)  e.g. %s" % (
            s, lst[0][0], lst[-1][0] + lst[-1][1], busy, len(lst), lst[len(lst) // 2][2]))
    # concurrency profile: time with k kernels in flight
    pts = []
    for n, s, ts, d in ev:
        pts.append((ts, 1)); pts.append((ts + d, -1))
    pts.sort()
    conc, last, hist = 0, 0.0, collections.Counter()
    for t, k in pts:
        hist[conc] += t - last
        last, conc = t, conc + k
    print("  time with k kernels in flight:", {k: round(v, 1) for k, v in sorted(hist.items())})
    # phases: split the step where the set of active streams changes (coarse: 250 us windows)
    win = 250.0
    nwin = int(end / win) + 1
    for w in range(nwin):
        a0, a1 = w * win, (w + 1) * win
        per = collections.Counter()
        names = collections.Counter()
        for n, s, ts, d in ev:
            o = min(ts + d, a1) - max(ts, a0)
            if o > 0:
                per[s] += o
                names[n] += o
        desc = " ".join("s%d:%3.0f%%" % (s, 100 * v / win) for s, v in sorted(per.items()))
        topn = ", ".join("%s %.0f" % (n[:28], v) for n, v in names.most_common(3))
        print("  [%6.0f-%6.0f] %-60s | %s" % (a0, a1, desc, topn))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="3stages")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--json", default="gpurun_out/trace.json")
    ap.add_argument("--read", default="")
    a = ap.parse_args()
    ev = json.load(open(a.read))["events"] if a.read else capture(a)
    summarize(ev)


if __name__ == "__main__":
    main()
