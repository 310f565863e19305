#!/usr/bin/env python
"""Run-to-run bimodality probe: in ONE process, capture the step graph several times and time each capture's replays.
Tells whether the ~1.8 % spread seen between bench processes comes from the capture / graph instantiation (varies within a
process) or from process-level state such as device addresses (constant within a process).

    python tools/graph_modes.py [--config 3stages] [--captures 5] [--steps 40]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="3stages")
    ap.add_argument("--captures", type=int, default=5)
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    import torch
    from bench import DEFAULT_BATCH
    from text2img_ekl_b200 import configs
    from text2img_ekl_b200.engine import GraphedStep
    from text2img_ekl_b200.synthetic import SyntheticLoader
    B = DEFAULT_BATCH[a.config]
    Trainer = configs.setup(a.config, batch=B)
    torch.manual_seed(0)
    tr = Trainer(None, None, 64)
    tr.setup()
    loader = SyntheticLoader(B, getattr(tr, "CLS_KIND", "index"), pool=1)
    for c in range(a.captures):
        gs = GraphedStep(tr, loader.pool[0])
        for _ in range(5):
            gs.replay()
        torch.cuda.synchronize()
        ms = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                gs.replay()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / a.steps)
        print("capture %d: %s ms/step  (in_flat @ %x)" % (c, " ".join("%.3f" % m for m in ms), gs._in_flat.data_ptr()), flush=True)
        gs.graph.reset()
        del gs
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
