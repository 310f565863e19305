"""Distribution of (our gradient deviation) / (bf16-storage oracle deviation) per parameter tensor, per network.
    python tools/grad_ratio.py [config ...] [--B n] [--repeat n]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="*", default=["coco"])
ap.add_argument("--B", type=int, default=4)
ap.add_argument("--repeat", type=int, default=1)
a = ap.parse_args()
from oracle import synth, ekl_oracle as O
from test_step_parity_gpu import build, rel

for name in a.configs:
    for rep in range(a.repeat):
        tr, oc, orc, orc16 = build(name, a.B)
        dev = tr.device
        b = synth.make_batch(oc, a.B, "it0")
        want = orc.step(**b)
        with O.storage("bf16"):
            w16 = orc16.step(**b)
        tr.train_step((b["imgs"], b["wrong_imgs"], b["embedding"], b["cls"], None),
                      noise=b["noise"].to(dev), eps=b["eps"].to(dev), seed=b["seed"].to(dev))
        torch.cuda.synchronize()
        nets = [("G", tr.netG, want["gradG"], w16["gradG"])] + [("D%d" % i, d, want["gradD"][i], w16["gradD"][i]) for i, d in enumerate(tr.netsD)]
        for tag, net, wg, fg in nets:
            rows = []
            for k, p in net.named_parameters():
                if k not in wg or k.endswith(("fc1.bias", "fc2.bias")):
                    continue
                r, f = rel(p.grad, wg[k]), rel(fg[k], wg[k])
                rows.append((r / max(f, 1e-12), r, f, k))
            rows.sort()
            q = [rows[int(len(rows) * x)][0] for x in (0.1, 0.5, 0.9)]
            print("%s rep%d %s: n=%d ratio q10/50/90 = %.2f %.2f %.2f | median ours %.3f floor %.3f | worst: %s" % (
                name, rep, tag, len(rows), q[0], q[1], q[2], np.median([r[1] for r in rows]), np.median([r[2] for r in rows]),
                " ; ".join("%s r=%.3f f=%.3f" % (k, r, f) for _, r, f, k in rows[-3:])), flush=True)
