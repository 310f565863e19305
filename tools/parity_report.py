"""Print every parity metric of the GPU training step against the CPU oracle (no asserts).
    python tools/parity_report.py [config ...] [--B n] [--iters n]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("configs", nargs="*", default=["splitz_cap_ca"])
ap.add_argument("--B", type=int, default=4)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
from oracle import synth
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_step_parity_gpu import build, rel

for name in a.configs:
    t0 = time.time()
    tr, oc, orc = build(name, a.B)
    from oracle import ekl_oracle as O
    _, _, orc16 = build(name, a.B)       # second oracle instance evaluated at the implementation's storage precision
    dev = tr.device
    print("=== %s B=%d (build %.1fs)" % (name, a.B, time.time() - t0), flush=True)
    for it in range(a.iters):
        b = synth.make_batch(oc, a.B, "it%d" % it)
        t0 = time.time(); want = orc.step(**b)
        with O.storage("bf16"):
            w16 = orc16.step(**b)
        t1 = time.time()
        errDs, errG = tr.train_step((b["imgs"], b["wrong_imgs"], b["embedding"], b["cls"], None),
                                    noise=b["noise"].to(dev), eps=b["eps"].to(dev), seed=b["seed"].to(dev))
        torch.cuda.synchronize(); t2 = time.time()
        print(" it%d oracle %.1fs gpu(eager) %.2fs" % (it, t1 - t0, t2 - t1))
        print("  cls exact:", bool(torch.equal(tr.real_cp.cpu(), want["real_cp"])))
        for i, (g, w) in enumerate(zip(tr.hcodes, want["h_codes"])):
            print("  h%d %.2e" % (i, rel(g, w)), end="")
        for i, (g, w) in enumerate(zip(tr.fake_imgs, want["fake_imgs"])):
            print("  img%d %.2e" % (i, rel(g, w)), end="")
        print()
        for i, (g, w) in enumerate(zip(errDs, want["errD"])):
            print("  errD%d got %s want %s" % (i, ["%.4f" % float(x) for x in g], ["%.4f" % float(x) for x in w]))
        print("  errG got %s want %s" % (["%.4f" % float(x) for x in errG], ["%.4f" % float(x) for x in want["errG"]]))
        for i, (g, w) in enumerate(zip(tr.engine.last_g_logits, want["g_logits"])):
            print("  glogit D%d: %s" % (i, " ".join("%.2e" % rel(g[q], w[q]) for q in range(len(w)))))
        for tag, ww in (("fp32", want), ("bf16-storage", w16)):
            rs = sorted([(rel(p.grad, ww["gradG"][k]), k) for k, p in tr.netG.named_parameters() if k in ww["gradG"]])
            print("  [vs %s oracle] gradG median %.2e worst: %s" % (tag, np.median([r for r, _ in rs]), " | ".join("%s %.2e" % (k, r) for r, k in rs[-4:])))
            for i, d in enumerate(tr.netsD):
                rs = sorted([(rel(p.grad, ww["gradD"][i][k]), k) for k, p in d.named_parameters() if k in ww["gradD"][i]])
                print("  [vs %s oracle] gradD%d median %.2e worst: %s" % (tag, i, np.median([r for r, _ in rs]), " | ".join("%s %.2e" % (k, r) for r, k in rs[-3:])))
            print("  [vs %s oracle] img %s  errG %.2e" % (tag, " ".join("%.2e" % rel(g, w) for g, w in zip(tr.fake_imgs, ww["fake_imgs"])),
                  rel(torch.stack([x.float() for x in errG]), ww["errG"])))
    for tag, net, sd in [("G", tr.netG, orc.sdG)] + [("D%d" % i, d, orc.sdDs[i]) for i, d in enumerate(tr.netsD)]:
        num = den = 0.0
        for k, v in net.state_dict().items():
            if v.is_floating_point() and "running" not in k:
                num += float((v.detach().float().cpu() - sd[k].detach()).pow(2).sum()); den += float(sd[k].detach().pow(2).sum())
        rm = max(rel(v, sd[k]) for k, v in net.state_dict().items() if "running_mean" in k or "running_var" in k)
        print("  params %s rel %.2e   running-stat worst rel %.2e" % (tag, (num / den) ** 0.5, rm))
