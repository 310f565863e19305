#!/usr/bin/env python
"""Per-layer timing of every conv / BatchNorm kernel call of one training step of a config, at the step's real shapes.

One eager step records the shape of every library call (ops.SHAPE_LOG); each unique shape is then timed in isolation
(CUDA events, L2 flushed between iterations) and the table is weighted by calls per step.

    python tools/layer_bench.py --config 3stages [--iters 5] [--only conv|bn]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="3stages")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    import torch
    from bench import DEFAULT_BATCH
    from text2img_ekl_b200 import _lib as L, configs, ops
    from text2img_ekl_b200.synthetic import SyntheticLoader
    B = a.batch or DEFAULT_BATCH[a.config]
    Trainer = configs.setup(a.config, batch=B)
    torch.manual_seed(0)
    tr = Trainer(None, None, 64)
    tr.setup()
    loader = SyntheticLoader(B, getattr(tr, "CLS_KIND", "index"), pool=1)
    tr.train_step(loader.pool[0])
    ops.SHAPE_LOG = []
    tr.train_step(loader.pool[0])
    torch.cuda.synchronize()
    log, ops.SHAPE_LOG = ops.SHAPE_LOG, None
    counts = collections.Counter(log)
    del tr
    torch.cuda.empty_cache()
    lib = L.lib()
    dev = torch.device("cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timeit(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return sorted(ts)[len(ts) // 2]

    rows = []
    for key, n in sorted(counts.items(), key=lambda kv: str(kv[0])):
        kind = key[0]
        if kind in ("fwd", "dgrad", "wgrad"):
            if a.only and a.only != "conv":
                continue
            _, fam, mode, b, H, W, Cin, Cout, group_b = key
            if fam != "conv_tc":
                continue
            K = 4 if mode == 2 else 3
            Ho, Wo = (2 * H, 2 * W) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
            conv = L.EklConv(mode, b, H, W, Cin, Cout, group_b, 0, 0, 0, 0, 0)
            x = torch.randn(b, H, W, Cin, device=dev).bfloat16()
            dy = torch.randn(b, Ho, Wo, Cout, device=dev).bfloat16()
            wm = torch.randn(Cout, K, K, Cin, device=dev) * 0.05
            wf = torch.empty(lib.ekl_conv_packed_elems(conv, 0), device=dev, dtype=torch.bfloat16)
            wd = torch.empty(lib.ekl_conv_packed_elems(conv, 1), device=dev, dtype=torch.bfloat16)
            L.check(lib.ekl_conv_pack(conv, L.ptr(wm), L.ptr(wf), L.ptr(wd), L.stream()))
            y = torch.empty(b, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
            dx = torch.empty_like(x)
            dw = torch.zeros(Cout, K, K, Cin, device=dev)
            stats = torch.zeros(max(b // group_b if group_b else 1, 1), 2, Cout, device=dev, dtype=torch.float64)
            if kind == "fwd":
                fn = lambda: L.check(lib.ekl_conv_fwd(conv, L.ptr(x), L.ptr(wf), L.ptr(y), L.ptr(stats), L.stream()))
            elif kind == "dgrad":
                fn = lambda: L.check(lib.ekl_conv_bwd_data(conv, L.ptr(dy), L.ptr(wd), L.ptr(dx), L.stream()))
            else:
                fn = lambda: L.check(lib.ekl_conv_bwd_weight(conv, L.ptr(x), L.ptr(dy), L.ptr(dw), L.stream()))
            t = timeit(fn)
            ref = 2.0 * b * Ho * Wo * Cout * Cin * K * K
            exe = ref / 2.25 if mode == 1 else ref
            nbytes = (x.numel() + dy.numel()) * 2 + (dw.numel() * 4 if kind == "wgrad" else wf.numel() * 2)
            rows.append((t * n, "%-5s m%d %3dx%4dx%4d %4d>%4d g%-3d" % (kind, mode, b, H, W, Cin, Cout, group_b), n, t,
                         "%6.0f TF/s ref %6.0f exe  %5.0f GB/s" % (ref / t / 1e6, exe / t / 1e6, nbytes / t / 1e3)))
            del x, dy, wm, wf, wd, y, dx, dw, stats
        elif kind in ("bn_fwd", "bn_bwd"):
            if a.only and a.only != "bn":
                continue
            _, M, Cy, groups, act, has_res = key
            Co = Cy // 2 if act == L.ACT_GLU else Cy
            y = torch.randn(M, Cy, device=dev).bfloat16()
            dout = torch.randn(M, Co, device=dev).bfloat16()
            res = torch.randn(M, Co, device=dev).bfloat16() if has_res else None
            gamma, beta = torch.ones(Cy, device=dev), torch.zeros(Cy, device=dev)
            mean, rstd = torch.zeros(groups, Cy, device=dev), torch.ones(groups, Cy, device=dev)
            out = torch.empty(M, Co, device=dev, dtype=torch.bfloat16)
            dy = torch.empty_like(y)
            sums = torch.zeros(max(int(lib.ekl_bn_bwd_scratch_doubles(M, Cy, groups, act)), 2), device=dev, dtype=torch.float64)
            dg, db = torch.zeros(Cy, device=dev), torch.zeros(Cy, device=dev)
            st = L.stream()
            if kind == "bn_fwd":
                fn = lambda: L.check(lib.ekl_bn_act_fwd(L.ptr(y), M, Cy, groups, None, 1e-5, 0.1, L.ptr(mean), L.ptr(rstd), None, None,
                                                        L.ptr(gamma), L.ptr(beta), act, L.ptr(res), L.ptr(out), st))
                nbytes = M * (Cy + Co + (Co if has_res else 0)) * 2
            else:
                fn = lambda: L.check(lib.ekl_bn_act_bwd(L.ptr(y), L.ptr(dout), M, Cy, groups, L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
                                                        L.ptr(beta), act, L.ptr(sums), L.ptr(dg), L.ptr(db), L.ptr(dy), st))
                nbytes = M * (Cy + Co + Cy) * 2          # minimal traffic: y and dout read once, dy written once
            t = timeit(fn)
            rows.append((t * n, "%-6s M=%8d Cy=%4d g%d act%d res%d" % (kind, M, Cy, groups, act, int(has_res)), n, t,
                         "%6.0f GB/s (minimal-traffic bytes)" % (nbytes / t / 1e3)))
            del y, dout, res, out, dy
    rows.sort(key=lambda r: -r[0])
    tot = sum(r[0] for r in rows)
    if a.json:
        import json
        json.dump([[name, n, t] for tt, name, n, t, rate in rows], open(a.json, "w"))
    print("config %s B=%d: isolated kernel time weighted by calls/step = %.2f ms" % (a.config, B, tot / 1e3))
    print("%-44s %3s %9s %9s  %s" % ("call", "n", "us each", "us/step", "rate"))
    for tt, name, n, t, rate in rows:
        print("%-44s %3d %9.1f %9.1f  %s" % (name, n, t, tt, rate))


if __name__ == "__main__":
    main()
