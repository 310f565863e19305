"""On-GPU numerical check of the conv kernel family through the C ABI; prints one line per case and keeps going
after failures (each group runs in its own subprocess so a trapped kernel cannot poison the others).

    python tools/kernel_check.py            # all groups
    python tools/kernel_check.py --group tc_fwd
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {  # mode -> list of (B, H, W, Cin, Cout)
    0: [(4, 8, 8, 64, 128), (2, 16, 16, 128, 64), (3, 4, 4, 64, 256), (2, 32, 32, 32, 64), (2, 64, 64, 16, 32),
        (8, 4, 4, 192, 512), (2, 8, 8, 320, 128), (3, 32, 32, 16, 64), (2, 32, 32, 32, 16), (2, 64, 64, 16, 16),
        # resident-filter kernel (conv_rw.cu): two N tiles, one tile per image, more tiles than SMs
        (2, 16, 16, 64, 128), (5, 16, 8, 64, 64), (6, 64, 64, 64, 64), (3, 32, 64, 32, 32),
        # 128-channel contraction on the resident-filter kernel (two channel blocks, 32-wide N tiles): forward / data-gradient
        (3, 32, 64, 128, 64), (4, 64, 64, 64, 128)],
    1: [(4, 4, 4, 128, 128), (2, 8, 8, 64, 64), (3, 16, 16, 64, 128), (2, 32, 32, 32, 32), (2, 4, 4, 256, 512),
        # sub-pixel plans on the resident-filter kernel (>= 2 x 148 tiles of 16x8 low-resolution pixels): one / two channel
        # blocks, SW128 / SW64 / SW32 rows, two N tiles
        (10, 64, 64, 64, 64), (5, 128, 128, 32, 32), (20, 32, 64, 128, 64), (3, 128, 128, 16, 16), (10, 64, 64, 64, 128)],
    2: [(4, 16, 16, 64, 128), (2, 8, 8, 128, 256), (3, 32, 32, 64, 64), (2, 8, 8, 512, 1024), (2, 64, 64, 32, 64),
        # data-gradients of these are sub-pixel plans on the resident-filter kernel (contraction 128 = two channel blocks / 64)
        (10, 128, 128, 64, 128), (6, 128, 128, 16, 64), (10, 128, 128, 128, 64)],
}
GROUPS = ["tc_fwd", "tc_dgrad", "tc_wgrad", "simt_fwd", "simt_dgrad", "simt_wgrad", "bn", "misc", "heads", "tc_split", "fold", "vc", "route"]


def run_group(group):
    import torch
    import torch.nn.functional as F
    from text2img_ekl_b200 import _lib as L
    lib = L.lib()
    L.check(lib.ekl_require_sm100())
    dev = torch.device("cuda")
    torch.manual_seed(0)
    nfail = 0

    def rel(a, b):
        return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))

    if group == "bn":
        return run_bn(torch, L, lib, dev, rel)
    if group == "misc":
        return run_misc(torch, L, lib, dev, rel)
    if group == "heads":
        return run_heads(torch, L, lib, dev, rel)
    if group == "tc_split":
        return run_split(torch, L, lib, dev, rel)
    if group == "fold":
        return run_fold(torch, L, lib, dev, rel)
    if group == "vc":
        return run_vc(torch, L, lib, dev, rel)
    if group == "route":
        return run_route(torch, L, lib, dev, rel)
    impl = L.IMPL_TC if group.startswith("tc") else L.IMPL_SIMT
    what = group.split("_")[1]
    for mode, shapes in SHAPES.items():
        for (B, H, W, Cin, Cout) in shapes:
            K = 4 if mode == 2 else 3
            Ho, Wo = (2 * H, 2 * W) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
            x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
            wm = (torch.randn(Cout, K, K, Cin, device=dev) / (K * K * Cin) ** 0.5).bfloat16().float()  # master, bf16-exact
            conv = L.EklConv(mode, B, H, W, Cin, Cout, 0, impl, 0, 0, 0)
            nf = lib.ekl_conv_packed_elems(conv, 0)
            nd = lib.ekl_conv_packed_elems(conv, 1)
            wf = torch.empty(nf, device=dev, dtype=torch.bfloat16)
            wd = torch.empty(nd, device=dev, dtype=torch.bfloat16)
            L.check(lib.ekl_conv_pack(conv, L.ptr(wm), L.ptr(wf), L.ptr(wd), L.stream()))
            # torch reference (fp32, NCHW)
            xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            wr = wm.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            xin = F.interpolate(xr, scale_factor=2, mode="nearest") if mode == 1 else xr
            yr = F.conv2d(xin, wr, stride=2 if mode == 2 else 1, padding=1)
            dyn = torch.randn(B, Ho, Wo, Cout, device=dev).bfloat16()
            yr.backward(dyn.float().permute(0, 3, 1, 2))
            tag = "%-10s mode%d B%d %dx%d Cin%d Cout%d" % (group, mode, B, H, W, Cin, Cout)
            try:
                if what == "fwd":
                    y = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
                    stats = None
                    if impl == L.IMPL_TC:
                        stats = torch.zeros(1, 2, Cout, device=dev, dtype=torch.float64)     # fp64 sums, accumulated
                    L.check(lib.ekl_conv_fwd(conv, L.ptr(x), L.ptr(wf), L.ptr(y), L.ptr(stats), L.stream()))
                    torch.cuda.synchronize()
                    e = rel(y.permute(0, 3, 1, 2), yr)
                    msg = "rel %.2e" % e
                    ok = e < 6e-3
                    if stats is not None:
                        s = stats[0].float()
                        yf = y.float().reshape(-1, Cout)
                        e1, e2 = rel(s[0], yf.sum(0)), rel(s[1], (yf * yf).sum(0))
                        msg += " stats %.1e %.1e" % (e1, e2)
                        ok = ok and e2 < 1e-3 and (e1 < 1e-2 or float((s[0] - yf.sum(0)).abs().max()) < 0.5)
                elif what == "dgrad":
                    dx = torch.full((B, H, W, Cin), float("nan"), device=dev, dtype=torch.bfloat16)
                    L.check(lib.ekl_conv_bwd_data(conv, L.ptr(dyn), L.ptr(wd), L.ptr(dx), L.stream()))
                    torch.cuda.synchronize()
                    e = rel(dx.permute(0, 3, 1, 2), xr.grad)
                    msg, ok = "rel %.2e" % e, e < 6e-3
                    if impl == L.IMPL_TC and lib.ekl_conv_dgrad_from_fwd(conv):
                        # same gradient from the FORWARD-packed filter (MN-major B operand)
                        dx2 = torch.full((B, H, W, Cin), float("nan"), device=dev, dtype=torch.bfloat16)
                        L.check(lib.ekl_conv_bwd_data_fw(conv, L.ptr(dyn), L.ptr(wf), L.ptr(dx2), None, L.stream()))
                        torch.cuda.synchronize()
                        e2 = rel(dx2.permute(0, 3, 1, 2), xr.grad)
                        msg += " from-fwd %.2e" % e2
                        ok = ok and e2 < 6e-3
                else:
                    dw = torch.zeros(Cout, K, K, Cin, device=dev)
                    L.check(lib.ekl_conv_bwd_weight(conv, L.ptr(x), L.ptr(dyn), L.ptr(dw), L.stream()))
                    torch.cuda.synchronize()
                    e = rel(dw.permute(0, 3, 1, 2), wr.grad)
                    msg, ok = "rel %.2e" % e, e < 6e-3
            except Exception as ex:  # noqa: BLE001
                msg, ok = "EXC %s" % str(ex)[:200], False
            print("%s %s %s" % ("PASS" if ok else "FAIL", tag, msg), flush=True)
            nfail += 0 if ok else 1
    return nfail


def run_bn(torch, L, lib, dev, rel):
    nfail = 0
    for (M, Cy, groups, act) in [(3 * 512, 128, 3, L.ACT_GLU), (2048, 64, 1, L.ACT_LRELU), (96, 32768, 1, L.ACT_GLU),
                                 (1536, 512, 3, L.ACT_LRELU), (4096, 64, 1, L.ACT_NONE), (640, 320, 1, L.ACT_GLU),
                                 (64, 512, 1, L.ACT_RELU), (3 * 384, 2048, 3, L.ACT_LRELU), (24, 32768, 1, L.ACT_GLU),
                                 (384, 512, 1, L.ACT_LRELU), (700, 64, 1, L.ACT_NONE)]:
        y = (torch.randn(M, Cy, device=dev) * 1.5 + 0.3).bfloat16()
        gamma = 1 + 0.1 * torch.randn(Cy, device=dev)
        beta = 0.1 * torch.randn(Cy, device=dev)
        Co = Cy // 2 if act == L.ACT_GLU else Cy
        res = torch.randn(M, Co, device=dev).bfloat16() if act == L.ACT_NONE else None
        dout = torch.randn(M, Co, device=dev).bfloat16()
        stats = torch.zeros(groups, 2, Cy, device=dev, dtype=torch.float64)          # fp64 sums, accumulated by col_stats
        mean = torch.full((groups, Cy), float("nan"), device=dev)
        rstd = torch.full((groups, Cy), float("nan"), device=dev)
        rm, rv = torch.zeros(Cy, device=dev), torch.ones(Cy, device=dev)
        out = torch.empty(M, Co, device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_col_stats(L.ptr(y), M, Cy, groups, L.ptr(stats), L.stream()))
        L.check(lib.ekl_bn_act_fwd(L.ptr(y), M, Cy, groups, L.ptr(stats), 1e-5, 0.1, L.ptr(mean), L.ptr(rstd), L.ptr(rm), L.ptr(rv),
                                   L.ptr(gamma), L.ptr(beta), act, L.ptr(res), L.ptr(out), L.stream()))
        small = M // groups <= 768
        sums = torch.zeros(max(int(lib.ekl_bn_bwd_scratch_doubles(M, Cy, groups, act)), 2), device=dev, dtype=torch.float64)
        dg, db = torch.zeros(Cy, device=dev), torch.zeros(Cy, device=dev)
        dy = torch.empty(M, Cy, device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_bn_act_bwd(L.ptr(y), L.ptr(dout), M, Cy, groups, L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(beta),
                                   act, L.ptr(sums), L.ptr(dg), L.ptr(db), L.ptr(dy), L.stream()))
        # inference path: the same kernel with mean / rstd given (sums == NULL) must reproduce the training output
        out_inf = torch.empty(M, Co, device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_bn_act_fwd(L.ptr(y), M, Cy, groups, None, 1e-5, 0.1, L.ptr(mean), L.ptr(rstd), None, None,
                                   L.ptr(gamma), L.ptr(beta), act, L.ptr(res), L.ptr(out_inf), L.stream()))
        torch.cuda.synchronize()
        # reference
        yr = y.float().requires_grad_(True)
        g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        rm2, rv2 = torch.zeros(Cy, device=dev), torch.ones(Cy, device=dev)
        outs = []
        for g in range(groups):
            sl = slice(g * M // groups, (g + 1) * M // groups)
            z = torch.nn.functional.batch_norm(yr[sl], rm2, rv2, g_, b_, True, 0.1, 1e-5)
            if act == L.ACT_GLU:
                o = z[:, :Co] * torch.sigmoid(z[:, Co:])
            elif act == L.ACT_LRELU:
                o = torch.nn.functional.leaky_relu(z, 0.2)
            elif act == L.ACT_RELU:
                o = torch.relu(z)
            else:
                o = z + res[sl].float()
            outs.append(o)
        o = torch.cat(outs)
        o.backward(dout.float())
        errs = dict(out=rel(out, o), dy=rel(dy, yr.grad), dgamma=rel(dg, g_.grad), dbeta=rel(db, b_.grad),
                    rmean=rel(rm, rm2), rvar=rel(rv, rv2), inference=rel(out_inf, out) * 10)
        ok = all(v < 1e-2 for v in errs.values())
        print("%s bn%s M%d C%d g%d act%d %s" % ("PASS" if ok else "FAIL", " (single-launch)" if small else "", M, Cy, groups, act,
                                               " ".join("%s %.1e" % kv for kv in errs.items())), flush=True)
        nfail += 0 if ok else 1
    return nfail


def run_split(torch, L, lib, dev, rel):
    """Split-K workspace path (ekl_conv_fwd_ws / ekl_conv_bwd_data_ws): few output tiles, long contraction."""
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    nfail = 0
    for (mode, B, H, W, Cin, Cout, gb) in [(0, 6, 4, 4, 512, 256, 2), (2, 8, 8, 8, 256, 512, 0), (0, 24, 4, 4, 1024, 512, 8),
                                           (0, 3, 4, 4, 640, 512, 0)]:
        K = 4 if mode == 2 else 3
        Ho, Wo = (H // 2, W // 2) if mode == 2 else (H, W)
        x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
        wm = (torch.randn(Cout, K, K, Cin, device=dev) / (K * K * Cin) ** 0.5).bfloat16().float()
        conv = L.EklConv(mode, B, H, W, Cin, Cout, gb, L.IMPL_TC, 0, 0, 0, 0)
        wf = torch.empty(lib.ekl_conv_packed_elems(conv, 0), device=dev, dtype=torch.bfloat16)
        wd = torch.empty(lib.ekl_conv_packed_elems(conv, 1), device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_conv_pack(conv, L.ptr(wm), L.ptr(wf), L.ptr(wd), L.stream()))
        xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        yr = F.conv2d(xr, wm.permute(0, 3, 1, 2).contiguous(), stride=2 if mode == 2 else 1, padding=1)
        dyn = torch.randn(B, Ho, Wo, Cout, device=dev).bfloat16()
        yr.backward(dyn.float().permute(0, 3, 1, 2))
        nf, nd = lib.ekl_conv_workspace_elems(conv, 0), lib.ekl_conv_workspace_elems(conv, 1)
        msg, ok = "ws fwd %d dgrad %d" % (nf, nd), nf > 0
        if nf > 0:
            ws = torch.zeros(nf, device=dev)
            for rep in range(2):                      # second pass proves the workspace was left zero
                y = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
                groups = B // gb if gb else 1
                stats = torch.zeros(groups, 2, Cout, device=dev, dtype=torch.float64)
                L.check(lib.ekl_conv_fwd_ws(conv, L.ptr(x), L.ptr(wf), L.ptr(y), L.ptr(stats), L.ptr(ws), L.stream()))
                torch.cuda.synchronize()
                e = rel(y.permute(0, 3, 1, 2), yr)
                sg = stats.float()
                yf = y.float().reshape(groups, -1, Cout)
                e2 = rel(sg[:, 1], (yf * yf).sum(1))
                ok = ok and e < 6e-3 and e2 < 1e-3 and float(ws.abs().max()) == 0.0
                msg += " | y %.1e stats %.1e" % (e, e2)
            # the same conv with the finishing pass fused with train-mode BatchNorm + LeakyReLU (one kernel) against the
            # unfused sequence ekl_conv_fwd_ws -> ekl_bn_act_fwd on the same inputs: y bit-identical, the rest to rounding
            if lib.ekl_conv_split_bn_fusable(conv, L.ACT_LRELU):
                groups = B // gb if gb else 1
                gamma, beta = 1 + 0.1 * torch.randn(Cout, device=dev), 0.1 * torch.randn(Cout, device=dev)
                ref = {}
                for fused in (0, 1):
                    yv = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
                    out = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
                    mean, rstd = torch.empty(groups, Cout, device=dev), torch.empty(groups, Cout, device=dev)
                    rm, rv = torch.zeros(Cout, device=dev), torch.ones(Cout, device=dev)
                    if fused:
                        aux = torch.zeros(int(lib.ekl_conv_split_bn_aux_floats(conv)), device=dev)
                        L.check(lib.ekl_conv_fwd_split_bn_act(conv, L.ptr(x), L.ptr(wf), L.ptr(ws), L.ptr(yv), 1e-5, 0.1, L.ptr(mean),
                                                              L.ptr(rstd), L.ptr(rm), L.ptr(rv), L.ptr(gamma), L.ptr(beta), L.ACT_LRELU,
                                                              L.ptr(out), L.ptr(aux), L.stream()))
                    else:
                        stats = torch.zeros(groups, 2, Cout, device=dev, dtype=torch.float64)
                        L.check(lib.ekl_conv_fwd_ws(conv, L.ptr(x), L.ptr(wf), L.ptr(yv), L.ptr(stats), L.ptr(ws), L.stream()))
                        L.check(lib.ekl_bn_act_fwd(L.ptr(yv), B * Ho * Wo, Cout, groups, L.ptr(stats), 1e-5, 0.1, L.ptr(mean),
                                                   L.ptr(rstd), L.ptr(rm), L.ptr(rv), L.ptr(gamma), L.ptr(beta), L.ACT_LRELU, None,
                                                   L.ptr(out), L.stream()))
                    torch.cuda.synchronize()
                    if fused:
                        same_y = bool(torch.equal(yv, ref["y"]))
                        errs = [rel(out, ref["out"]), rel(mean, ref["mean"]), rel(rstd, ref["rstd"]), rel(rm, ref["rm"]), rel(rv, ref["rv"])]
                        zero = float(ws.abs().max()) == 0.0 and float(aux[:Cout // 32].abs().max()) == 0.0
                        ok = ok and same_y and zero and errs[0] < 6e-3 and max(errs[1:]) < 1e-4
                        msg += " | fused-bn y==%s out %.1e mean %.1e rstd %.1e run %.1e %.1e" % ((same_y,) + tuple(errs))
                    else:
                        ref = dict(y=yv, out=out, mean=mean, rstd=rstd, rm=rm, rv=rv)
        if nd > 0:
            ws = torch.zeros(nd, device=dev)
            dx = torch.full((B, H, W, Cin), float("nan"), device=dev, dtype=torch.bfloat16)
            L.check(lib.ekl_conv_bwd_data_ws(conv, L.ptr(dyn), L.ptr(wd), L.ptr(dx), L.ptr(ws), L.stream()))
            torch.cuda.synchronize()
            e = rel(dx.permute(0, 3, 1, 2), xr.grad)
            ok = ok and e < 6e-3 and float(ws.abs().max()) == 0.0
            msg += " | dx %.1e" % e
            if lib.ekl_conv_dgrad_from_fwd(conv):
                dx2 = torch.full((B, H, W, Cin), float("nan"), device=dev, dtype=torch.bfloat16)
                L.check(lib.ekl_conv_bwd_data_fw(conv, L.ptr(dyn), L.ptr(wf), L.ptr(dx2), L.ptr(ws), L.stream()))
                torch.cuda.synchronize()
                e2 = rel(dx2.permute(0, 3, 1, 2), xr.grad)
                ok = ok and e2 < 6e-3 and float(ws.abs().max()) == 0.0
                msg += " from-fwd %.1e" % e2
        print("%s tc_split mode%d B%d %dx%d %d>%d gb%d %s" % ("PASS" if ok else "FAIL", mode, B, H, W, Cin, Cout, gb, msg), flush=True)
        nfail += 0 if ok else 1
    return nfail


def run_heads(torch, L, lib, dev, rel):
    """Fused discriminator heads + losses (ops.dhead_dots / ops.d_loss) against the reference formulation in torch:
    sigmoid(conv4x4/s4) heads, nn.BCELoss on constant labels, ce_loss(log_softmax) (cub:60-65, 423-448)."""
    import torch.nn.functional as F
    from text2img_ekl_b200 import ops
    torch.backends.cudnn.allow_tf32 = False          # the torch reference below must be true fp32
    nfail = 0
    for (G, B, C, E1) in [(3, 4, 512, 201), (1, 24, 512, 91), (3, 32, 64, 201)]:
        GB = G * B
        x = torch.randn(GB, 4, 4, C, device=dev).bfloat16().requires_grad_(True)
        h = torch.randn(GB, 4, 4, C, device=dev).bfloat16().requires_grad_(True)
        ws = [(torch.randn(1, C, 4, 4, device=dev) * 0.02).contiguous(memory_format=torch.channels_last).requires_grad_(True)
              for _ in range(2)]
        bs = [torch.randn(1, device=dev).requires_grad_(True) for _ in range(2)]
        cls = (torch.randn(GB, E1, device=dev) * 3).requires_grad_(True)
        cp0 = torch.softmax(torch.randn(B, E1, device=dev), 1)
        cp1 = torch.zeros(B, E1, device=dev); cp1[:, -1] = 1
        tm, tu, ct = ((1, 0, 0), (1, 1, 0), (0, -1, 1)) if G == 3 else ((1,), (1,), (0,))
        coeff = 0.7
        lu, lm = ops.dhead_dots(x, h, ws[0], bs[0], ws[1], bs[1])
        losses, pm, pu, logp = ops.d_loss(lm, lu, cls, cp0, cp1, G, B, tm, tu, ct, coeff)
        (losses[0] * 1.5).backward()
        got = [x.grad, h.grad, ws[0].grad, bs[0].grad, ws[1].grad, bs[1].grad, cls.grad]
        # reference
        xr, hr = x.detach().float().requires_grad_(True), h.detach().float().requires_grad_(True)
        wr = [w.detach().clone().requires_grad_(True) for w in ws]
        br = [b.detach().clone().requires_grad_(True) for b in bs]
        cr = cls.detach().clone().requires_grad_(True)
        pu_r = torch.sigmoid(F.conv2d(xr.permute(0, 3, 1, 2), wr[0], br[0], stride=4)).view(-1)
        pm_r = torch.sigmoid(F.conv2d(hr.permute(0, 3, 1, 2), wr[1], br[1], stride=4)).view(-1)
        lq = F.log_softmax(cr, 1)
        bce = torch.nn.BCELoss()
        m = u = k = 0
        for g in range(G):
            sl = slice(g * B, (g + 1) * B)
            m = m + bce(pm_r[sl], torch.full((B,), float(tm[g]), device=dev))
            u = u + coeff * bce(pu_r[sl], torch.full((B,), float(tu[g]), device=dev))
            if ct[g] >= 0:
                k = k + (-(cp0 if ct[g] == 0 else cp1) * lq[sl]).sum() / B
        tot = m + u + k
        (tot * 1.5).backward()
        want = [xr.grad, hr.grad, wr[0].grad, br[0].grad, wr[1].grad, br[1].grad, cr.grad]
        errs = [rel(losses, torch.stack([tot, m, u, k])), rel(pm, pm_r), rel(pu, pu_r), rel(logp, lq)] + \
               [rel(a, b) for a, b in zip(got, want)]
        tol = [1e-4, 1e-4, 1e-4, 1e-5, 6e-3, 6e-3, 5e-4, 5e-4, 5e-4, 5e-4, 1e-4]
        ok = all(e < t for e, t in zip(errs, tol))
        print("%s heads G%d B%d C%d E%d %s" % ("PASS" if ok else "FAIL", G, B, C, E1, " ".join("%.1e" % e for e in errs)), flush=True)
        nfail += 0 if ok else 1
    # fused flat Adam (optim.FlatAdam) against torch.optim.Adam, 3 steps, mixed layouts
    from text2img_ekl_b200.optim import FlatAdam
    torch.manual_seed(1)
    net_a = torch.nn.Sequential(torch.nn.Conv2d(16, 32, 3, bias=False), torch.nn.BatchNorm2d(32), torch.nn.Linear(7, 5)).to(dev)
    net_a[0].weight.data = net_a[0].weight.data.contiguous(memory_format=torch.channels_last)
    import copy
    net_b = copy.deepcopy(net_a)
    oa = FlatAdam(net_a.parameters(), lr=2e-4, betas=(0.5, 0.999))
    ob = torch.optim.Adam(net_b.parameters(), lr=2e-4, betas=(0.5, 0.999))
    worst = 0.0
    for it in range(3):
        for pa, pb in zip(net_a.parameters(), net_b.parameters()):
            g = torch.randn_like(pb) * (10.0 ** (it - 1))
            pb.grad = g.clone()
            if pa.grad is None:
                pa.grad = g.clone()
            else:
                pa.grad.copy_(g)
        oa.step(); ob.step()
        worst = max([worst] + [rel(pa, pb) for pa, pb in zip(net_a.parameters(), net_b.parameters())])
    sh = net_a[0].weight._ekl_shadow
    ok = worst < 1e-6 and bool((sh == net_a[0].weight.detach().permute(0, 2, 3, 1).reshape(-1).bfloat16()).all())
    print("%s flat_adam 3 steps worst rel %.1e (bf16 shadow exact: %s)" % ("PASS" if ok else "FAIL", worst, ok), flush=True)
    nfail += 0 if ok else 1
    for (B, D) in [(24, 128), (32, 128), (5, 17)]:
        # fused reparameterisation + KL (model.py:145-152; cub:54-58) on strided halves of one tensor
        x = torch.randn(B, 2 * D, device=dev).requires_grad_(True)
        eps = torch.randn(B, D, device=dev)
        c, std, kl = ops.reparam_kl(x[:, :D], x[:, D:], eps)
        gc, gs = torch.randn_like(c), torch.randn_like(std)
        ((c * gc).sum() + (std * gs).sum() + 2.0 * kl).backward()
        xr = x.detach().clone().requires_grad_(True)
        mu, lv = xr[:, :D], xr[:, D:]
        sr = torch.exp(0.5 * lv)
        cr = eps * sr + mu
        klr = -0.5 * torch.mean(1 + lv - mu.pow(2) - lv.exp())
        ((cr * gc).sum() + (sr * gs).sum() + 2.0 * klr).backward()
        errs = [rel(c, cr), rel(std, sr), abs(float(kl) - float(klr)) / abs(float(klr)), rel(x.grad, xr.grad)]
        ok = all(e < 1e-5 for e in errs)
        print("%s reparam_kl B%d D%d %s" % ("PASS" if ok else "FAIL", B, D, " ".join("%.1e" % e for e in errs)), flush=True)
        nfail += 0 if ok else 1
    return nfail


def run_misc(torch, L, lib, dev, rel):
    """cat(tile(code), x) forward / backward and LeakyReLU-from-output backward."""
    nfail = 0
    for (B, H, W, Cc, Cx) in [(3, 64, 64, 128, 64), (2, 128, 128, 128, 32), (5, 4, 4, 256, 512), (4, 4, 4, 128, 1024),
                              (2, 16, 16, 8, 8)]:
        code = torch.randn(B, Cc, device=dev)
        x = torch.randn(B, H, W, Cx, device=dev).bfloat16()
        out = torch.full((B, H, W, Cc + Cx), float("nan"), device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_cat_code(L.ptr(code), Cc, L.ptr(x), Cx, B, H * W, L.ptr(out), L.stream()))
        want = torch.cat((code.bfloat16().view(B, 1, 1, Cc).expand(B, H, W, Cc), x), 3)
        ok = bool((out == want).all())
        dcat = torch.randn(B, H, W, Cc + Cx, device=dev).bfloat16()
        dcode = torch.zeros(B, Cc, device=dev)
        dx = torch.full((B, H, W, Cx), float("nan"), device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_cat_code_bwd(L.ptr(dcat), Cc, Cx, B, H * W, L.ptr(dcode), L.ptr(dx), L.stream()))
        torch.cuda.synchronize()
        e = rel(dcode, dcat[..., :Cc].float().sum((1, 2)))
        ok = ok and bool((dx == dcat[..., Cc:]).all()) and e < 1e-5
        print("%s cat_code B%d %dx%d Cc%d Cx%d dcode rel %.1e" % ("PASS" if ok else "FAIL", B, H, W, Cc, Cx, e), flush=True)
        nfail += 0 if ok else 1
    for (G, B, H, W) in [(3, 2, 64, 64), (1, 3, 16, 128), (2, 1, 256, 256)]:
        xs = [torch.randn(B, 3, H, W, device=dev) for _ in range(G)]
        out = torch.full((G * B, H // 2, W // 2, 16), float("nan"), device=dev, dtype=torch.bfloat16)
        p = [L.ptr(t) for t in xs] + [None] * (3 - G)
        L.check(lib.ekl_img_s2d(p[0], p[1], p[2], G, B, H, W, L.ptr(out), L.stream()))
        x = torch.cat(xs, 0)
        want = torch.zeros(G * B, H // 2, W // 2, 16, device=dev)
        want[..., :12] = torch.nn.functional.pixel_unshuffle(x, 2).permute(0, 2, 3, 1)
        ok = bool((out == want.bfloat16()).all())
        d = torch.randn(B, H // 2, W // 2, 16, device=dev).bfloat16()
        dx = torch.full((B, 3, H, W), float("nan"), device=dev)
        L.check(lib.ekl_img_s2d_bwd(L.ptr(d), B, H, W, L.ptr(dx), L.stream()))
        wantd = torch.nn.functional.pixel_shuffle(d[..., :12].float().permute(0, 3, 1, 2), 2)
        ok = ok and bool((dx == wantd).all())
        print("%s img_s2d G%d B%d %dx%d" % ("PASS" if ok else "FAIL", G, B, H, W), flush=True)
        nfail += 0 if ok else 1
    # (the folded jointConv is checked end to end in the `fold` group)
    for (B, H, W, C) in [(2, 64, 64, 16), (3, 16, 32, 8)]:
        y = torch.randn(B, H, W, C, device=dev).bfloat16()
        img = torch.full((B, 3, H, W), float("nan"), device=dev)
        L.check(lib.ekl_head_tanh_fwd(L.ptr(y), B, H * W, C, L.ptr(img), L.stream()))
        yr = y[..., :3].float().permute(0, 3, 1, 2).requires_grad_(True)
        want = torch.tanh(yr)
        dimg = torch.randn(B, 3, H, W, device=dev)
        want.backward(dimg)
        dy = torch.full((B, H, W, C), float("nan"), device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_head_tanh_bwd(L.ptr(y), L.ptr(dimg), B, H * W, C, L.ptr(dy), L.stream()))
        e1, e2 = rel(img, want), rel(dy[..., :3].permute(0, 3, 1, 2), yr.grad)
        ok = e1 < 1e-5 and e2 < 6e-3 and bool((dy[..., 3:] == 0).all())
        print("%s head_tanh B%d %dx%d C%d img %.1e dy %.1e" % ("PASS" if ok else "FAIL", B, H, W, C, e1, e2), flush=True)
        nfail += 0 if ok else 1
    for act, fn in ((L.ACT_LRELU, lambda t: torch.nn.functional.leaky_relu(t, 0.2)), (L.ACT_TANH, torch.tanh)):
        B, H, W, Cin, Cout = 2, 32, 32, 32, 16 if act == L.ACT_TANH else 64
        x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
        wm = (torch.randn(Cout, 3, 3, Cin, device=dev) / (9 * Cin) ** 0.5).bfloat16().float()
        conv = L.EklConv(0, B, H, W, Cin, Cout, 0, L.IMPL_TC, 0, 0, act, 0)
        wf = torch.empty(lib.ekl_conv_packed_elems(conv, 0), device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_conv_pack(conv, L.ptr(wm), L.ptr(wf), None, L.stream()))
        y = torch.full((B, H, W, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_conv_fwd(conv, L.ptr(x), L.ptr(wf), L.ptr(y), None, L.stream()))
        yr = fn(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wm.permute(0, 3, 1, 2), padding=1))
        e = rel(y.permute(0, 3, 1, 2), yr)
        ok = e < 6e-3
        print("%s tc conv epilogue act%d rel %.2e" % ("PASS" if ok else "FAIL", act, e), flush=True)
        nfail += 0 if ok else 1
    for n in (8 * 1000, 8 * 123457):
        o = torch.randn(n, device=dev).bfloat16()
        d = torch.randn(n, device=dev).bfloat16()
        dx = torch.empty_like(o)
        L.check(lib.ekl_lrelu_bwd(L.ptr(o), L.ptr(d), L.ptr(dx), n, L.stream()))
        want = torch.where(o.float() > 0, d.float(), 0.2 * d.float()).bfloat16()
        ok = bool((dx == want).all())
        print("%s lrelu_bwd n%d" % ("PASS" if ok else "FAIL", n), flush=True)
        nfail += 0 if ok else 1
    return nfail


def run_vc(torch, L, lib, dev, rel):
    """ekl_linear_bn_relu_fwd / _bwd (VC_NET hidden layers) against F.linear + batch_norm + relu in fp32, train and eval."""
    import torch.nn.functional as F
    from text2img_ekl_b200 import ops
    nfail = 0
    for (B, K, N) in [(24, 1325, 512), (24, 512, 256), (32, 228, 512), (64, 1215, 512), (64, 512, 256), (4, 300, 512), (33, 100, 64)]:
        lin, bn = torch.nn.Linear(K, N).to(dev), torch.nn.BatchNorm1d(N).to(dev)
        with torch.no_grad():
            bn.weight.normal_(1.0, 0.1); bn.bias.normal_(0, 0.1)
        lin_r, bn_r = torch.nn.Linear(K, N).to(dev), torch.nn.BatchNorm1d(N).to(dev)
        lin_r.load_state_dict(lin.state_dict()); bn_r.load_state_dict(bn.state_dict())
        x = torch.randn(B, K, device=dev, requires_grad=True)
        xr = x.detach().clone().requires_grad_(True)
        dh = torch.randn(B, N, device=dev)
        h = ops.linear_bn_relu(x, lin, bn)
        h.backward(dh)
        saved = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        hr = F.relu(bn_r(lin_r(xr)))
        hr.backward(dh)
        torch.backends.cuda.matmul.allow_tf32 = saved
        bn.eval(); bn_r.eval()
        with torch.no_grad():
            he, her = ops.linear_bn_relu(x.detach(), lin, bn), F.relu(bn_r(lin_r(xr.detach())))
        torch.cuda.synchronize()
        errs = dict(h=rel(h, hr), dx=rel(x.grad, xr.grad), dW=rel(lin.weight.grad, lin_r.weight.grad),
                    dgamma=rel(bn.weight.grad, bn_r.weight.grad), dbeta=rel(bn.bias.grad, bn_r.bias.grad),
                    rmean=rel(bn.running_mean, bn_r.running_mean), rvar=rel(bn.running_var, bn_r.running_var), eval=rel(he, her),
                    count=abs(int(bn.num_batches_tracked) - int(bn_r.num_batches_tracked)))
        ok = all(v < 2e-4 for v in errs.values())
        print("%s linear_bn_relu B%d K%d N%d %s" % ("PASS" if ok else "FAIL", B, K, N, " ".join("%s %.1e" % kv for kv in errs.items())), flush=True)
        nfail += 0 if ok else 1
    return nfail


def run_route(torch, L, lib, dev, rel):
    """ekl_caps_route_fwd / _bwd (discriminator capsule class head) against the materialised-prior restatement
    oracle/capsule_ref.py in fp32 (autograd for the gradients), through capsule.CapsuleLinear.forward / forward_norm."""
    from oracle import capsule_ref as R
    from text2img_ekl_b200 import capsule
    nfail = 0
    saved = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    for (B, O, K, scale) in [(32, 201, 512, 1.0), (96, 201, 512, 1.0), (4, 201, 512, 3.0), (5, 37, 64, 2.0), (3, 224, 128, 1.0), (2, 20, 64, 1.0),
                             (3, 100, 256, 0.3)]:
        mod = capsule.CapsuleLinear(out_capsules=O, in_length=K, out_length=16).to(dev)
        x = (torch.randn(B, 16, K, device=dev) * scale).requires_grad_(True)
        xr = x.detach().clone().requires_grad_(True)
        wr = mod.weight.detach().clone().requires_grad_(True)
        assert lib.ekl_caps_route_supported(16, O, 16, 3)
        launches0 = ops_count()
        n = mod.forward_norm(x)
        used = ops_count() - launches0
        gn = torch.randn_like(n)
        n.backward(gn)
        nr = R.capsule_linear(xr, wr).norm(dim=-1)
        nr.backward(gn)
        e = dict(norm=rel(n, nr), dx=rel(x.grad, xr.grad), dW=rel(mod.weight.grad, wr.grad))
        # the vector output (forward()) with its own gradient
        mod.weight.grad = None
        x2 = x.detach().clone().requires_grad_(True)
        v = mod(x2)
        gv = torch.randn_like(v)
        v.backward(gv)
        xr2, wr2 = x.detach().clone().requires_grad_(True), mod.weight.detach().clone().requires_grad_(True)
        vr = R.capsule_linear(xr2, wr2)
        vr.backward(gv)
        torch.cuda.synchronize()
        e.update(v=rel(v, vr), dx_v=rel(x2.grad, xr2.grad), dW_v=rel(mod.weight.grad, wr2.grad))
        ok = all(val < 2e-4 for val in e.values()) and used == 1          # forward: one routing launch
        print("%s caps_route B%d O%d K%d x%.1f launches %d %s" % ("PASS" if ok else "FAIL", B, O, K, scale, used,
                                                                 " ".join("%s %.1e" % kv for kv in e.items())), flush=True)
        nfail += 0 if ok else 1
    torch.backends.cuda.matmul.allow_tf32 = saved
    return nfail


def ops_count():
    from text2img_ekl_b200 import ops
    return ops.LAUNCHES[0]


def run_fold(torch, L, lib, dev, rel):
    """Folded jointConv (ekl_code_bias9_fwd / ekl_border_sums9 / ekl_code_bias9_bwd + a conv over a channel window of the
    master filter) and the image head (3 real filters in a 16-wide tile) through the autograd layer, against the
    reference formulation conv3x3(cat(tile(code), h)) / conv3x3(h) in fp32."""
    import torch.nn.functional as F
    from text2img_ekl_b200 import ops
    nfail = 0
    for (B, H, W, ngf, ef) in [(3, 16, 16, 64, 128), (2, 32, 32, 32, 128), (4, 16, 8, 64, 256), (2, 64, 64, 64, 128)]:
        N = 2 * ngf
        code = torch.randn(B, ef, device=dev).requires_grad_(True)
        h = torch.randn(B, H, W, ngf, device=dev).bfloat16().requires_grad_(True)
        wq = (torch.randn(N, ngf + ef, 3, 3, device=dev) / (9 * (ngf + ef)) ** 0.5)
        wq[:, ef:] = wq[:, ef:].bfloat16().float()                      # the conv part is exact in bf16, the code part stays fp32
        w = wq.contiguous(memory_format=torch.channels_last).requires_grad_(True)
        spec = ops.ConvSpec(ops.S1, ngf, N, impl=L.IMPL_TC, w_cin_total=ngf + ef, w_cin_off=ef)
        bias9 = ops.code_bias9(code, w)
        y, stats = ops.conv_bias9(h, w, bias9, spec, want_stats=True)
        dy = torch.randn(B, H, W, N, device=dev).bfloat16()
        y.backward(dy)
        torch.cuda.synchronize()
        cr, hr, wr = code.detach().clone().requires_grad_(True), h.detach().float().requires_grad_(True), wq.clone().requires_grad_(True)
        xin = torch.cat((cr.view(B, ef, 1, 1).expand(B, ef, H, W), hr.permute(0, 3, 1, 2)), 1)
        yr = F.conv2d(xin, wr, padding=1)
        yr.backward(dy.float().permute(0, 3, 1, 2))
        yf = y.detach().float().reshape(-1, N)
        errs = dict(y=rel(y.permute(0, 3, 1, 2), yr), dh=rel(h.grad, hr.grad),
                    dcode=rel(code.grad, cr.grad), dw_code=rel(w.grad[:, :ef], wr.grad[:, :ef]), dw_h=rel(w.grad[:, ef:], wr.grad[:, ef:]),
                    s1=rel(stats.view(2, N)[0].float(), yf.sum(0)), s2=rel(stats.view(2, N)[1].float(), (yf * yf).sum(0)))
        tol = dict(y=6e-3, dh=6e-3, dcode=2e-3, dw_code=2e-3, dw_h=2e-3, s1=1e-2, s2=1e-3)
        ok = all(errs[k] < tol[k] for k in errs)
        print("%s fold B%d %dx%d ngf%d ef%d %s" % ("PASS" if ok else "FAIL", B, H, W, ngf, ef, " ".join("%s %.1e" % kv for kv in errs.items())), flush=True)
        nfail += 0 if ok else 1
    for (B, H, W, ngf) in [(2, 64, 64, 64), (3, 32, 32, 32), (2, 64, 64, 16)]:
        h = torch.randn(B, H, W, ngf, device=dev).bfloat16().requires_grad_(True)
        wq = (torch.randn(3, ngf, 3, 3, device=dev) / (9 * ngf) ** 0.5).bfloat16().float()
        w = wq.contiguous(memory_format=torch.channels_last).requires_grad_(True)
        spec = ops.ConvSpec(ops.S1, ngf, 16, impl=L.IMPL_TC, w_cout_valid=3)
        y, _ = ops.conv(h, w, spec)
        dy = torch.zeros(B, H, W, 16, device=dev, dtype=torch.bfloat16)
        dy[..., :3] = torch.randn(B, H, W, 3, device=dev).bfloat16()
        y.backward(dy)
        torch.cuda.synchronize()
        hr, wr = h.detach().float().requires_grad_(True), wq.clone().requires_grad_(True)
        yr = F.conv2d(hr.permute(0, 3, 1, 2), wr, padding=1)
        yr.backward(dy[..., :3].float().permute(0, 3, 1, 2))
        errs = dict(y=rel(y[..., :3].permute(0, 3, 1, 2), yr), pad=float(y[..., 3:].float().abs().max()),
                    dh=rel(h.grad, hr.grad), dw=rel(w.grad, wr.grad))
        ok = errs["y"] < 6e-3 and errs["pad"] == 0.0 and errs["dh"] < 6e-3 and errs["dw"] < 2e-3
        print("%s head window B%d %dx%d ngf%d %s" % ("PASS" if ok else "FAIL", B, H, W, ngf, " ".join("%s %.1e" % kv for kv in errs.items())), flush=True)
        nfail += 0 if ok else 1
    return nfail


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        sys.exit(1 if run_group(a.group) else 0)
    bad = 0
    for g in ([a.group] if a.group else GROUPS):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--group", g], timeout=300)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = -9
            print("TIMEOUT group %s" % g, flush=True)
        print("== group %s exit %d" % (g, rc), flush=True)
        bad += rc != 0
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
