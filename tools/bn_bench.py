"""Timing of the BatchNorm+activation kernels at real layer shapes (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from text2img_ekl_b200 import _lib as L

ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
CUSTOM = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]     # M,Cy,groups,act
SHAPES = [("custom",) + c for c in CUSTOM] or [("s3_up", 24 * 256 * 256, 32, 1, L.ACT_GLU), ("s3_res", 24 * 128 * 128, 64, 1, L.ACT_GLU),
          ("s2_up4", 32 * 64 * 64, 128, 1, L.ACT_GLU), ("d_l2", 3 * 24 * 64 * 64, 128, 3, L.ACT_LRELU),
          ("res_none", 24 * 128 * 128, 32, 1, L.ACT_NONE), ("tail", 3 * 24 * 16, 1024, 3, L.ACT_LRELU)]
lib = L.lib()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(ITERS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


for name, M, Cy, groups, act in SHAPES:
    Co = Cy // 2 if act == L.ACT_GLU else Cy
    y = torch.randn(M, Cy, device=dev).bfloat16()
    dout = torch.randn(M, Co, device=dev).bfloat16()
    res = torch.randn(M, Co, device=dev).bfloat16() if act == L.ACT_NONE else None
    gamma, beta = torch.ones(Cy, device=dev), torch.zeros(Cy, device=dev)
    part = torch.zeros(groups, 2, Cy, device=dev, dtype=torch.float64)          # fp64 statistics sums (accumulated)
    mean, rstd = torch.empty(groups, Cy, device=dev), torch.empty(groups, Cy, device=dev)
    out = torch.empty(M, Co, device=dev, dtype=torch.bfloat16)
    dy = torch.empty_like(y)
    sums = torch.zeros(max(int(lib.ekl_bn_bwd_scratch_doubles(M, Cy, groups, act)), 2), device=dev, dtype=torch.float64)
    dg, db = torch.zeros(Cy, device=dev), torch.zeros(Cy, device=dev)
    st = L.stream()
    t_stats = timeit(lambda: L.check(lib.ekl_col_stats(L.ptr(y), M, Cy, groups, L.ptr(part), st)))
    t_fwd = timeit(lambda: L.check(lib.ekl_bn_act_fwd(L.ptr(y), M, Cy, groups, L.ptr(part), 1e-5, 0.1, L.ptr(mean), L.ptr(rstd), None,
                                                      None, L.ptr(gamma), L.ptr(beta), act, L.ptr(res), L.ptr(out), st)))
    t_bwd = timeit(lambda: L.check(lib.ekl_bn_act_bwd(L.ptr(y), L.ptr(dout), M, Cy, groups, L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
                                                      L.ptr(beta), act, L.ptr(sums), L.ptr(dg), L.ptr(db), L.ptr(dy), st)))
    b_fwd = M * (Cy + Co + (Co if res is not None else 0)) * 2
    b_bwd = M * (Cy + Co + Cy) * 2          # minimal traffic: y and dout read once, dy written once
    print("%-9s M=%8d Cy=%4d g%d act%d | stats %7.1f us %6.0f GB/s | fwd %7.1f us %6.0f GB/s | bwd %7.1f us %6.0f GB/s (minimal-traffic bytes)"
          % (name, M, Cy, groups, act, t_stats, M * Cy * 2 / t_stats / 1e3, t_fwd, b_fwd / t_fwd / 1e3, t_bwd, b_bwd / t_bwd / 1e3), flush=True)
