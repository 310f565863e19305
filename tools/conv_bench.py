"""Per-layer timing of the conv kernels at the real layer shapes of a config (CUDA events, L2 flushed between
iterations).  Prints executed and reference-count TFLOP/s per direction."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from text2img_ekl_b200 import _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ONLY = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 8
# (name, mode, B, H, W, Cin, Cout)
LAYERS = [("up1", 1, B, 4, 4, 1024, 1024), ("up2", 1, B, 8, 8, 512, 512), ("up3", 1, B, 16, 16, 256, 256),
          ("up4", 1, B, 32, 32, 128, 128), ("joint2", 0, B, 64, 64, 320, 128), ("res2a", 0, B, 64, 64, 64, 128),
          ("res2b", 0, B, 64, 64, 64, 64), ("up5", 1, B, 64, 64, 64, 64),
          ("d64_2", 2, 3 * B, 32, 32, 64, 128), ("d64_3", 2, 3 * B, 16, 16, 128, 256), ("d64_4", 2, 3 * B, 8, 8, 256, 512),
          ("d128_2", 2, 3 * B, 64, 64, 64, 128), ("d128_5", 2, 3 * B, 8, 8, 512, 1024), ("d128_6", 0, 3 * B, 4, 4, 1024, 512),
          ("djoint", 0, 3 * B, 4, 4, 768, 512)]


def main():
    lib = L.lib()
    L.check(lib.ekl_require_sm100())
    dev = torch.device("cuda")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    print("%-8s %5s %14s | %8s %8s %8s (us) | %7s %7s %7s (TF/s exec) | ref-count TF/s fwd" % ("layer", "mode", "shape", "fwd", "dgrad", "wgrad", "fwd", "dgrad", "wgrad"))
    for name, mode, b, H, W, Cin, Cout in LAYERS:
        if ONLY and name not in ONLY:
            continue
        K = 4 if mode == 2 else 3
        Ho, Wo = (2 * H, 2 * W) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
        conv = L.EklConv(mode, b, H, W, Cin, Cout, 0, 0, 0, 0, 0, 0)
        x = torch.randn(b, H, W, Cin, device=dev).bfloat16()
        dy = torch.randn(b, Ho, Wo, Cout, device=dev).bfloat16()
        wm = torch.randn(Cout, K, K, Cin, device=dev) * 0.05
        wf = torch.empty(lib.ekl_conv_packed_elems(conv, 0), device=dev, dtype=torch.bfloat16)
        wd = torch.empty(lib.ekl_conv_packed_elems(conv, 1), device=dev, dtype=torch.bfloat16)
        L.check(lib.ekl_conv_pack(conv, L.ptr(wm), L.ptr(wf), L.ptr(wd), L.stream()))
        y = torch.empty(b, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
        dx = torch.empty_like(x)
        dw = torch.zeros(Cout, K, K, Cin, device=dev)
        stats = torch.zeros(4, 2, Cout, device=dev, dtype=torch.float64)
        fns = [lambda: lib.ekl_conv_fwd(conv, L.ptr(x), L.ptr(wf), L.ptr(y), L.ptr(stats), L.stream()),
               lambda: lib.ekl_conv_bwd_data(conv, L.ptr(dy), L.ptr(wd), L.ptr(dx), L.stream()),
               lambda: lib.ekl_conv_bwd_weight(conv, L.ptr(x), L.ptr(dy), L.ptr(dw), L.stream())]
        times = []
        for fn in fns:
            for _ in range(3):
                L.check(fn())
            ts = []
            for _ in range(ITERS):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); L.check(fn()); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            times.append(sorted(ts)[len(ts) // 2])
        ref_flop = 2.0 * b * Ho * Wo * Cout * Cin * K * K
        exe_flop = ref_flop / 2.25 if mode == 1 else ref_flop
        tf = [exe_flop / (t * 1e-6) / 1e12 for t in times]
        print("%-8s %5d %14s | %8.1f %8.1f %8.1f      | %7.1f %7.1f %7.1f            | %7.1f" % (
            name, mode, "%dx%dx%d>%d" % (H, W, Cin, Cout), times[0], times[1], times[2], tf[0], tf[1], tf[2], ref_flop / (times[0] * 1e-6) / 1e12))


main()
