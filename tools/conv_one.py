#!/usr/bin/env python
"""Time (and let ncu capture) one conv kernel call at one shape:  conv_one.py MODE B H W CIN COUT fwd|dgrad|wgrad [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from text2img_ekl_b200 import _lib as L

mode, B, H, W, Cin, Cout = [int(v) for v in sys.argv[1:7]]
kind = sys.argv[7]
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 5
lib = L.lib()
dev = torch.device("cuda")
K = 4 if mode == 2 else 3
Ho, Wo = (2 * H, 2 * W) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
conv = L.EklConv(mode, B, H, W, Cin, Cout, B, 0, 0, 0, 0, 0)
x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
dy = torch.randn(B, Ho, Wo, Cout, device=dev).bfloat16()
wm = torch.randn(Cout, K, K, Cin, device=dev) * 0.05
wf = torch.empty(lib.ekl_conv_packed_elems(conv, 0), device=dev, dtype=torch.bfloat16)
wd = torch.empty(lib.ekl_conv_packed_elems(conv, 1), device=dev, dtype=torch.bfloat16)
L.check(lib.ekl_conv_pack(conv, L.ptr(wm), L.ptr(wf), L.ptr(wd), L.stream()))
y = torch.empty(B, Ho, Wo, Cout, device=dev, dtype=torch.bfloat16)
dx = torch.empty_like(x)
dw = torch.zeros(Cout, K, K, Cin, device=dev)
stats = torch.zeros(4, 2, Cout, device=dev, dtype=torch.float64)
fn = {"fwd": lambda: L.check(lib.ekl_conv_fwd(conv, L.ptr(x), L.ptr(wf), L.ptr(y), L.ptr(stats), L.stream())),
      "dgrad": lambda: L.check(lib.ekl_conv_bwd_data(conv, L.ptr(dy), L.ptr(wd), L.ptr(dx), L.stream())),
      "wgrad": lambda: L.check(lib.ekl_conv_bwd_weight(conv, L.ptr(x), L.ptr(dy), L.ptr(dw), L.stream()))}[kind]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for _ in range(2):
    fn()
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
t = sorted(ts)[len(ts) // 2]
ref = 2.0 * B * Ho * Wo * Cout * Cin * K * K
print("%s mode%d %dx%dx%d %d>%d : %.1f us  %.0f TF/s (ref count)" % (kind, mode, B, H, W, Cin, Cout, t, ref / t / 1e6))
