"""Device side of the reference's input pipeline (datasets.py:43-68 `get_imgs`).

The reference's loader builds, per sample and on the host, a pyramid of the cropped / flipped final-size image: every
stage below the last gets `transforms.Scale(imsize[i])` (a PIL BILINEAR resize), every level `ToTensor()` +
`Normalize((0.5,)*3, (0.5,)*3)`, and the batch then crosses PCIe as BRANCH_NUM fp32 tensors per image set.  Here the
loader only has to deliver the final-size uint8 crops ([B, S, S, 3], 12x fewer bytes than the fp32 pyramid for three
stages); `image_pyramid` builds all levels on the device, bit-exactly equal to what PIL + torchvision produce
(include/ekl_b200.h: ekl_img_pyramid_level).  Decoding, cropping and flipping stay host-side I/O (out of scope).
"""
import math

import torch

from . import _lib as L
from . import ops

PRECISION_BITS = 32 - 8 - 2       # PIL libImaging/Resample.c: fixed-point weights of the 8-bit path

_TABLES = {}


def pil_bilinear_tables(in_size, out_size):
    """PIL's coefficient tables for a BILINEAR resize in_size -> out_size, computed the way Resample.c does (double
    arithmetic, support scaled by the down-scale factor, weights normalised, then round-half-away to 22 fractional bits).
    Returns (bounds int32 [out, 2] = (first source index, count), kk int32 [out, ksize])."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale > 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = torch.zeros(out_size, 2, dtype=torch.int32)
    kk = torch.zeros(out_size, ksize, dtype=torch.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        n = xmax - xmin
        w = []
        for x in range(n):
            v = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - v if v < 1.0 else 0.0)
        ww = 0.0
        for v in w:                        # left-to-right double additions, as in the C loop
            ww += v
        for x in range(n):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS)) if v >= 0 else int(-0.5 + v * (1 << PRECISION_BITS))
        bounds[xx, 0], bounds[xx, 1] = xmin, n
    return bounds, kk


def image_pyramid(crops_u8, sizes):
    """crops_u8: uint8 [B, S, S, 3] on the device (the loader's final-size crops, PIL layout); sizes: stage sizes in
    ascending order, the last equal to S (cfg.TREE.BASE_SIZE * SCALE**i).  -> list of fp32 [B, 3, s, s] in [-1, 1],
    equal bit for bit to `normalize(transforms.Scale(s)(img))` / `normalize(img)` of datasets.py:60-66."""
    assert crops_u8.dtype == torch.uint8 and crops_u8.is_cuda and crops_u8.dim() == 4 and crops_u8.shape[3] == 3
    crops_u8 = crops_u8.contiguous()
    B, S = crops_u8.shape[0], crops_u8.shape[1]
    assert crops_u8.shape[2] == S and sizes[-1] == S
    dev = crops_u8.device
    out = []
    for s in sizes:
        img = torch.empty(B, 3, s, s, device=dev, dtype=torch.float32)
        if s == S:
            L.check(L.lib().ekl_img_pyramid_level(L.ptr(crops_u8), B, S, s, None, None, 0, None, L.ptr(img), L.stream()))
            ops._count()
        else:
            key = (S, s, dev)
            if key not in _TABLES:
                b, k = pil_bilinear_tables(S, s)
                _TABLES[key] = (b.to(dev), k.to(dev))
            b, k = _TABLES[key]
            tmp = torch.empty(B * S * s * 3, device=dev, dtype=torch.uint8)
            L.check(L.lib().ekl_img_pyramid_level(L.ptr(crops_u8), B, S, s, L.ptr(b), L.ptr(k), k.shape[1], L.ptr(tmp), L.ptr(img),
                                                  L.stream()))
            ops._count(2)
        out.append(img)
    return out


def stage_images(imgs, num_Ds, device):
    """The image half of `prepare_data` (trainer.py:266-290, cub:295-320).  imgs is either what the reference's loader
    delivers -- a list of BRANCH_NUM fp32 [B, 3, s, s] tensors, moved to the device as they are -- or ONE uint8
    [B, S, S, 3] tensor of final-size crops, from which the pyramid is built on the device."""
    if torch.is_tensor(imgs) and imgs.dtype == torch.uint8:
        from .miscc.config import cfg
        S = imgs.shape[1]
        sizes = [cfg.TREE.BASE_SIZE * 2 ** i for i in range(num_Ds)]       # datasets.py:93-96: the loader doubles per level
        if sizes[-1] != S:                                                 # TREE.SCALE != 2: the stages the generator emits
            sizes = [cfg.TREE.BASE_SIZE * cfg.TREE.SCALE ** i for i in range(num_Ds)]
        if sizes[-1] != S:
            raise ValueError("uint8 crops are %dx%d but the last stage is %d" % (S, S, sizes[-1]))
        return image_pyramid(imgs.to(device, non_blocking=True), sizes)
    return [imgs[i].to(device, non_blocking=True) for i in range(num_Ds)]
