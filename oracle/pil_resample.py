"""CPU restatement of the image pyramid of the reference's loader -- TEST INFRASTRUCTURE.

datasets.py:43-68 `get_imgs`: the cropped / flipped final-size image is resized to every lower stage size with
`transforms.Scale(imsize[i])` = PIL `Image.resize((s, s), Image.BILINEAR)` and each level goes through
`ToTensor()` + `Normalize((0.5,)*3, (0.5,)*3)`.  PIL's 8-bit resize (libImaging/Resample.c, third-party, Pillow; the
version in this image is pinned by tests/test_oracle_golden.py::test_pyramid_restatement_equals_pil) is restated here in
integer arithmetic:
  * precompute_coeffs: per output pixel the window [xmin, xmin+n) and the normalised triangle weights, filter support
    scaled by the down-scale factor (anti-aliasing);
  * normalize_coeffs_8bpc: weights -> fixed point with 22 fractional bits, round half away from zero;
  * one horizontal then one vertical pass, each accumulating pixel*weight from 1 << 21 and clipping (acc >> 22) to
    [0, 255] -- the intermediate image is uint8.
Bit-exact against PIL for uint8 RGB images; the product kernel (csrc/img_ops.cu ekl_img_pyramid_level) is checked against
this on the GPU.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bilinear_coeffs(in_size, out_size):
    """-> (bounds int32 [out,2] = (xmin, n), kk int32 [out, ksize]) exactly as Resample.c computes them (double math)."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale > 1.0 else 1.0
    support = 1.0 * filterscale                          # bilinear filter support = 1.0
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(n):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0.0 else v
            w[x] = 1.0 - v if v < 1.0 else 0.0
            ww += w[x]
        for x in range(n):
            if ww != 0.0:
                w[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, n)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """one resampling pass of a uint8 [H, W, C] image along `axis` (1 = horizontal, 0 = vertical)."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)               # resampled axis first
    out = np.zeros((bounds.shape[0],) + src.shape[1:], np.int64)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_bilinear_u8(img, size):
    """PIL Image.resize((size, size), BILINEAR) of a uint8 [H, W, 3] array."""
    h, w = img.shape[:2]
    if w != size:
        b, k = bilinear_coeffs(w, size)
        img = _pass(img, b, k, 1)
    if h != size:
        b, k = bilinear_coeffs(h, size)
        img = _pass(img, b, k, 0)
    return img


def to_normalised_chw(img_u8):
    """ToTensor() + Normalize((0.5,)*3, (0.5,)*3): uint8 HWC -> float32 CHW in [-1, 1], in fp32 arithmetic like torch."""
    t = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)
    return (t - np.float32(0.5)) / np.float32(0.5)


def pyramid(img_u8, sizes):
    """datasets.py:60-66: every level but the last is a bilinear resize of the final-size image; -> list of float32 CHW."""
    out = []
    for i, s in enumerate(sizes):
        lvl = resize_bilinear_u8(img_u8, s) if i < len(sizes) - 1 else img_u8
        out.append(to_normalised_chw(lvl))
    return out
