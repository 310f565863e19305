"""Host logic of the conv family, no GPU: the gather plans (tap tables, parity views, packed-filter sums) that
the tcgen05 / SIMT kernels execute are emulated in torch on the CPU from ekl_conv_plan_dump() and compared with
F.conv2d / autograd for forward and data-gradient, and the weight-gradient scatter is checked the same way."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from text2img_ekl_b200 import _lib as L


def dump(conv, dgrad):
    buf = (C.c_int * 1024)()
    L.check(L.lib().ekl_conv_plan_dump(conv, dgrad, buf, 1024))
    h = list(buf[:11])
    nvar, ntaps = h[0], h[1]
    taps = [[tuple(buf[11 + (v * ntaps + t) * 8: 11 + (v * ntaps + t) * 8 + 8]) for t in range(ntaps)] for v in range(nvar)]
    return dict(nvar=nvar, ntaps=ntaps, n_a=h[2], m=(h[3], h[4], h[5]), Cin=h[6], N=h[7], transposed=h[8], KK=h[9] * h[10]), taps


def parity_views(t, n):
    """t [B,H,W,C] -> list of n views (n=1: [t]; n=4: (ph,pw) parity classes)."""
    if n == 1:
        return [t]
    return [t[:, ph::2, pw::2, :] for ph in range(2) for pw in range(2)]


def shifted(v, dh, dw):
    """out[b,h,w] = v[b,h+dh,w+dw] with zeros outside."""
    B, H, W, Cc = v.shape
    out = torch.zeros_like(v)
    hs, he = max(0, -dh), min(H, H - dh)
    ws, we = max(0, -dw), min(W, W - dw)
    if hs < he and ws < we:
        out[:, hs:he, ws:we] = v[:, hs + dh:he + dh, ws + dw:we + dw]
    return out


def pack(wm, info, taps):
    """master [Cout,KK,Cin] -> packed [nvar][N][ntaps][K]."""
    Cout, KK, Cin = wm.shape
    out = torch.zeros(info["nvar"], info["N"], info["ntaps"], info["Cin"], dtype=wm.dtype)
    for v in range(info["nvar"]):
        for t in range(info["ntaps"]):
            tp = taps[v][t]
            acc = sum(wm[:, tp[4 + i], :] for i in range(tp[3]))          # [Cout, Cin]
            out[v, :, t, :] = acc.t() if info["transposed"] else acc
    return out


def run_gather(A, out_shape, info, taps, wp):
    """emulate out_v[p,n] = sum_t sum_k A_map[p+(dh,dw),k] * Wp[v][n][t][k]; returns full-resolution output."""
    views = parity_views(A, info["n_a"])
    out = torch.zeros(out_shape, dtype=A.dtype)
    outs = parity_views(out, info["nvar"])
    for v in range(info["nvar"]):
        acc = 0
        for t in range(info["ntaps"]):
            m, dh, dw = taps[v][t][:3]
            acc = acc + torch.einsum("bhwk,nk->bhwn", shifted(views[m], dh, dw), wp[v, :, t, :])
        outs[v].copy_(acc)
    return out


CASES = [(0, 2, 6, 5, 4, 6), (1, 2, 3, 4, 5, 3), (2, 2, 8, 6, 3, 5), (0, 1, 4, 4, 8, 8), (1, 1, 4, 4, 2, 2), (2, 3, 4, 4, 4, 2),
         # edge shapes: single row / column maps, one channel, one sample, strongly non-square, the 2x2 -> 1x1 stride-2 case
         (0, 1, 1, 7, 3, 2), (0, 2, 5, 1, 1, 4), (0, 1, 1, 1, 2, 3), (1, 1, 1, 1, 3, 2), (1, 2, 1, 5, 2, 1), (1, 1, 7, 2, 1, 3),
         (2, 1, 2, 2, 3, 2), (2, 2, 2, 10, 1, 3), (2, 1, 12, 2, 2, 1)]


@pytest.mark.parametrize("mode,B,H,W,Cin,Cout", CASES)
def test_plan_forward_and_dgrad_and_wgrad(mode, B, H, W, Cin, Cout):
    torch.manual_seed(mode * 7 + B)
    K = 4 if mode == 2 else 3
    Ho, Wo = (2 * H, 2 * W) if mode == 1 else ((H // 2, W // 2) if mode == 2 else (H, W))
    conv = L.EklConv(mode, B, H, W, Cin, Cout, 0, 1, 0, 0, 0)
    x = torch.randn(B, H, W, Cin, dtype=torch.float64)
    wm = torch.randn(Cout, K * K, Cin, dtype=torch.float64)
    xr = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = wm.view(Cout, K, K, Cin).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    xin = F.interpolate(xr, scale_factor=2, mode="nearest") if mode == 1 else xr
    yr = F.conv2d(xin, wr, stride=2 if mode == 2 else 1, padding=1)
    dy = torch.randn(B, Ho, Wo, Cout, dtype=torch.float64)
    yr.backward(dy.permute(0, 3, 1, 2))
    # forward
    info, taps = dump(conv, 0)
    y = run_gather(x, (B, Ho, Wo, Cout), info, taps, pack(wm, info, taps).double())
    assert torch.allclose(y.permute(0, 3, 1, 2), yr, atol=1e-9)
    # weight gradient through the forward plan: dW[co][src][ci] += sum_p dY_v[p,co] * A_map[p+(dh,dw),ci]
    views = parity_views(x, info["n_a"])
    dys = parity_views(dy, info["nvar"])
    dw = torch.zeros(Cout, K * K, Cin, dtype=torch.float64)
    for v in range(info["nvar"]):
        for t in range(info["ntaps"]):
            m, dh, dw_, nsrc = taps[v][t][:4]
            contrib = torch.einsum("bhwn,bhwk->nk", dys[v], shifted(views[m], dh, dw_))
            for i in range(nsrc):
                dw[:, taps[v][t][4 + i], :] += contrib
    assert torch.allclose(dw.view(Cout, K, K, Cin).permute(0, 3, 1, 2), wr.grad, atol=1e-9)
    # data gradient
    info, taps = dump(conv, 1)
    assert info["transposed"] == 1 and info["N"] == Cin and info["Cin"] == Cout
    dx = run_gather(dy, (B, H, W, Cin), info, taps, pack(wm, info, taps).double())
    assert torch.allclose(dx.permute(0, 3, 1, 2), xr.grad, atol=1e-9)


def test_library_exports_every_declared_symbol():
    import re, os
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ekl_b200.h")).read()
    names = set(re.findall(r"\b(ekl_[a-z0-9_]+)\s*\(", hdr))
    lib = L.lib()
    for n in sorted(names):
        assert hasattr(lib, n), n
        assert n in L.SIGNATURES, "binding missing for " + n
    assert lib.ekl_version() == 100


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.lib().ekl_require_sm100() != 0
    assert L.lib().ekl_last_error()


def test_space_to_depth_filter_equals_conv4x4_stride2():
    """The discriminator stem identity the tcgen05 path relies on (model._Encode16): conv4x4/s2/p1 over a 3-channel
    image == conv3x3/s1/p1 over its space-to-depth image with the re-indexed filter (fp32 on the CPU, exact algebra)."""
    from text2img_ekl_b200 import model
    torch.manual_seed(0)
    x = torch.randn(2, 3, 16, 12)
    w = torch.randn(8, 3, 4, 4)
    idx, mask = model._s2d_filter_index()
    w2 = (w.reshape(8, 48)[:, idx] * mask).view(8, 3, 3, 16).permute(0, 3, 1, 2)
    xs = torch.zeros(2, 16, 8, 6)
    for c in range(3):
        for ph in range(2):
            for pw in range(2):
                xs[:, (c * 2 + ph) * 2 + pw] = x[:, c, ph::2, pw::2]
    a = F.conv2d(x, w, stride=2, padding=1)
    b = F.conv2d(xs, w2, stride=1, padding=1)
    assert torch.allclose(a, b, atol=1e-5)
    assert int(mask.sum()) == 48           # every master-filter entry appears exactly once


def test_joint_conv_fold_identity():
    """conv3x3(cat(tile(c), h)) == conv3x3(h, Wx) + bias9[b, border class] with bias9 = VALID @ (Wc . c) -- the algebra
    behind NEXT_STAGE_G._joint / ekl_code_bias9_fwd / ekl_conv_fwd_bias9, in fp32 on the CPU.  VALID[q, t] = 1 iff tap
    t = (kh, kw) of a 3x3 / pad-1 conv reads inside the map for a pixel of border class q = 3*rc + cc (the predicate
    csrc/codefold.cu::tap_valid evaluates)."""
    valid = torch.zeros(9, 9)
    ok = lambda cls, k: not ((cls == 0 and k == 0) or (cls == 2 and k == 2))
    for rc in range(3):
        for cc in range(3):
            for kh in range(3):
                for kw in range(3):
                    valid[rc * 3 + cc, kh * 3 + kw] = float(ok(rc, kh) and ok(cc, kw))
    torch.manual_seed(0)
    B, Cc, Cx, N, H, W = 2, 5, 4, 6, 7, 5
    c, h = torch.randn(B, Cc), torch.randn(B, Cx, H, W)
    w = torch.randn(N, Cc + Cx, 3, 3)
    ref = F.conv2d(torch.cat((c.view(B, Cc, 1, 1).expand(B, Cc, H, W), h), 1), w, padding=1)
    T = torch.einsum("bc,nckl->bkln", c, w[:, :Cc]).reshape(B, 9, N)
    bias9 = torch.einsum("qt,btn->bqn", valid, T)
    cls_h = torch.tensor([0] + [1] * (H - 2) + [2])
    cls_w = torch.tensor([0] + [1] * (W - 2) + [2])
    q = cls_h.view(H, 1) * 3 + cls_w.view(1, W)                      # [H, W]
    got = F.conv2d(h, w[:, Cc:], padding=1) + bias9[:, q].permute(0, 3, 1, 2)
    assert torch.allclose(ref, got, atol=1e-4)


def test_split_k_and_operand_planning_host_logic():
    """Host-side planning that needs no GPU: which plans get a split-K workspace, which data-gradients can read the
    forward-packed filter (include/ekl_b200.h)."""
    lib = L.lib()
    mk = lambda mode, B, H, W, Cin, Cout, gb=0: L.EklConv(mode, B, H, W, Cin, Cout, gb, L.IMPL_TC, 0, 0, 0, L.W_KRSC)
    # 4x4 discriminator tail, batch 24: 3 output tiles of 128 rows, contraction 9*2048 -> split
    tail = mk(L.S1, 24, 4, 4, 2048, 1024, 24)
    assert lib.ekl_conv_workspace_elems(tail, 0) == 24 * 16 * 1024
    assert lib.ekl_conv_workspace_elems(tail, 1) == 24 * 16 * 2048
    assert lib.ekl_conv_route(tail, 0) == 2 and lib.ekl_conv_route(tail, 1) == 2          # split-K + finishing pass
    # a layer that fills the machine is never split
    big = mk(L.DOWN2, 72, 64, 64, 128, 256, 24)
    assert lib.ekl_conv_workspace_elems(big, 0) == 0 and lib.ekl_conv_workspace_elems(big, 1) == 0
    assert lib.ekl_conv_route(big, 0) == 0
    # small-channel 3x3 layers run on the resident-filter kernel: no workspace, transposed operand still packed
    res = mk(L.S1, 24, 128, 128, 32, 64, 24)
    assert lib.ekl_conv_workspace_elems(res, 0) == 0 and lib.ekl_conv_route(res, 0) == 1
    assert lib.ekl_conv_dgrad_from_fwd(res) == 0
    # stride-1 / stride-2 convs with 64-multiple channels: data-gradient straight from the forward-packed filter
    assert lib.ekl_conv_dgrad_from_fwd(tail) == 1 and lib.ekl_conv_dgrad_from_fwd(big) == 1
    assert lib.ekl_conv_dgrad_from_fwd(mk(L.UP2, 24, 8, 8, 512, 512)) == 0          # pre-summed taps need their own pack
    # sub-pixel plans with a short contraction and a few tiles per SM run on the resident-filter kernel as well: the
    # data-gradient of the 64 -> 128 discriminator conv (contraction 128) and the small-channel up-convs of the generator
    d1 = mk(L.DOWN2, 72, 128, 128, 64, 128, 24)
    assert lib.ekl_conv_route(d1, 1) == 1 and lib.ekl_conv_dgrad_from_fwd(d1) == 0 and lib.ekl_conv_route(d1, 0) == 0
    assert lib.ekl_conv_route(mk(L.UP2, 24, 128, 128, 32, 32, 24), 0) == 1
    assert lib.ekl_conv_route(mk(L.UP2, 24, 64, 64, 64, 64, 24), 0) == 1
    assert lib.ekl_conv_route(mk(L.UP2, 24, 32, 32, 128, 128, 24), 0) == 0          # 192 tiles: stays on the CTA-pair kernel
    assert lib.ekl_conv_route(mk(L.DOWN2, 72, 32, 32, 64, 128, 24), 1) == 0        # too few tiles per SM
    from text2img_ekl_b200 import ops
    s1, s2 = ops.ConvSpec(L.DOWN2, 64, 128), ops.ConvSpec(L.DOWN2, 128, 256)
    s1.w_layout = s2.w_layout = L.W_KRSC
    assert not s1.dgrad_from_fwd() and s2.dgrad_from_fwd()        # the host keeps a transposed operand for the former


def test_batchnorm_scratch_and_fused_finishing_eligibility_on_host():
    """Host-side answers of the round-2 BatchNorm entry points (no kernel launch): the backward scratch is one sum per
    128-byte line for the streaming passes and empty for the single-launch small layers; the fused split-K finishing +
    BatchNorm kernel is offered exactly for split plans with <= 768 pixels per group and Cout % 32 == 0."""
    lib = L.lib()
    spread = 16
    assert lib.ekl_bn_bwd_scratch_doubles(24 * 64 * 64, 128, 1, L.ACT_GLU) == 1 * 2 * 128 * spread
    assert lib.ekl_bn_bwd_scratch_doubles(72 * 64 * 64, 128, 3, L.ACT_LRELU) == 3 * 2 * 128 * spread
    assert lib.ekl_bn_bwd_scratch_doubles(72 * 16, 1024, 3, L.ACT_LRELU) == 0            # 384 rows per group: one launch
    assert lib.ekl_bn_bwd_scratch_doubles(10, 128, 3, L.ACT_LRELU) == -1                  # rows do not divide into groups
    mk = lambda mode, B, H, W, Cin, Cout, gb=0: L.EklConv(mode, B, H, W, Cin, Cout, gb, L.IMPL_TC, 0, 0, 0, L.W_KRSC)
    tail = mk(L.S1, 72, 4, 4, 2048, 1024, 24)                                              # split, 3 groups of 384 pixels
    assert lib.ekl_conv_workspace_elems(tail, 0) > 0
    assert lib.ekl_conv_split_bn_fusable(tail, L.ACT_LRELU) == 1 and lib.ekl_conv_split_bn_fusable(tail, L.ACT_GLU) == 0
    assert lib.ekl_conv_split_bn_aux_floats(tail) == 1024 // 32 + 3 * 1024
    big = mk(L.DOWN2, 72, 64, 64, 128, 256, 24)                                            # fills the machine: never split
    assert lib.ekl_conv_split_bn_fusable(big, L.ACT_LRELU) == 0
    up = mk(L.UP2, 24, 4, 4, 1024, 1024, 24)                                               # 4 variants: not a split plan
    assert lib.ekl_conv_split_bn_fusable(up, L.ACT_GLU) == 0


def test_bn_counters_single_vector_add():
    """engine.BnCounters: every num_batches_tracked becomes a view into one int64 tensor, the per-call increments are
    tallied on the host and added once; state_dict keys / values stay those of nn.BatchNorm."""
    from text2img_ekl_b200.engine import BnCounters
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.BatchNorm2d(4))
    net[1].num_batches_tracked.fill_(7)
    bc = BnCounters([net])
    assert net[1]._ekl_counted and int(net[1].num_batches_tracked) == 7
    net[1]._ekl_calls += 3          # what ops.bn_act does per call (groups = 3)
    net[2]._ekl_calls += 1
    bc.flush()
    bc.flush()                      # nothing pending: no-op
    sd = net.state_dict()
    assert int(sd["1.num_batches_tracked"]) == 10 and int(sd["2.num_batches_tracked"]) == 1
    assert sd["1.num_batches_tracked"].dtype == torch.int64 and sd["1.num_batches_tracked"].dim() == 0


def test_generation_image_writers(tmp_path):
    """evaluate()'s output side (cub:733-775): uint8 conversion truncates like the reference's .byte(); the tile sheet
    equals torchvision's save_image(nrow=10, normalize=True) pixels; file names follow the reference's pattern."""
    import numpy as np
    from PIL import Image
    from torchvision.utils import make_grid
    from text2img_ekl_b200 import cub_trainer_splitz_cap_ca as T
    g = torch.Generator().manual_seed(3)
    imgs = torch.rand(13, 3, 16, 16, generator=g) * 2.4 - 1.2          # beyond [-1, 1]: the clamp matters
    ref = imgs.add(1).div(2).mul(255).clamp(0, 255).byte().permute(0, 2, 3, 1)
    assert torch.equal(T.to_uint8_nhwc(imgs), ref)
    want = make_grid(imgs, nrow=10, padding=2, normalize=True).mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)
    assert torch.equal(T.image_grid_uint8(imgs, nrow=10), want)
    keys = ["001.Bird/a_%d" % i for i in range(4)]
    T.condGANTrainer.save_singleimages(None, imgs[:4], keys, str(tmp_path), "test", 2, torch.tensor([5, 6, 7, 8]), 16, 0)
    f = tmp_path / "single_samples" / "001.Bird" / "a_1_16_class6_sid2_nid0.png"
    assert f.exists() and np.array_equal(np.asarray(Image.open(f)), ref[1].numpy())
    T.condGANTrainer.save_superimages(None, [imgs[:4], imgs[4:8], imgs[8:12]], keys, str(tmp_path), "test", 16)
    sheet = np.asarray(Image.open(tmp_path / "super" / "test" / "001.Bird" / "a_3_16.png"))
    assert sheet.shape == (2 + 18, 2 + 3 * 18, 3)
    assert np.array_equal(sheet, T.image_grid_uint8(torch.stack([imgs[3], imgs[7], imgs[11]])).numpy())


def test_evaluate_control_flow_with_stub_generator(tmp_path, monkeypatch):
    """evaluate() (cub:776-911) host logic on the CPU with a stand-in generator: snapshot loading with the 'module.'
    prefix, one noise draw per batch, one generation per sentence embedding (at most 10), zero-based class one-hots,
    single-PNG and tile-sheet outputs, image count.  (The real generator under evaluate() runs in the GPU suite.)"""
    import numpy as np
    from PIL import Image
    from text2img_ekl_b200 import configs, cub_trainer_splitz_cap_ca as T, model
    from text2img_ekl_b200.miscc.config import cfg
    Trainer = configs.setup("splitz_cap_ca", batch=3)
    calls = []

    class StubG(model.COND_G_NET_CATZ_CA):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.w = torch.nn.Parameter(torch.ones(1))

        def forward(self, noise, sen, cls=None, cls_prior=None, eps=None, seed=None):
            calls.append((noise.clone(), sen.clone(), None if cls is None else cls.clone(), self.training))
            return [sen[:, :3].reshape(-1, 3, 1, 1) * self.w], 0, 0, 0, 0, 0, 0

        def image(self, hcodes):
            h = torch.tanh(hcodes[0])
            return [h.expand(-1, 3, 64, 64).contiguous(), (h.expand(-1, 3, 128, 128) * torch.linspace(-1, 1, 128)).contiguous()]

    monkeypatch.setattr(T, "build_G", lambda use_cap=None: (StubG(), None))
    monkeypatch.setattr(model, "to_kernel_layout", lambda net: net)
    torch.save({"module.w": torch.full((1,), 2.0)}, tmp_path / "netG_epoch7.pth")
    g = torch.Generator().manual_seed(1)
    batches = [([torch.zeros(3, 3, 64, 64)], torch.randn(3, 12, cfg.TEXT.DIMENSION, generator=g),
                torch.tensor([1, 5, 200]), ["c%d/k%d" % (b, i) for i in range(3)]) for b in range(2)]
    ev = object.__new__(Trainer)
    ev.device, ev.data_loader, ev.num_batches, ev.batch_size = torch.device("cpu"), batches, 2, 3
    try:
        cfg.TRAIN.NET_G = str(tmp_path / "netG_epoch7.pth")
        cfg.TEST.B_EXAMPLE = False
        n = ev.evaluate("test", save_dir=str(tmp_path / "o1"))
        assert n == 2 * 10 * 3 and len(calls) == 20
        assert all(not c[3] for c in calls)                                        # cfg.TEST.EVAL_MODE -> eval()
        assert all(torch.equal(calls[i][0], calls[0][0]) for i in range(10)) and not torch.equal(calls[10][0], calls[0][0])
        assert torch.equal(calls[3][1], batches[0][1][:, 3]) and torch.equal(calls[13][1], batches[1][1][:, 3])
        assert calls[0][2].shape == (3, cfg.GAN.ENTITY_DIM) and calls[0][2].argmax(1).tolist() == [0, 4, 199]
        files = sorted(p.name for p in (tmp_path / "o1" / "single_samples" / "c0").glob("*.png"))
        assert len(files) == 30 and "k1_128_class4_sid9_nid0.png" in files
        assert np.asarray(Image.open(tmp_path / "o1" / "single_samples" / "c1" / "k2_128_class199_sid0_nid0.png")).shape == (128, 128, 3)
        cfg.TEST.B_EXAMPLE = True
        assert ev.evaluate("test", save_dir=str(tmp_path / "o2")) == 60
        sheet = np.asarray(Image.open(tmp_path / "o2" / "super" / "test" / "c0" / "k0_128.png"))
        assert sheet.shape == (2 + 130, 2 + 10 * 130, 3)
        assert "Testset_evalmode_fixednoise_epoch7_" in ev._eval_save_dir()
    finally:
        cfg.TRAIN.NET_G = ""
        cfg.TEST.B_EXAMPLE = True


def test_flat_adam_state_dict_round_trip_and_no_cpu_step():
    """optim.FlatAdam's host logic: parameters / moments become views into flat buffers without changing shapes, layouts
    or values; its state_dict is interchangeable with torch.optim.Adam's in both directions and loading copies INTO the
    flat buffers (addresses stay valid for a captured graph); step() on CPU tensors fails loudly (no CPU path)."""
    import io
    from text2img_ekl_b200 import _lib
    from text2img_ekl_b200.optim import FlatAdam
    mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Linear(8, 3))
    torch.manual_seed(0)
    net, net2 = mk(), mk()
    net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last)
    before = [p.detach().clone() for p in net.parameters()]
    strides = [p.stride() for p in net.parameters()]
    opt = FlatAdam(net.parameters(), lr=2e-4, betas=(0.5, 0.999))
    for p, b, st in zip(net.parameters(), before, strides):
        assert torch.equal(p, b) and p.stride() == st
        assert opt.flat_p.data_ptr() <= p.data_ptr() < opt.flat_p.data_ptr() + 4 * opt.n
    assert all(o % 4 == 0 for o in opt.offsets)                       # 16-byte aligned fp32 views
    opt.exp_avg.normal_()
    opt.exp_avg_sq.uniform_()
    opt.state_dev[0] = 7
    buf = io.BytesIO()
    torch.save(opt.state_dict(), buf)
    buf.seek(0)
    sd = torch.load(buf)
    adam = torch.optim.Adam(net2.parameters(), lr=1e-3, betas=(0.9, 0.99))
    adam.load_state_dict(sd)                                           # FlatAdam -> torch.optim.Adam
    for p, q in zip(opt.plist, net2.parameters()):
        assert float(adam.state[q]["step"]) == 7.0
        assert torch.equal(adam.state[q]["exp_avg"], opt.state[p]["exp_avg"])
        assert torch.equal(adam.state[q]["exp_avg_sq"], opt.state[p]["exp_avg_sq"])
    opt2 = FlatAdam(mk().parameters(), lr=5e-4)
    ptrs = (opt2.exp_avg.data_ptr(), opt2.exp_avg_sq.data_ptr())
    opt2.load_state_dict(adam.state_dict())                            # torch.optim.Adam -> FlatAdam, in place
    assert (opt2.exp_avg.data_ptr(), opt2.exp_avg_sq.data_ptr()) == ptrs and float(opt2.state_dev[0]) == 7.0
    for p, q in zip(opt.plist, opt2.plist):
        assert torch.equal(opt2.state[q]["exp_avg"], opt.state[p]["exp_avg"])
        assert torch.equal(opt2.state[q]["exp_avg_sq"], opt.state[p]["exp_avg_sq"])
    assert opt2.param_groups[0]["lr"] == 2e-4 and tuple(opt2.param_groups[0]["betas"]) == (0.5, 0.999)
    with pytest.raises(ValueError):
        FlatAdam(torch.nn.Linear(2, 2).parameters()).load_state_dict(adam.state_dict())
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(_lib.EklError):
        opt.step()


def test_color_consistency_term_formula_and_gradient_routing():
    """engine.StepEngine.color_consistency on CPU tensors (the statistics fall back to the torch formula there):
    sum over adjacent stages of coeff * (MSE(mu_i, mu_{i-1}) + 5 * MSE(cov_i, cov_{i-1})) with the LOWER stage detached
    -- stage 0 receives no gradient, the middle stage only through its 'upper' role."""
    import types
    from oracle import ekl_oracle as O
    from text2img_ekl_b200.engine import StepEngine
    g = torch.Generator().manual_seed(2)
    imgs = [torch.tanh(torch.randn(3, 3, 8 << i, 8 << i, generator=g)).requires_grad_(True) for i in range(3)]
    eng = types.SimpleNamespace(color_coeff=2.0, last_color=[])
    total = StepEngine.color_consistency(eng, imgs)
    want = 0.0
    for i in (1, 2):
        (m1, c1), (m2, c2) = O.mean_covariance(imgs[i].detach()), O.mean_covariance(imgs[i - 1].detach())
        want += 2.0 * float(((m1 - m2) ** 2).mean()) + 2.0 * 5 * float(((c1 - c2) ** 2).mean())
    assert abs(float(total.detach()) - want) < 1e-6 * max(1.0, abs(want)) and len(eng.last_color) == 2
    total.backward()
    assert imgs[0].grad is None and imgs[1].grad is not None and imgs[2].grad is not None
    # the middle stage's gradient is that of its pair with stage 0 only (its role under stage 2 is detached)
    mid = imgs[1].detach().clone().requires_grad_(True)
    m1, c1 = O.mean_covariance(mid)
    m2, c2 = O.mean_covariance(imgs[0].detach())
    (2.0 * ((m1 - m2) ** 2).mean() + 10.0 * ((c1 - c2) ** 2).mean()).backward()
    assert torch.allclose(imgs[1].grad, mid.grad, rtol=1e-4, atol=1e-9)


def test_public_loss_helpers_match_oracle_on_cpu():
    """The reference-named module functions the trainers export (KL_loss, ce_loss, compute_mean_covariance, onehot:
    cub:33-65, 322-331) against the oracle's restatements, which tests/test_oracle_golden.py pins to the real reference."""
    from oracle import ekl_oracle as O
    from text2img_ekl_b200 import cub_trainer_splitz_cap_ca as T
    g = torch.Generator().manual_seed(8)
    mu, lv = torch.randn(6, 128, generator=g), 0.3 * torch.randn(6, 128, generator=g)
    assert torch.allclose(T.KL_loss(mu, lv), O.kl_loss(mu, lv), rtol=1e-6)
    logq = torch.log_softmax(torch.randn(6, 201, generator=g), 1)
    p = torch.softmax(torch.randn(6, 201, generator=g), 1)
    assert torch.allclose(T.ce_loss(logq, p), O.ce_loss(logq, p), rtol=1e-6)
    img = torch.rand(3, 3, 16, 16, generator=g) * 2 - 1
    (m, c), (mo, co) = T.compute_mean_covariance(img), O.mean_covariance(img)
    assert m.shape == (3, 3, 1, 1) and c.shape == (3, 3, 3) and torch.equal(m, mo) and torch.equal(c, co)
    cls = torch.randint(0, 200, (9,), generator=g)
    assert torch.equal(T.onehot(cls, 201), O.onehot(cls, 201))                   # bit-exact class targets


def test_product_state_dicts_equal_the_reference_golden_lists():
    """Drop-in boundary: the package's own modules, built from each BASELINE config, expose exactly the
    reference's state_dict keys, shapes AND registration order -- compared with the lists tests/golden/*.npz recorded
    from the real reference (`shapes/G`, `shapes/D<i>`, `gnet/shapes`, `dnet<res>/shapes`), so reference snapshots load."""
    import glob
    import os
    import numpy as np
    from text2img_ekl_b200 import configs, cub_trainer_splitz_cap_ca as T, model
    from text2img_ekl_b200.miscc.config import cfg
    fmt = lambda net: ["%s|%s" % (k, ",".join(map(str, v.shape))) for k, v in net.state_dict().items()]
    seen = 0
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "step_*.npz"))):
        name, b, w = os.path.basename(path)[5:-4].rsplit("_", 2)
        gold = np.load(path)
        Trainer = configs.setup(name, batch=int(b[1:]), width=int(w[1:]))
        if Trainer.KIND == "catz_ca":                   # cub flavour: config 4 and its conditioning variants
            netG = T.build_G(g_class=Trainer.G_CLASS)[0]
        else:
            cond_dim = cfg.TEXT.DIMENSION + (cfg.GAN.ENTITY_DIM + 1 if Trainer.COND == "txt+cls" else 0)
            netG = model.COND_G_NET(cond_dim, model.get_shareGs(cfg.GAN.GF_DIM), use_cap=cfg.TRAIN.G_CAPSULE)
        assert fmt(netG) == list(gold["shapes/G"]), path
        for i, d in enumerate(T.build_Ds()):
            assert fmt(d) == list(gold["shapes/D%d" % i]), (path, i)
        seen += 1
    assert seen >= 12             # 5 BASELINE configs + 2 full-width cases + 5 conditioning variants
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "modules.npz"))
    configs.setup("3stages", batch=4, width=8)
    assert fmt(model.G_NET(model.get_shareGs(cfg.GAN.GF_DIM))) == list(gold["gnet/shapes"])
    for res, D in ((64, model.D_NET64), (128, model.D_NET128), (256, model.D_NET256)):
        assert fmt(D()) == list(gold["dnet%d/shapes" % res])


def test_reference_yml_plus_overrides_equals_the_shipped_resolved_yml():
    import os
    """configs.setup builds each config from the REFERENCE's own cfg/*.yml + the documented override dict when the
    reference tree is present, and from the package's resolved copy otherwise (the GPU box): both routes must give the
    identical cfg, for the five BASELINE configs and every conditioning variant."""
    import copy
    from text2img_ekl_b200 import configs
    from text2img_ekl_b200.miscc.config import cfg
    if not os.path.isdir(os.path.join(configs.REF_ROOT, "cfg")):
        pytest.skip("reference tree not present")
    for name in list(configs.RESOLVED) + list(configs.VARIANTS):
        assert configs.yml_path(name, True)[0].startswith(configs.REF_ROOT)
        configs.setup(name, prefer_reference=True)
        a = copy.deepcopy(cfg)
        configs.setup(name, prefer_reference=False)
        assert a == cfg, name
    configs.setup("splitz_scale4_sum")
    assert cfg.TREE.SCALE == 4 and cfg.TRAIN.CAT_Z == "sum"
    assert configs.setup("catz_exchange").G_CLASS == "COND_G_NET_CATZ" and cfg.TRAIN.EXCHANGE is True


def test_async_loss_log_consumes_in_order_and_never_blocks(tmp_path):
    """miscc.losslog.AsyncLossLog (the sync-free replacement of cub:457-460's per-100-iteration `.item()` scalars):
    samples are taken only on due iterations, land in the JSONL file with their iteration number, a full ring drops
    (and counts) instead of waiting, and a tensorboardX-style writer receives add_scalar calls."""
    import json
    from text2img_ekl_b200.miscc.losslog import AsyncLossLog

    class W:
        def __init__(self):
            self.calls = []

        def add_scalar(self, k, v, step):
            self.calls.append((k, round(v, 4), step))
    w = W()
    path = str(tmp_path / "log" / "scalars.jsonl")
    log = AsyncLossLog(path, every=100, slots=2, writer=w)
    assert log.due(0) and log.due(300) and not log.due(150)
    for count in (0, 100, 200):
        assert log.push(count, ["D_loss0", "D_loss1", "G_loss"], torch.tensor([1.0, 2.0, 3.0]) + count)
    assert log.flush() == 1 and log.dropped == 0          # the earlier samples were consumed by the following push's poll
    rows = [json.loads(ln) for ln in open(path)]
    assert [r["count"] for r in rows] == [0, 100, 200] and rows[2]["G_loss"] == 203.0
    assert ("D_loss1", 102.0, 100) in w.calls and len(w.calls) == 9
    # a ring whose slots are all in flight drops instead of blocking
    for s in log.slots:
        s["busy"], s["tag"] = True, (7, ["x"], True)
        s["ev"] = type("E", (), {"query": lambda self: False, "synchronize": lambda self: None})()
        s["buf"] = torch.zeros(4)
    assert log.push(300, ["x"], torch.tensor([1.0])) is False and log.dropped == 1
    assert log.flush() == 2


def test_discriminator_trunk_marks_fire_deepest_first():
    """model._DBase._trunk drops an ops.grad_mark behind its blocks (no-op unless a gradient reducer is registered
    for the network, parallel.GradReducer): with stand-in blocks on the CPU the block order of all three trunk depths is
    unchanged, and registered marks fire in backward order, deepest block first."""
    from text2img_ekl_b200 import model, ops

    class Blk(torch.nn.Module):
        def __init__(self, tag, log):
            super().__init__()
            self.tag, self.log, self.w = tag, log, torch.nn.Parameter(torch.ones(1))

        def forward(self, x, groups):
            self.log.append(self.tag)
            return x * self.w + groups

    for names in (["img_code_s16"], ["img_code_s16", "img_code_s32", "img_code_s32_1"],
                  ["img_code_s16", "img_code_s32", "img_code_s64", "img_code_s64_1", "img_code_s64_2"]):
        order, fired = [], []
        d = model._DBase()
        for n in names:
            setattr(d, n, Blk(n, order))
        x = torch.ones(2, 3, 4, 4, requires_grad=True)
        y = d._trunk(x, 1)                                   # no registration: plain pass-through
        assert order == names and y.shape == (2, 4, 4, 3)
        ops.GRAD_MARKS[id(d)] = lambda after: fired.append(after.tag)
        try:
            d._trunk(x, 1).sum().backward()
        finally:
            ops.GRAD_MARKS.clear()
        marked = [n for n in names if n not in ("img_code_s32_1", "img_code_s64_2")]
        assert fired == marked[::-1], (names, fired)


def test_step_engine_update_paths_run_on_cpu_stand_ins(monkeypatch):
    """Host-side control flow of engine.StepEngine that every step takes -- construction (flat gradient buffers, BN
    counters, no gradient reducers in a single process), the discriminator update (_d_update -> optimiser) and the
    generator update (_g_step: loss -> backward -> join -> optimiser -> counters) -- exercised with tiny CPU stand-in
    networks and torch.optim.Adam, so that a Python-level slip in these paths shows up before a GPU run."""
    from text2img_ekl_b200 import configs, ops
    from text2img_ekl_b200.engine import StepEngine
    configs.setup("3stages", batch=4)
    torch.manual_seed(0)
    netG = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.BatchNorm1d(4))
    netsD = [torch.nn.Linear(4, 1), torch.nn.Linear(4, 1)]
    optG = torch.optim.Adam(netG.parameters(), lr=1e-2)
    optsD = [torch.optim.Adam(d.parameters(), lr=1e-2) for d in netsD]
    eng = StepEngine(netG, netsD, optG, optsD, "cond")
    assert eng.redD == [None, None] and eng.redG is None and eng.parallel_d and not ops.GRAD_MARKS
    x = torch.randn(6, 4)
    # discriminator update: gradients are views of the flat buffer, _d_update applies them
    w0 = netsD[1].weight.detach().clone()
    eng.gradsD[1].zero()
    netsD[1](x).square().mean().backward()
    assert netsD[1].weight.grad.data_ptr() == eng.gradsD[1].flat.data_ptr() and float(eng.gradsD[1].flat.abs().sum()) > 0
    eng._d_update(1)
    assert not torch.equal(netsD[1].weight.detach(), w0)
    # generator update with a stand-in loss
    g0 = netG[0].weight.detach().clone()
    monkeypatch.setattr(eng, "g_loss", lambda real_cp, per_d=None: (netsD[0](netG(x)).square().mean(),) + (torch.zeros(()),) * 3)
    d0 = netsD[0].weight.detach().clone()
    res = eng.g_step(None)
    assert len(res) == 4 and not torch.equal(netG[0].weight.detach(), g0)
    assert torch.equal(netsD[0].weight.detach(), d0) and netsD[0].weight.requires_grad      # D frozen during, restored after
    assert int(netG[1].num_batches_tracked) == 1


def test_abi_rejects_bad_arguments_with_status_and_message():
    """SURVEY 8b error contract of the C ABI: int status (negative = invalid argument), never an exception or a launch,
    and a thread-local message from ekl_last_error().  Argument checks run before any CUDA call, so they are testable
    without a device."""
    lib = L.lib()
    p = C.c_void_p(256)                      # a non-null dummy address: must never be dereferenced on these paths

    def rejected(rc, needle):
        msg = lib.ekl_last_error().decode()
        assert rc < 0 and needle in msg, (rc, msg)

    rejected(lib.ekl_color_stats_fwd(None, 2, 64, None, None, None, None), "null pointer")
    rejected(lib.ekl_color_stats_fwd(p, 2, 63, p, p, p, None), "multiple of 4")
    rejected(lib.ekl_color_stats_bwd(p, p, None, None, 0, 64, p, None), "multiple of 4")
    rejected(lib.ekl_head_tanh_fwd(p, 2, 64, 7, p, None), "C % 8")
    rejected(lib.ekl_img_s2d(p, None, None, 4, 2, 8, 8, p, None), "bad arguments")
    rejected(lib.ekl_img_s2d(p, None, None, 1, 2, 7, 8, p, None), "bad arguments")
    rejected(lib.ekl_lrelu_bwd(p, p, p, 12, None), "n % 8")
    rejected(lib.ekl_cat_code(p, 12, p, 64, 2, 16, p, None), "channels % 8")
    bad_mode = L.EklConv(7, 2, 8, 8, 16, 16, 0, L.IMPL_TC, 0, 0, 0)
    rejected(lib.ekl_conv_fwd(bad_mode, p, p, p, None, None), "bad conv mode")
    odd = L.EklConv(L.DOWN2, 2, 7, 8, 16, 16, 0, L.IMPL_TC, 0, 0, 0)
    rejected(lib.ekl_conv_fwd(odd, p, p, p, None, None), "even H, W")
    ok = L.EklConv(L.S1, 2, 8, 8, 16, 16, 0, L.IMPL_TC, 0, 0, 0)
    rejected(lib.ekl_conv_fwd(ok, None, p, p, None, None), "null pointer")
    rejected(lib.ekl_conv_bwd_data(ok, p, None, p, None), "null pointer")
