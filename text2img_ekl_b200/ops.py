"""torch.autograd glue over the C ABI (include/ekl_b200.h).

Internal activation layout is NHWC bf16 ([B,H,W,C] contiguous).  Parameters stay fp32 in the reference's
shapes; conv filters are stored channels_last so that their memory is the [Cout][KH][KW][Cin] master layout
the kernels read and the weight-gradient kernel accumulates into (`.grad` is written in place, no autograd
accumulation pass).  No op here has a non-CUDA fallback.
"""
import os
import weakref

import torch

from . import _lib as L

S1, UP2, DOWN2 = L.S1, L.UP2, L.DOWN2
ACT_NONE, ACT_GLU, ACT_LRELU, ACT_RELU, ACT_TANH = L.ACT_NONE, L.ACT_GLU, L.ACT_LRELU, L.ACT_RELU, L.ACT_TANH
BN_EPS, BN_MOM = 1e-5, 0.1          # nn.BatchNorm defaults (model.py:91)

# counts launches of this library's kernels (bench.py reports it as gpu_launches)
LAUNCHES = [0]


def _count(n=1):
    LAUNCHES[0] += n


# bench.py's kernel_roofline(): when a list, every library call appends (accounting name, reference-count flops,
# executed flops, algorithmic bytes).  Device times come from CUPTI records of the graph replay, not from here.
ACCOUNT = None
# tools/layer_bench.py: when a list, every conv / BatchNorm call appends its shape key
SHAPE_LOG = None


# parallel.TailAllreduce (opt-in, N > 1): id(network) -> callback(module).  A network's forward drops a mark behind a
# block; when autograd has produced the gradient of the marked activation, every parameter registered AFTER that block
# has its final gradient, and the callback may start reducing that tail of the flat gradient buffer.
GRAD_MARKS = {}


def grad_mark(x, owner, after):
    cb = GRAD_MARKS.get(id(owner)) if GRAD_MARKS else None
    if cb is None or not x.requires_grad:
        return x

    def fire(grad, after=after, cb=cb):
        cb(after)                  # returns None: the gradient passes through unchanged
    x.register_hook(fire)
    return x


class StatsArena:
    """Zero-initialised fp64 scratch for the BatchNorm statistics of one training step.

    The conv epilogues ACCUMULATE per-channel sums into a [groups][2][C] double buffer (fp64 red.global.add), the
    normalise pass reads them directly; the backward pass does the same with its two reductions.  Every such buffer must
    be zero before its producer runs.  Instead of one fill kernel per layer and direction, the step engine calls reset()
    once per step (one memset for all layers of all networks) and every layer takes the next slice.  Slices are handed
    out in host program order, which is the same every step, so a captured CUDA graph sees fixed addresses.  Without a
    reset in the current step (modules used stand-alone) take() falls back to a fresh torch.zeros."""

    def __init__(self):
        self.buf, self.off, self.want, self.live = None, 0, 0, False

    def reset(self, device):
        need = max(self.want, 1 << 16)
        if self.buf is None or self.buf.device != device or self.buf.numel() < need:
            self.buf = torch.zeros(need + need // 4, dtype=torch.float64, device=device)
        else:
            self.buf.zero_()
        self.off, self.want, self.live = 0, 0, True
        _count()

    def take(self, n, device):
        n = (n + 1) // 2 * 2                       # 16-byte aligned slices
        self.want += n
        if self.live and self.buf.device == device and self.off + n <= self.buf.numel():
            v = self.buf[self.off:self.off + n]
            self.off += n
            return v
        self.live = False                          # exhausted: this step finishes on fresh buffers, the next reset() grows
        return torch.zeros(n, dtype=torch.float64, device=device)


def zeros_f32(n, device):
    """n zero-initialised floats from the step's scratch arena (see StatsArena)."""
    return ARENA.take((n + 1) // 2, device).view(torch.float32)[:n]


ARENA = StatsArena()


def _log(*key):
    if SHAPE_LOG is not None:
        SHAPE_LOG.append(key)


def _acct(name, flops=(0.0, 0.0), nbytes=0.0):
    if ACCOUNT is not None:
        ACCOUNT.append((name, float(flops[0]), float(flops[1]), float(nbytes)))


def _conv_flops(spec, B, H, W):
    """(reference-count flops, executed flops) of one conv pass.  Reference count = the dense conv the reference runs
    (SURVEY 8d): 2*B*Ho*Wo*Cout*Cin*KH*KW on the upsampled grid for up-convs, with the layer's real channel / tap
    counts (spec.ref = (cin, cout, taps) where they differ from what the kernel contracts: folded code channels, padded
    image-head outputs, the space-to-depth stem).  Executed = what the kernel multiplies (sub-pixel up-convs: 16 taps per
    2x2 output block instead of 36)."""
    Ho, Wo = spec.out_hw(H, W)
    k = 16 if spec.mode == DOWN2 else 9
    rcin, rcout, rk = spec.ref or (spec.cin, spec.cout, k)
    ref = 2.0 * B * Ho * Wo * rcout * rcin * rk
    exe = 2.0 * B * Ho * Wo * spec.cout * spec.cin * k / (2.25 if spec.mode == UP2 else 1.0)
    return ref, exe


def _route(spec, c, dgrad):
    """accounting name of the kernel family a conv call runs on (include/ekl_b200.h: ekl_conv_route)."""
    if ACCOUNT is None:
        return ""
    if spec.impl != L.IMPL_TC:
        return "conv_simt"
    return "conv_tc:" + ("generic", "rw", "split")[L.lib().ekl_conv_route(c, dgrad)]


def _grad_buffer(p):
    """fp32 gradient tensor of parameter p with p's memory layout, created zeroed on first use."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
    return p.grad


# id(parameter) -> ConvSpecs packed from it.  Fused optimisers update parameters WITHOUT bumping their autograd
# version counter, so optimiser steps must invalidate the packed operands explicitly (mark_dirty; installed as an
# optimizer post-step hook by define_optimizers).
_SPECS_OF = {}


def mark_dirty(params):
    for p in params:
        for spec in _SPECS_OF.get(id(p), ()):
            spec._dirty = True


class ConvSpec:
    """One convolution layer's kernel-side state: packed bf16 filter operands, refreshed when the fp32 master
    parameter changes (optimizer step / load_state_dict)."""

    def __init__(self, mode, cin, cout, impl=L.IMPL_TC, x_fmt=0, y_fmt=0, act=ACT_NONE, w_cin_total=0, w_cin_off=0,
                 w_cout_valid=0):
        self.mode, self.cin, self.cout, self.impl = mode, cin, cout, impl
        self.x_fmt, self.y_fmt, self.act = x_fmt, y_fmt, act
        # window of the master filter (include/ekl_b200.h: ekl_conv.w_cin_total / w_cin_off / w_cout_valid)
        self.w_cin_total, self.w_cin_off, self.w_cout_valid = w_cin_total, w_cin_off, w_cout_valid
        self.ref = None                  # (cin, cout, taps) of the reference layer when they differ (flop accounting only)
        self._ver, self._dirty = None, False
        self._weight_ref = None                         # prepack(): the leaf parameter packed from
        self._ready_f = self._ready_d = None            # events of a side-stream refresh of the forward / data-gradient operand
        self._stale = 0                                 # operands to repack: bit 0 forward, bit 1 data-gradient
        self.w_layout = L.W_KCRS
        self.w_fwd = self.w_dgrad = self._fwd_buf = None
        self._ws = {}

    def conv(self, B, H, W, group_b=0):
        return L.EklConv(self.mode, B, H, W, self.cin, self.cout, group_b, self.impl, self.x_fmt, self.y_fmt, self.act,
                         self.w_layout, self.w_cin_total, self.w_cin_off, self.w_cout_valid)

    def workspace(self, c, dgrad, device):
        """fp32 split-K workspace of this layer for descriptor c (None when the plan does not split).  Allocated once
        per (shape, direction), zero-filled; the kernels leave it zero."""
        key = (c.B, c.H, c.W, c.group_b, dgrad)
        if key not in self._ws:
            n = L.lib().ekl_conv_workspace_elems(c, dgrad) if self.impl == L.IMPL_TC else 0
            self._ws[key] = torch.zeros(n, device=device, dtype=torch.float32) if n > 0 else None
        return self._ws[key]

    def out_hw(self, H, W):
        return (2 * H, 2 * W) if self.mode == UP2 else ((H // 2, W // 2) if self.mode == DOWN2 else (H, W))

    def bind(self, weight):
        """Detect the master filter's memory layout (channels_last storage is the fast, coalesced one)."""
        if weight.is_contiguous(memory_format=torch.channels_last):
            self.w_layout = L.W_KRSC
        elif weight.is_contiguous():
            self.w_layout = L.W_KCRS
        else:
            raise L.EklError("conv filter must be contiguous or channels_last")

    def dgrad_from_fwd(self):
        """The data-gradient kernel can read the forward-packed filter as an MN-major operand (no transposed pack):
        stride-1 / stride-2 convs with 64-multiple channel counts that never take the resident-filter path (its
        data-gradient contraction is Cout <= 128; conv_rw.cu: ekl_rw_supported).  Shape-independent on purpose: the
        operand a layer keeps packed must not depend on the batch it happens to see."""
        return (self.impl == L.IMPL_TC and self.mode != UP2 and self.w_layout == L.W_KRSC and self.cin % 64 == 0
                and self.cout % 64 == 0 and self.cout > 128 and self.x_fmt == 0 and self.y_fmt == 0)

    def _shadow(self, weight):
        """bf16 shadow slice maintained by optim.FlatAdam, usable as the packed forward operand when that operand is a
        plain cast of the master in memory order: stride-1 / stride-2 plans (one variant, one source tap per packed
        tap) of a channels_last leaf parameter."""
        sh = getattr(weight, "_ekl_shadow", None)
        if sh is None or not weight.is_leaf or self.mode == UP2 or self.w_layout != L.W_KRSC or self.impl != L.IMPL_TC:
            return None
        if self.w_cin_total or self.w_cout_valid:          # a window of the master is not a plain cast of it
            return None
        return sh

    def packed(self, weight, need=3):
        """(forward operand, data-gradient operand) of `weight`, refreshed if the master changed.  need: bit 0 = the
        caller reads the forward operand, bit 1 = the data-gradient operand; only what is needed is (re)packed / waited
        for, so a forward pass never waits for the transposed packs of a side-stream refresh (prepack)."""
        if (need & 1) and self._ready_f is not None:        # refreshed on a side stream (prepack): order after it
            torch.cuda.current_stream().wait_event(self._ready_f)
            self._ready_f = None
        if (need & 2) and self._ready_d is not None:
            torch.cuda.current_stream().wait_event(self._ready_d)
            self._ready_d = None
        key = (weight._version, weight.data_ptr())
        external = key != self._ver or not weight.is_leaf     # derived (e.g. zero-padded) filters are rebuilt every step
        lib = L.lib()
        if external or self._dirty:
            self.bind(weight)
            if weight.is_leaf:
                lst = _SPECS_OF.setdefault(id(weight), [])
                if self not in lst:
                    lst.append(self)
                    self._weight_ref = weakref.ref(weight)
            c = self.conv(1, 4, 4)
            shadow = self._shadow(weight)
            self._stale = 3
            if shadow is not None:
                if external:                       # parameter changed outside the fused optimiser (init / load_state_dict)
                    shadow.copy_(weight.detach().permute(0, 2, 3, 1).reshape(-1))
                self.w_fwd = shadow
                self._stale &= ~1                  # the optimiser's bf16 shadow IS the forward operand
            else:
                if self._fwd_buf is None:
                    self._fwd_buf = torch.empty(lib.ekl_conv_packed_elems(c, 0), device=weight.device, dtype=torch.bfloat16)
                self.w_fwd = self._fwd_buf
            if self.dgrad_from_fwd():
                self._stale &= ~2                  # the data-gradient kernel reads the forward operand
            elif self.w_dgrad is None:
                self.w_dgrad = torch.empty(lib.ekl_conv_packed_elems(c, 1), device=weight.device, dtype=torch.bfloat16)
            self._ver, self._dirty = key, False
        do = self._stale & need
        if do:
            c = self.conv(1, 4, 4)
            w_fwd_arg = self.w_fwd if do & 1 else None
            w_dgrad_arg = self.w_dgrad if do & 2 else None
            nb = weight.numel() * 4 + ((w_fwd_arg.numel() if w_fwd_arg is not None else 0) +
                                       (w_dgrad_arg.numel() if w_dgrad_arg is not None else 0)) * 2
            _acct("pack_weights", nbytes=nb)
            L.check(lib.ekl_conv_pack(c, L.ptr(weight), L.ptr(w_fwd_arg), L.ptr(w_dgrad_arg), L.stream()))
            _count((w_fwd_arg is not None) + (w_dgrad_arg is not None))
            self._stale &= ~do
        return self.w_fwd, self.w_dgrad


_PACK_STREAMS = {}


def prepack(params, key=0):
    """Refresh, on a side stream, the packed operands of every convolution whose master filter changed (optimiser step)
    instead of lazily in front of each convolution's first use: the ~50 small pack kernels of a network leave the
    critical path and overlap whatever the issuing stream does next.  Consumers order themselves after the side stream
    at their first use (ConvSpec.packed).  Returns the side stream if anything was issued (the caller joins it before a
    capture ends), else None."""
    todo = []
    for p in params:
        for spec in _SPECS_OF.get(id(p), ()):
            w = spec._weight_ref() if spec._weight_ref is not None else None
            if spec._dirty and w is not None and spec._ready_f is None and spec._ready_d is None:
                todo.append((spec, w))
    if not todo:
        return None
    main = torch.cuda.current_stream()
    side = _PACK_STREAMS.get((main.device, key))
    if side is None:
        side = _PACK_STREAMS[(main.device, key)] = torch.cuda.Stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        # forward operands first (few: up-convs and filter windows; the rest read the optimiser's shadow), with their own
        # event: the network's next FORWARD pass waits for these only, its backward for the transposed operands
        for spec, w in todo:
            spec.packed(w, 1)
        ev_f = torch.cuda.Event()
        ev_f.record(side)
        for spec, w in todo:
            spec.packed(w, 2)
        ev_d = torch.cuda.Event()
        ev_d.record(side)
    for spec, _ in todo:
        spec._ready_f, spec._ready_d = ev_f, ev_d
    return side


def _run_dgrad(spec, c, weight, dy, dx):
    """dx = conv^T(dy) through whichever operand the layer keeps: the forward-packed filter read as an MN-major operand
    (no transposed pack), else the packed data-gradient operand; split-K workspace when the plan has one."""
    lib = L.lib()
    ws = spec.workspace(c, 1, dx.device)
    if spec.dgrad_from_fwd():
        w_fwd, _ = spec.packed(weight, 1)
        L.check(lib.ekl_conv_bwd_data_fw(c, L.ptr(dy), L.ptr(w_fwd), L.ptr(dx), L.ptr(ws), L.stream()))
    else:
        _, w_dgrad = spec.packed(weight, 2)
        if ws is not None:
            L.check(lib.ekl_conv_bwd_data_ws(c, L.ptr(dy), L.ptr(w_dgrad), L.ptr(dx), L.ptr(ws), L.stream()))
        else:
            L.check(lib.ekl_conv_bwd_data(c, L.ptr(dy), L.ptr(w_dgrad), L.ptr(dx), L.stream()))
    _count(2 if ws is not None else 1)


class _Conv(torch.autograd.Function):
    """y = conv(x) (+ per-tile BatchNorm partial statistics).  Backward: tcgen05 dgrad + wgrad."""

    @staticmethod
    def forward(ctx, x, weight, spec, group_b, want_stats, skip_wgrad, fuse_bn=None):
        """fuse_bn = (bn module, groups, act) or None: when the plan splits along K and the layer is followed by train-mode
        BatchNorm + activation, the finishing pass and the normalise pass are ONE kernel (ekl_conv_fwd_split_bn_act); the
        extra outputs (out, mean, rstd) are handed to _BnAct, which then launches nothing in its forward."""
        ctx.set_materialize_grads(False)      # no zero-filled "gradient" of the statistics output
        lib = L.lib()
        if spec.x_fmt == L.FMT_NCHW_F32:
            B, _, H, W = x.shape
            assert x.dtype == torch.float32 and x.is_contiguous()
        else:
            B, H, W, _ = x.shape
            assert x.dtype == torch.bfloat16 and x.is_contiguous()
        Ho, Wo = spec.out_hw(H, W)
        w_fwd, _ = spec.packed(weight, 1)
        c = spec.conv(B, H, W, group_b)
        if spec.y_fmt == L.FMT_NCHW_F32:
            y = torch.empty(B, spec.cout, Ho, Wo, device=x.device, dtype=torch.float32)
        else:
            y = torch.empty(B, Ho, Wo, spec.cout, device=x.device, dtype=torch.bfloat16)
        stats = None
        ws = spec.workspace(c, 0, x.device) if spec.act == ACT_NONE else None
        if want_stats and spec.impl == L.IMPL_TC:
            groups = B // group_b if (group_b > 0 and B % group_b == 0) else 1
            stats = ARENA.take(groups * 2 * spec.cout, x.device)          # [groups][2][Cout] fp64 sums, zero on entry
        fam = "conv_tc" if spec.impl == L.IMPL_TC else "conv_simt"
        _log("fwd", fam, spec.mode, B, H, W, spec.cin, spec.cout, group_b)
        _acct(_route(spec, c, 0), _conv_flops(spec, B, H, W), x.numel() * x.element_size() + y.numel() * y.element_size())
        pre = (None, None, None)
        if ws is not None and fuse_bn is not None and lib.ekl_conv_split_bn_fusable(c, fuse_bn[2]):
            bn, groups, act = fuse_bn
            C = spec.cout
            mean = torch.empty(groups, C, device=x.device, dtype=torch.float32)
            rstd = torch.empty(groups, C, device=x.device, dtype=torch.float32)
            out = torch.empty_like(y)
            aux = zeros_f32(int(lib.ekl_conv_split_bn_aux_floats(c)), x.device)
            _acct("bn", nbytes=y.numel() * 4)
            L.check(lib.ekl_conv_fwd_split_bn_act(c, L.ptr(x), L.ptr(w_fwd), L.ptr(ws), L.ptr(y), BN_EPS, BN_MOM, L.ptr(mean),
                                                  L.ptr(rstd), L.ptr(bn.running_mean), L.ptr(bn.running_var), L.ptr(bn.weight),
                                                  L.ptr(bn.bias), act, L.ptr(out), L.ptr(aux), L.stream()))
            _count(2)
            pre, stats = (out, mean, rstd), None
        elif ws is not None:
            L.check(lib.ekl_conv_fwd_ws(c, L.ptr(x), L.ptr(w_fwd), L.ptr(y), L.ptr(stats), L.ptr(ws), L.stream()))
            _count(2)
        else:
            L.check(lib.ekl_conv_fwd(c, L.ptr(x), L.ptr(w_fwd), L.ptr(y), L.ptr(stats), L.stream()))
            _count(1)
        ctx.dims = (B, H, W)
        if spec.act != ACT_NONE and spec.impl == L.IMPL_TC:
            ctx.save_for_backward(x, weight, y)        # fused epilogue activation: its derivative needs the output
        else:
            ctx.save_for_backward(x, weight)
        ctx.spec, ctx.c, ctx.skip_wgrad, ctx.w_leaf = spec, c, skip_wgrad, weight.is_leaf
        ctx.mark_non_differentiable(*[t for t in (stats,) + pre if t is not None])
        return (y, stats) + pre

    @staticmethod
    def backward(ctx, dy, _dstats, _dout=None, _dmean=None, _drstd=None):
        if dy is None:
            return None, None, None, None, None, None, None
        lib = L.lib()
        x, weight = ctx.saved_tensors[:2]
        spec, c = ctx.spec, ctx.c
        fam = "conv_tc" if spec.impl == L.IMPL_TC else "conv_simt"
        dy = dy.contiguous()
        if len(ctx.saved_tensors) == 3:
            y = ctx.saved_tensors[2]
            if spec.act == ACT_LRELU:                   # dpre = dy * LeakyReLU'(0.2), evaluated from the output
                dpre = torch.empty_like(y)
                L.check(lib.ekl_lrelu_bwd(L.ptr(y), L.ptr(dy), L.ptr(dpre), y.numel(), L.stream()))
                _count()
                dy = dpre
            elif spec.act == ACT_TANH:
                dy = dy * (1.0 - y.float() ** 2).to(dy.dtype)
        dx = None
        want_w = ctx.needs_input_grad[1] and not ctx.skip_wgrad
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _log("dgrad", fam, spec.mode, *ctx.dims, spec.cin, spec.cout, 0)
            _acct(_route(spec, c, 1), _conv_flops(spec, *ctx.dims), dx.numel() * dx.element_size() + dy.numel() * dy.element_size())
            _run_dgrad(spec, c, weight, dy, dx)
        dw = None
        if want_w:
            if ctx.w_leaf:
                buf = _grad_buffer(weight)          # parameters: accumulate in place, autograd sees no gradient
            else:
                buf = dw = torch.zeros_like(weight, memory_format=torch.preserve_format)
            _log("wgrad", fam, spec.mode, *ctx.dims, spec.cin, spec.cout, 0)
            _acct("conv_wgrad" if spec.impl == L.IMPL_TC else "conv_simt", _conv_flops(spec, *ctx.dims),
                  x.numel() * x.element_size() + dy.numel() * dy.element_size())
            L.check(lib.ekl_conv_bwd_weight(c, L.ptr(x), L.ptr(dy), L.ptr(buf), L.stream()))
            _count()
        return dx, dw, None, None, None, None, None


def conv(x, weight, spec, group_b=0, want_stats=False, skip_wgrad=False):
    return _Conv.apply(x, weight, spec, group_b, want_stats, skip_wgrad)[:2]


_SPLIT_BN = os.environ.get("EKL_SPLIT_BN", "1") != "0"      # fused finishing + BatchNorm forward kernel after split-K convs


def conv_bn_act(x, weight, spec, bn, groups, act, residual=None):
    """conv -> train-mode BatchNorm -> activation (+ residual).  Layers whose conv plan splits along K run the finishing
    pass and the normalise pass as one kernel (see _Conv.forward); everything else is conv() + bn_act()."""
    B = x.shape[0]
    fuse = (bn, groups, act) if (_SPLIT_BN and bn.training and residual is None and spec.impl == L.IMPL_TC and x.is_cuda) else None
    y, stats, out, mean, rstd = _Conv.apply(x, weight, spec, B // groups, bn.training, False, fuse)
    return bn_act(y, stats, bn, groups, act, residual, pre=(out, mean, rstd) if out is not None else None)


def border_class_sums(dy):
    """[B,H,W,N] bf16 -> fp32 [B,9,N]: sums of dy over the 9 border classes of ekl_conv_fwd_bias9 (class = 3*rc + cc),
    i.e. the gradient of its bias9 argument; one pass over dy (include/ekl_b200.h: ekl_border_sums9)."""
    B, H, W, N = dy.shape
    S = zeros_f32(B * 9 * N, dy.device).view(B, 9, N)
    L.check(L.lib().ekl_border_sums9(L.ptr(dy), B, H, W, N, L.ptr(S), L.stream()))
    _count()
    return S


class _CodeBias9(torch.autograd.Function):
    """bias9[b,q,n] = sum over the taps t inside the map for border class q of sum_c code[b,c] * W[n, c, t]: the tiled
    condition-code channels of a jointConv (model.py:403, 411-414) as a per-sample bias (include/ekl_b200.h:
    ekl_code_bias9_fwd / _bwd).  weight: the FULL master filter [N, ef + ngf, 3, 3] (channels_last storage), whose first
    `ef` input channels are the code channels; its gradient over those channels is accumulated in place."""

    @staticmethod
    def forward(ctx, code, weight):
        B, ef = code.shape
        N, Ctot = weight.shape[0], weight.shape[1]
        assert weight.is_contiguous(memory_format=torch.channels_last) and code.dtype == torch.float32
        code = code.contiguous()
        bias9 = torch.empty(B, 9, N, device=code.device, dtype=torch.float32)
        L.check(L.lib().ekl_code_bias9_fwd(L.ptr(code), L.ptr(weight), B, ef, Ctot, N, L.ptr(bias9), L.stream()))
        _count()
        ctx.save_for_backward(code, weight)
        return bias9

    @staticmethod
    def backward(ctx, S):
        code, weight = ctx.saved_tensors
        B, ef = code.shape
        N, Ctot = weight.shape[0], weight.shape[1]
        S = S.contiguous()
        dcode = zeros_f32(B * ef, code.device).view(B, ef) if ctx.needs_input_grad[0] else None
        dw = None
        buf = None
        if ctx.needs_input_grad[1]:
            if weight.is_leaf:
                buf = _grad_buffer(weight)              # accumulated in place (same channels_last memory as the master)
            else:
                buf = dw = torch.zeros_like(weight, memory_format=torch.preserve_format)
        L.check(L.lib().ekl_code_bias9_bwd(L.ptr(S), L.ptr(code), L.ptr(weight), B, ef, Ctot, N, L.ptr(dcode), L.ptr(buf), L.stream()))
        _count()
        return dcode, dw


def code_bias9(code, weight):
    return _CodeBias9.apply(code, weight)


class _ConvBias9(torch.autograd.Function):
    """y = conv3x3(x, weight) + bias9[b, border class(h, w), :]  (include/ekl_b200.h: ekl_conv_fwd_bias9)."""

    @staticmethod
    def forward(ctx, x, weight, bias9, spec, want_stats):
        ctx.set_materialize_grads(False)
        lib = L.lib()
        B, H, W, _ = x.shape
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and spec.mode == S1 and spec.impl == L.IMPL_TC
        w_fwd, _ = spec.packed(weight, 1)
        c = spec.conv(B, H, W, 0)
        y = torch.empty(B, H, W, spec.cout, device=x.device, dtype=torch.bfloat16)
        stats = ARENA.take(2 * spec.cout, x.device) if want_stats else None
        bias9 = bias9.float().contiguous()
        _log("fwd", "conv_tc", spec.mode, B, H, W, spec.cin, spec.cout, 0)
        _acct(_route(spec, c, 0), _conv_flops(spec, B, H, W), x.numel() * 2 + y.numel() * 2)
        L.check(lib.ekl_conv_fwd_bias9(c, L.ptr(x), L.ptr(w_fwd), L.ptr(bias9), L.ptr(y), L.ptr(stats), L.stream()))
        _count()
        ctx.dims = (B, H, W)
        ctx.save_for_backward(x, weight)
        ctx.spec, ctx.c, ctx.w_leaf = spec, c, weight.is_leaf
        ctx.mark_non_differentiable(*([stats] if stats is not None else []))
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats):
        if dy is None:
            return None, None, None, None, None
        lib = L.lib()
        x, weight = ctx.saved_tensors
        spec, c = ctx.spec, ctx.c
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            _log("dgrad", "conv_tc", spec.mode, *ctx.dims, spec.cin, spec.cout, 0)
            _acct(_route(spec, c, 1), _conv_flops(spec, *ctx.dims), dx.numel() * 2 + dy.numel() * 2)
            _run_dgrad(spec, c, weight, dy, dx)
        if ctx.needs_input_grad[1]:
            buf = _grad_buffer(weight) if ctx.w_leaf else torch.zeros_like(weight, memory_format=torch.preserve_format)
            dw = None if ctx.w_leaf else buf
            _log("wgrad", "conv_tc", spec.mode, *ctx.dims, spec.cin, spec.cout, 0)
            _acct("conv_wgrad", _conv_flops(spec, *ctx.dims), x.numel() * 2 + dy.numel() * 2)
            L.check(lib.ekl_conv_bwd_weight(c, L.ptr(x), L.ptr(dy), L.ptr(buf), L.stream()))
            _count()
        if ctx.needs_input_grad[2]:
            db = border_class_sums(dy)
        return dx, dw, db, None, None


def conv_bias9(x, weight, bias9, spec, want_stats=False):
    return _ConvBias9.apply(x, weight, bias9, spec, want_stats)


class _BnAct(torch.autograd.Function):
    """Train-mode BatchNorm (per-group batch statistics) + GLU / LeakyReLU / ReLU / identity (+ residual)."""

    @staticmethod
    def forward(ctx, y, stats, gamma, beta, running_mean, running_var, groups, act, residual, skip_pgrad, pre=None):
        lib = L.lib()
        C = y.shape[-1]
        M = y.numel() // C
        dev = y.device
        if pre is not None:
            # out / mean / rstd (and the running-statistics update) were produced by the conv's fused finishing kernel
            out, mean, rstd = pre
            _log("bn_fwd", M, C, groups, act, False)
            ctx.save_for_backward(y, mean, rstd, gamma, beta)
            ctx.groups, ctx.act, ctx.has_res, ctx.skip_pgrad = groups, act, False, skip_pgrad
            return out.view_as(out)
        if stats is None:
            stats = ARENA.take(groups * 2 * C, dev)
            _acct("bn", nbytes=M * C * 2)
            L.check(lib.ekl_col_stats(L.ptr(y), M, C, groups, L.ptr(stats), L.stream()))
            _count()
        _log("bn_fwd", M, C, groups, act, residual is not None)
        mean = torch.empty(groups, C, device=dev, dtype=torch.float32)
        rstd = torch.empty(groups, C, device=dev, dtype=torch.float32)
        Co = C // 2 if act == ACT_GLU else C
        out = torch.empty(*y.shape[:-1], Co, device=dev, dtype=torch.bfloat16)
        # algorithmic bytes (minimal-traffic model): y read once, out written once (+ the residual read)
        _acct("bn", nbytes=M * (C + Co + (Co if residual is not None else 0)) * 2)
        L.check(lib.ekl_bn_act_fwd(L.ptr(y), M, C, groups, L.ptr(stats), BN_EPS, BN_MOM, L.ptr(mean), L.ptr(rstd),
                                   L.ptr(running_mean), L.ptr(running_var), L.ptr(gamma), L.ptr(beta), act, L.ptr(residual),
                                   L.ptr(out), L.stream()))
        _count()
        ctx.save_for_backward(y, mean, rstd, gamma, beta)
        ctx.groups, ctx.act, ctx.has_res, ctx.skip_pgrad = groups, act, residual is not None, skip_pgrad
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = L.lib()
        y, mean, rstd, gamma, beta = ctx.saved_tensors
        C = y.shape[-1]
        M = y.numel() // C
        dout = dout.contiguous()
        nsum = lib.ekl_bn_bwd_scratch_doubles(M, C, ctx.groups, ctx.act)
        sums = ARENA.take(nsum, y.device) if nsum > 0 else None    # zero on entry: the two backward reductions land here
        dy = torch.empty_like(y)
        pg = gamma.requires_grad and not ctx.skip_pgrad
        Co = C // 2 if ctx.act == ACT_GLU else C
        _log("bn_bwd", M, C, ctx.groups, ctx.act, ctx.has_res)
        # algorithmic bytes (minimal-traffic model): y and dout read once, dy written once (the kernel makes two passes)
        _acct("bn", nbytes=M * (2 * C + Co) * 2)
        L.check(lib.ekl_bn_act_bwd(L.ptr(y), L.ptr(dout), M, C, ctx.groups, L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
                                   L.ptr(beta), ctx.act, L.ptr(sums),
                                   L.ptr(_grad_buffer(gamma)) if pg else None, L.ptr(_grad_buffer(beta)) if pg else None,
                                   L.ptr(dy), L.stream()))
        _count(2)
        return dy, None, None, None, None, None, None, None, (dout if ctx.has_res else None), None, None


def bn_act(y, stats, bn, groups, act, residual=None, skip_pgrad=False, pre=None):
    """bn: an nn.BatchNorm{1,2}d module holding weight/bias/running stats (state_dict-compatible)."""
    if not bn.training:
        # inference: normalise with the running statistics (no autograd; the evaluate() path)
        lib = L.lib()
        C = y.shape[-1]
        M = y.numel() // C
        mean = bn.running_mean.detach().view(1, C).contiguous()
        rstd = torch.rsqrt(bn.running_var.detach() + bn.eps).view(1, C).contiguous()
        Co = C // 2 if act == ACT_GLU else C
        out = torch.empty(*y.shape[:-1], Co, device=y.device, dtype=torch.bfloat16)
        L.check(lib.ekl_bn_act_fwd(L.ptr(y), M, C, 1, None, BN_EPS, BN_MOM, L.ptr(mean), L.ptr(rstd), None, None,
                                   L.ptr(bn.weight), L.ptr(bn.bias), act, L.ptr(residual), L.ptr(out), L.stream()))
        _count()
        return out
    if getattr(bn, "_ekl_counted", False):
        bn._ekl_calls += groups          # flushed once per step for the whole network (engine.BnCounters)
    elif bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(groups)
    return _BnAct.apply(y, stats, bn.weight, bn.bias, bn.running_mean, bn.running_var, groups, act, residual, skip_pgrad, pre)


class _LReluFromOut(torch.autograd.Function):
    """identity forward on an already-activated tensor; backward = LeakyReLU'(0.2) evaluated from the output."""

    @staticmethod
    def forward(ctx, out):
        ctx.save_for_backward(out)
        return out.view_as(out)

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        dout = dout.contiguous()
        dx = torch.empty_like(out)
        L.check(L.lib().ekl_lrelu_bwd(L.ptr(out), L.ptr(dout), L.ptr(dx), out.numel(), L.stream()))
        _count()
        return dx


def lrelu_from_out(out):
    return _LReluFromOut.apply(out)


class _CatCode(torch.autograd.Function):
    """cat(tile(c_code over HxW), x) along channels (model.py:411-414 / 956-959), NHWC bf16."""

    @staticmethod
    def forward(ctx, code, x):
        B, H, W, Cx = x.shape
        Cc = code.shape[1]
        code = code.float().contiguous()
        out = torch.empty(B, H, W, Cc + Cx, device=x.device, dtype=torch.bfloat16)
        L.check(L.lib().ekl_cat_code(L.ptr(code), Cc, L.ptr(x), Cx, B, H * W, L.ptr(out), L.stream()))
        _count()
        ctx.dims = (B, H, W, Cc, Cx)
        return out

    @staticmethod
    def backward(ctx, dcat):
        B, H, W, Cc, Cx = ctx.dims
        dcat = dcat.contiguous()
        dcode = torch.zeros(B, Cc, device=dcat.device, dtype=torch.float32)
        dx = torch.empty(B, H, W, Cx, device=dcat.device, dtype=torch.bfloat16)
        L.check(L.lib().ekl_cat_code_bwd(L.ptr(dcat), Cc, Cx, B, H * W, L.ptr(dcode), L.ptr(dx), L.stream()))
        _count()
        return (dcode if ctx.needs_input_grad[0] else None), (dx if ctx.needs_input_grad[1] else None)


def cat_code(code, x):
    return _CatCode.apply(code, x)


class _ImgS2D(torch.autograd.Function):
    """Space-to-depth of up to three NCHW fp32 image batches into one NHWC bf16 [G*B, H/2, W/2, 16] tensor
    (include/ekl_b200.h: ekl_img_s2d); backward = inverse map of the (single) image batch that needs a gradient."""

    @staticmethod
    def forward(ctx, *imgs):
        B, C, H, W = imgs[0].shape
        assert C == 3 and len(imgs) <= 3
        imgs = [t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous() for t in imgs]
        out = torch.empty(len(imgs) * B, H // 2, W // 2, 16, device=imgs[0].device, dtype=torch.bfloat16)
        p = [L.ptr(t) for t in imgs] + [None] * (3 - len(imgs))
        L.check(L.lib().ekl_img_s2d(p[0], p[1], p[2], len(imgs), B, H, W, L.ptr(out), L.stream()))
        _count()
        ctx.dims = (B, H, W, len(imgs))
        return out

    @staticmethod
    def backward(ctx, dout):
        B, H, W, n = ctx.dims
        dout = dout.contiguous()
        grads = []
        for g in range(n):
            if not ctx.needs_input_grad[g]:
                grads.append(None)
                continue
            dx = torch.empty(B, 3, H, W, device=dout.device, dtype=torch.float32)
            L.check(L.lib().ekl_img_s2d_bwd(L.ptr(dout[g * B:(g + 1) * B]), B, H, W, L.ptr(dx), L.stream()))
            _count()
            grads.append(dx)
        return tuple(grads)


def img_s2d(*imgs):
    return _ImgS2D.apply(*imgs)


class _HeadTanh(torch.autograd.Function):
    """NHWC bf16 [B,H,W,C] conv output (3 real channels) -> tanh -> NCHW fp32 image [B,3,H,W] (model.py:433-436)."""

    @staticmethod
    def forward(ctx, y):
        B, H, W, C = y.shape
        img = torch.empty(B, 3, H, W, device=y.device, dtype=torch.float32)
        L.check(L.lib().ekl_head_tanh_fwd(L.ptr(y), B, H * W, C, L.ptr(img), L.stream()))
        _count()
        ctx.save_for_backward(y)
        return img

    @staticmethod
    def backward(ctx, dimg):
        (y,) = ctx.saved_tensors
        B, H, W, C = y.shape
        dimg = dimg.float().contiguous()
        dy = torch.empty_like(y)
        L.check(L.lib().ekl_head_tanh_bwd(L.ptr(y), L.ptr(dimg), B, H * W, C, L.ptr(dy), L.stream()))
        _count()
        return dy


def head_tanh(y):
    return _HeadTanh.apply(y)


def _flat_head_weight(w):
    """[1,C,4,4] head filter -> ([16*C] fp32 in NHWC (kh,kw,c) order, shares_memory)."""
    if w.is_contiguous(memory_format=torch.channels_last):
        return w.permute(0, 2, 3, 1).reshape(-1), True
    return w.permute(0, 2, 3, 1).reshape(-1).contiguous(), False


class _DHeadDots(torch.autograd.Function):
    """Raw (pre-sigmoid) uncond / match logits of a discriminator: one 16*8ndf-long dot per sample and head
    (model.py:886-888, 935-952).  x_code / h_c: [GB,4,4,C] NHWC bf16."""

    @staticmethod
    def forward(ctx, x_code, h_c, w_u, b_u, w_m, b_m):
        GB = x_code.shape[0]
        K = x_code.numel() // GB
        fu, ctx.inplace_u = _flat_head_weight(w_u)
        fm, ctx.inplace_m = _flat_head_weight(w_m)
        lu = torch.empty(GB, device=x_code.device, dtype=torch.float32)
        lm = torch.empty(GB, device=x_code.device, dtype=torch.float32)
        L.check(L.lib().ekl_dhead_dots(L.ptr(x_code), L.ptr(h_c), L.ptr(fu), L.ptr(b_u), L.ptr(fm), L.ptr(b_m), GB, K,
                                       L.ptr(lu), L.ptr(lm), L.stream()))
        _count()
        ctx.save_for_backward(x_code, h_c, w_u, b_u, w_m, b_m)
        return lu, lm

    @staticmethod
    def backward(ctx, g_u, g_m):
        x_code, h_c, w_u, b_u, w_m, b_m = ctx.saved_tensors
        GB = x_code.shape[0]
        K = x_code.numel() // GB
        fu, _ = _flat_head_weight(w_u)
        fm, _ = _flat_head_weight(w_m)
        g_u, g_m = g_u.contiguous(), g_m.contiguous()
        dx = torch.empty_like(x_code) if ctx.needs_input_grad[0] else None
        dh = torch.empty_like(h_c) if ctx.needs_input_grad[1] else None
        outs = [None] * 4
        bufs = []
        for i, (p, flat_ok) in enumerate(((w_u, ctx.inplace_u), (b_u, True), (w_m, ctx.inplace_m), (b_m, True))):
            if not ctx.needs_input_grad[2 + i]:
                bufs.append(None)
            elif p.is_leaf and flat_ok:
                bufs.append(_grad_buffer(p))                    # accumulated in place (same memory order)
            else:
                t = torch.zeros(p.numel(), device=p.device, dtype=torch.float32)
                bufs.append(t)
                outs[i] = t
        L.check(L.lib().ekl_dhead_dots_bwd(L.ptr(x_code), L.ptr(h_c), L.ptr(fu), L.ptr(fm), L.ptr(g_u), L.ptr(g_m), GB, K,
                                           L.ptr(dx), L.ptr(dh), L.ptr(bufs[0]), L.ptr(bufs[1]), L.ptr(bufs[2]),
                                           L.ptr(bufs[3]), L.stream()))
        _count()
        for i, p in enumerate((w_u, b_u, w_m, b_m)):
            if outs[i] is not None:
                g = outs[i]
                outs[i] = g.view(1, 4, 4, -1).permute(0, 3, 1, 2) if p.dim() == 4 else g.view(p.shape)
        return (dx, dh) + tuple(outs)


def dhead_dots(x_code, h_c, w_u, b_u, w_m, b_m):
    return _DHeadDots.apply(x_code, h_c, w_u, b_u, w_m, b_m)


def _int3(vals):
    import ctypes as C
    v = list(vals) + [0] * (3 - len(vals))
    return (C.c_int * 3)(*v)


class _DLoss(torch.autograd.Function):
    """BCE (constant labels per group) + soft-target CE of one discriminator pass, fused (include/ekl_b200.h:
    ekl_dloss_fwd).  Returns (losses[4], p_match, p_uncond, log_softmax(cls)); only losses[0] is differentiable."""

    @staticmethod
    def forward(ctx, lm, lu, cls, cp0, cp1, cfg):
        ctx.set_materialize_grads(False)
        groups, B, t_match, t_uncond, cls_tgt, coeff = cfg
        GB, E1 = cls.shape
        assert GB == groups * B
        lm, lu, cls = lm.contiguous(), lu.contiguous(), cls.contiguous()
        dev = cls.device
        losses = torch.empty(4, device=dev, dtype=torch.float32)
        pm = torch.empty(GB, device=dev, dtype=torch.float32)
        pu = torch.empty(GB, device=dev, dtype=torch.float32)
        logp = torch.empty(GB, E1, device=dev, dtype=torch.float32)
        cp0 = cp0.float().contiguous()
        cp1 = cp1.float().contiguous() if cp1 is not None else None
        L.check(L.lib().ekl_dloss_fwd(groups, B, E1, _int3(t_match), _int3(t_uncond), _int3(cls_tgt), float(coeff),
                                      L.ptr(lm), L.ptr(lu), L.ptr(cls), L.ptr(cp0), L.ptr(cp1), L.ptr(losses), L.ptr(pm),
                                      L.ptr(pu), L.ptr(logp), L.stream()))
        _count()
        ctx.cfg = cfg
        ctx.save_for_backward(pm, pu, logp, cp0, cp1)
        ctx.mark_non_differentiable(pm, pu, logp)
        return losses, pm, pu, logp

    @staticmethod
    def backward(ctx, go, *_):
        if go is None:
            return None, None, None, None, None, None
        groups, B, t_match, t_uncond, cls_tgt, coeff = ctx.cfg
        pm, pu, logp, cp0, cp1 = ctx.saved_tensors
        GB, E1 = logp.shape
        go = go.contiguous()
        gm, gu, gcls = torch.empty_like(pm), torch.empty_like(pu), torch.empty_like(logp)
        L.check(L.lib().ekl_dloss_bwd(groups, B, E1, _int3(t_match), _int3(t_uncond), _int3(cls_tgt), float(coeff), L.ptr(go),
                                      L.ptr(pm), L.ptr(pu), L.ptr(logp), L.ptr(cp0), L.ptr(cp1), L.ptr(gm), L.ptr(gu),
                                      L.ptr(gcls), L.stream()))
        _count()
        return gm, gu, gcls, None, None, None


def d_loss(lm, lu, cls, cp0, cp1, groups, B, t_match, t_uncond, cls_tgt, uncond_coeff):
    return _DLoss.apply(lm, lu, cls, cp0, cp1, (groups, B, tuple(t_match), tuple(t_uncond), tuple(cls_tgt), uncond_coeff))


class _ReparamKL(torch.autograd.Function):
    """c = eps*exp(0.5*logvar) + mu, std, and KL(mu, logvar) in one kernel (include/ekl_b200.h: ekl_reparam_kl_fwd).
    mu / logvar: fp32 [B,D] views with unit column stride (e.g. the two halves of a GLU output)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        ctx.set_materialize_grads(False)
        B, D = mu.shape
        assert mu.dtype == torch.float32 and logvar.dtype == torch.float32 and mu.stride(1) == 1 and logvar.stride(1) == 1
        eps = eps.float().contiguous()
        c, std = torch.empty(B, D, device=mu.device), torch.empty(B, D, device=mu.device)
        kl = torch.empty((), device=mu.device)
        L.check(L.lib().ekl_reparam_kl_fwd(L.ptr(mu), mu.stride(0), L.ptr(logvar), logvar.stride(0), L.ptr(eps), B, D, L.ptr(c),
                                           L.ptr(std), L.ptr(kl), L.stream()))
        _count()
        ctx.save_for_backward(mu, logvar, eps)
        return c, std, kl

    @staticmethod
    def backward(ctx, dc, dstd, dkl):
        if dc is None and dstd is None and dkl is None:
            return None, None, None
        mu, logvar, eps = ctx.saved_tensors
        B, D = mu.shape
        dmu, dlv = torch.empty(B, D, device=mu.device), torch.empty(B, D, device=mu.device)
        dc = dc.contiguous() if dc is not None else None
        dstd = dstd.contiguous() if dstd is not None else None
        dkl = dkl.contiguous() if dkl is not None else None
        L.check(L.lib().ekl_reparam_kl_bwd(L.ptr(mu), mu.stride(0), L.ptr(logvar), logvar.stride(0), L.ptr(eps), B, D, L.ptr(dc),
                                           L.ptr(dstd), L.ptr(dkl), L.ptr(dmu), L.ptr(dlv), L.stream()))
        _count()
        return dmu, dlv, None


def reparam_kl(mu, logvar, eps):
    return _ReparamKL.apply(mu, logvar, eps)


class _LinearBnRelu(torch.autograd.Function):
    """h = ReLU(BatchNorm1d(x W^T + b)) in train mode over a batch of <= 64 rows (VC_NET's hidden layers, model.py:169-181;
    include/ekl_b200.h: ekl_linear_bn_relu_fwd / _bwd).  Parameter gradients are accumulated in place."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var):
        B, K = x.shape
        N = weight.shape[0]
        x = x.float().contiguous()
        h = torch.empty(B, N, device=x.device)
        xhat = torch.empty(B, N, device=x.device)
        rstd = torch.empty(N, device=x.device)
        L.check(L.lib().ekl_linear_bn_relu_fwd(L.ptr(x), L.ptr(weight), L.ptr(bias), L.ptr(gamma), L.ptr(beta), L.ptr(running_mean),
                                               L.ptr(running_var), B, K, N, BN_EPS, BN_MOM, 1, L.ptr(h), L.ptr(xhat), L.ptr(rstd),
                                               L.stream()))
        _count()
        ctx.save_for_backward(x, weight, bias, gamma, beta, h, xhat, rstd)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, weight, bias, gamma, beta, h, xhat, rstd = ctx.saved_tensors
        B, K = x.shape
        N = weight.shape[0]
        dh = dh.float().contiguous()
        dy = torch.empty(B, N, device=x.device)

        def buf(p, need):
            if not need or p is None:
                return None, None
            if p.is_leaf:
                return _grad_buffer(p), None
            t = torch.zeros_like(p)
            return t, t
        bw, rw = buf(weight, ctx.needs_input_grad[1])
        bb, rb = buf(bias, ctx.needs_input_grad[2])
        bg, rg = buf(gamma, ctx.needs_input_grad[3])
        be, re_ = buf(beta, ctx.needs_input_grad[4])
        L.check(L.lib().ekl_linear_bn_relu_bwd(L.ptr(dh), L.ptr(h), L.ptr(xhat), L.ptr(rstd), L.ptr(gamma), L.ptr(x), B, K, N, L.ptr(bw),
                                               L.ptr(bb), L.ptr(bg), L.ptr(be), L.ptr(dy), L.stream()))
        _count()
        dx = dy @ weight if ctx.needs_input_grad[0] else None        # one library GEMM, only where the input needs it
        return dx, rw, rb, rg, re_, None, None


def linear_bn_relu(x, linear, bn):
    """linear: nn.Linear, bn: nn.BatchNorm1d (state_dict-compatible holders of the parameters / running statistics)."""
    if not bn.training:
        B, K = x.shape
        N = linear.weight.shape[0]
        x = x.float().contiguous()
        h = torch.empty(B, N, device=x.device)
        L.check(L.lib().ekl_linear_bn_relu_fwd(L.ptr(x), L.ptr(linear.weight), L.ptr(linear.bias), L.ptr(bn.weight), L.ptr(bn.bias),
                                               L.ptr(bn.running_mean), L.ptr(bn.running_var), B, K, N, bn.eps, BN_MOM, 0, L.ptr(h), None,
                                               None, L.stream()))
        _count()
        return h
    if getattr(bn, "_ekl_counted", False):
        bn._ekl_calls += 1
    elif bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return _LinearBnRelu.apply(x, linear.weight, linear.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)


class _ColorStats(torch.autograd.Function):
    """Per-image channel mean [B,3,1,1] and covariance [B,3,3] of an fp32 NCHW image batch (cub:33-52
    compute_mean_covariance; include/ekl_b200.h: ekl_color_stats_fwd / _bwd): one read of the images forward, one read +
    one write backward."""

    @staticmethod
    def forward(ctx, img):
        ctx.set_materialize_grads(False)
        B, C, H, W = img.shape
        mean = torch.empty(B, 3, 1, 1, device=img.device)
        cov = torch.empty(B, 3, 3, device=img.device)
        scratch = torch.empty(B * 9, device=img.device, dtype=torch.float64)
        L.check(L.lib().ekl_color_stats_fwd(L.ptr(img), B, H * W, L.ptr(scratch), L.ptr(mean), L.ptr(cov), L.stream()))
        _count(2)
        ctx.save_for_backward(img, mean)
        return mean, cov

    @staticmethod
    def backward(ctx, dmean, dcov):
        if dmean is None and dcov is None:
            return None
        img, mean = ctx.saved_tensors
        B, C, H, W = img.shape
        dmean = dmean.float().contiguous() if dmean is not None else None
        dcov = dcov.float().contiguous() if dcov is not None else None
        dimg = torch.empty_like(img)
        L.check(L.lib().ekl_color_stats_bwd(L.ptr(img), L.ptr(mean), L.ptr(dmean), L.ptr(dcov), B, H * W, L.ptr(dimg), L.stream()))
        _count()
        return dimg


def color_stats_supported(img):
    return (img.is_cuda and img.dim() == 4 and img.shape[1] == 3 and img.dtype == torch.float32 and img.is_contiguous()
            and (img.shape[2] * img.shape[3]) % 4 == 0)


def color_stats(img):
    return _ColorStats.apply(img)
