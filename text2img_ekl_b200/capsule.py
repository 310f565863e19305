"""CapsuleLinear -- stand-in for the un-vendored third-party `capsule_layer.modules.CapsuleLinear` the reference
imports (model.py:12; call sites :248,290,301,943,1082).  *Parity unpinned*: the package is absent from the
reference tree and its version is not pinned anywhere, so the arithmetic follows the published routing-by-agreement
algorithm as restated (and documented) in oracle/capsule_ref.py; this module is checked against that restatement.

Shared-weight mode only (what the reference uses): weight [out_capsules, out_length, in_length], input
[B, in_capsules, in_length] -> [B, out_capsules, out_length].

The [B, O, I, L] prior tensor (201 MB at B=32 for the generator stem) is never formed.  On the GPU the
small-in_length case (the generator stem, in_length 8) runs on the capsule kernels of csrc/capsule.cu (agreement
logits / softmax over out-capsules kept on chip).  The wide case (the discriminator class head: in_length 512, 16
in-capsules, 201 out-capsules of length 16) is the opposite regime -- there the prior is small (206 KB per sample) -- so
it is one library GEMM for the prior and ONE routing kernel per direction (csrc/capsule_route.cu); other shapes run the
reduced algebra below as GEMMs through torch.einsum.  Because
prior[b,o,i,:] = W[o] x[b,i], every routing quantity lives in in_length space:
    logit[b,o,i] = <x[b,i], u[b,o]>,  u[b,o] = W[o]^T (sum of earlier v[b,o])
    s[b,o]       = W[o] y[b,o],       y[b,o] = sum_i softmax_o(logit)[b,o,i] x[b,i]
(property pinned on the CPU by tests/test_oracle_golden.py::test_capsule_reduced_algebra_equals_materialised_priors).
"""
import torch
import torch.nn as nn

EPS = 1e-8


def squash(s):
    n2 = (s * s).sum(dim=-1, keepdim=True)
    return s * (n2 / (1.0 + n2) / torch.sqrt(n2 + EPS))


# ------------------------------------------------------------------ CUDA kernels (small in_length: the generator stem)
def _k():
    from . import _lib as L
    return L, L.lib()


def _count(n=1):
    from . import ops
    ops._count(n)


class _ProjU(torch.autograd.Function):
    """u[b,o,:] = W[o]^T v[b,o,:]"""

    @staticmethod
    def forward(ctx, W, v):
        L, lib = _k()
        B, O, Lh = v.shape
        K = W.shape[2]
        W, v = W.contiguous(), v.contiguous()
        u = torch.empty(B, O, K, device=v.device)
        L.check(lib.ekl_caps_proj_u(L.ptr(W), L.ptr(v), B, O, Lh, K, L.ptr(u), L.stream()))
        _count()
        ctx.save_for_backward(W, v)
        return u

    @staticmethod
    def backward(ctx, gu):
        L, lib = _k()
        W, v = ctx.saved_tensors
        B, O, Lh = v.shape
        K = W.shape[2]
        gu = gu.contiguous()
        gv = torch.empty_like(v)
        L.check(lib.ekl_caps_proj_s(L.ptr(W), L.ptr(gu), B, O, Lh, K, L.ptr(gv), None, L.stream()))      # gv = W gu
        gW = torch.zeros_like(W)
        L.check(lib.ekl_caps_outer(L.ptr(v), L.ptr(gu), B, O, Lh, K, L.ptr(gW), L.stream()))
        _count(2)
        return gW, gv


class _SSquash(torch.autograd.Function):
    """v = squash(W[o] y[b,o,:])"""

    @staticmethod
    def forward(ctx, W, y):
        L, lib = _k()
        B, O, K = y.shape
        Lh = W.shape[1]
        W, y = W.contiguous(), y.contiguous()
        s = torch.empty(B, O, Lh, device=y.device)
        v = torch.empty(B, O, Lh, device=y.device)
        L.check(lib.ekl_caps_proj_s(L.ptr(W), L.ptr(y), B, O, Lh, K, L.ptr(s), L.ptr(v), L.stream()))
        _count()
        ctx.save_for_backward(W, y, s)
        return v

    @staticmethod
    def backward(ctx, gv):
        L, lib = _k()
        W, y, s = ctx.saved_tensors
        B, O, K = y.shape
        Lh = W.shape[1]
        gv = gv.contiguous()
        gs, gy = torch.empty_like(s), torch.empty_like(y)
        L.check(lib.ekl_caps_squash_bwd(L.ptr(W), L.ptr(s), L.ptr(gv), B, O, Lh, K, L.ptr(gs), L.ptr(gy), L.stream()))
        gW = torch.zeros_like(W)
        L.check(lib.ekl_caps_outer(L.ptr(gs), L.ptr(y), B, O, Lh, K, L.ptr(gW), L.stream()))
        _count(2)
        return gW, gy


class _Agree(torch.autograd.Function):
    """y[b,o,:] = sum_i softmax_o(<x[b,i], u[b,o]>) x[b,i]   (routing-by-agreement step; logits stay on chip)"""

    @staticmethod
    def forward(ctx, x, u):
        L, lib = _k()
        B, I, K = x.shape
        O = u.shape[1]
        x, u = x.contiguous(), u.contiguous()
        y = torch.empty(B, O, K, device=x.device)
        M, Z = torch.empty(B, I, device=x.device), torch.empty(B, I, device=x.device)
        L.check(lib.ekl_caps_agree_fwd(L.ptr(x), L.ptr(u), B, I, O, K, L.ptr(y), L.ptr(M), L.ptr(Z), L.stream()))
        _count()
        ctx.save_for_backward(x, u, M, Z)
        return y

    @staticmethod
    def backward(ctx, gy):
        L, lib = _k()
        x, u, M, Z = ctx.saved_tensors
        B, I, K = x.shape
        O = u.shape[1]
        gy = gy.contiguous()
        gu, gx = torch.empty_like(u), torch.empty_like(x)
        L.check(lib.ekl_caps_agree_bwd(L.ptr(x), L.ptr(u), L.ptr(M), L.ptr(Z), L.ptr(gy), B, I, O, K, L.ptr(gu), L.ptr(gx),
                                       L.stream()))
        _count()
        return gx, gu


class _Route(torch.autograd.Function):
    """prior [B,I,O,L] -> (v [B,O,L], |v| [B,O]): the whole routing in one kernel per direction (csrc/capsule_route.cu;
    the discriminator class head, where in_length is wide and the prior small)."""

    @staticmethod
    def forward(ctx, prior, iters):
        L, lib = _k()
        B, I, O, Lh = prior.shape
        prior = prior.contiguous()
        v = torch.empty(B, O, Lh, device=prior.device)
        n = torch.empty(B, O, device=prior.device)
        L.check(lib.ekl_caps_route_fwd(L.ptr(prior), B, I, O, Lh, iters, L.ptr(v), L.ptr(n), L.stream()))
        _count()
        ctx.save_for_backward(prior)
        ctx.iters = iters
        return v, n

    @staticmethod
    def backward(ctx, gv, gn):
        L, lib = _k()
        prior, = ctx.saved_tensors
        B, I, O, Lh = prior.shape
        gp = torch.empty_like(prior)
        gv = gv.contiguous() if gv is not None else None
        gn = gn.contiguous() if gn is not None else None
        L.check(lib.ekl_caps_route_bwd(L.ptr(prior), L.ptr(gv) if gv is not None else None, L.ptr(gn) if gn is not None else None,
                                       B, I, O, Lh, ctx.iters, L.ptr(gp), L.stream()))
        _count()
        return gp, None


def _route_wide(x, weight, num_iterations):
    """(v, |v|) through the materialised-prior kernels, or None when the shape is outside their regime."""
    if not (x.is_cuda and weight.dtype == torch.float32):
        return None
    B, I, K = x.shape
    O, Lh = weight.shape[0], weight.shape[1]
    L, lib = _k()
    if K < 64 or not lib.ekl_caps_route_supported(I, O, Lh, num_iterations):
        return None
    prior = torch.matmul(x.reshape(B * I, K), weight.reshape(O * Lh, K).t()).view(B, I, O, Lh)
    return _Route.apply(prior, num_iterations)


def _dynamic_routing_cuda(x, weight, num_iterations):
    B, I, K = x.shape
    O = weight.shape[0]
    y = (x.sum(dim=1, keepdim=True) / O).expand(B, O, K)          # iteration 0: zero logits -> uniform coupling 1/O
    v = _SSquash.apply(weight, y)
    vsum = v
    for _ in range(1, num_iterations):
        v = _SSquash.apply(weight, _Agree.apply(x, _ProjU.apply(weight, vsum)))
        vsum = vsum + v
    return v


def capsule_linear(x, weight, routing_type="dynamic", num_iterations=3):
    x = x.float()
    B, O = x.shape[0], weight.shape[0]
    if routing_type == "dynamic" and x.is_cuda and weight.dtype == torch.float32:
        L, lib = _k()
        if lib.ekl_caps_supported(x.shape[1], x.shape[2], O, weight.shape[1]):
            return _dynamic_routing_cuda(x, weight, num_iterations)
        routed = _route_wide(x, weight, num_iterations)
        if routed is not None:
            return routed[0]
    if routing_type == "dynamic":
        vsum = x.new_zeros(B, O, weight.shape[1])
        v = None
        for r in range(num_iterations):
            if r == 0:
                # logits are zero: softmax over O is uniform 1/O
                y = (x.sum(dim=1, keepdim=True) / O).expand(B, O, x.shape[2])
            else:
                u = torch.einsum("olk,bol->bok", weight, vsum)
                c = torch.softmax(torch.einsum("bik,bok->boi", x, u), dim=1)
                y = torch.einsum("boi,bik->bok", c, x)
            v = squash(torch.einsum("olk,bok->bol", weight, y))
            vsum = vsum + v
        return v
    if routing_type == "k_means":
        out = torch.einsum("olk,bk->bol", weight, x.mean(dim=1))
        for r in range(num_iterations):
            u = torch.einsum("olk,bol->bok", weight, torch.nn.functional.normalize(out, dim=-1))
            c = torch.softmax(torch.einsum("bik,bok->boi", x, u), dim=1)
            out = torch.einsum("olk,bok->bol", weight, torch.einsum("boi,bik->bok", c, x))
        return out
    raise ValueError(routing_type)


class CapsuleLinear(nn.Module):
    ROUTING_TYPE = "dynamic"
    NUM_ITERATIONS = 3

    def __init__(self, out_capsules, in_length, out_length, in_capsules=None, share_weight=True, routing_type=None,
                 num_iterations=None, **kwargs):
        super().__init__()
        if in_capsules is not None or not share_weight:
            raise ValueError("only the shared-weight mode used by the reference is implemented")
        self.out_capsules, self.in_length, self.out_length = out_capsules, in_length, out_length
        self.routing_type = routing_type or self.ROUTING_TYPE
        self.num_iterations = num_iterations or self.NUM_ITERATIONS
        self.weight = nn.Parameter(torch.empty(out_capsules, out_length, in_length))
        nn.init.xavier_uniform_(self.weight)
        self.bias = None          # read by the reference's weights_init (class name contains "Linear")

    def forward(self, input):
        return capsule_linear(input, self.weight, self.routing_type, self.num_iterations)

    def forward_norm(self, input):
        """forward(input).norm(dim=-1) (model.py:969-970), the norm fused into the routing kernel where that runs."""
        if self.routing_type == "dynamic" and input.is_cuda:
            routed = _route_wide(input.float(), self.weight, self.num_iterations)
            if routed is not None:
                return routed[1]
        return self.forward(input).norm(dim=-1)

    def extra_repr(self):
        return "out_capsules=%d, in_length=%d, out_length=%d, routing=%s x%d" % (
            self.out_capsules, self.in_length, self.out_length, self.routing_type, self.num_iterations)
