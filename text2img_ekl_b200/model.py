"""Drop-in for the reference's model.py: the same class names, constructor signatures, forward signatures /
return tuples and state_dict keys + shapes (reference file:line cited per class), with the forward / backward
work done by the sm_100a kernels in libekl_b200.so (ops.py) instead of cuDNN / ATen.

Tensor conventions at this boundary
  * every module accepts NCHW-shaped tensors of any dtype / memory format (what a reference user passes);
  * feature maps are returned NCHW-shaped but stored channels_last in bf16 (a zero-copy view of the NHWC buffers
    the kernels use); images (GET_IMAGE_G) and all logits / codes / losses are fp32 like the reference's;
  * constructors read the global `cfg` at construction time, exactly like the reference.
RNG injection (SURVEY 8b extension): CA_NET / VC_NET / the G nets take optional eps / seed tensors; by default
they draw as the reference does (CA eps on the device, VC seed on the host).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import capsule
from . import ops
from .miscc.config import cfg


# ------------------------------------------------------------------------------------------- helpers
def to_nhwc(x):
    """NCHW-shaped tensor (any layout / dtype) -> contiguous NHWC bf16 (no copy if it already is one)."""
    x = x.permute(0, 2, 3, 1)
    if x.dtype != torch.bfloat16:
        x = x.to(torch.bfloat16)
    return x.contiguous()


def to_public(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2)


def _tc_ok(cin, cout):
    return cin % 16 == 0 and cout % 32 == 0


def _spec(mode, cin, cout, **kw):
    return ops.ConvSpec(mode, cin, cout, impl=L.IMPL_TC if _tc_ok(cin, cout) else L.IMPL_SIMT, **kw)


def _conv_bn_act(x, conv_mod, bn_mod, spec, act, groups=1, residual=None):
    return ops.conv_bn_act(x, conv_mod.weight, spec, bn_mod, groups, act, residual)


class Reshape(nn.Module):          # model.py:50-56
    def __init__(self, *args):
        super().__init__()
        self.shape = args

    def forward(self, x):
        return x.view(self.shape)


class Permute(nn.Module):          # model.py:59-65
    def __init__(self, *args):
        super().__init__()
        self.shape = args

    def forward(self, x):
        return x.permute(self.shape).contiguous()


class GLU(nn.Module):              # model.py:68-76 (first half * sigmoid(second half) on dim 1)
    def forward(self, x):
        nc = x.size(1)
        assert nc % 2 == 0, "channels dont divide 2!"
        nc //= 2
        return x[:, :nc] * torch.sigmoid(x[:, nc:])


def conv3x3(in_planes, out_planes):   # model.py:79-82
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=1, padding=1, bias=False)


# ------------------------------------------------------------------------------------------- G blocks
class _UpBlock(nn.Sequential):
    """model.py:87-94 upBlock: Upsample(x2 nearest) -> conv3x3 -> BN -> GLU as ONE conv kernel (4 sub-pixel 2x2 convs,
    BN partial statistics in its epilogue) + one fused normalise+GLU pass.  Children keep the reference indices 0-3."""

    def __init__(self, cin, cout):
        super().__init__(nn.Upsample(scale_factor=2, mode="nearest"), conv3x3(cin, cout * 2),
                         nn.BatchNorm2d(cout * 2), GLU())
        self._spec = _spec(ops.UP2, cin, cout * 2)

    def forward(self, x):
        return to_public(_conv_bn_act(to_nhwc(x), self[1], self[2], self._spec, ops.ACT_GLU))


def upBlock(in_planes, out_planes):
    return _UpBlock(in_planes, out_planes)


class _Block3x3Glu(nn.Sequential):
    """model.py:98-104 Block3x3_relu: conv3x3 -> BN -> GLU."""

    def __init__(self, cin, cout):
        super().__init__(conv3x3(cin, cout * 2), nn.BatchNorm2d(cout * 2), GLU())
        self._spec = _spec(ops.S1, cin, cout * 2)

    def forward(self, x):
        return to_public(_conv_bn_act(to_nhwc(x), self[0], self[1], self._spec, ops.ACT_GLU))


def Block3x3_relu(in_planes, out_planes):
    return _Block3x3Glu(in_planes, out_planes)


class ResBlock(nn.Module):
    """model.py:107-123: conv-BN-GLU-conv-BN + identity; the skip add is fused into the second normalise pass."""

    def __init__(self, channel_num):
        super().__init__()
        c = channel_num
        self.block = nn.Sequential(conv3x3(c, c * 2), nn.BatchNorm2d(c * 2), GLU(), conv3x3(c, c), nn.BatchNorm2d(c))
        self._s0, self._s3 = _spec(ops.S1, c, c * 2), _spec(ops.S1, c, c)

    def forward(self, x):
        x = to_nhwc(x)
        b = self.block
        h = _conv_bn_act(x, b[0], b[1], self._s0, ops.ACT_GLU)
        return to_public(_conv_bn_act(h, b[3], b[4], self._s3, ops.ACT_NONE, residual=x))


class CA_NET(nn.Module):
    """model.py:126-157 conditioning augmentation: fc(+bias) -> GLU -> (mu, logvar) -> c = eps*std + mu."""

    def __init__(self, cond_dim=None):
        super().__init__()
        self.t_dim = cfg.TEXT.DIMENSION
        self.ef_dim = cfg.GAN.EMBEDDING_DIM
        self.fc = nn.Linear(cond_dim if cond_dim is not None else self.t_dim, self.ef_dim * 4, bias=True)
        self.relu = GLU()

    def encode(self, text_embedding):
        x = self.relu(self.fc(text_embedding))
        return x[:, :self.ef_dim], x[:, self.ef_dim:]

    def reparametrize(self, mu, logvar, eps=None):
        std = torch.exp(0.5 * logvar)
        if eps is None:
            eps = torch.randn_like(std)          # reference draws on the device (model.py:147-150)
        return eps * std + mu, std

    def forward(self, text_embedding, eps=None):
        mu, logvar = self.encode(text_embedding)
        if mu.is_cuda:
            if eps is None:
                eps = torch.randn_like(mu)          # reference draws on the device (model.py:147-150)
            # one fused kernel: reparameterisation + the KL term the trainer takes of (mu, logvar) (cub:54-58)
            c_code, std, self.last_kl = ops.reparam_kl(mu, logvar, eps)
            self.last_kl_of = (mu, logvar)
            return c_code, mu, logvar, std
        c_code, std = self.reparametrize(mu, logvar, eps)
        return c_code, mu, logvar, std


class VC_NET(nn.Module):
    """model.py:160-201: cat(noise, cond) -> fc1-BN-ReLU -> fc2-BN-ReLU -> (fc31, fc32) -> c = seed*std + mu."""

    def __init__(self, cond_dim):
        super().__init__()
        self.cond_dim = cond_dim
        self.noise_dim = cfg.GAN.Z_DIM
        self.manifd_dim = cfg.GAN.MANIFD_DIM
        self.threshold = -1
        self.fc1 = nn.Linear(self.cond_dim + self.noise_dim, 512)
        self.bn_fc1 = nn.BatchNorm1d(512)
        self.fc2 = nn.Linear(512, 256)
        self.bn_fc2 = nn.BatchNorm1d(256)
        self.fc31 = nn.Linear(256, self.manifd_dim)
        self.fc32 = nn.Linear(256, self.manifd_dim)

    def encode(self, x):
        if x.is_cuda and x.shape[0] <= 64:
            # Linear + BatchNorm1d + ReLU as one kernel per layer and direction (ops.linear_bn_relu)
            h = ops.linear_bn_relu(x, self.fc1, self.bn_fc1)
            h = ops.linear_bn_relu(h, self.fc2, self.bn_fc2)
        else:
            h = F.relu(self.bn_fc1(self.fc1(x)))
            h = F.relu(self.bn_fc2(self.fc2(h)))
        return self.fc31(h), self.fc32(h)

    def reparameterize(self, mu, logvar, seed):
        std = torch.exp(0.5 * logvar)
        return seed * std + mu, std

    def forward(self, noise, cond, seed=None):
        x = torch.cat((noise, cond), 1)
        self.bs = x.shape[0]
        mu, logvar = self.encode(x)
        if seed is None:
            if (not self.training) and self.threshold > 0:
                from scipy.stats import truncnorm as tn        # model.py:194-195
                seed = torch.tensor(tn.rvs(-self.threshold, self.threshold, size=self.bs * self.manifd_dim),
                                    dtype=torch.float).view(self.bs, self.manifd_dim)
            else:
                seed = torch.randn(self.bs, self.manifd_dim)     # HOST draw, like the reference (model.py:192)
            seed = seed.to(mu.device)
        if mu.is_cuda:
            c, std, self.last_kl = ops.reparam_kl(mu, logvar, seed)
            self.last_kl_of = (mu, logvar)
            return c, mu, logvar, std
        c, std = self.reparameterize(mu, logvar, seed)
        return c, mu, logvar, std


def _stem_bn_glu(x_f32, bn, ngf):
    """[B, ngf*32] fp32 -> BatchNorm1d -> GLU -> [B,4,4,ngf] NHWC bf16 (model.py:215-216,225)."""
    B = x_f32.shape[0]
    h = ops.bn_act(x_f32.to(torch.bfloat16).contiguous(), None, bn, 1, ops.ACT_GLU)       # [B, ngf*16]
    return h.view(B, ngf, 4, 4).permute(0, 2, 3, 1).contiguous()


class _InitStageBase(nn.Module):
    def _make_ups(self, ngf):
        self.upsample1 = upBlock(ngf, ngf // 2)
        self.upsample2 = upBlock(ngf // 2, ngf // 4)
        self.upsample3 = upBlock(ngf // 4, ngf // 8)
        self.upsample4 = upBlock(ngf // 8, ngf // 16)

    def _ups(self, x):
        return self.upsample4(self.upsample3(self.upsample2(self.upsample1(to_public(x)))))


class COND_INIT_STAGE_G(_InitStageBase):
    """model.py:204-235: Linear(no bias) -> BN1d -> GLU -> [B,ngf,4,4] -> 4 upBlocks."""

    def __init__(self, ngf):
        super().__init__()
        self.in_dim = cfg.GAN.MANIFD_DIM * 2 if cfg.TRAIN.CAT_Z == "concat" else cfg.GAN.MANIFD_DIM
        self.gf_dim = ngf
        self.fc = nn.Sequential(nn.Linear(self.in_dim, ngf * 4 * 4 * 2, bias=False), nn.BatchNorm1d(ngf * 4 * 4 * 2), GLU())
        self._make_ups(ngf)

    def forward(self, ac_x):
        x = _stem_bn_glu(self.fc[0](ac_x), self.fc[1], self.gf_dim)
        return self._ups(ops.grad_mark(x, self, self.fc))          # everything after the stem Linear is final past this point


class COND_INIT_STAGE_G_withCap(_InitStageBase):
    """model.py:238-277: cat(z, noise) -> [B,-1,8] -> CapsuleLinear(ngf capsules of length 32) -> BN1d -> GLU -> ..."""

    def __init__(self, ngf):
        super().__init__()
        self.in_dim = cfg.GAN.MANIFD_DIM
        self.gf_dim = ngf
        self.bs = cfg.TRAIN.BATCH_SIZE
        self.fc_cap = nn.Sequential(
            Reshape(self.bs, -1, 8),
            capsule.CapsuleLinear(out_capsules=ngf, in_length=8, out_length=4 * 4 * 2, in_capsules=None),
            Reshape(-1, ngf * 4 * 4 * 2), nn.BatchNorm1d(ngf * 4 * 4 * 2), GLU())
        self._make_ups(ngf)

    def forward(self, z, noise=None):
        if noise is not None:
            z = torch.cat((z, noise), 1)
        caps = self.fc_cap[1](z.view(z.shape[0], -1, 8))
        x = _stem_bn_glu(caps.reshape(z.shape[0], -1), self.fc_cap[3], self.gf_dim)
        return self._ups(ops.grad_mark(x, self, self.fc_cap))


class COND_INIT_STAGE_G_Exchange_Cap(_InitStageBase):
    """model.py:280-333: two capsule stems (ngf capsules of length 16) on the two halves of z, concatenated."""

    def __init__(self, ngf):
        super().__init__()
        self.in_dim = cfg.GAN.MANIFD_DIM
        self.gf_dim = ngf
        self.bs = cfg.TRAIN.BATCH_SIZE

        def stem():
            return nn.Sequential(
                Reshape(self.bs, -1, 8),
                capsule.CapsuleLinear(out_capsules=(ngf // 2) * 2, in_length=8, out_length=4 * 4, in_capsules=None),
                Reshape(-1, (ngf // 2) * 4 * 4 * 2), nn.BatchNorm1d((ngf // 2) * 4 * 4 * 2), GLU())
        self.fc_cap, self.fc_cap1 = stem(), stem()
        self._make_ups(ngf)

    def forward(self, z):
        B = z.shape[0]
        outs = []
        for zz, st in ((z[:, :self.in_dim], self.fc_cap), (z[:, self.in_dim:], self.fc_cap1)):
            caps = st[1](zz.contiguous().view(B, -1, 8))
            outs.append(_stem_bn_glu(caps.reshape(B, -1), st[3], self.gf_dim // 2))
        return self._ups(torch.cat(outs, dim=3))


class INIT_STAGE_G(_InitStageBase):
    """model.py:336-376 (StackGAN++ stem): in = cat(c_code, z_code) when conditioned."""

    def __init__(self, ngf):
        super().__init__()
        self.gf_dim = ngf
        self.in_dim = cfg.GAN.Z_DIM + cfg.GAN.EMBEDDING_DIM if cfg.GAN.B_CONDITION else cfg.GAN.Z_DIM
        self.fc = nn.Sequential(nn.Linear(self.in_dim, ngf * 4 * 4 * 2, bias=False), nn.BatchNorm1d(ngf * 4 * 4 * 2), GLU())
        self._make_ups(ngf)

    def forward(self, z_code, c_code=None):
        in_code = torch.cat((c_code, z_code), 1) if cfg.GAN.B_CONDITION and c_code is not None else z_code
        return self._ups(_stem_bn_glu(self.fc[0](in_code), self.fc[1], self.gf_dim))


class NEXT_STAGE_G(nn.Module):
    """model.py:379-423: cat(tile(c_code), h) -> jointConv -> R_NUM ResBlocks -> upBlock (-> upBlock if SCALE 4).

    The tiled code channels are spatially constant, so the jointConv over cat(tile(c), h) is computed as a conv over the
    h channels only (a window of the master filter's input channels) plus a per-sample bias with 9 border variants
    (ops.code_bias9: bias9[b, q] = sum of the code's tap responses T[b, t] = Wc[:, :, t] c[b] over the taps that fall
    inside the map for border class q): the [B, ef+ngf, H, W] tensor is never formed, the conv contraction drops from
    9*(ef+ngf) to 9*ngf, and the filter gradient lands in place in the two channel ranges of the one master gradient."""

    def __init__(self, ngf, num_residual=None):
        super().__init__()
        self.gf_dim = ngf
        if cfg.GAN.B_CONDITION:
            self.ef_dim = cfg.GAN.EMBEDDING_DIM * 2 if cfg.TRAIN.CAT_Z == "concat" else cfg.GAN.EMBEDDING_DIM
        else:
            self.ef_dim = cfg.GAN.Z_DIM
        self.num_residual = cfg.GAN.R_NUM if num_residual is None else num_residual
        self.jointConv = Block3x3_relu(ngf + self.ef_dim, ngf)
        self.residual = nn.Sequential(*[ResBlock(ngf) for _ in range(self.num_residual)])
        self.upsample = upBlock(ngf, ngf // 2)
        if cfg.TREE.SCALE == 4:
            self.upsample2 = upBlock(ngf // 2, ngf // 4)
        self._fold = ngf % 16 == 0 and (2 * ngf) % 32 == 0 and self.ef_dim % 4 == 0 and 2 * ngf <= 512
        if self._fold:
            # the conv proper contracts the h channels only: a window [ef, ef + ngf) of the master filter's input channels
            self._spec_x = ops.ConvSpec(ops.S1, ngf, 2 * ngf, impl=L.IMPL_TC, w_cin_total=ngf + self.ef_dim, w_cin_off=self.ef_dim)
            self._spec_x.ref = (ngf + self.ef_dim, 2 * ngf, 9)          # the reference convolves the tiled code channels too

    def _joint(self, h_code, c_code):
        x = to_nhwc(h_code)
        c = c_code.view(-1, self.ef_dim)
        conv_mod, bn_mod = self.jointConv[0], self.jointConv[1]
        w = conv_mod.weight                                              # [2ngf, ef+ngf, 3, 3], code channels first
        if not (self._fold and x.shape[1] >= 2 and x.shape[2] >= 2 and w.is_contiguous(memory_format=torch.channels_last)):
            return self.jointConv(to_public(ops.cat_code(c, x)))
        bias9 = ops.code_bias9(c.float(), w)                             # [B, 9, 2ngf]: the code channels' contribution
        y, stats = ops.conv_bias9(x, w, bias9, self._spec_x, want_stats=bn_mod.training)
        return to_public(ops.bn_act(y, stats, bn_mod, 1, ops.ACT_GLU))

    def forward(self, h_code, c_code):
        out = self.upsample(self.residual(self._joint(h_code, c_code)))
        if cfg.TREE.SCALE == 4:
            out = self.upsample2(out)
        return out


class _TanhFromOut(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out):
        ctx.save_for_backward(out)
        return out.view_as(out)

    @staticmethod
    def backward(ctx, d):
        (out,) = ctx.saved_tensors
        return d * (1.0 - out * out)


class GET_IMAGE_G(nn.Module):
    """model.py:426-437: conv3x3(ngf -> 3) + tanh.  Runs on the tcgen05 conv kernel with the 3 output channels
    zero-padded to one 16-wide N tile (the layer is HBM-bound: it reads h_code once); one fused pass applies tanh in
    fp32 to the 3 real channels and writes the reference's NCHW fp32 image (tanh' is evaluated from the
    pre-activation in backward: a bf16 tanh output would lose 1 - y^2 near saturation)."""
    PAD = 16

    def __init__(self, ngf):
        super().__init__()
        self.gf_dim = ngf
        self.img = nn.Sequential(conv3x3(ngf, 3), nn.Tanh())
        if ngf % 16 == 0:
            # 3 real filters in a 16-wide output tile: the packed rows beyond are zero, their gradients are dropped
            self._spec = ops.ConvSpec(ops.S1, ngf, self.PAD, impl=L.IMPL_TC, w_cout_valid=3)
            self._spec.ref = (ngf, 3, 9)
        else:
            self._spec = ops.ConvSpec(ops.S1, ngf, 3, impl=L.IMPL_SIMT, y_fmt=L.FMT_NCHW_F32, act=ops.ACT_TANH)

    def forward(self, h_code):
        w = self.img[0].weight
        if self._spec.impl == L.IMPL_SIMT:
            y, _ = ops.conv(to_nhwc(h_code), w, self._spec)
            return _TanhFromOut.apply(y)
        if not w.is_contiguous(memory_format=torch.channels_last):
            w = w.contiguous(memory_format=torch.channels_last)
        y, _ = ops.conv(to_nhwc(h_code), w, self._spec)                       # [B,H,W,16] bf16 pre-activation
        return ops.head_tanh(y)


def get_shareGs(gf_dim):            # model.py:439-451
    share_Gs = []
    if cfg.TREE.BRANCH_NUM > 0:
        share_Gs.append(GET_IMAGE_G(gf_dim))
    if cfg.TREE.BRANCH_NUM > 1:
        share_Gs.append(GET_IMAGE_G(gf_dim // cfg.TREE.SCALE))
    if cfg.TREE.BRANCH_NUM > 2:
        share_Gs.append(GET_IMAGE_G(gf_dim // cfg.TREE.SCALE ** 2))
    return share_Gs


class _GBase(nn.Module):
    """Registration order = forward order: conditioning nets (registered by the subclass BEFORE _build_stages), then the
    stages.  The subclasses mark the condition code (ops.grad_mark, a no-op unless the step engine registered a callback):
    when ITS gradient arrives -- after the backward of every stage, the code feeds them all -- every parameter registered
    after the conditioning nets is final, so their optimiser update / gradient exchange overlaps the conditioning nets'
    backward (engine.TailUpdate).  The stem marks its Linear / capsule output the same way."""

    def _build_stages(self, share_Gs, h_net1):
        if cfg.TREE.BRANCH_NUM > 0:
            self.h_net1 = h_net1
            self.img_net1 = share_Gs[0]
        if cfg.TREE.BRANCH_NUM > 1:
            self.h_net2 = NEXT_STAGE_G(self.gf_dim)
            self.img_net2 = share_Gs[1]
        if cfg.TREE.BRANCH_NUM > 2:
            self.h_net3 = NEXT_STAGE_G(self.gf_dim // cfg.TREE.SCALE)
            self.img_net3 = share_Gs[2]

    def _later_stages(self, h_code1, c_code):
        h_codes = [h_code1]
        if cfg.TREE.BRANCH_NUM > 1:
            h_codes.append(self.h_net2(h_codes[-1], c_code))
        if cfg.TREE.BRANCH_NUM > 2:
            h_codes.append(self.h_net3(h_codes[-1], c_code))
        return h_codes

    def image(self, hcodes):                 # model.py:547-563 / 649-665 / 728-744
        return [getattr(self, "img_net%d" % (i + 1))(h) for i, h in enumerate(hcodes[:cfg.TREE.BRANCH_NUM])]

    def get_image(self, entity_hcodes, sen_hcodes):   # model.py:529-545: element-wise product of two code sets
        return [getattr(self, "img_net%d" % (i + 1))(entity_hcodes[i] * sen_hcodes[i])
                for i in range(min(cfg.TREE.BRANCH_NUM, len(entity_hcodes)))]


class COND_G_NET_CATZ_CA(_GBase):
    """model.py:455-563: split-z generator: c = cat|product|sum(CA_NET(sentence), VC_NET(noise, class))."""

    def __init__(self, sen_dim, cls_dim, share_Gs, use_cap=False, cat="concat", exchange=False):
        super().__init__()
        self.gf_dim = cfg.GAN.GF_DIM
        self.ca_net1 = CA_NET()
        self.vc_net2 = VC_NET(cls_dim)
        self.cat, self.exchange = cat, exchange
        self.cls_prior = torch.zeros(cfg.TRAIN.BATCH_SIZE, cfg.GAN.MANIFD_DIM)      # test-time prior buffer (:464)
        if use_cap:
            h1 = COND_INIT_STAGE_G_Exchange_Cap(self.gf_dim * 16) if exchange else COND_INIT_STAGE_G_withCap(self.gf_dim * 16)
        else:
            h1 = COND_INIT_STAGE_G(self.gf_dim * 16)
        self._build_stages(share_Gs, h1)

    def forward(self, noise, sen, cls=None, cls_prior=None, eps=None, seed=None):
        c_code1, mu1, logvar1, std1 = self.ca_net1(sen, eps)
        if self.training or (not cfg.TEST.CLS_PRIOR):
            c_code2, mu2, logvar2, std2 = self.vc_net2(noise, cls, seed)
        elif cls_prior is not None:
            c_code2, mu2, logvar2, std2 = cls_prior, 0, 0, 0
        else:
            self.cls_prior = self.cls_prior.to(sen.device)
            c_code2, mu2, logvar2, std2 = self.cls_prior.normal_(0, 1), 0, 0, 0
        if self.exchange or self.cat == "concat":
            c_code = torch.cat((c_code1, c_code2), 1)
        elif self.cat == "product":
            c_code = c_code1 * c_code2
        else:
            c_code = c_code1 + c_code2
        c_code = ops.grad_mark(c_code, self, self.vc_net2)        # see _GBase._build_stages
        if isinstance(self.h_net1, COND_INIT_STAGE_G_withCap):
            h_code1 = self.h_net1(c_code, noise)          # model.py:512
        else:
            h_code1 = self.h_net1(c_code)                 # (the reference passes 2 args and fails here: SURVEY A#2)
        return self._later_stages(h_code1, c_code), mu1, mu2, logvar1, logvar2, std1, std2


class COND_G_NET_CATZ(_GBase):
    """model.py:567-665: as above with two VC_NETs."""

    def __init__(self, sen_dim, cls_dim, share_Gs, use_cap=False, cat="concat", exchange=False):
        super().__init__()
        self.gf_dim = cfg.GAN.GF_DIM
        self.vc_net1 = VC_NET(sen_dim)
        self.vc_net2 = VC_NET(cls_dim)
        self.cat, self.exchange = cat, exchange
        if use_cap:
            h1 = COND_INIT_STAGE_G_Exchange_Cap(self.gf_dim * 16) if exchange else COND_INIT_STAGE_G_withCap(self.gf_dim * 16)
        else:
            h1 = COND_INIT_STAGE_G(self.gf_dim * 16)
        self._build_stages(share_Gs, h1)

    def forward(self, noise, sen, cls, seed1=None, seed2=None):
        c_code1, mu1, logvar1, std1 = self.vc_net1(noise, sen, seed1)
        c_code2, mu2, logvar2, std2 = self.vc_net2(noise, cls, seed2)
        if self.exchange or self.cat == "concat":
            c_code = torch.cat((c_code1, c_code2), 1)
        elif self.cat == "product":
            c_code = c_code1 * c_code2
        else:
            c_code = c_code1 + c_code2
        c_code = ops.grad_mark(c_code, self, self.vc_net2)
        return self._later_stages(self.h_net1(c_code), c_code), mu1, mu2, logvar1, logvar2, std1, std2


class COND_G_NET(_GBase):
    """model.py:669-744: c = VC_NET(noise, cond)."""

    def __init__(self, cond_dim, share_Gs, use_cap=False):
        super().__init__()
        self.gf_dim = cfg.GAN.GF_DIM
        self.vc_net = VC_NET(cond_dim)
        h1 = COND_INIT_STAGE_G_withCap(self.gf_dim * 16) if use_cap else COND_INIT_STAGE_G(self.gf_dim * 16)
        self._build_stages(share_Gs, h1)

    def forward(self, noise, cond, seed=None):
        c_code, mu, logvar, std = self.vc_net(noise, cond, seed)
        c_code = ops.grad_mark(c_code, self, self.vc_net)
        return self._later_stages(self.h_net1(c_code), c_code), mu, logvar, std


class G_NET(_GBase):
    """model.py:747-808 (StackGAN++ G).  The reference's forward unpacks 3 of CA_NET's 4 returns and fails
    (model.py:769 vs :157); here it returns (h_codes, mu, logvar) as that line intends."""

    def __init__(self, share_Gs):
        super().__init__()
        self.gf_dim = cfg.GAN.GF_DIM
        if cfg.GAN.B_CONDITION:
            self.ca_net = CA_NET()
        self._build_stages(share_Gs, INIT_STAGE_G(self.gf_dim * 16))

    def forward(self, z_code, text_embedding=None, eps=None):
        if cfg.GAN.B_CONDITION and text_embedding is not None:
            c_code, mu, logvar, _ = self.ca_net(text_embedding, eps)
        else:
            c_code, mu, logvar = z_code, None, None
        return self._later_stages(self.h_net1(z_code, c_code), c_code), mu, logvar


# ------------------------------------------------------------------------------------------- D blocks
class _Block3x3Lrelu(nn.Sequential):
    """model.py:812-818 Block3x3_leakRelu: conv3x3 -> BN -> LeakyReLU(0.2)."""

    def __init__(self, cin, cout):
        super().__init__(conv3x3(cin, cout), nn.BatchNorm2d(cout), nn.LeakyReLU(0.2, inplace=True))
        self._spec = _spec(ops.S1, cin, cout)

    def forward(self, x, groups=1):
        return to_public(_conv_bn_act(to_nhwc(x), self[0], self[1], self._spec, ops.ACT_LRELU, groups))


def Block3x3_leakRelu(in_planes, out_planes):
    return _Block3x3Lrelu(in_planes, out_planes)


class _DownBlock(nn.Sequential):
    """model.py:822-828 downBlock: conv4x4 s2 p1 -> BN -> LeakyReLU(0.2)."""

    def __init__(self, cin, cout):
        super().__init__(nn.Conv2d(cin, cout, 4, 2, 1, bias=False), nn.BatchNorm2d(cout), nn.LeakyReLU(0.2, inplace=True))
        self._spec = _spec(ops.DOWN2, cin, cout)

    def forward(self, x, groups=1):
        return to_public(_conv_bn_act(to_nhwc(x), self[0], self[1], self._spec, ops.ACT_LRELU, groups))


def downBlock(in_planes, out_planes):
    return _DownBlock(in_planes, out_planes)


def _s2d_filter_index():
    """Index / mask tensors mapping the conv4x4-s2-p1 filter [n, c, kh, kw] (flattened c*16 + kh*4 + kw) onto the
    3x3 filter over space-to-depth channels ch = (c*2 + ph)*2 + pw (include/ekl_b200.h: ekl_img_s2d):
    input row 2*ho - 1 + kh is s2d row ho + di with parity ph for (di, ph) = (-1,1), (0,0), (0,1), (1,0) <-> kh 0..3."""
    k_of = {(-1, 1): 0, (0, 0): 1, (0, 1): 2, (1, 0): 3}
    idx = torch.zeros(3, 3, 16, dtype=torch.long)
    mask = torch.zeros(3, 3, 16)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            for c in range(3):
                for ph in range(2):
                    for pw in range(2):
                        if (di, ph) in k_of and (dj, pw) in k_of:
                            idx[di + 1, dj + 1, (c * 2 + ph) * 2 + pw] = c * 16 + k_of[(di, ph)] * 4 + k_of[(dj, pw)]
                            mask[di + 1, dj + 1, (c * 2 + ph) * 2 + pw] = 1.0
    return idx.reshape(-1), mask.reshape(-1)


class _Encode16(nn.Sequential):
    """model.py:832-850 encode_image_by_16times; children keep indices 0..10.  The first conv (Cin = 3, HBM-bound)
    runs on the tcgen05 kernel as a 3x3 conv over the space-to-depth image (12 -> 16 channels, 4x fewer pixels) with
    LeakyReLU fused into its epilogue; `x` may be a tuple of image batches (real, wrong, fake) which are gathered by
    the space-to-depth kernel instead of torch.cat.  Widths that are not multiples of 32 use the SIMT conv."""

    def __init__(self, ndf):
        super().__init__(
            nn.Conv2d(3, ndf, 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 2), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(ndf * 2, ndf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 4), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 8), nn.LeakyReLU(0.2, inplace=True))
        self._tc0 = ndf % 32 == 0
        if self._tc0:
            self._s0 = ops.ConvSpec(ops.S1, 16, ndf, impl=L.IMPL_TC, act=ops.ACT_LRELU)
            self._s0.ref = (3, ndf, 16)              # conv4x4 s2 over 3 channels, counted on its own output grid
            self._s2d = None
        else:
            self._s0 = ops.ConvSpec(ops.DOWN2, 3, ndf, impl=L.IMPL_SIMT, x_fmt=L.FMT_NCHW_F32, act=ops.ACT_LRELU)
        self._s = [_spec(ops.DOWN2, ndf, ndf * 2), _spec(ops.DOWN2, ndf * 2, ndf * 4), _spec(ops.DOWN2, ndf * 4, ndf * 8)]

    def forward(self, x, groups=1):
        imgs = list(x) if isinstance(x, (list, tuple)) else [x]
        w = self[0].weight
        if self._tc0:
            if self._s2d is None or self._s2d[0].device != w.device:
                self._s2d = tuple(t.to(w.device) for t in _s2d_filter_index())
            idx, mask = self._s2d
            n = w.shape[0]
            w2 = (w.reshape(n, 48)[:, idx] * mask).view(n, 3, 3, 16).permute(0, 3, 1, 2)     # [n,16,3,3], KRSC memory
            h, _ = ops.conv(ops.img_s2d(*imgs), w2, self._s0)                                 # LeakyReLU in the epilogue
        else:
            x = imgs[0] if len(imgs) == 1 else torch.cat(imgs, 0)
            if x.dtype != torch.float32 or not x.is_contiguous():
                x = x.float().contiguous()
            y, _ = ops.conv(x, w, self._s0)
            h = ops.lrelu_from_out(y)
        for i, (ci, bi) in enumerate(((2, 3), (5, 6), (8, 9))):
            h = _conv_bn_act(h, self[ci], self[bi], self._s[i], ops.ACT_LRELU, groups)
        return to_public(h)


def encode_image_by_16times(ndf):
    return _Encode16(ndf)


def _head_logit(head, feat_nhwc):
    """nn.Conv2d(8ndf, 1, 4, stride 4) + Sigmoid on a 4x4 map == one 8192-long dot per sample (model.py:886-888)."""
    B = feat_nhwc.shape[0]
    w = head[0].weight.permute(0, 2, 3, 1).reshape(-1)            # (kh, kw, c) order of the NHWC feature map
    return torch.sigmoid(feat_nhwc.reshape(B, -1).float() @ w + head[0].bias)


class _DBase(nn.Module):
    def _trunk(self, x_var, groups):
        # ops.grad_mark: no-op unless a tail-first gradient all-reduce is registered for this network (parallel.py)
        x = ops.grad_mark(self.img_code_s16(x_var, groups), self, self.img_code_s16)
        if hasattr(self, "img_code_s32"):
            x = ops.grad_mark(self.img_code_s32(x, groups), self, self.img_code_s32)
        if hasattr(self, "img_code_s64"):
            x = ops.grad_mark(self.img_code_s64(x, groups), self, self.img_code_s64)
            x = ops.grad_mark(self.img_code_s64_1(x, groups), self, self.img_code_s64_1)
            x = self.img_code_s64_2(x, groups)
        elif hasattr(self, "img_code_s32_1"):
            x = self.img_code_s32_1(x, groups)
        return to_nhwc(x)

    def _make_trunk(self, ndf, res):
        self.img_code_s16 = encode_image_by_16times(ndf)
        if res >= 128:
            self.img_code_s32 = downBlock(ndf * 8, ndf * 16)
        if res >= 256:
            self.img_code_s64 = downBlock(ndf * 16, ndf * 32)
            self.img_code_s64_1 = Block3x3_leakRelu(ndf * 32, ndf * 16)
            self.img_code_s64_2 = Block3x3_leakRelu(ndf * 16, ndf * 8)
        elif res >= 128:
            self.img_code_s32_1 = Block3x3_leakRelu(ndf * 16, ndf * 8)

    def _cond_logit(self, x_code, c_code, groups):
        c = c_code.view(-1, self.ef_dim)
        if groups > 1:
            c = c.repeat(groups, 1)
        h_c = self.jointConv(to_public(ops.cat_code(c, x_code)), groups)
        return _head_logit(self.logits, to_nhwc(h_c))


def _logit_head(ndf):
    return nn.Sequential(nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=4), nn.Sigmoid())


_FC_TRANSPOSED = os.environ.get("EKL_FC_T", "1") != "0"


class _JointD(_DBase):
    """model.py:918-977 / 1054-1121 / 1206-1257 JOINT_D_NET{64,128,256}: returns [match[B], real[B], cp[B,E+1]].
    `groups` > 1 = several equally-sized batches stacked along dim 0 with independent BatchNorm statistics (the
    reference's separate real / wrong / fake calls, cub_trainer_splitz_cap_ca.py:418-420) in one pass."""
    RES = 64

    def __init__(self, use_cap=False):
        super().__init__()
        self.df_dim = cfg.GAN.DF_DIM
        if self.RES == 256:
            self.ef_dim = cfg.GAN.EMBEDDING_DIM                        # ignores CAT_Z (model.py:1210)
            use_cap = False
        else:
            self.ef_dim = cfg.GAN.EMBEDDING_DIM * 2 if cfg.TRAIN.CAT_Z == "concat" else cfg.GAN.EMBEDDING_DIM
        self.entity_num = cfg.GAN.ENTITY_DIM
        self.use_cap = use_cap
        ndf = self.df_dim
        self._make_trunk(ndf, self.RES)
        self.jointConv = Block3x3_leakRelu(ndf * 8 + self.ef_dim, ndf * 8)
        self.logits = _logit_head(ndf)
        if use_cap:
            self.fc_ac_cap = nn.Sequential(
                capsule.CapsuleLinear(out_capsules=self.entity_num + 1, in_length=ndf * 8, out_length=16, in_capsules=None))
        else:
            self.fc_ac = nn.Linear(ndf * 8 * 4 * 4, self.entity_num + 1)
        self.uncond_logits = _logit_head(ndf)

    def heads_raw(self, x_var, c_code, groups=1):
        """(match logit [B], uncond logit [B], class logits [B,E+1]) BEFORE sigmoid / log_softmax: what the fused
        loss kernel consumes (engine.StepEngine); forward() applies the reference's output activations."""
        x_code = self._trunk(x_var, groups)                            # [B,4,4,8ndf] NHWC bf16
        c = c_code.view(-1, self.ef_dim)
        if groups > 1:
            c = c.repeat(groups, 1)
        h_c = to_nhwc(self.jointConv(to_public(ops.cat_code(c, x_code)), groups))
        lu, lm = ops.dhead_dots(x_code, h_c, self.uncond_logits[0].weight, self.uncond_logits[0].bias,
                                self.logits[0].weight, self.logits[0].bias)
        B = x_code.shape[0]
        if self.use_cap:
            cls = self.fc_ac_cap[0].forward_norm(x_code.reshape(B, 16, self.df_dim * 8).float())   # == permute(0,2,3,1).view
            #                                                                  (:967-968), then .norm(dim=-1) (:969-970)
        else:
            flat = x_code.permute(0, 3, 1, 2).reshape(B, -1).float()                 # NCHW flatten order (:974)
            if flat.is_cuda and _FC_TRANSPOSED:
                # the same Linear as W x^T: its [E+1, B] result and both gradient GEMMs have 16-byte aligned leading
                # dimensions, whereas [B, 201] (ld 201) sends cuBLAS to its unaligned 64x64 kernel -- 43-89 us per call on
                # the discriminator branch's critical path
                cls = torch.addmm(self.fc_ac.bias.unsqueeze(1), self.fc_ac.weight, flat.t()).t()
            else:
                cls = self.fc_ac(flat)
        return lm, lu, cls

    def forward(self, x_var, c_code, groups=1):
        lm, lu, cls = self.heads_raw(x_var, c_code, groups)
        return [torch.sigmoid(lm), torch.sigmoid(lu), F.log_softmax(cls, dim=1)]


class JOINT_D_NET64(_JointD):
    RES = 64


class JOINT_D_NET128(_JointD):
    RES = 128


class JOINT_D_NET256(_JointD):
    RES = 256

    def __init__(self):
        super().__init__(use_cap=False)


class _PlainD(_DBase):
    """model.py:874-914 / 1006-1050 / 1154-1202 D_NET{64,128,256}: returns [cond[B], uncond[B]] ([cond] if unconditioned)."""
    RES = 64

    def __init__(self):
        super().__init__()
        self.df_dim = cfg.GAN.DF_DIM
        self.ef_dim = cfg.GAN.EMBEDDING_DIM
        ndf = self.df_dim
        self._make_trunk(ndf, self.RES)
        self.logits = _logit_head(ndf)
        if cfg.GAN.B_CONDITION:
            self.jointConv = Block3x3_leakRelu(ndf * 8 + self.ef_dim, ndf * 8)
            self.uncond_logits = _logit_head(ndf)

    def forward(self, x_var, c_code=None, groups=1):
        x_code = self._trunk(x_var, groups)
        if cfg.GAN.B_CONDITION and c_code is not None:
            output = self._cond_logit(x_code, c_code, groups)
        else:
            output = _head_logit(self.logits, x_code)
        if cfg.GAN.B_CONDITION:
            return [output.view(-1), _head_logit(self.uncond_logits, x_code).view(-1)]
        return [output.view(-1)]


class D_NET64(_PlainD):
    RES = 64


class D_NET128(_PlainD):
    RES = 128


class D_NET256(_PlainD):
    RES = 256


def to_kernel_layout(net):
    """Store every 4-D conv filter channels_last ([Cout][KH][KW][Cin] in memory): coalesced for the pack and
    weight-gradient kernels.  Shapes, state_dict keys and values are unchanged."""
    for p in net.parameters():
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
            if p.grad is not None:
                p.grad = p.grad.contiguous(memory_format=torch.channels_last)
    return net
