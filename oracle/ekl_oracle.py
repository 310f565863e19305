"""CPU oracle for the G+D training step -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A functional torch-fp32 restatement of the reference arithmetic.  Networks are flat
``{state_dict key: tensor}`` dicts whose key names and shapes are exactly those of the reference
modules, so weights move between the reference, this oracle and the CUDA implementation by name.
Every function cites the reference lines it follows (paths relative to /root/reference;
``cub:`` = cub_trainer_splitz_cap_ca.py).

Parity: pinned against the real reference by tests/test_oracle_golden.py (fixtures made by
oracle/gen_golden.py), except capsule routing which is "parity unpinned" (oracle/capsule_ref.py).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import capsule_ref

BN_EPS = 1e-5      # nn.BatchNorm default, model.py:91
BN_MOM = 0.1

# --------------------------------------------------------------------------- storage-precision model
# The reference computes in fp32.  The B200 implementation keeps fp32 accumulation / statistics / parameters but STORES
# feature maps, their gradients and the conv filter operands in bf16.  GAN gradients are chaotic w.r.t. such 2^-9
# perturbations (tests/test_bf16_sensitivity.py measures it on this fp32 oracle), so gradient parity of the CUDA path is
# judged against this same oracle evaluated with rounding at exactly the tensors the implementation stores in bf16
# (STORAGE = "bf16"); forward outputs and losses are judged against the plain fp32 oracle (STORAGE = "fp32").
STORAGE = "fp32"


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _q(x):
    """a tensor the implementation stores in bf16 (value and its gradient)."""
    return _RoundBF16.apply(x) if STORAGE == "bf16" else x


def _qw(w):
    """a conv filter operand (bf16 copy of the fp32 master; the gradient stays fp32)."""
    return w + (w.bfloat16().float() - w).detach() if STORAGE == "bf16" else w


class storage:
    """with storage("bf16"): ... -- evaluate the oracle at the implementation's storage precision."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        global STORAGE
        self.prev, STORAGE = STORAGE, self.mode

    def __exit__(self, *a):
        global STORAGE
        STORAGE = self.prev


# nn.Module.train() / .eval() of the reference networks: evaluate() builds netG and calls netG.eval() when
# cfg.TEST.EVAL_MODE (cub:789-793), so BatchNorm normalises with the running statistics and updates nothing.
TRAINING = True


class eval_mode:
    """with eval_mode(): ... -- the oracle's networks in nn.Module.eval() state (generation path, cub:776-911)."""

    def __enter__(self):
        global TRAINING
        self.prev, TRAINING = TRAINING, False

    def __exit__(self, *a):
        global TRAINING
        TRAINING = self.prev


@dataclass
class OracleCfg:
    """The subset of miscc/config.py:13-77 the step reads."""
    BRANCH_NUM: int = 2            # TREE.BRANCH_NUM
    SCALE: int = 2                 # TREE.SCALE
    BATCH_SIZE: int = 32           # TRAIN.BATCH_SIZE
    CAT_Z: str = "concat"          # TRAIN.CAT_Z
    EXCHANGE: bool = False         # TRAIN.EXCHANGE
    G_CAPSULE: bool = False
    D_CAPSULE: bool = False
    KL: float = 2.0                # TRAIN.COEFF.KL
    UNCOND_LOSS: float = 1.0       # TRAIN.COEFF.UNCOND_LOSS
    LR_G: float = 2e-4
    LR_D: float = 2e-4
    EMBEDDING_DIM: int = 128       # GAN.EMBEDDING_DIM
    DF_DIM: int = 64
    GF_DIM: int = 64
    Z_DIM: int = 100
    R_NUM: int = 2
    B_CONDITION: bool = True
    ENTITY_DIM: int = 200
    MANIFD_DIM: int = 128
    TEXT_DIM: int = 1024           # TEXT.DIMENSION
    # which generator assembly / step flavour (SURVEY 8 "config resolution")
    G_KIND: str = "catz_ca"        # catz_ca | catz | cond | gnet
    COND: str = "txt+cls"          # cond-G input: 'txt+cls' (trainer.py:525) or 'txt' (cub:571)
    CLS_KIND: str = "index"        # 'index' (birds, cub:303-304,556-557) | 'multihot' (coco, trainer.py:518)
    ROUTING: str = "dynamic"
    ROUTING_ITERS: int = 3

    def g_ef_dim(self):            # model.py:383-389
        if self.B_CONDITION:
            return self.EMBEDDING_DIM * 2 if self.CAT_Z == "concat" else self.EMBEDDING_DIM
        return self.Z_DIM

    def d_ef_dim(self, res):       # model.py:922-924,1058-1060 ; JOINT_D_NET256 ignores CAT_Z (:1210)
        if res == 256:
            return self.EMBEDDING_DIM
        return self.EMBEDDING_DIM * 2 if self.CAT_Z == "concat" else self.EMBEDDING_DIM


# --------------------------------------------------------------------------- building blocks
def glu(x):
    """model.py:68-76: first half * sigmoid(second half) along dim 1."""
    nc = x.size(1) // 2
    return x[:, :nc] * torch.sigmoid(x[:, nc:])


def _bn(x, sd, p, training=None):
    """nn.BatchNorm{1,2}d: train mode (batch statistics, running stats updated in place) or, under eval_mode(), the
    running statistics."""
    training = TRAINING if training is None else training
    rm, rv = sd.get(p + ".running_mean"), sd.get(p + ".running_var")
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training, BN_MOM, BN_EPS)
    if training and (p + ".num_batches_tracked") in sd:
        sd[p + ".num_batches_tracked"] += 1
    return y


def up_block(x, sd, p):
    """model.py:87-94: nearest x2 -> conv3x3 (no bias) -> BN -> GLU.  Sequential idx 1,2."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = _q(F.conv2d(x, _qw(sd[p + ".1.weight"]), padding=1))
    return _q(glu(_bn(x, sd, p + ".2")))


def block3x3_glu(x, sd, p):
    """model.py:98-104 Block3x3_relu: conv3x3 -> BN -> GLU.  Sequential idx 0,1."""
    x = _q(F.conv2d(x, _qw(sd[p + ".0.weight"]), padding=1))
    return _q(glu(_bn(x, sd, p + ".1")))


def res_block(x, sd, p):
    """model.py:107-123: conv-BN-GLU-conv-BN + identity.  block idx 0,1,3,4."""
    y = _q(F.conv2d(x, _qw(sd[p + ".block.0.weight"]), padding=1))
    y = _q(glu(_bn(y, sd, p + ".block.1")))
    y = _q(F.conv2d(y, _qw(sd[p + ".block.3.weight"]), padding=1))
    y = _bn(y, sd, p + ".block.4")
    return _q(y + x)


def ca_net(text, sd, p, cfg, eps):
    """model.py:126-157 CA_NET: fc(+bias) -> GLU -> (mu, logvar); c = eps*exp(0.5 logvar)+mu."""
    x = glu(F.linear(text, sd[p + ".fc.weight"], sd[p + ".fc.bias"]))
    ef = cfg.EMBEDDING_DIM
    mu, logvar = x[:, :ef], x[:, ef:]
    std = torch.exp(0.5 * logvar)
    return eps * std + mu, mu, logvar, std


def vc_net(noise, cond, sd, p, seed):
    """model.py:160-201 VC_NET: cat(noise,cond) -> fc1-BN-ReLU -> fc2-BN-ReLU -> fc31/fc32; reparam."""
    x = torch.cat((noise, cond), 1)
    h = F.relu(_bn(F.linear(x, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]), sd, p + ".bn_fc1"))
    h = F.relu(_bn(F.linear(h, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"]), sd, p + ".bn_fc2"))
    mu = F.linear(h, sd[p + ".fc31.weight"], sd[p + ".fc31.bias"])
    logvar = F.linear(h, sd[p + ".fc32.weight"], sd[p + ".fc32.bias"])
    std = torch.exp(0.5 * logvar)
    return seed * std + mu, mu, logvar, std


def _up4(x, sd, p):
    for i in (1, 2, 3, 4):          # model.py:227-233
        x = up_block(x, sd, "%s.upsample%d" % (p, i))
    return x


def init_stage_fc(code, sd, p, ngf):
    """model.py:204-235 COND_INIT_STAGE_G / :336-376 INIT_STAGE_G: Linear(no bias)-BN1d-GLU-view-4 upBlocks."""
    x = _q(F.linear(code, sd[p + ".fc.0.weight"]))
    x = _q(glu(_bn(x, sd, p + ".fc.1")))
    return _up4(x.view(-1, ngf, 4, 4), sd, p)


def init_stage_cap(z, noise, sd, p, ngf, cfg):
    """model.py:238-277 COND_INIT_STAGE_G_withCap: cat(z,noise) -> [bs,-1,8] -> CapsuleLinear(ngf caps, len 32)
    -> [-1, ngf*32] -> BN1d -> GLU -> view [B,ngf,4,4] -> 4 upBlocks.  fc_cap idx 1 (capsule), 3 (BN)."""
    if noise is not None:
        z = torch.cat((z, noise), 1)
    x = z.view(cfg.BATCH_SIZE, -1, 8)
    x = capsule_ref.capsule_linear(x, sd[p + ".fc_cap.1.weight"], cfg.ROUTING, cfg.ROUTING_ITERS)
    x = _q(x.reshape(-1, ngf * 4 * 4 * 2))
    x = _q(glu(_bn(x, sd, p + ".fc_cap.3")))
    return _up4(x.view(-1, ngf, 4, 4), sd, p)


def init_stage_exchange_cap(z, sd, p, ngf, cfg):
    """model.py:280-333 COND_INIT_STAGE_G_Exchange_Cap: two capsule stems (ngf caps of len 16) on the z halves."""
    half = cfg.MANIFD_DIM
    outs = []
    for zz, q in ((z[:, :half].contiguous(), ".fc_cap"), (z[:, half:].contiguous(), ".fc_cap1")):
        x = zz.view(cfg.BATCH_SIZE, -1, 8)
        x = capsule_ref.capsule_linear(x, sd[p + q + ".1.weight"], cfg.ROUTING, cfg.ROUTING_ITERS)
        x = _q(x.reshape(-1, (ngf // 2) * 4 * 4 * 2))
        x = _q(glu(_bn(x, sd, p + q + ".3")))
        outs.append(x.view(-1, ngf // 2, 4, 4))
    return _up4(torch.cat(outs, 1), sd, p)


def next_stage(h, c, sd, p, cfg):
    """model.py:379-423 NEXT_STAGE_G: tile c, cat((c,h)), jointConv, R_NUM ResBlocks, upBlock(s)."""
    s = h.size(2)
    cc = _q(c).view(-1, c.size(1), 1, 1).repeat(1, 1, s, s)
    x = block3x3_glu(torch.cat((cc, h), 1), sd, p + ".jointConv")
    for i in range(cfg.R_NUM):
        x = res_block(x, sd, "%s.residual.%d" % (p, i))
    x = up_block(x, sd, p + ".upsample")
    if cfg.SCALE == 4:
        x = up_block(x, sd, p + ".upsample2")
    return x


def get_image(h, sd, p):
    """model.py:426-437 GET_IMAGE_G: conv3x3(ngf->3) + tanh."""
    return torch.tanh(_q(F.conv2d(h, _qw(sd[p + ".img.0.weight"]), padding=1)))


def _stages(c_code, h1, sd, cfg):
    hs = [h1]
    if cfg.BRANCH_NUM > 1:
        hs.append(next_stage(hs[-1], c_code, sd, "h_net2", cfg))
    if cfg.BRANCH_NUM > 2:
        hs.append(next_stage(hs[-1], c_code, sd, "h_net3", cfg))
    return hs


def g_forward_catz_ca(sd, cfg, noise, sen, cls, eps, seed):
    """model.py:482-527 COND_G_NET_CATZ_CA.forward (training branch)."""
    c1, mu1, lv1, std1 = ca_net(sen, sd, "ca_net1", cfg, eps)
    c2, mu2, lv2, std2 = vc_net(noise, cls, sd, "vc_net2", seed)
    if cfg.CAT_Z == "concat" or cfg.EXCHANGE:
        c = torch.cat((c1, c2), 1)
    elif cfg.CAT_Z == "product":
        c = c1 * c2
    else:
        c = c1 + c2
    ngf = cfg.GF_DIM * 16
    if cfg.G_CAPSULE and cfg.EXCHANGE:
        h1 = init_stage_exchange_cap(c, sd, "h_net1", ngf, cfg)
    elif cfg.G_CAPSULE:
        h1 = init_stage_cap(c, noise, sd, "h_net1", ngf, cfg)        # model.py:512
    else:
        raise TypeError("reference defect: COND_INIT_STAGE_G.forward takes one arg (SURVEY app. A #2)")
    return _stages(c, h1, sd, cfg), mu1, mu2, lv1, lv2, std1, std2


def g_forward_catz(sd, cfg, noise, sen, cls, seed1, seed2):
    """model.py:592-628 COND_G_NET_CATZ.forward: both codes come from VC_NETs; h_net1 takes c_code alone (:614)."""
    c1, mu1, lv1, std1 = vc_net(noise, sen, sd, "vc_net1", seed1)
    c2, mu2, lv2, std2 = vc_net(noise, cls, sd, "vc_net2", seed2)
    if cfg.EXCHANGE or cfg.CAT_Z == "concat":
        c = torch.cat((c1, c2), 1)
    elif cfg.CAT_Z == "product":
        c = c1 * c2
    else:
        c = c1 + c2
    ngf = cfg.GF_DIM * 16
    if cfg.G_CAPSULE and cfg.EXCHANGE:
        h1 = init_stage_exchange_cap(c, sd, "h_net1", ngf, cfg)
    elif cfg.G_CAPSULE:
        h1 = init_stage_cap(c, None, sd, "h_net1", ngf, cfg)
    else:
        h1 = init_stage_fc(c, sd, "h_net1", ngf)
    return _stages(c, h1, sd, cfg), mu1, mu2, lv1, lv2, std1, std2


def g_forward_cond(sd, cfg, noise, cond, seed):
    """model.py:687-708 COND_G_NET.forward."""
    c, mu, lv, std = vc_net(noise, cond, sd, "vc_net", seed)
    ngf = cfg.GF_DIM * 16
    if cfg.G_CAPSULE:
        h1 = init_stage_cap(c, None, sd, "h_net1", ngf, cfg)
    else:
        h1 = init_stage_fc(c, sd, "h_net1", ngf)
    return _stages(c, h1, sd, cfg), mu, lv, std


def g_forward_gnet(sd, cfg, z, text, eps):
    """model.py:747-790 G_NET with its sub-modules called explicitly (forward itself is broken, app. A #1):
    CA_NET -> INIT_STAGE_G(cat(c, z)) (model.py:359-361) -> NEXT_STAGE_G..."""
    c, mu, lv, std = ca_net(text, sd, "ca_net", cfg, eps)
    h1 = init_stage_fc(torch.cat((c, z), 1), sd, "h_net1", cfg.GF_DIM * 16)
    return _stages(c, h1, sd, cfg), mu, lv, std


def g_images(hs, sd):
    """model.py:547-563 .image(hcodes)."""
    return [get_image(h, sd, "img_net%d" % (i + 1)) for i, h in enumerate(hs)]


# --------------------------------------------------------------------------- discriminators
def _down(x, sd, pconv, pbn=None):
    """conv4x4 s2 p1 no bias (+BN) + LeakyReLU(0.2): model.py:822-850."""
    x = F.conv2d(x, _qw(sd[pconv + ".weight"]), stride=2, padding=1)
    if pbn is not None:
        x = _bn(_q(x), sd, pbn)
    return _q(F.leaky_relu(x, 0.2))


def _b3_lrelu(x, sd, p):
    """model.py:812-818 Block3x3_leakRelu."""
    x = _q(F.conv2d(x, _qw(sd[p + ".0.weight"]), padding=1))
    return _q(F.leaky_relu(_bn(x, sd, p + ".1"), 0.2))


def d_trunk(x, sd, res):
    """encode_image_by_16times (model.py:832-850, Sequential idx 0 | 2,3 | 5,6 | 8,9) + per-resolution tail
    (model.py:1097-1098, 1238-1242)."""
    p = "img_code_s16"
    x = _down(_q(x), sd, p + ".0")
    x = _down(x, sd, p + ".2", p + ".3")
    x = _down(x, sd, p + ".5", p + ".6")
    x = _down(x, sd, p + ".8", p + ".9")
    if res >= 128:
        x = _down(x, sd, "img_code_s32.0", "img_code_s32.1")
    if res >= 256:
        x = _down(x, sd, "img_code_s64.0", "img_code_s64.1")
        x = _b3_lrelu(x, sd, "img_code_s64_1")
        x = _b3_lrelu(x, sd, "img_code_s64_2")
    elif res >= 128:
        x = _b3_lrelu(x, sd, "img_code_s32_1")
    return x


def _cond_logit(x_code, c_code, sd):
    """tile c, cat((c,x)), jointConv, conv4x4/s4(+bias)+sigmoid: model.py:956-962."""
    cc = _q(c_code).view(-1, c_code.size(1), 1, 1).repeat(1, 1, 4, 4)
    h = _b3_lrelu(torch.cat((cc, x_code), 1), sd, "jointConv")
    return torch.sigmoid(F.conv2d(h, sd["logits.0.weight"], sd["logits.0.bias"], stride=4)).view(-1)


def d_joint_forward(x, c_code, sd, cfg, res, use_cap):
    """model.py:953-977 / 1095-1121 / 1237-1257 JOINT_D_NET{64,128,256}.forward -> [match, real, cp]."""
    x_code = d_trunk(x, sd, res)
    match = _cond_logit(x_code, c_code, sd)
    real = torch.sigmoid(F.conv2d(x_code, sd["uncond_logits.0.weight"], sd["uncond_logits.0.bias"], stride=4)).view(-1)
    if use_cap:
        xc = x_code.permute(0, 2, 3, 1).contiguous().view(-1, 16, cfg.DF_DIM * 8)     # model.py:967-968
        out = capsule_ref.capsule_linear(xc, sd["fc_ac_cap.0.weight"], cfg.ROUTING, cfg.ROUTING_ITERS)
        cp = F.log_softmax(out.norm(dim=-1), dim=1)
    else:
        flat = x_code.reshape(-1, cfg.DF_DIM * 8 * 16)
        cp = F.log_softmax(F.linear(flat, sd["fc_ac.weight"], sd["fc_ac.bias"]), dim=1)
    return [match, real, cp]


def d_plain_forward(x, c_code, sd, cfg, res):
    """model.py:896-914 / 1030-1050 / 1180-1202 D_NET{64,128,256}.forward -> [cond, uncond]."""
    x_code = d_trunk(x, sd, res)
    if c_code is not None:
        out = _cond_logit(x_code, c_code, sd)
    else:
        out = torch.sigmoid(F.conv2d(x_code, sd["logits.0.weight"], sd["logits.0.bias"], stride=4)).view(-1)
    unc = torch.sigmoid(F.conv2d(x_code, sd["uncond_logits.0.weight"], sd["uncond_logits.0.bias"], stride=4)).view(-1)
    return [out, unc]


# --------------------------------------------------------------------------- losses
def kl_loss(mu, logvar):
    """cub:54-58: -0.5 * mean(1 + logvar - mu^2 - exp(logvar))."""
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


def ce_loss(logq, p):
    """cub:60-65: -sum(p*logq)/B."""
    return -torch.sum(p * logq) / p.shape[0]


def mean_covariance(img):
    """cub:33-52 compute_mean_covariance."""
    b, c, h, w = img.shape
    mu = img.mean(2, keepdim=True).mean(3, keepdim=True)
    d = (img - mu).view(b, c, h * w)
    return mu, torch.bmm(d, d.transpose(1, 2)) / (h * w)


def onehot(cls, n):
    """cub:322-331."""
    out = torch.zeros(cls.shape[0], n, device=cls.device)
    out[torch.arange(cls.shape[0], device=cls.device), cls] = 1
    return out


def bce(p, target_value):
    """nn.BCELoss (mean, log clamped at -100) against a constant target (cub:423-431).  Evaluated in fp32 outside any
    autocast region (torch refuses BCE on autocast inputs; bench.py's bf16-autocast eager baseline runs this port)."""
    with torch.autocast(device_type=p.device.type, enabled=False):
        p = p.float()
        return F.binary_cross_entropy(p, torch.full_like(p, float(target_value)))


def d_loss(real, wrong, fake, real_cp, fake_cp, cfg):
    """cub:423-450 / trainer.py:396-422."""
    e_match = bce(real[0], 1) + bce(wrong[0], 0) + bce(fake[0], 0)
    if len(real) > 1 and cfg.UNCOND_LOSS > 0:
        u = cfg.UNCOND_LOSS
        e_unc = u * bce(real[1], 1) + u * bce(wrong[1], 1) + u * bce(fake[1], 0)     # wrong -> REAL (cub:430)
        if len(real) > 2:
            e_cls = ce_loss(real[2], real_cp) + ce_loss(fake[2], fake_cp)
        else:       # two-head D_NET* (model.py:874-914): no class head; the reference has no loss assembly for these
            e_cls = torch.zeros((), device=real[0].device)      # (app. A #14) -- match + uncond, trainer.py:408-410
        return e_match + e_unc + e_cls, e_match, e_unc, e_cls
    z = torch.zeros((), device=real[0].device)
    return bce(real[0], 1) + 0.5 * (bce(wrong[0], 0) + bce(fake[0], 0)), e_match, z, z


# --------------------------------------------------------------------------- whole step
def _leaves(sd):
    out = {}
    for k, v in sd.items():
        if torch.is_floating_point(v) and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
            out[k] = v
    return out


D_RES = (64, 128, 256)


@dataclass
class OracleTrainer:
    """State + one training step, restating cub:493-608 (G_KIND 'catz_ca') and trainer.py:464-545 ('cond')."""
    cfg: OracleCfg
    sdG: Dict[str, torch.Tensor]
    sdDs: List[Dict[str, torch.Tensor]]
    d_res: List[int] = field(default_factory=list)
    d_joint: bool = True

    def __post_init__(self):
        c = self.cfg
        if not self.d_res:
            self.d_res = [64, 128 if c.SCALE == 2 else 256, 256][: len(self.sdDs)]      # cub:144-154
        self.pG = _leaves(self.sdG)
        self.pDs = [_leaves(sd) for sd in self.sdDs]
        self.optG = torch.optim.Adam(list(self.pG.values()), lr=c.LR_G, betas=(0.5, 0.999))   # cub:199-215
        self.optDs = [torch.optim.Adam(list(p.values()), lr=c.LR_D, betas=(0.5, 0.999)) for p in self.pDs]

    def D(self, i, x, c_code):
        c = self.cfg
        if self.d_joint:
            use_cap = c.D_CAPSULE and self.d_res[i] != 256
            return d_joint_forward(x, c_code, self.sdDs[i], c, self.d_res[i], use_cap)
        return d_plain_forward(x, c_code, self.sdDs[i], c, self.d_res[i])

    def step(self, imgs, wrong_imgs, embedding, cls, noise, eps=None, seed=None):
        """One hot-loop iteration.  `cls` is the loader's tensor (1-based int64 [B] for birds, float
        multi-hot [B,E+1] for coco); noise/eps/seed are the injected RNG draws."""
        c = self.cfg
        out = {}
        E = c.ENTITY_DIM
        B = embedding.shape[0]
        # (0) prepare_data + label tensors: cub:295-331,556-557 | trainer.py:518
        if c.CLS_KIND == "index":
            cls0 = cls.long() - 1
            cls_onehot, real_cp = onehot(cls0, E), onehot(cls0, E + 1)
            cls_multi = real_cp
        else:
            cls_multi = cls.float()
            real_cp = cls_multi / cls_multi.sum(1).view(-1, 1)
            cls_onehot = cls_multi
        fake_cp = torch.zeros(B, E + 1, device=embedding.device)
        fake_cp[:, -1] = 1                                                                # cub:520-521
        out["real_cp"], out["fake_cp"], out["cls_onehot"] = real_cp, fake_cp, cls_onehot
        # (1) generate: cub:567-587
        if c.G_KIND in ("catz_ca", "catz"):
            if c.G_KIND == "catz":          # eps plays the role of the sentence VC_NET's seed (first host draw)
                hs, mu1, mu2, lv1, lv2, std1, std2 = g_forward_catz(self.sdG, c, noise, embedding, cls_onehot, eps, seed)
            else:
                hs, mu1, mu2, lv1, lv2, std1, std2 = g_forward_catz_ca(self.sdG, c, noise, embedding, cls_onehot, eps, seed)
            mu = torch.cat((mu1, mu2), 1) if c.CAT_Z == "concat" else (mu1 * mu2 if c.CAT_Z == "product" else mu1 + mu2)
            kls = [(mu1, lv1), (mu2, lv2)]
        elif c.G_KIND == "cond":
            cond = torch.cat((embedding, cls_multi), 1) if c.COND == "txt+cls" else embedding   # trainer.py:525 | cub:571
            hs, mu, lv, std = g_forward_cond(self.sdG, c, noise, cond, seed)
            kls = [(mu, lv)]
        else:
            raise ValueError(c.G_KIND)
        fakes = g_images(hs, self.sdG)
        out["h_codes"], out["fake_imgs"], out["mu"] = hs, fakes, mu
        # (2) D updates: cub:404-461
        out["errD"], out["d_logits"], out["gradD"] = [], [], []
        for i in range(len(self.sdDs)):
            self.optDs[i].zero_grad(set_to_none=True)
            real = self.D(i, imgs[i], mu.detach())
            wrong = self.D(i, wrong_imgs[i], mu.detach())
            fake = self.D(i, fakes[i].detach(), mu.detach())
            errD, e_m, e_u, e_c = d_loss(real, wrong, fake, real_cp, fake_cp, c)
            errD.backward()
            out["gradD"].append({k: p.grad.clone() for k, p in self.pDs[i].items() if p.grad is not None})
            self.optDs[i].step()
            out["errD"].append(torch.stack([errD.detach(), e_m.detach(), e_u.detach(), e_c.detach()]))
            out["d_logits"].append([[t.detach() for t in real], [t.detach() for t in wrong], [t.detach() for t in fake]])
        # (3) G update through the UPDATED Ds: cub:463-490,604-608
        self.optG.zero_grad(set_to_none=True)
        e_match = e_unc = e_cls = 0
        out["g_logits"] = []
        for i in range(len(self.sdDs)):
            o = self.D(i, fakes[i], mu)
            e_match = e_match + bce(o[0], 1)
            if len(o) > 1 and c.UNCOND_LOSS > 0:
                e_unc = e_unc + c.UNCOND_LOSS * bce(o[1], 1)
                if len(o) > 2:
                    e_cls = e_cls + ce_loss(o[2], real_cp)
            out["g_logits"].append([t.detach() for t in o])
        kl = [kl_loss(m, l) for m, l in kls]
        errG = e_match + e_unc + e_cls + sum(kl) * c.KL
        errG.backward()
        out["gradG"] = {k: p.grad.clone() for k, p in self.pG.items() if p.grad is not None}
        self.optG.step()
        z = torch.zeros((), device=errG.device)
        out["errG"] = torch.stack([errG.detach(), (e_match + z).detach(), (e_unc + z).detach(), (e_cls + z).detach()]
                                  + [k.detach() for k in kl])
        return out
