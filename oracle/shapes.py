"""state_dict key -> shape specs of the reference networks, restated from the constructors
(model.py file:line cited per block).  Pinned against the real reference by the `shapes/*` entries
of tests/golden/*.npz.  TEST INFRASTRUCTURE."""
from collections import OrderedDict

import torch


def _bn(sd, p, c):
    sd[p + ".weight"] = (c,)
    sd[p + ".bias"] = (c,)
    sd[p + ".running_mean"] = (c,)
    sd[p + ".running_var"] = (c,)
    sd[p + ".num_batches_tracked"] = ()


def _up(sd, p, cin, cout):                      # model.py:87-94
    sd[p + ".1.weight"] = (cout * 2, cin, 3, 3)
    _bn(sd, p + ".2", cout * 2)


def _vc(sd, p, cond_dim, c):                    # model.py:160-174
    def lin(n, o, i):
        sd["%s.%s.weight" % (p, n)] = (o, i)
        sd["%s.%s.bias" % (p, n)] = (o,)
    lin("fc1", 512, cond_dim + c.Z_DIM)
    _bn(sd, p + ".bn_fc1", 512)
    lin("fc2", 256, 512)
    _bn(sd, p + ".bn_fc2", 256)
    lin("fc31", c.MANIFD_DIM, 256)
    lin("fc32", c.MANIFD_DIM, 256)


def _init_stage(sd, p, ngf, c, in_dim, cap, exchange=False):   # model.py:204-376
    if cap and exchange:
        for q in (".fc_cap", ".fc_cap1"):
            sd[p + q + ".1.weight"] = (ngf, 16, 8)
            _bn(sd, p + q + ".3", (ngf // 2) * 32)
    elif cap:
        sd[p + ".fc_cap.1.weight"] = (ngf, 32, 8)
        _bn(sd, p + ".fc_cap.3", ngf * 32)
    else:
        sd[p + ".fc.0.weight"] = (ngf * 32, in_dim)
        _bn(sd, p + ".fc.1", ngf * 32)
    for i, (a, b) in enumerate(((ngf, ngf // 2), (ngf // 2, ngf // 4), (ngf // 4, ngf // 8), (ngf // 8, ngf // 16))):
        _up(sd, "%s.upsample%d" % (p, i + 1), a, b)


def _next_stage(sd, p, ngf, c):                 # model.py:379-407
    ef = c.g_ef_dim()
    sd[p + ".jointConv.0.weight"] = (ngf * 2, ngf + ef, 3, 3)
    _bn(sd, p + ".jointConv.1", ngf * 2)
    for i in range(c.R_NUM):
        q = "%s.residual.%d.block" % (p, i)
        sd[q + ".0.weight"] = (ngf * 2, ngf, 3, 3)
        _bn(sd, q + ".1", ngf * 2)
        sd[q + ".3.weight"] = (ngf, ngf, 3, 3)
        _bn(sd, q + ".4", ngf)
    _up(sd, p + ".upsample", ngf, ngf // 2)
    if c.SCALE == 4:
        _up(sd, p + ".upsample2", ngf // 2, ngf // 4)


def g_shapes(c, kind=None, cond_dim=None):
    """Ordered like the reference's module registration order (model.py:455-480, 669-685, 747-765)."""
    kind = kind or c.G_KIND
    sd = OrderedDict()
    gf = c.GF_DIM
    if kind == "catz_ca":
        sd["ca_net1.fc.weight"] = (c.EMBEDDING_DIM * 4, c.TEXT_DIM)
        sd["ca_net1.fc.bias"] = (c.EMBEDDING_DIM * 4,)
        _vc(sd, "vc_net2", c.ENTITY_DIM, c)
        in_dim = c.MANIFD_DIM * 2 if c.CAT_Z == "concat" else c.MANIFD_DIM
        _init_stage(sd, "h_net1", gf * 16, c, in_dim, c.G_CAPSULE, c.EXCHANGE)
    elif kind == "catz":                           # model.py:567-590 COND_G_NET_CATZ: two VC_NETs
        _vc(sd, "vc_net1", c.TEXT_DIM, c)
        _vc(sd, "vc_net2", c.ENTITY_DIM, c)
        in_dim = c.MANIFD_DIM * 2 if c.CAT_Z == "concat" else c.MANIFD_DIM
        _init_stage(sd, "h_net1", gf * 16, c, in_dim, c.G_CAPSULE, c.EXCHANGE)
    elif kind == "cond":
        _vc(sd, "vc_net", cond_dim, c)
        in_dim = c.MANIFD_DIM * 2 if c.CAT_Z == "concat" else c.MANIFD_DIM
        _init_stage(sd, "h_net1", gf * 16, c, in_dim, c.G_CAPSULE)
    elif kind == "gnet":
        sd["ca_net.fc.weight"] = (c.EMBEDDING_DIM * 4, c.TEXT_DIM)
        sd["ca_net.fc.bias"] = (c.EMBEDDING_DIM * 4,)
        _init_stage(sd, "h_net1", gf * 16, c, c.Z_DIM + c.EMBEDDING_DIM, False)
    sd["img_net1.img.0.weight"] = (3, gf, 3, 3)
    if c.BRANCH_NUM > 1:
        _next_stage(sd, "h_net2", gf, c)
        sd["img_net2.img.0.weight"] = (3, gf // c.SCALE, 3, 3)
    if c.BRANCH_NUM > 2:
        _next_stage(sd, "h_net3", gf // c.SCALE, c)
        sd["img_net3.img.0.weight"] = (3, gf // c.SCALE ** 2, 3, 3)
    return sd


def d_shapes(c, res, joint=True, use_cap=False):
    """model.py:832-850 trunk + :874-1257 heads, in registration order."""
    sd = OrderedDict()
    ndf = c.DF_DIM
    p = "img_code_s16"
    sd[p + ".0.weight"] = (ndf, 3, 4, 4)
    for idx, (a, b) in ((2, (ndf, ndf * 2)), (5, (ndf * 2, ndf * 4)), (8, (ndf * 4, ndf * 8))):
        sd["%s.%d.weight" % (p, idx)] = (b, a, 4, 4)
        _bn(sd, "%s.%d" % (p, idx + 1), b)
    if res >= 128:
        sd["img_code_s32.0.weight"] = (ndf * 16, ndf * 8, 4, 4)
        _bn(sd, "img_code_s32.1", ndf * 16)
    if res >= 256:
        sd["img_code_s64.0.weight"] = (ndf * 32, ndf * 16, 4, 4)
        _bn(sd, "img_code_s64.1", ndf * 32)
        sd["img_code_s64_1.0.weight"] = (ndf * 16, ndf * 32, 3, 3)
        _bn(sd, "img_code_s64_1.1", ndf * 16)
        sd["img_code_s64_2.0.weight"] = (ndf * 8, ndf * 16, 3, 3)
        _bn(sd, "img_code_s64_2.1", ndf * 8)
    elif res >= 128:
        sd["img_code_s32_1.0.weight"] = (ndf * 8, ndf * 16, 3, 3)
        _bn(sd, "img_code_s32_1.1", ndf * 8)
    ef = c.d_ef_dim(res) if joint else c.EMBEDDING_DIM

    def _logits(name):
        sd[name + ".0.weight"] = (1, ndf * 8, 4, 4)
        sd[name + ".0.bias"] = (1,)

    def _joint():
        sd["jointConv.0.weight"] = (ndf * 8, ndf * 8 + ef, 3, 3)
        _bn(sd, "jointConv.1", ndf * 8)
    if joint:
        _joint()
        _logits("logits")
        if use_cap and res != 256:
            sd["fc_ac_cap.0.weight"] = (c.ENTITY_DIM + 1, 16, ndf * 8)
        else:
            sd["fc_ac.weight"] = (c.ENTITY_DIM + 1, ndf * 8 * 16)
            sd["fc_ac.bias"] = (c.ENTITY_DIM + 1,)
        _logits("uncond_logits")
    else:
        _logits("logits")
        if c.B_CONDITION:
            _joint()
            _logits("uncond_logits")
    return sd


def make_state_dict(shapes, tag):
    """Allocate + detfill a state_dict from a shapes spec."""
    from . import detfill
    sd = OrderedDict()
    for k, s in shapes.items():
        sd[k] = torch.zeros(s, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
    with torch.no_grad():
        detfill.fill_state_dict(sd, tag)
    return sd
