"""Drive the REAL reference (/root/reference) on CPU -- build-container only, TEST INFRASTRUCTURE.

Used by oracle/gen_golden.py (to make tests/golden fixtures that pin the oracle) and by bench.py's
`--impl reference` / cpu_baseline legs *when /root/reference exists*.  Nothing here copies reference
code: the reference modules are imported where they lie, behind stand-ins for the packages missing from
this image (oracle/shims: easydict, capsule_layer [parity unpinned], tensorboardX, tensorboard,
tensorflow, inception_score) and with `.cuda()` made a no-op for CPU runs (the reference calls it
unconditionally at model.py:465 and cub_trainer_splitz_cap_ca.py:520).
"""
import contextlib
import copy
import os
import sys

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("EKL_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_state = {}


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "model.py"))


def import_reference():
    """Import reference modules once; returns a namespace dict."""
    if _state:
        return _state
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo, REF_ROOT, _SHIMS):        # shims end up first: they shadow the TF1-only inception_score.py
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import yaml
    if not getattr(yaml, "_ekl_patched", False):            # config.py:119 calls yaml.load(f) without Loader
        _orig = yaml.load
        yaml.load = lambda stream, Loader=None: _orig(stream, Loader=Loader or yaml.SafeLoader)
        yaml._ekl_patched = True
    # CPU: .cuda() no-ops
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.set_device = lambda *a, **k: None
    from miscc import config as ref_config
    _state["config"] = ref_config
    _state["cfg"] = ref_config.cfg
    _state["cfg_defaults"] = copy.deepcopy(ref_config.cfg)
    import model as ref_model
    import cub_trainer_splitz_cap_ca as ref_cub
    import trainer as ref_trainer
    _state.update(model=ref_model, cub=ref_cub, trainer=ref_trainer)
    return _state


def set_cfg(name, batch=None, gf=None, df=None):
    """Reset the reference's global cfg to defaults, merge the yml, apply the documented overrides."""
    from .configs import CONFIGS
    R = import_reference()
    cfg, spec = R["cfg"], CONFIGS[name]

    def _reset(dst, src):
        for k, v in src.items():
            if isinstance(v, dict):
                _reset(dst[k], v)
            else:
                dst[k] = copy.deepcopy(v)
    _reset(cfg, R["cfg_defaults"])
    R["config"].cfg_from_file(os.path.join(REF_ROOT, "cfg", spec["yml"]))
    cfg.CUDA = False
    cfg.TRAIN.BATCH_SIZE = batch or spec["batch"]
    for k, v in spec["over"].items():
        a, b = k.split(".")
        cfg[a][b] = v
    if gf:
        cfg.GAN.GF_DIM = gf
    if df:
        cfg.GAN.DF_DIM = df
    return cfg


@contextlib.contextmanager
def rng_tape(normal_draws, randn_draws):
    """Replay recorded draws: Tensor.normal_() calls pop from normal_draws (cub:567 noise, model.py:148-150
    CA eps), torch.randn calls pop from randn_draws (model.py:192 VC seed)."""
    normal_draws, randn_draws = list(normal_draws), list(randn_draws)
    o_normal, o_randn = torch.Tensor.normal_, torch.randn

    def normal_(self, *a, **k):
        src = normal_draws.pop(0)
        assert tuple(src.shape) == tuple(self.shape), (src.shape, self.shape)
        return self.copy_(src)

    def randn(*size, **k):
        src = randn_draws.pop(0)
        shp = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert tuple(src.shape) == shp, (src.shape, shp)
        return src.clone()
    torch.Tensor.normal_, torch.randn = normal_, randn
    try:
        yield
    finally:
        torch.Tensor.normal_, torch.randn = o_normal, o_randn
        assert not normal_draws and not randn_draws, "unused RNG draws"


def build_nets(name, batch=None, gf=None, df=None):
    """Construct reference G and Ds for a resolved config.  Returns (cfg, netG, netsD)."""
    from .configs import CONFIGS
    R = import_reference()
    cfg = set_cfg(name, batch, gf, df)
    m, spec = R["model"], CONFIGS[name]
    share = m.get_shareGs(cfg.GAN.GF_DIM)
    if spec["ocfg"]["G_KIND"] == "catz_ca":
        netG = m.COND_G_NET_CATZ_CA(cfg.TEXT.DIMENSION, cfg.GAN.ENTITY_DIM, share, use_cap=cfg.TRAIN.G_CAPSULE,
                                    cat=cfg.TRAIN.CAT_Z, exchange=cfg.TRAIN.EXCHANGE)          # cub:130
    elif spec["ocfg"]["G_KIND"] == "catz":
        netG = m.COND_G_NET_CATZ(cfg.TEXT.DIMENSION, cfg.GAN.ENTITY_DIM, share, use_cap=cfg.TRAIN.G_CAPSULE,
                                 cat=cfg.TRAIN.CAT_Z, exchange=cfg.TRAIN.EXCHANGE)             # model.py:567
    else:
        cd = cfg.TEXT.DIMENSION + (cfg.GAN.ENTITY_DIM + 1 if spec["ocfg"]["COND"] == "txt+cls" else 0)
        netG = m.COND_G_NET(cd, share, use_cap=cfg.TRAIN.G_CAPSULE)                            # trainer.py:116 | cub:135
    netsD = [m.JOINT_D_NET64(use_cap=cfg.TRAIN.D_CAPSULE)]
    if cfg.TREE.BRANCH_NUM > 1:                                                              # cub:150-154
        netsD.append(m.JOINT_D_NET128(use_cap=cfg.TRAIN.D_CAPSULE) if cfg.TREE.SCALE == 2 else m.JOINT_D_NET256())
    if cfg.TREE.BRANCH_NUM > 2:
        netsD.append(m.JOINT_D_NET256())
    return cfg, netG, netsD


class RefStepper:
    """Runs the reference's own step functions (train_joint_Dnet / loss_joint_Gnet) on one batch."""

    def __init__(self, name, netG, netsD):
        from .configs import CONFIGS
        R = import_reference()
        self.R, self.cfg, self.spec = R, R["cfg"], CONFIGS[name]
        self.kind = self.spec["ocfg"]["G_KIND"]
        mod = R["cub"] if self.kind in ("catz_ca", "catz") else R["trainer"]
        self.mod = mod
        t = mod.condGANTrainer.__new__(mod.condGANTrainer)      # skip __init__ (mkdirs, set_device)
        cfg = self.cfg
        t.batch_size = cfg.TRAIN.BATCH_SIZE
        t.netG, t.netsD, t.num_Ds = netG, netsD, len(netsD)
        t.optimizerG, t.optimizersD = mod.define_optimizers(netG, netsD)
        t.criterion = nn.BCELoss()
        t.CE = mod.ce_loss
        t.real_labels = torch.ones(t.batch_size)
        t.fake_labels = torch.zeros(t.batch_size)
        t.fake_cp = torch.zeros(t.batch_size, cfg.GAN.ENTITY_DIM + 1)
        t.fake_cp[:, -1] = 1
        t.summary_writer = None
        self.t = t
        self.noise = torch.zeros(t.batch_size, cfg.GAN.Z_DIM)

    def step(self, imgs, wrong_imgs, embedding, cls, noise, eps=None, seed=None):
        t, cfg, out = self.t, self.cfg, {}
        data = ([i.clone() for i in imgs], [i.clone() for i in wrong_imgs], embedding.clone(), cls.clone(), None)
        t.imgs_tcpu, t.real_imgs, t.wrong_imgs, t.txt_embedding, t.cls_label = t.prepare_data(data)
        count = 1            # count % 100 != 0 -> skips the .data[0] logging path (trainer.py:433)
        if self.kind in ("catz_ca", "catz"):
            t.cls_onehot = t.onehot(t.cls_label, cfg.GAN.ENTITY_DIM)
            t.real_cp = t.onehot(t.cls_label, cfg.GAN.ENTITY_DIM + 1)
            # draws: catz_ca = noise.normal_, CA eps (device normal_), VC seed (host randn); catz = noise, two host randn
            tape = rng_tape([noise, eps], [seed]) if self.kind == "catz_ca" else rng_tape([noise], [eps, seed])
            with tape:
                self.noise.data.normal_(0, 1)
                (t.hcodes, t.mu1, t.mu2, t.logvar1, t.logvar2, t.std1, t.std2) = \
                    t.netG(self.noise, t.txt_embedding, t.cls_onehot)
            if cfg.TRAIN.CAT_Z == "concat":
                t.mu = torch.cat((t.mu1, t.mu2), 1)
            elif cfg.TRAIN.CAT_Z == "product":
                t.mu = t.mu1 * t.mu2
            else:
                t.mu = t.mu1 + t.mu2
        else:
            if self.spec["ocfg"]["CLS_KIND"] == "multihot":
                t.real_cp = t.cls_label / (torch.sum(t.cls_label, 1).view(-1, 1))          # trainer.py:518
                t.cond_info = torch.cat((t.txt_embedding, t.cls_label), 1)                   # trainer.py:525
            else:   # config 3: birds index labels, cond = text only (cub:303-304,557,571)
                c0 = t.cls_label.long() - 1
                t.real_cp = t.onehot(c0, cfg.GAN.ENTITY_DIM + 1)
                t.cond_info = t.txt_embedding
            with rng_tape([noise], [seed]):
                self.noise.data.normal_(0, 1)
                t.hcodes, t.mu, t.logvar, t.std = t.netG(self.noise, t.cond_info)
        t.fake_imgs = t.netG.image(t.hcodes)
        out["real_cp"], out["mu"] = t.real_cp, t.mu
        out["h_codes"], out["fake_imgs"] = t.hcodes, t.fake_imgs
        out["errD"], out["gradD"], out["d_logits"] = [], [], []
        # hook D forwards to record logits
        for i in range(t.num_Ds):
            rec = []
            h = t.netsD[i].register_forward_hook(lambda m, a, o, rec=rec: rec.append([x.detach() for x in o]))
            if self.kind in ("catz_ca", "catz"):
                res = t.train_joint_Dnet(i, count)
            else:
                res = self._old_train_joint_Dnet(i, count)
            h.remove()
            out["errD"].append(torch.stack([r.detach() for r in res]))
            out["d_logits"].append(rec)
            out["gradD"].append({k: p.grad.clone() for k, p in t.netsD[i].named_parameters() if p.grad is not None})
        t.netG.zero_grad()
        rec = []
        hooks = [d.register_forward_hook(lambda m, a, o, rec=rec: rec.append([x.detach() for x in o])) for d in t.netsD]
        res = t.loss_joint_Gnet(count)
        for h in hooks:
            h.remove()
        res[0].backward()
        out["gradG"] = {k: p.grad.clone() for k, p in t.netG.named_parameters() if p.grad is not None}
        t.optimizerG.step()
        out["g_logits"] = rec
        out["errG"] = torch.stack([torch.as_tensor(r, dtype=torch.float32).detach() for r in res])
        return out

    def _old_train_joint_Dnet(self, i, count):
        """trainer.py:380-437 builds fake_cp on the host and calls `.cuda()` on it (a no-op here)."""
        return self.t.train_joint_Dnet(i, count)
