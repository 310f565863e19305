"""Drop-in for the reference's cub_trainer_splitz_cap_ca.py (split-z + capsule + CA_NET flavour): same module
functions (`load_network`, `define_optimizers`, `KL_loss`, `ce_loss`, `compute_mean_covariance`, `weights_init`) and
the same `condGANTrainer` methods (`prepare_data`, `onehot`, `train_joint_Dnet`, `loss_joint_Gnet`, `train`), driving
the B200 kernels through engine.StepEngine.  One process per GPU: `nn.DataParallel` (cub:139,163) is replaced by
torch.distributed (NCCL) gradient all-reduce when WORLD_SIZE > 1; BatchNorm statistics stay per replica.
Out of scope: Inception-score evaluation, tensorboard image summaries.
"""
import os
import time

import torch
import torch.nn as nn
import torch.optim as optim

from . import model
from . import ops
from .engine import (KL_loss, StepEngine, ce_loss, compute_mean_covariance, onehot)  # noqa: F401  (reference names)
from .miscc.config import cfg
from .miscc.utils import mkdir_p
from .parallel import make_allreduce

USE_CLS = True
SPLIT_Z = True


def weights_init(m):
    """cub:67-77: orthogonal(gain 1) for Conv / Linear class names (incl. CapsuleLinear), BN gamma ~ N(1, 0.02), beta 0."""
    classname = m.__class__.__name__
    if classname.find("Conv") != -1 and hasattr(m, "weight"):
        nn.init.orthogonal_(m.weight.data, 1.0)
    elif classname.find("BatchNorm") != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    elif classname.find("Linear") != -1:
        nn.init.orthogonal_(m.weight.data, 1.0)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def copy_G_params(net):
    return [p.data.clone() for p in net.parameters()]


def load_params(net, new_param):
    for p, new_p in zip(net.parameters(), new_param):
        p.data.copy_(new_p)


def build_G():
    shareGs = model.get_shareGs(cfg.GAN.GF_DIM)
    if USE_CLS and SPLIT_Z:
        netG = model.COND_G_NET_CATZ_CA(cfg.TEXT.DIMENSION, cfg.GAN.ENTITY_DIM, shareGs, use_cap=cfg.TRAIN.G_CAPSULE,
                                        cat=cfg.TRAIN.CAT_Z, exchange=cfg.TRAIN.EXCHANGE)            # cub:130
    else:
        netG = model.COND_G_NET(cfg.TEXT.DIMENSION, shareGs, use_cap=cfg.TRAIN.G_CAPSULE)           # cub:135
    return netG, shareGs


def build_Ds(allow_three=False):
    netsD = []
    if cfg.TREE.BRANCH_NUM > 0:
        netsD.append(model.JOINT_D_NET64(use_cap=cfg.TRAIN.D_CAPSULE))
    if cfg.TREE.BRANCH_NUM > 1:
        netsD.append(model.JOINT_D_NET128(use_cap=cfg.TRAIN.D_CAPSULE) if cfg.TREE.SCALE == 2 else model.JOINT_D_NET256())
    if cfg.TREE.BRANCH_NUM > 2:
        # the reference asserts 'br3 todo' here (cub:156); the 3-stage config composes JOINT_D_NET256 (cfg 2)
        netsD.append(model.JOINT_D_NET256())
    return netsD


def load_network(gpus, device=None):
    """cub:113-196.  Returns (netG, shareGs, netsD, num_Ds, count)."""
    device = device or (torch.device("cuda", gpus[0]) if gpus else torch.device("cuda"))
    netG, shareGs = build_G()
    netG.apply(weights_init)
    netsD = build_Ds()
    for d in netsD:
        d.apply(weights_init)
    count = 0
    if cfg.TRAIN.NET_G != "":
        state_dict = torch.load(cfg.TRAIN.NET_G, map_location="cpu")
        netG.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()})
        name = os.path.basename(cfg.TRAIN.NET_G)
        digits = "".join(ch for ch in name[name.rfind("_") + 1:name.rfind(".")] if ch.isdigit())
        count = int(digits) + 1 if digits else 0
    if cfg.TRAIN.NET_D != "":
        for i, d in enumerate(netsD):
            sd = torch.load("%s%d.pth" % (cfg.TRAIN.NET_D, i), map_location="cpu")
            d.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in sd.items()})
    netG.to(device)
    model.to_kernel_layout(netG)
    for d in netsD:
        d.to(device)
        model.to_kernel_layout(d)
    return netG, shareGs, netsD, len(netsD), count


def define_optimizers(netG, netsD=()):
    """cub:199-215: Adam(lr 2e-4, betas (0.5, 0.999)) per network.  On the GPU: optim.FlatAdam, the same update as one
    fused kernel per network over flat parameter / gradient / moment buffers (graph-capturable)."""
    kw = dict(betas=(0.5, 0.999))
    if next(netG.parameters()).is_cuda:
        from .optim import FlatAdam
        optimizersD = [FlatAdam(d.parameters(), lr=cfg.TRAIN.DISCRIMINATOR_LR, **kw) for d in netsD]
        optimizerG = FlatAdam(netG.parameters(), lr=cfg.TRAIN.GENERATOR_LR, **kw)
        return optimizerG, optimizersD
    optimizersD = [optim.Adam(d.parameters(), lr=cfg.TRAIN.DISCRIMINATOR_LR, **kw) for d in netsD]
    optimizerG = optim.Adam(netG.parameters(), lr=cfg.TRAIN.GENERATOR_LR, **kw)
    return optimizerG, optimizersD


class condGANTrainer(object):
    KIND = "catz_ca"
    COND = "txt+cls"

    def __init__(self, output_dir, data_loader, imsize):
        if cfg.TRAIN.FLAG and output_dir:
            self.model_dir = os.path.join(output_dir, "Model")
            self.image_dir = os.path.join(output_dir, "Image")
            self.log_dir = os.path.join(output_dir, "Log")
            for d in (self.model_dir, self.image_dir, self.log_dir):
                mkdir_p(d)
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        s_gpus = str(cfg.GPU_ID).split(",")
        self.gpus = [local_rank] if "LOCAL_RANK" in os.environ else [int(ix) for ix in s_gpus]
        self.num_gpus = len(self.gpus)
        self.device = torch.device("cuda", self.gpus[0])
        torch.cuda.set_device(self.device)
        self.batch_size = cfg.TRAIN.BATCH_SIZE
        self.max_epoch = cfg.TRAIN.MAX_EPOCH
        self.snapshot_interval = cfg.TRAIN.SNAPSHOT_INTERVAL
        self.data_loader = data_loader
        self.num_batches = len(data_loader) if data_loader is not None else 0
        self.engine = None

    # ------------------------------------------------------------------ data
    def prepare_data(self, data):
        """cub:295-320: zero-base the class index and move the batch to the device."""
        imgs, w_imgs, t_embedding, cls, _ = data
        cls = cls.long() - 1
        dev = self.device
        real_vimgs = [imgs[i].to(dev, non_blocking=True) for i in range(self.num_Ds)]
        wrong_vimgs = [w_imgs[i].to(dev, non_blocking=True) for i in range(self.num_Ds)]
        return imgs, real_vimgs, wrong_vimgs, t_embedding.to(dev, non_blocking=True), cls.to(dev, non_blocking=True)

    def onehot(self, cls_vec, n):
        return onehot(cls_vec, n)

    # ------------------------------------------------------------------ setup (cub:494-537)
    def setup(self):
        self.netG, self.shareGs, self.netsD, self.num_Ds, start_count = load_network(self.gpus, self.device)
        self.optimizerG, self.optimizersD = define_optimizers(self.netG, self.netsD)
        self.criterion = nn.BCELoss()
        self.CE = ce_loss
        B = self.batch_size
        self.real_labels = torch.ones(B, device=self.device)
        self.fake_labels = torch.zeros(B, device=self.device)
        self.fake_cp = torch.zeros(B, cfg.GAN.ENTITY_DIM + 1, device=self.device)
        self.fake_cp[:, -1] = 1
        self.noise = torch.zeros(B, cfg.GAN.Z_DIM, device=self.device)
        self.engine = StepEngine(self.netG, self.netsD, self.optimizerG, self.optimizersD, self.KIND, self.COND,
                                 allreduce=make_allreduce())
        return start_count

    # ------------------------------------------------------------------ reference-named step pieces
    def labels(self):
        """cub:556-557 (birds): one-hot class condition [B,E] and class target [B,E+1]."""
        self.cls_onehot = self.onehot(self.cls_label, cfg.GAN.ENTITY_DIM)
        self.real_cp = self.onehot(self.cls_label, cfg.GAN.ENTITY_DIM + 1)
        return self.cls_onehot

    def generate(self, eps=None, seed=None):
        e = self.engine
        self.fake_imgs = e.generate(self.noise, self.txt_embedding, self.labels(), eps, seed)
        self.hcodes, self.mu = e.hcodes, e.mu
        for k in ("mu1", "mu2", "logvar1", "logvar2", "std1", "std2", "logvar", "std"):
            if hasattr(e, k):
                setattr(self, k, getattr(e, k))

    def train_joint_Dnet(self, idx, count):
        """cub:404-461 -> (errD, errD_match, errD_uncond, errD_cls)."""
        return self.engine.d_step(idx, self.real_imgs[idx], self.wrong_imgs[idx], self.real_cp, self.fake_cp)

    def loss_joint_Gnet(self, count):
        """cub:463-490 -> (errG_total, match, uncond, cls, kl_sen, kl_cls)."""
        return self.engine.g_loss(self.real_cp)

    def train_step(self, data, count=1, noise=None, eps=None, seed=None):
        """One iteration of the hot loop (cub:552-608)."""
        self.imgs_tcpu, self.real_imgs, self.wrong_imgs, self.txt_embedding, self.cls_label = self.prepare_data(data)
        if noise is None:
            self.noise.normal_(0, 1)
        else:
            self.noise.copy_(noise)
        self.generate(eps, seed)
        # the discriminator updates are independent: engine.d_steps runs them as parallel stream branches
        errDs = self.engine.d_steps(self.real_imgs, self.wrong_imgs, self.real_cp, self.fake_cp)
        errG = self.engine.g_step(self.real_cp)
        return errDs, errG

    EAGER_STEPS_BEFORE_CAPTURE = 2

    def _loop_step(self, data, nxt, count):
        """One iteration of train()'s loop.  On a GPU the first EAGER_STEPS_BEFORE_CAPTURE batches run eagerly (they
        initialise the library handles and buffers that must exist before a capture), then the whole step is captured
        once (engine.GraphedStep; the capture itself executes nothing) and every following full-size batch is one graph
        replay with the next batch's upload hidden behind it.  Batches of another shape (a loader without drop_last;
        the reference uses drop_last=True, main.py:135) take the eager path.  EKL_GRAPH=0 keeps everything eager."""
        if self.device.type == "cuda" and os.environ.get("EKL_GRAPH", "1") != "0":
            gs = getattr(self, "_graphed", None)
            if gs is None and getattr(self, "_eager_done", 0) >= self.EAGER_STEPS_BEFORE_CAPTURE:
                from .engine import GraphedStep
                gs = self._graphed = GraphedStep(self, data, warmup=0)
            if gs is not None and gs.accepts(data):
                return gs.step(data, nxt if nxt is not None and gs.accepts(nxt) else None)
        self._eager_done = getattr(self, "_eager_done", 0) + 1
        return self.train_step(data, count)

    def train(self):
        """cub:492-672.  On a GPU the loop runs on a side stream: autograd pins each parameter's gradient accumulation to
        the stream of its first use, and work pinned to the legacy default stream cannot be captured into a graph."""
        if getattr(self, "device", None) is None or torch.device(self.device).type != "cuda":
            return self._train_loop()
        cur, side = torch.cuda.current_stream(), torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._train_loop()
        cur.wait_stream(side)

    def _train_loop(self):
        start_count = self.setup()
        count = start_count
        start_epoch = start_count // max(self.num_batches, 1)
        for epoch in range(start_epoch, self.max_epoch):
            start_t = time.time()
            errDs = errG = None
            it = iter(self.data_loader)
            data = next(it, None)
            while data is not None:
                nxt = next(it, None)
                errDs, errG = self._loop_step(data, nxt, count)
                count += 1
                data = nxt
            if errG is None:
                break
            end_t = time.time()
            tot = [sum(e[k] for e in errDs).item() for k in range(4)]
            print("[%d/%d][BN=%d][%d stages] Loss_D_all: %.2f match: %.2f uncond: %.2f cls: %.2f | "
                  "Loss_G_all: %.2f match: %.2f uncond: %.2f cls: %.2f KL: %s Time: %.2fs"
                  % (epoch, self.max_epoch, self.num_batches, self.num_Ds, tot[0], tot[1], tot[2], tot[3],
                     errG[0].item(), float(errG[1]), float(errG[2]), float(errG[3]),
                     " ".join("%.3f" % float(k) for k in errG[4:]), end_t - start_t))
            if hasattr(self, "model_dir") and (epoch % self.snapshot_interval == 0 or epoch > 199):
                # keys carry the DataParallel "module." prefix like the reference's snapshots (cub:662-667)
                sd = {"module." + k: v for k, v in self.netG.state_dict().items()}
                torch.save(sd, "%s/netG_epoch%d.pth" % (self.model_dir, epoch))
