"""Drop-in for the reference's trainer.py (the older twin: COND_G_NET conditioned on cat(text, multi-hot class),
single (mu, logvar); trainer.py:107-564).  Shares the B200 step engine with cub_trainer_splitz_cap_ca.py.
`COND = 'txt'` + index labels reproduces the capsule-only composition of cfg/birds_2stgs_onlycapsule.yml
(cub_trainer_splitz_cap_ca.py:135,571 with USE_CLS False)."""
import torch

from . import cub_trainer_splitz_cap_ca as _cub
from . import model
from .cub_trainer_splitz_cap_ca import (KL_loss, ce_loss, compute_mean_covariance, copy_G_params, define_optimizers,  # noqa: F401
                                        load_params, weights_init)
from .datasets import stage_images
from .miscc.config import cfg


def load_network(gpus, device=None, cond="txt+cls", plain_d=False):
    """trainer.py:107-160 with entity_netG = COND_G_NET(E + 1 + text) (:116) as the generator."""
    device = device or (torch.device("cuda", gpus[0]) if gpus else torch.device("cuda"))
    shareGs = model.get_shareGs(cfg.GAN.GF_DIM)
    cond_dim = cfg.TEXT.DIMENSION + (cfg.GAN.ENTITY_DIM + 1 if cond == "txt+cls" else 0)
    netG = model.COND_G_NET(cond_dim, shareGs, use_cap=cfg.TRAIN.G_CAPSULE)
    netG.apply(weights_init)
    netsD = _cub.build_Ds(plain_d)
    for d in netsD:
        d.apply(weights_init)
    count = _cub.load_snapshots(netG, netsD)           # trainer.py:138-152: cfg.TRAIN.NET_G / NET_D resume
    netG.to(device)
    model.to_kernel_layout(netG)
    for d in netsD:
        d.to(device)
        model.to_kernel_layout(d)
    return netG, shareGs, netsD, len(netsD), count


class condGANTrainer(_cub.condGANTrainer):
    KIND = "cond"
    COND = "txt+cls"
    CLS_KIND = "multihot"          # coco-style float multi-hot labels (datasets.py:337-344); 'index' = birds

    def prepare_data(self, data):
        """trainer.py:266-290: the class tensor is used as delivered (no zero-basing)."""
        imgs, w_imgs, t_embedding, cls, _ = data
        if self.CLS_KIND == "index":
            cls = cls.long() - 1
        dev = self.device
        real_vimgs = stage_images(imgs, self.num_Ds, dev)       # fp32 pyramid as delivered, or uint8 crops -> device pyramid
        wrong_vimgs = stage_images(w_imgs, self.num_Ds, dev)
        return imgs, real_vimgs, wrong_vimgs, t_embedding.to(dev, non_blocking=True), cls.to(dev, non_blocking=True)

    def setup(self):
        self.netG, self.shareGs, self.netsD, self.num_Ds, start_count = load_network(self.gpus, self.device, self.COND, self.PLAIN_D)
        self._replicate()
        self.optimizerG, self.optimizersD = define_optimizers(self.netG, self.netsD)
        B = self.batch_size
        self.fake_cp = torch.zeros(B, cfg.GAN.ENTITY_DIM + 1, device=self.device)
        self.fake_cp[:, -1] = 1
        self.noise = torch.zeros(B, cfg.GAN.Z_DIM, device=self.device)
        from .engine import StepEngine
        self.engine = StepEngine(self.netG, self.netsD, self.optimizerG, self.optimizersD, self.KIND, self.COND)
        return start_count

    def labels(self):
        if self.CLS_KIND == "multihot":
            self.real_cp = self.cls_label / torch.sum(self.cls_label, 1).view(-1, 1)          # trainer.py:518
            return self.cls_label
        self.real_cp = self.onehot(self.cls_label, cfg.GAN.ENTITY_DIM + 1)
        return self.real_cp
