"""The five BASELINE configs as this package resolves them (SURVEY section 8 'config resolution'): yml file +
trainer flavour.  `setup(name)` resets the global cfg, merges the yml and returns the trainer class to use."""
import os

from .miscc.config import cfg, cfg_from_file, reset_cfg

CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")

# name -> (yml, trainer flavour, condition input, class-label kind)
RESOLVED = {
    "catcls": ("b200_catcls.yml", "trainer", "txt+cls", "multihot"),
    "3stages": ("b200_3stages.yml", "trainer", "txt+cls", "multihot"),
    "onlycapsule": ("b200_onlycapsule.yml", "trainer", "txt", "index"),
    "splitz_cap_ca": ("b200_splitz_cap_ca.yml", "cub", None, "index"),
    "coco": ("b200_coco.yml", "trainer", "txt+cls", "multihot"),
}


def setup(name, batch=None, width=None):
    yml, flavour, cond, cls_kind = RESOLVED[name]
    reset_cfg()
    cfg_from_file(os.path.join(CFG_DIR, yml))
    if batch:
        cfg.TRAIN.BATCH_SIZE = batch
    if width:
        cfg.GAN.GF_DIM = cfg.GAN.DF_DIM = width
    if flavour == "cub":
        from .cub_trainer_splitz_cap_ca import condGANTrainer
        return condGANTrainer
    from . import trainer as T

    class _Trainer(T.condGANTrainer):
        COND = cond
        CLS_KIND = cls_kind
    _Trainer.__name__ = "condGANTrainer"
    return _Trainer
