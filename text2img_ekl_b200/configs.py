"""The five BASELINE configs (and the conditioning variants of SURVEY 8f row 2) as this package resolves them
(SURVEY section 8 'config resolution'): the reference's own cfg/*.yml + a documented override dict + trainer flavour.

`setup(name)` resets the global cfg, merges the yml, applies the overrides and returns the trainer class to use.
The yml is the reference's file (<EKL_REFERENCE_ROOT or /root/reference>/cfg/<yml>) whenever that tree is present;
the package ships resolved copies (cfg/b200_*.yml = reference yml + the same overrides) for machines without the
reference -- tests/test_plan_host.py checks that both routes give the identical cfg.
"""
import os

from .miscc.config import cfg, cfg_from_dict, cfg_from_file, reset_cfg

CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")
REF_ROOT = os.environ.get("EKL_REFERENCE_ROOT", "/root/reference")

# Overrides on top of the reference yml that make the reference classes compose (SURVEY 8 "config resolution"; four of
# the five ymls do not run in the reference as written).  BATCH_SIZE is the per-GPU batch BASELINE.json names.
_SUM = {"TRAIN": {"CAT_Z": "sum"}}            # default 'concat' doubles in_dim / ef_dim of every non-split-z generator
# name -> (reference yml, shipped resolved copy, overrides, trainer flavour, condition input, class-label kind)
RESOLVED = {
    "catcls": ("birds_2stgs_catcls.yml", "b200_catcls.yml", {"TRAIN": {"CAT_Z": "sum", "BATCH_SIZE": 24}},
               "trainer", "txt+cls", "multihot"),
    "3stages": ("birds_3stages.yml", "b200_3stages.yml", _SUM, "trainer", "txt+cls", "multihot"),
    "onlycapsule": ("birds_2stgs_onlycapsule.yml", "b200_onlycapsule.yml",
                    {"TRAIN": {"CAT_Z": "sum", "G_CAPSULE": True, "D_CAPSULE": True, "BATCH_SIZE": 32}}, "trainer", "txt", "index"),
    "splitz_cap_ca": ("birds_2stg_splitz_cap_ca.realcls.yml", "b200_splitz_cap_ca.yml", {}, "cub", None, "index"),
    "coco": ("coco_2stgs.yml", "b200_coco.yml", _SUM, "trainer", "txt+cls", "multihot"),
}
# conditioning variants of config 4 (one override each; oracle/configs.py pins the same set against the reference)
_C4 = ("birds_2stg_splitz_cap_ca.realcls.yml", "b200_splitz_cap_ca.yml")
VARIANTS = {
    "splitz_cat_sum": _C4 + ({"TRAIN": {"CAT_Z": "sum"}}, "cub", None, "index"),                      # model.py:500-505
    "splitz_cat_product": _C4 + ({"TRAIN": {"CAT_Z": "product"}}, "cub", None, "index"),
    "splitz_scale4_sum": _C4 + ({"TRAIN": {"CAT_Z": "sum"}, "TREE": {"SCALE": 4}}, "cub", None, "index"),   # model.py:406-407
    "catz_exchange": _C4 + ({"TRAIN": {"EXCHANGE": True}}, "cub_catz", None, "index"),                # model.py:280-333,567
    "catz_plain": _C4 + ({"TRAIN": {"G_CAPSULE": False, "D_CAPSULE": False}}, "cub_catz", None, "index"),
    # config 2's generator with the StackGAN++ two-head D_NET64/128/256 (model.py:874-1202): match + uncond losses only
    "3stages_dnet": ("birds_3stages.yml", "b200_3stages.yml", {"TRAIN": {"CAT_Z": "sum"}}, "trainer_plain_d", "txt+cls", "multihot"),
}


def yml_path(name, prefer_reference=True):
    """(path, overrides to apply after merging it)."""
    ref_yml, shipped, over = (RESOLVED.get(name) or VARIANTS[name])[:3]
    ref = os.path.join(REF_ROOT, "cfg", ref_yml)
    if prefer_reference and os.path.isfile(ref):
        return ref, over
    # the shipped copy already contains the BASELINE config's overrides (applying them again changes nothing); a variant
    # adds its own on top
    return os.path.join(CFG_DIR, shipped), over


def setup(name, batch=None, width=None, prefer_reference=True):
    flavour, cond, cls_kind = (RESOLVED.get(name) or VARIANTS[name])[3:]
    path, over = yml_path(name, prefer_reference)
    reset_cfg()
    cfg_from_file(path)
    cfg_from_dict(over)
    if batch:
        cfg.TRAIN.BATCH_SIZE = batch
    if width:
        cfg.GAN.GF_DIM = cfg.GAN.DF_DIM = width
    if flavour in ("cub", "cub_catz"):
        from . import cub_trainer_splitz_cap_ca as T
        if flavour == "cub":
            return T.condGANTrainer

        class _CatzTrainer(T.condGANTrainer):
            G_CLASS = "COND_G_NET_CATZ"          # model.py:567: two VC_NETs (no shipped reference trainer builds it)
        _CatzTrainer.__name__ = "condGANTrainer"
        return _CatzTrainer
    from . import trainer as T

    class _Trainer(T.condGANTrainer):
        COND = cond
        CLS_KIND = cls_kind
        PLAIN_D = flavour == "trainer_plain_d"
    _Trainer.__name__ = "condGANTrainer"
    return _Trainer
