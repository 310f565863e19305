"""Global configuration, same surface as the reference's miscc/config.py: a module-level attribute dict `cfg`
whose defaults are the reference's (miscc/config.py:13-77) and `cfg_from_file(path)` which merges a YAML file
with the reference's strictness (unknown key -> KeyError, type mismatch -> ValueError; miscc/config.py:80-121).
Model classes read `cfg` at construction time exactly like the reference's do.
"""
import copy

import numpy as np


class AttrDict(dict):
    """dict with attribute access; nested dicts are converted on assignment (easydict-compatible subset)."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


_DEFAULTS = {
    "DATASET_NAME": "birds", "EMBEDDING_TYPE": "cnn-rnn", "CONFIG_NAME": "", "DATA_DIR": "",
    "GPU_ID": "0", "CUDA": True, "WORKERS": 6,
    "TREE": {"BRANCH_NUM": 3, "BASE_SIZE": 64, "SCALE": 2},
    "TEST": {"B_EXAMPLE": True, "SAMPLE_NUM": 30000, "EVAL_MODE": True, "G_CAPSULE": False, "CLS_PRIOR": False},
    "TRAIN": {"BATCH_SIZE": 64, "VIS_COUNT": 64, "MAX_EPOCH": 600, "SNAPSHOT_INTERVAL": 2000,
              "DISCRIMINATOR_LR": 2e-4, "GENERATOR_LR": 2e-4, "FLAG": True, "NET_G": "", "ENTITY_NET_G": "",
              "NET_D": "", "ENTITY_NET_D": "", "BIG_EVAL": False, "G_CAPSULE": False, "D_CAPSULE": False,
              "CAT_Z": "concat", "EXCHANGE": False, "GENERAL_IS": False,
              "COEFF": {"KL": 2.0, "UNCOND_LOSS": 0.0, "COLOR_LOSS": 0.0}},
    "GAN": {"EMBEDDING_DIM": 128, "DF_DIM": 64, "GF_DIM": 64, "Z_DIM": 100, "NETWORK_TYPE": "default", "R_NUM": 2,
            "B_CONDITION": False, "ENTITY_DIM": 200, "MANIFD_DIM": 128},
    "TEXT": {"DIMENSION": 1024},
}

cfg = AttrDict(_DEFAULTS)
__C = cfg


def reset_cfg():
    """Restore the defaults in place (the reference has no such helper; tests switch configs in one process)."""
    fresh = AttrDict(copy.deepcopy(_DEFAULTS))
    cfg.clear()
    for k, v in fresh.items():
        cfg[k] = v
    return cfg


def _merge_a_into_b(a, b):
    if not isinstance(a, dict):
        return
    for k, v in a.items():
        if k not in b:
            raise KeyError("{} is not a valid config key".format(k))
        old = b[k]
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            v = AttrDict(v)
        if type(old) is not type(v):
            if isinstance(old, np.ndarray):
                v = np.array(v, dtype=old.dtype)
            else:
                raise ValueError("Type mismatch ({} vs. {}) for config key: {}".format(type(old), type(v), k))
        if isinstance(v, AttrDict):
            try:
                _merge_a_into_b(v, old)
            except Exception:
                print("Error under config key: {}".format(k))
                raise
        else:
            b[k] = v


def cfg_from_file(filename):
    """Load a YAML config file and merge it into `cfg` (unknown keys and type changes are errors)."""
    import yaml
    with open(filename, "r") as f:
        loaded = yaml.safe_load(f)
    _merge_a_into_b(AttrDict(loaded or {}), cfg)


def cfg_from_dict(d):
    _merge_a_into_b(AttrDict(d), cfg)
