"""Deterministic synthetic batches in the loader's layout (datasets.py:346) -- TEST INFRASTRUCTURE."""
import numpy as np
import torch

from . import detfill


def make_batch(c, B, tag="b0"):
    """c: OracleCfg.  Returns dict(imgs, wrong_imgs, embedding, cls, noise, eps, seed) of CPU fp32/int64 tensors.
    imgs/wrong ~ U(-1,1) [B,3,64*2^i,64*2^i]; embedding ~N(0,1) [B,1024]; birds cls in 1..E (1-based, cub:303-304);
    coco cls = sparse multi-hot over E columns, column E set iff none (datasets.py:337-344)."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    imgs, wrong = [], []
    for i in range(c.BRANCH_NUM):
        s = 64 * (c.SCALE ** i)
        imgs.append(t(detfill.uniform("%s:img%d" % (tag, i), (B, 3, s, s))))
        wrong.append(t(detfill.uniform("%s:wrong%d" % (tag, i), (B, 3, s, s))))
    emb = t(detfill.normalish(tag + ":emb", (B, c.TEXT_DIM)))
    if c.CLS_KIND == "index":
        cls = t(detfill.randint(tag + ":cls", (B,), 1, c.ENTITY_DIM + 1))
    else:
        m = (detfill.uniform(tag + ":mh", (B, c.ENTITY_DIM + 1), 0, 1) < 0.03).astype(np.float32)
        m[:, -1] = 0
        m[m.sum(1) == 0, -1] = 1
        cls = t(m)
    return dict(imgs=imgs, wrong_imgs=wrong, embedding=emb, cls=cls,
                noise=t(detfill.normalish(tag + ":noise", (B, c.Z_DIM))),
                eps=t(detfill.normalish(tag + ":eps", (B, c.EMBEDDING_DIM))),
                seed=t(detfill.normalish(tag + ":seed", (B, c.MANIFD_DIM))))
