"""Synthetic batches in the loader's layout (datasets.py:346): (imgs[list], wrong_imgs[list], embedding, cls, keys).
SURVEY 8d: per rank torch.Generator().manual_seed(1234 + rank); imgs ~ U(-1,1) fp32 [B,3,64*2^i,64*2^i];
embedding ~ N(0,1) [B,1024]; birds cls ~ randint(1, E+1) int64 (1-based); coco cls = Bernoulli(0.03) multi-hot over E
columns with column E set iff none is."""
import torch

from .miscc.config import cfg


def make_batch(B, cls_kind, gen, pin=False):
    imgs, wrong = [], []
    for i in range(cfg.TREE.BRANCH_NUM):
        s = cfg.TREE.BASE_SIZE * (cfg.TREE.SCALE ** i)
        imgs.append(torch.rand(B, 3, s, s, generator=gen) * 2 - 1)
        wrong.append(torch.rand(B, 3, s, s, generator=gen) * 2 - 1)
    emb = torch.randn(B, cfg.TEXT.DIMENSION, generator=gen)
    E = cfg.GAN.ENTITY_DIM
    if cls_kind == "index":
        cls = torch.randint(1, E + 1, (B,), generator=gen)
    else:
        m = (torch.rand(B, E + 1, generator=gen) < 0.03).float()
        m[:, -1] = 0
        m[m.sum(1) == 0, -1] = 1
        cls = m
    if pin and torch.cuda.is_available():
        imgs = [t.pin_memory() for t in imgs]
        wrong = [t.pin_memory() for t in wrong]
        emb, cls = emb.pin_memory(), cls.pin_memory()
    return imgs, wrong, emb, cls, None


class SyntheticLoader:
    """A fixed pool of pinned host batches, cycled: stands in for the DataLoader in benchmarks."""

    def __init__(self, B, cls_kind, rank=0, pool=4, length=1 << 30):
        gen = torch.Generator().manual_seed(1234 + rank)
        self.pool = [make_batch(B, cls_kind, gen, pin=True) for _ in range(pool)]
        self.length = length

    def __len__(self):
        return self.length

    def __iter__(self):
        i = 0
        while i < self.length:
            yield self.pool[i % len(self.pool)]
            i += 1
