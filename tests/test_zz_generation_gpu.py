"""GPU tests of the generation path and of the reference's file-based resume.
Collected last (file name) so that a problem here cannot mask the parity tests under `pytest -x`."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _TestSetLoader:
    """Batches in the test-split layout evaluate() consumes (datasets.py test mode: several sentence embeddings per
    image): (imgs[list], t_embeddings [B, n_sentences, T], cls 1-based int64 [B], keys)."""

    def __init__(self, n_batches, B, n_sent, T, E):
        g = torch.Generator().manual_seed(9)
        self.batches = [([torch.zeros(B, 3, 64, 64)], torch.randn(B, n_sent, T, generator=g),
                         torch.randint(1, E + 1, (B,), generator=g), ["%03d.Class/img_%d_%d" % (k, k, i) for i in range(B)])
                        for k in range(n_batches)]

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


def test_save_model_load_network_and_evaluate(tmp_path, monkeypatch):
    """(1) save_model writes the files cfg.TRAIN.NET_G / NET_D name (cub:218-228); load_network reads them back
    ('module.' prefix stripped, count parsed from the file name, cub:171-184).  (2) evaluate() (cub:776-911) generates
    the final-stage image for every sentence embedding of every test batch through the kernels and writes single PNGs
    or per-sample tile sheets."""
    from PIL import Image
    from text2img_ekl_b200 import configs, cub_trainer_splitz_cap_ca as T
    from text2img_ekl_b200.miscc.config import cfg
    from text2img_ekl_b200.synthetic import SyntheticLoader
    torch.backends.cuda.matmul.allow_tf32 = False
    monkeypatch.setenv("EKL_GRAPH", "0")
    Trainer = configs.setup("splitz_cap_ca", batch=4)
    loader = SyntheticLoader(4, "index", pool=2, length=2)
    tr = Trainer(None, loader, 64)
    tr.max_epoch = 1
    tr.train()                                     # running statistics differ from their initial values
    dev = tr.device
    T.save_model(tr.netG, None, tr.netsD, 3, str(tmp_path))
    try:
        cfg.TRAIN.NET_G, cfg.TRAIN.NET_D = str(tmp_path / "netG_3.pth"), str(tmp_path / "netD")
        netG2, _, netsD2, num_Ds, count = T.load_network([dev.index or 0], dev)
        assert count == 4 and num_Ds == len(tr.netsD)
        for a, b in zip([netG2] + list(netsD2), [tr.netG] + list(tr.netsD)):
            sa, sb = a.state_dict(), b.state_dict()
            assert list(sa) == list(sb)
            for k in sa:
                assert torch.equal(sa[k].cpu(), sb[k].cpu()), k
        cfg.TEST.G_CAPSULE = cfg.TRAIN.G_CAPSULE
        test_loader = _TestSetLoader(2, 4, 3, cfg.TEXT.DIMENSION, cfg.GAN.ENTITY_DIM)
        ev = Trainer(None, test_loader, 64)
        cfg.TEST.B_EXAMPLE = False
        out = tmp_path / "single"
        assert ev.evaluate("test", save_dir=str(out)) == 2 * 3 * 4
        pngs = sorted((out / "single_samples").rglob("*.png"))
        assert len(pngs) == 24 and "_128_class" in pngs[0].name and "_nid0" in pngs[0].name
        a0 = np.asarray(Image.open(pngs[0]))
        assert a0.shape == (128, 128, 3) and a0.std() > 0
        cfg.TEST.B_EXAMPLE = True
        out = tmp_path / "sheets"
        assert ev.evaluate("test", save_dir=str(out)) == 24
        sheets = sorted((out / "super" / "test").rglob("*.png"))
        assert len(sheets) == 8 and np.asarray(Image.open(sheets[0])).shape == (2 + 130, 2 + 3 * 130, 3)
    finally:
        cfg.TRAIN.NET_G = cfg.TRAIN.NET_D = ""
        cfg.TEST.B_EXAMPLE, cfg.TEST.G_CAPSULE = True, False
