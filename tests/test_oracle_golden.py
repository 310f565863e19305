"""Pin the CPU oracle (oracle/ekl_oracle.py) against fixtures produced by the REAL reference
(oracle/gen_golden.py -> tests/golden/*.npz): every loss, logit, image/h_code checksum, every gradient
checksum and the post-Adam parameter norms, for all five resolved BASELINE configs and the conditioning variants of
SURVEY 8f row 2 (CAT_Z sum / product, TREE.SCALE 4, COND_G_NET_CATZ with the exchange capsule stem and with the Linear
stem)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import configs, shapes, summary, synth
from oracle.ekl_oracle import OracleTrainer
from oracle import ekl_oracle as O
from oracle import detfill

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "step_*.npz")))
RTOL, ATOL = 2e-4, 1e-6      # fp32 CPU arithmetic on a different host / thread count than the generator


def _parse(path):
    base = os.path.basename(path)[5:-4]
    name, b, w = base.rsplit("_", 2)
    return name, int(b[1:]), int(w[1:])


def _shape_list(sh):
    return ["%s|%s" % (k, ",".join(map(str, v))) for k, v in sh.items()]


def build_oracle(name, B, width):
    oc = configs.oracle_cfg(name, batch=B, gf=width, df=width)
    gsh = shapes.g_shapes(oc, cond_dim=configs.cond_dim(oc))
    res = [64, 128 if oc.SCALE == 2 else 256, 256][: oc.BRANCH_NUM]                    # cub:144-154
    dsh = [shapes.d_shapes(oc, r, joint=True, use_cap=oc.D_CAPSULE) for r in res]
    sdG = shapes.make_state_dict(gsh, "G")
    sdDs = [shapes.make_state_dict(s, "D%d" % i) for i, s in enumerate(dsh)]
    return oc, gsh, dsh, OracleTrainer(oc, sdG, sdDs)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p) for p in CASES])
def test_step_matches_reference(path):
    name, B, width = _parse(path)
    gold = np.load(path)
    oc, gsh, dsh, tr = build_oracle(name, B, width)
    # state_dict key names / shapes / order are the reference's
    assert _shape_list(gsh) == list(gold["shapes/G"])
    for i, s in enumerate(dsh):
        assert _shape_list(s) == list(gold["shapes/D%d" % i])
    iters = 1 + max(int(k.split("/")[0][2:]) for k in gold.files if k.startswith("it"))
    for it in range(iters):
        out = tr.step(**synth.make_batch(oc, B, "it%d" % it))
        params = {"G": tr.sdG}
        params.update({"D%d" % i: sd for i, sd in enumerate(tr.sdDs)})
        got = summary.summarize(out, params)
        keys = [k for k in gold.files if k.startswith("it%d/" % it)]
        assert len(keys) == len(got)
        for k in keys:
            g, r = got[k.split("/", 1)[1]], gold[k]
            scale = max(1.0, float(np.abs(r).max()))
            np.testing.assert_allclose(g, r, rtol=RTOL, atol=ATOL * scale, err_msg=k)


def test_modules_match_reference():
    gold = np.load(os.path.join(GOLD, "modules.npz"))
    oc = configs.oracle_cfg("3stages", batch=4, gf=8, df=8)
    gsh = shapes.g_shapes(oc, kind="gnet")
    assert _shape_list(gsh) == list(gold["gnet/shapes"])
    sd = shapes.make_state_dict(gsh, "GN")
    b = synth.make_batch(oc, 4, "gn")
    hs, mu, lv, std = O.g_forward_gnet(sd, oc, b["noise"], b["embedding"], b["eps"])
    imgs = O.g_images(hs, sd)
    for i in range(3):
        np.testing.assert_allclose(summary.tsum("h%d" % i, hs[i]), gold["gnet/h%d" % i], rtol=RTOL, atol=1e-5)
        np.testing.assert_allclose(summary.tsum("img%d" % i, imgs[i]), gold["gnet/img%d" % i], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(summary.tsum("mu", mu), gold["gnet/mu"], rtol=RTOL, atol=1e-5)
    for res in (64, 128, 256):
        dsh = shapes.d_shapes(oc, res, joint=False)
        assert _shape_list(dsh) == list(gold["dnet%d/shapes" % res])
        sdd = shapes.make_state_dict(dsh, "DP%d" % res)
        x = torch.from_numpy(detfill.uniform("dp:x%d" % res, (4, 3, res, res)))
        cc = torch.from_numpy(detfill.normalish("dp:c%d" % res, (4, oc.EMBEDDING_DIM)))
        o = O.d_plain_forward(x, cc, sdd, oc, res)
        np.testing.assert_allclose(o[0].detach().numpy(), gold["dnet%d/cond" % res], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(o[1].detach().numpy(), gold["dnet%d/uncond" % res], rtol=RTOL, atol=1e-6)
    mu = torch.from_numpy(detfill.normalish("l:mu", (6, 128)))
    lv = torch.from_numpy(detfill.normalish("l:lv", (6, 128))) * 0.3
    np.testing.assert_allclose(float(O.kl_loss(mu, lv)), gold["loss/kl"][0], rtol=1e-5)
    logq = torch.log_softmax(torch.from_numpy(detfill.normalish("l:q", (6, 201))), 1)
    p = torch.softmax(torch.from_numpy(detfill.normalish("l:p", (6, 201))), 1)
    np.testing.assert_allclose(float(O.ce_loss(logq, p)), gold["loss/ce"][0], rtol=1e-5)
    img = torch.from_numpy(detfill.uniform("l:img", (3, 3, 16, 16)))
    mean, cov = O.mean_covariance(img)
    np.testing.assert_allclose(mean.reshape(-1).numpy(), gold["loss/mean"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(cov.reshape(-1).numpy(), gold["loss/cov"], rtol=1e-5, atol=1e-7)
    cls = torch.from_numpy(detfill.randint("l:cls", (9,), 0, 200))
    oh = O.onehot(cls, 201)
    assert np.array_equal(oh.argmax(1).numpy().astype(np.float64), gold["loss/onehot"])      # bit-exact indices
    assert float(oh.sum()) == gold["loss/onehot_sum"][0]


def test_capsule_reduced_algebra_equals_materialised_priors():
    """The routing never needs the [B,O,I,L] priors: logits/weighted sums reduce to in_length-space algebra.
    (property used by the CUDA kernel; checked here against the materialised restatement)."""
    from oracle import capsule_ref
    x = torch.from_numpy(detfill.normalish("cap:x", (3, 12, 8)))
    w = torch.from_numpy(detfill.normalish("cap:w", (20, 6, 8))) * 0.3
    ref = capsule_ref.capsule_linear(x, w, "dynamic", 3)
    vsum = torch.zeros(3, 20, 6)
    for r in range(3):
        u = torch.einsum("olk,bol->bok", w, vsum)
        logit = torch.einsum("bik,bok->boi", x, u)
        c = torch.softmax(logit, dim=1)
        y = torch.einsum("boi,bik->bok", c, x)
        v = capsule_ref.squash(torch.einsum("olk,bok->bol", w, y))
        vsum = vsum + v
    np.testing.assert_allclose(v.numpy(), ref.numpy(), rtol=1e-4, atol=1e-6)


def test_pyramid_restatement_equals_pil():
    """oracle/pil_resample.py restates PIL's 8-bit BILINEAR resize + ToTensor/Normalize (datasets.py:43-68); pinned here
    against the Pillow / torchvision of this image, bit for bit."""
    Image = pytest.importorskip("PIL.Image")
    from oracle import pil_resample as R
    rng = np.random.default_rng(5)
    for S, sizes in ((256, (64, 128, 256)), (128, (64, 128)), (76, (64, 76)), (100, (37, 100))):
        img = rng.integers(0, 256, (S, S, 3), dtype=np.uint8)
        img[: S // 4, : S // 3] = 255                       # saturated block: exercises the clip
        img[S // 2:, S // 2:] //= 16
        levels = R.pyramid(img, sizes)
        for s, lv in zip(sizes, levels):
            ref = np.asarray(Image.fromarray(img).resize((s, s), Image.BILINEAR)) if s != S else img
            want = ((torch.from_numpy(ref.copy()).permute(2, 0, 1).float().div(255) - 0.5) / 0.5).numpy()
            assert lv.shape == (3, s, s) and lv.dtype == np.float32
            assert np.array_equal(lv, want), (S, s, np.abs(lv - want).max())


def test_pyramid_host_tables_equal_restatement():
    """the product's host-side coefficient tables (text2img_ekl_b200/datasets.py) are the oracle's, integer for integer."""
    from oracle import pil_resample as R
    from text2img_ekl_b200 import datasets as D
    for a, b in ((256, 128), (256, 64), (128, 64), (76, 64), (100, 37), (64, 128), (299, 64)):
        bo, ko = R.bilinear_coeffs(a, b)
        bt, kt = D.pil_bilinear_tables(a, b)
        assert np.array_equal(bo, bt.numpy()) and np.array_equal(ko, kt.numpy()), (a, b)
