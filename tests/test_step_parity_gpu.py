"""GPU parity of the whole G+D training step against the CPU oracle, through the package's public trainer API
(which reaches the kernels through the C ABI).  Identical weights (loaded by state_dict key), identical synthetic
batch and injected noise / eps / seed.  Tolerances: the north star's BF16 bound, relative L2 <= 1e-2 per tensor for
images / logits / losses; gradients are compared per tensor with the bound written below; class-target indices
bit-exact."""
import numpy as np
import pytest
import torch

from oracle import configs as ocfg, shapes, synth
from oracle.ekl_oracle import OracleTrainer

pytestmark = pytest.mark.gpu

TOL_OUT = 1e-2        # stage-1 images, logits, losses (north star: rel-L2 <= 1e-2)
TOL_DEEP = 2e-2       # stage-2/3 images: 13-19 bf16 conv+BN layers deep, BatchNorm over a batch of only 2-4 samples
TOL_GRAD = 3e-2       # per-tensor gradient rel-L2 (bf16 activations through >= 10 train-mode BN layers)
TOL_GRAD_MEDIAN = 1e-2


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def build(name, B):
    from text2img_ekl_b200 import configs
    Trainer = configs.setup(name, batch=B)
    tr = Trainer(None, None, 64)
    tr.setup()
    oc = ocfg.oracle_cfg(name, batch=B)
    gsh = shapes.g_shapes(oc, cond_dim=ocfg.cond_dim(oc))
    dsh = [shapes.d_shapes(oc, r, joint=True, use_cap=oc.D_CAPSULE) for r in [64, 128, 256][: oc.BRANCH_NUM]]
    sdG = shapes.make_state_dict(gsh, "G")
    sdDs = [shapes.make_state_dict(s, "D%d" % i) for i, s in enumerate(dsh)]
    tr.netG.load_state_dict(sdG)
    for d, sd in zip(tr.netsD, sdDs):
        d.load_state_dict(sd)
    return tr, oc, OracleTrainer(oc, sdG, sdDs)


CASES = [("splitz_cap_ca", 4), ("catcls", 4), ("onlycapsule", 4), ("coco", 4), ("3stages", 2)]


@pytest.mark.parametrize("name,B", CASES)
def test_training_step_matches_oracle(name, B):
    torch.backends.cuda.matmul.allow_tf32 = False
    tr, oc, orc = build(name, B)
    dev = tr.device
    report = []
    for it in range(2):
        b = synth.make_batch(oc, B, "it%d" % it)
        want = orc.step(**b)
        data = (b["imgs"], b["wrong_imgs"], b["embedding"], b["cls"], None)
        grads_before = None
        errDs, errG = tr.train_step(data, noise=b["noise"].to(dev), eps=b["eps"].to(dev), seed=b["seed"].to(dev))
        torch.cuda.synchronize()
        # class targets: bit-exact indices
        assert torch.equal(tr.real_cp.argmax(1).cpu(), want["real_cp"].argmax(1))
        if oc.CLS_KIND == "index":
            assert torch.equal(tr.real_cp.cpu(), want["real_cp"])
        # images
        for i, (g, w) in enumerate(zip(tr.fake_imgs, want["fake_imgs"])):
            r = rel(g, w)
            report.append(("it%d img%d" % (it, i), r))
            assert r <= (TOL_OUT if i == 0 else TOL_DEEP), (name, it, "img", i, r)
        # losses
        for i, (g, w) in enumerate(zip(errDs, want["errD"])):
            r = rel(torch.stack([x.float() for x in g]), w)
            report.append(("it%d errD%d" % (it, i), r))
            assert r <= TOL_OUT, (name, it, "errD", i, r, [float(x) for x in g], w.tolist())
        r = rel(torch.stack([x.float() for x in errG]), want["errG"])
        report.append(("it%d errG" % it, r))
        assert r <= TOL_OUT, (name, it, "errG", r, [float(x) for x in errG], want["errG"].tolist())
        # generator-step logits of every D
        for i, (g, w) in enumerate(zip(tr.engine.last_g_logits, want["g_logits"])):
            for q in range(len(w)):
                r = rel(g[q], w[q])
                report.append(("it%d glogit%d_%d" % (it, i, q), r))
                assert r <= TOL_OUT, (name, it, "g_logits", i, q, r)
        # gradients of G (the D gradients were consumed by the in-step optimiser update; G's are still in .grad)
        rs = []
        for k, p in tr.netG.named_parameters():
            if k in want["gradG"]:
                rs.append((rel(p.grad, want["gradG"][k]), k))
        worst = max(rs)
        med = float(np.median([r for r, _ in rs]))
        report.append(("it%d gradG worst %s" % (it, worst[1]), worst[0]))
        report.append(("it%d gradG median" % it, med))
        assert worst[0] <= TOL_GRAD, (name, it, worst)
        assert med <= TOL_GRAD_MEDIAN, (name, it, med)
        # gradients of every D from its own update (kept in the flat buffers; the G step does not touch them)
        for i, d in enumerate(tr.netsD):
            rs = [(rel(p.grad, want["gradD"][i][k]), k) for k, p in d.named_parameters() if k in want["gradD"][i]]
            worst = max(rs)
            report.append(("it%d gradD%d worst %s" % (it, i, worst[1]), worst[0]))
            report.append(("it%d gradD%d median" % (it, i), float(np.median([r for r, _ in rs]))))
            assert worst[0] <= TOL_GRAD, (name, it, i, worst)
    # parameters after two optimiser steps
    for tag, net, sd in [("G", tr.netG, orc.sdG)] + [("D%d" % i, d, orc.sdDs[i]) for i, d in enumerate(tr.netsD)]:
        num = den = 0.0
        for k, v in net.state_dict().items():
            if v.is_floating_point() and "running" not in k:
                num += float((v.detach().float().cpu() - sd[k].detach()).pow(2).sum())
                den += float(sd[k].detach().pow(2).sum())
        r = (num / den) ** 0.5
        report.append(("params " + tag, r))
        assert r <= 3e-3, (name, tag, r)     # Adam's first steps move every weight by ~lr*sign(g): sign flips of ~0 grads dominate
    print("\n" + "\n".join("%-50s %.3e" % kv for kv in report))
