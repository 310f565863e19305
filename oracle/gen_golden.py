"""Generate tests/golden/*.npz by running the REAL reference here (build container only).

    python -m oracle.gen_golden            # from the repo root; needs /root/reference

For each resolved BASELINE config the reference nets are built by the reference's own constructors,
their weights overwritten by oracle.detfill (so tests can regenerate them without the reference), and
the reference's own train_joint_Dnet / loss_joint_Gnet / Adam steps are run on oracle.synth batches
with replayed RNG draws.  Stored: losses, logits, checksums + sampled entries of images, h_codes and
every gradient, and post-step parameter norms.  TEST INFRASTRUCTURE.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import configs, detfill, ref_harness, summary, synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (config, batch, gf/df width, iterations)
CASES = [(n, 4, 8, 2) for n in configs.CONFIGS if n not in configs.NO_REFERENCE_STEP] + [("splitz_cap_ca", 4, 64, 1), ("catcls", 2, 64, 1)]
# the five BASELINE configs + the conditioning variants of SURVEY 8f row 2 (configs.CONFIGS lists both)


def run_case(name, B, width, iters):
    torch.manual_seed(0)
    cfg, netG, netsD = ref_harness.build_nets(name, batch=B, gf=width, df=width)
    oc = configs.oracle_cfg(name, batch=B, gf=width, df=width)
    with torch.no_grad():
        detfill.fill_state_dict(netG.state_dict(), "G")
        for i, d in enumerate(netsD):
            detfill.fill_state_dict(d.state_dict(), "D%d" % i)
    shapes = {"G": {k: tuple(v.shape) for k, v in netG.state_dict().items()}}
    for i, d in enumerate(netsD):
        shapes["D%d" % i] = {k: tuple(v.shape) for k, v in d.state_dict().items()}
    rs = ref_harness.RefStepper(name, netG, netsD)
    flat = {}
    for it in range(iters):
        b = synth.make_batch(oc, B, "it%d" % it)
        out = rs.step(**b)
        params = {"G": netG.state_dict()}
        params.update({"D%d" % i: d.state_dict() for i, d in enumerate(netsD)})
        for k, v in summary.summarize(out, params).items():
            flat["it%d/%s" % (it, k)] = v
    for tag, sh in shapes.items():
        flat["shapes/" + tag] = np.array(["%s|%s" % (k, ",".join(map(str, v))) for k, v in sh.items()])
    return flat


def module_cases():
    """Module-level known answers: G_NET sub-module composition, D_NET64/128/256, loss helpers."""
    R = ref_harness.import_reference()
    m, cub = R["model"], R["cub"]
    flat = {}
    cfg = ref_harness.set_cfg("3stages", batch=4, gf=8, df=8)
    cfg.TRAIN.CAT_Z = "sum"
    oc = configs.oracle_cfg("3stages", batch=4, gf=8, df=8)
    # G_NET (forward itself is broken: model.py:769) -> call sub-modules as SURVEY 8 prescribes
    g = m.G_NET(m.get_shareGs(cfg.GAN.GF_DIM))
    with torch.no_grad():
        detfill.fill_state_dict(g.state_dict(), "GN")
    b = synth.make_batch(oc, 4, "gn")
    with ref_harness.rng_tape([b["eps"]], []):
        c, mu, lv, std = g.ca_net(b["embedding"])
    h1 = g.h_net1(b["noise"], c)
    h2 = g.h_net2(h1, c)
    h3 = g.h_net3(h2, c)
    for i, (h, net) in enumerate(((h1, g.img_net1), (h2, g.img_net2), (h3, g.img_net3))):
        flat["gnet/h%d" % i] = summary.tsum("h%d" % i, h)
        flat["gnet/img%d" % i] = summary.tsum("img%d" % i, net(h))
    flat["gnet/mu"] = summary.tsum("mu", mu)
    flat["gnet/shapes"] = np.array(["%s|%s" % (k, ",".join(map(str, v.shape))) for k, v in g.state_dict().items()])
    # plain D_NETs
    for res, cls in ((64, m.D_NET64), (128, m.D_NET128), (256, m.D_NET256)):
        d = cls()
        with torch.no_grad():
            detfill.fill_state_dict(d.state_dict(), "DP%d" % res)
        x = torch.from_numpy(detfill.uniform("dp:x%d" % res, (4, 3, res, res)))
        cc = torch.from_numpy(detfill.normalish("dp:c%d" % res, (4, cfg.GAN.EMBEDDING_DIM)))
        o = d(x, cc)
        flat["dnet%d/cond" % res] = o[0].detach().double().numpy()
        flat["dnet%d/uncond" % res] = o[1].detach().double().numpy()
        flat["dnet%d/shapes" % res] = np.array(["%s|%s" % (k, ",".join(map(str, v.shape))) for k, v in d.state_dict().items()])
    # loss helpers
    mu = torch.from_numpy(detfill.normalish("l:mu", (6, 128)))
    lv = torch.from_numpy(detfill.normalish("l:lv", (6, 128))) * 0.3
    flat["loss/kl"] = np.array([float(cub.KL_loss(mu.clone(), lv.clone()))])
    logq = torch.log_softmax(torch.from_numpy(detfill.normalish("l:q", (6, 201))), 1)
    p = torch.softmax(torch.from_numpy(detfill.normalish("l:p", (6, 201))), 1)
    flat["loss/ce"] = np.array([float(cub.ce_loss(logq, p))])
    img = torch.from_numpy(detfill.uniform("l:img", (3, 3, 16, 16)))
    mean, cov = cub.compute_mean_covariance(img)
    flat["loss/mean"] = mean.double().reshape(-1).numpy()
    flat["loss/cov"] = cov.double().reshape(-1).numpy()
    t = cub.condGANTrainer.__new__(cub.condGANTrainer)
    cls = torch.from_numpy(detfill.randint("l:cls", (9,), 0, 200))
    flat["loss/onehot"] = t.onehot(cls, 201).argmax(1).double().numpy()
    flat["loss/onehot_sum"] = np.array([float(t.onehot(cls, 201).sum())])
    return flat


def main():
    """python -m oracle.gen_golden [--only name,name,...]  (--only: just those step cases, modules.npz untouched)"""
    os.makedirs(GOLD, exist_ok=True)
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    for name, B, width, iters in CASES:
        if only is not None and name not in only:
            continue
        flat = run_case(name, B, width, iters)
        path = os.path.join(GOLD, "step_%s_b%d_w%d.npz" % (name, B, width))
        np.savez_compressed(path, **flat)
        print(path, len(flat), os.path.getsize(path))
    if only is not None:
        return
    flat = module_cases()
    path = os.path.join(GOLD, "modules.npz")
    np.savez_compressed(path, **flat)
    print(path, len(flat), os.path.getsize(path))


if __name__ == "__main__":
    main()
