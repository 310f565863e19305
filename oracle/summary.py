"""Flatten a step's outputs into small named float64 vectors (fixture format) -- TEST INFRASTRUCTURE."""
import numpy as np
import torch

from . import detfill

NSAMP = 24


def tsum(name, t):
    """[sum, l2, NSAMP sampled entries at deterministic positions]."""
    a = t.detach().double().reshape(-1).cpu().numpy()
    idx = detfill.randint("samp:" + name, (NSAMP,), 0, max(a.size, 1)) % max(a.size, 1)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum())], a[idx]])


def summarize(out, params=None):
    """out: dict from OracleTrainer.step / RefStepper.step -> {key: float64 vector}."""
    s = {}
    for i, e in enumerate(out["errD"]):
        s["errD%d" % i] = e.double().numpy()
    s["errG"] = out["errG"].double().numpy()
    s["mu"] = tsum("mu", out["mu"])
    s["real_cp_argmax"] = out["real_cp"].argmax(1).double().numpy()
    for i, (h, f) in enumerate(zip(out["h_codes"], out["fake_imgs"])):
        s["h%d" % i] = tsum("h%d" % i, h)
        s["img%d" % i] = tsum("img%d" % i, f)
    for i, trio in enumerate(out["d_logits"]):
        for j, lg in enumerate(trio):
            for q, t in enumerate(lg):
                s["dlog%d_%d_%d" % (i, j, q)] = t.double().reshape(-1).numpy() if t.dim() == 1 else tsum("dl", t)
    for i, lg in enumerate(out["g_logits"]):
        for q, t in enumerate(lg):
            s["glog%d_%d" % (i, q)] = t.double().reshape(-1).numpy() if t.dim() == 1 else tsum("gl", t)
    for k, g in out["gradG"].items():
        s["gG:" + k] = tsum(k, g)[:6]
    for i, gd in enumerate(out["gradD"]):
        for k, g in gd.items():
            s["gD%d:%s" % (i, k)] = tsum(k, g)[:6]
    if params is not None:
        for tag, sd in params.items():
            tot = 0.0
            for k, v in sd.items():
                if torch.is_floating_point(v):
                    tot += float(v.detach().double().pow(2).sum())
            s["pnorm:" + tag] = np.array([np.sqrt(tot)])
    return s
