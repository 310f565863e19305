"""How far can ANY bf16-storage implementation be from the fp32 reference?  (CPU, no GPU.)

The fp32 oracle is evaluated twice on the same step: plainly, and with bf16 rounding applied only at the tensors the
CUDA implementation stores in bf16 (feature maps, their gradients, conv filter operands; accumulation, BatchNorm
statistics and parameters stay fp32).  Forward quantities move by ~1e-2 (the north-star bound), but parameter
gradients of this GAN step move by 5-30 %: train-mode BatchNorm over tiny batches, GLU gates and the saturating
discriminator losses amplify 2^-9 perturbations chaotically.  Per-tensor rel-L2 <= 1e-2 on gradients is therefore not
a property an implementation can have; tests/test_step_parity_gpu.py bounds the CUDA path's gradient deviation by
this measured floor instead, and checks every backward kernel in isolation at 1e-2 (tests/test_kernels_gpu.py)."""
import numpy as np
import torch

from oracle import configs as ocfg, shapes, synth
from oracle import ekl_oracle as O


def run(name, B, mode):
    oc = ocfg.oracle_cfg(name, batch=B)
    gsh = shapes.g_shapes(oc, cond_dim=ocfg.cond_dim(oc))
    dsh = [shapes.d_shapes(oc, r, True, oc.D_CAPSULE) for r in [64, 128, 256][: oc.BRANCH_NUM]]
    tr = O.OracleTrainer(oc, shapes.make_state_dict(gsh, "G"), [shapes.make_state_dict(s, "D%d" % i) for i, s in enumerate(dsh)])
    with O.storage(mode):
        return tr.step(**synth.make_batch(oc, B, "it0"))


def rel(a, b):
    return float((a.detach() - b.detach()).norm() / (b.detach().norm() + 1e-30))


def test_bf16_storage_floor_on_fp32_oracle():
    a, b = run("catcls", 4, "fp32"), run("catcls", 4, "bf16")
    img = [rel(x, y) for x, y in zip(b["fake_imgs"], a["fake_imgs"])]
    assert max(img) < 2e-2                                     # forward: within the bf16 bound
    assert rel(b["errG"], a["errG"]) < 1e-2
    g = np.median([rel(b["gradG"][k], a["gradG"][k]) for k in a["gradG"]])
    d = [np.median([rel(b["gradD"][i][k], a["gradD"][i][k]) for k in a["gradD"][i]]) for i in range(2)]
    print("bf16-storage floor: images %s, gradG median %.2f, gradD median %s" % (img, g, d))
    assert g > 5e-2 and min(d) > 1e-2                          # gradients: far beyond 1e-2 with NO implementation involved
    assert g < 1.0                                             # ...but still correlated (not garbage)


def test_storage_mode_is_off_by_default_and_reversible():
    assert O.STORAGE == "fp32"
    with O.storage("bf16"):
        assert O.STORAGE == "bf16"
        x = torch.tensor([1.0 + 2 ** -10])
        assert float(O._q(x)) == 1.0
    assert O.STORAGE == "fp32" and float(O._q(torch.tensor([1.0 + 2 ** -10]))) != 1.0
