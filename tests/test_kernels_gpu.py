"""Every CUDA kernel family in isolation, through the C ABI, against a torch fp32 reference of the same op on the
same (bf16-representable) inputs: rel-L2 <= 6e-3 (one bf16 output rounding) for conv forward / data-gradient and
BatchNorm+activation forward / backward, <= 1e-4 for fp32 outputs (weight gradients, BN parameter gradients,
running statistics).  tcgen05 and SIMT implementations are both checked, so each also cross-checks the other."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("group", ["tc_fwd", "tc_dgrad", "tc_wgrad", "simt_fwd", "simt_dgrad", "simt_wgrad", "bn", "misc", "heads", "tc_split", "fold", "vc", "route"])
def test_kernel_group(group, capsys):
    import kernel_check
    nfail = kernel_check.run_group(group)
    out = capsys.readouterr().out
    print(out)
    assert nfail == 0, out


def test_capsule_reduced_algebra_matches_materialised_restatement_on_gpu():
    from oracle import capsule_ref
    from text2img_ekl_b200 import capsule
    torch.manual_seed(0)
    for (B, I, K, O, Lh) in [(4, 48, 8, 1024, 32), (6, 16, 512, 201, 16)]:
        x = torch.randn(B, I, K, device="cuda")
        w = (torch.randn(O, Lh, K, device="cuda") / K ** 0.5).requires_grad_(True)
        w2 = w.detach().clone().requires_grad_(True)
        for routing in ("dynamic", "k_means"):
            a = capsule.capsule_linear(x, w, routing, 3)
            b = capsule_ref.capsule_linear(x, w2, routing, 3)
            assert float((a - b).norm() / b.norm()) < 1e-4
            g = torch.randn_like(a)
            (ga,) = torch.autograd.grad(a, w, g)
            (gb,) = torch.autograd.grad(b, w2, g)
            assert float((ga - gb).norm() / gb.norm()) < 1e-3


def test_fails_loudly_without_library(monkeypatch):
    from text2img_ekl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libekl_b200.so")
    with pytest.raises(_lib.EklError):
        _lib.lib()
