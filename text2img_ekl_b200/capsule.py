"""CapsuleLinear -- stand-in for the un-vendored third-party `capsule_layer.modules.CapsuleLinear` the reference
imports (model.py:12; call sites :248,290,301,943,1082).  *Parity unpinned*: the package is absent from the
reference tree and its version is not pinned anywhere, so the arithmetic follows the published routing-by-agreement
algorithm as restated (and documented) in oracle/capsule_ref.py; this module is checked against that restatement.

Shared-weight mode only (what the reference uses): weight [out_capsules, out_length, in_length], input
[B, in_capsules, in_length] -> [B, out_capsules, out_length].

The [B, O, I, L] prior tensor (201 MB at B=32 for the generator stem) is never formed.  Because
prior[b,o,i,:] = W[o] x[b,i], every routing quantity lives in in_length space:
    logit[b,o,i] = <x[b,i], u[b,o]>,  u[b,o] = W[o]^T (sum of earlier v[b,o])
    s[b,o]       = W[o] y[b,o],       y[b,o] = sum_i softmax_o(logit)[b,o,i] x[b,i]
(property pinned on the CPU by tests/test_oracle_golden.py::test_capsule_reduced_algebra_equals_materialised_priors).
"""
import torch
import torch.nn as nn

EPS = 1e-8


def squash(s):
    n2 = (s * s).sum(dim=-1, keepdim=True)
    return s * (n2 / (1.0 + n2) / torch.sqrt(n2 + EPS))


def capsule_linear(x, weight, routing_type="dynamic", num_iterations=3):
    x = x.float()
    B, O = x.shape[0], weight.shape[0]
    if routing_type == "dynamic":
        vsum = x.new_zeros(B, O, weight.shape[1])
        v = None
        for r in range(num_iterations):
            if r == 0:
                # logits are zero: softmax over O is uniform 1/O
                y = (x.sum(dim=1, keepdim=True) / O).expand(B, O, x.shape[2])
            else:
                u = torch.einsum("olk,bol->bok", weight, vsum)
                c = torch.softmax(torch.einsum("bik,bok->boi", x, u), dim=1)
                y = torch.einsum("boi,bik->bok", c, x)
            v = squash(torch.einsum("olk,bok->bol", weight, y))
            vsum = vsum + v
        return v
    if routing_type == "k_means":
        out = torch.einsum("olk,bk->bol", weight, x.mean(dim=1))
        for r in range(num_iterations):
            u = torch.einsum("olk,bol->bok", weight, torch.nn.functional.normalize(out, dim=-1))
            c = torch.softmax(torch.einsum("bik,bok->boi", x, u), dim=1)
            out = torch.einsum("olk,bok->bol", weight, torch.einsum("boi,bik->bok", c, x))
        return out
    raise ValueError(routing_type)


class CapsuleLinear(nn.Module):
    ROUTING_TYPE = "dynamic"
    NUM_ITERATIONS = 3

    def __init__(self, out_capsules, in_length, out_length, in_capsules=None, share_weight=True, routing_type=None,
                 num_iterations=None, **kwargs):
        super().__init__()
        if in_capsules is not None or not share_weight:
            raise ValueError("only the shared-weight mode used by the reference is implemented")
        self.out_capsules, self.in_length, self.out_length = out_capsules, in_length, out_length
        self.routing_type = routing_type or self.ROUTING_TYPE
        self.num_iterations = num_iterations or self.NUM_ITERATIONS
        self.weight = nn.Parameter(torch.empty(out_capsules, out_length, in_length))
        nn.init.xavier_uniform_(self.weight)
        self.bias = None          # read by the reference's weights_init (class name contains "Linear")

    def forward(self, input):
        return capsule_linear(input, self.weight, self.routing_type, self.num_iterations)

    def extra_repr(self):
        return "out_capsules=%d, in_length=%d, out_length=%d, routing=%s x%d" % (
            self.out_capsules, self.in_length, self.out_length, self.routing_type, self.num_iterations)
