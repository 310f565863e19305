"""CapsuleLinear restatement -- TEST INFRASTRUCTURE, *PARITY UNPINNED*.

The reference imports `capsule_layer.modules.CapsuleLinear` (model.py:12; call sites
model.py:248,290,301,943,1082), PyPI package "CapsuleLayer" (github leftthomas/CapsuleLayer).
That package is NOT vendored in /root/reference, no version is pinned anywhere in the
reference (no requirements/setup/lock file) and it is not installed here, so this file
restates the published algorithm (Sabour et al. 2017 routing-by-agreement, as exposed by
that package's `routing_type='dynamic'`; and its `'k_means'` variant) from the constraints
the reference's own call sites impose:

  * shared-weight mode (in_capsules=None): weight [out_capsules, out_length, in_length]
    (state-dict shapes [1024,32,8] and [201,16,512] are implied by model.py:248,943);
  * input [B, in_capsules, in_length] -> ONE tensor [B, out_capsules, out_length]
    (it sits inside nn.Sequential before Reshape, model.py:248-251; .norm(dim=-1) of it is
    taken at model.py:969-970);
  * the module has a `.bias` attribute (read by weights_init, cub_trainer_splitz_cap_ca.py:74-77,
    because the class name contains "Linear"); here bias is None (no bias term).

Algorithm (routing_type='dynamic', num_iterations=3):
    prior[b,o,i,:] = weight[o] @ x[b,i,:]                      # [B,O,I,L]
    logit[b,o,i]   = 0
    repeat num_iterations times (r = 0..n-1):
        c      = softmax(logit, over the OUT-capsule axis o)   # each in-capsule distributes itself
        s[b,o] = sum_i c[b,o,i] * prior[b,o,i,:]
        v[b,o] = squash(s[b,o]) = |s|^2/(1+|s|^2) * s/|s|
        if r < n-1: logit[b,o,i] += <prior[b,o,i,:], v[b,o,:]>
    return v
routing_type='k_means' (cosine similarity): out = mean_i prior; repeat: logit = <prior, normalise(out)>,
c = softmax(logit over o), out = sum_i c*prior; no squash.
No reference test pins any value at this boundary => parity for capsule configs is defined
against THIS restatement.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

EPS = 1e-8


def squash(s, dim=-1):
    n2 = (s * s).sum(dim=dim, keepdim=True)
    return s * (n2 / (1.0 + n2) / torch.sqrt(n2 + EPS))


def capsule_linear(x, weight, routing_type="dynamic", num_iterations=3):
    """x [B,I,K], weight [O,L,K] -> [B,O,L]; materialises priors exactly as the package does."""
    prior = torch.einsum("olk,bik->boil", weight, x)  # [B,O,I,L]
    if routing_type == "dynamic":
        logit = prior.new_zeros(prior.shape[:3])
        v = None
        for r in range(num_iterations):
            c = F.softmax(logit, dim=1)
            s = (c.unsqueeze(-1) * prior).sum(dim=2)
            v = squash(s)
            if r != num_iterations - 1:
                logit = logit + (prior * v.unsqueeze(2)).sum(dim=-1)
        return v
    elif routing_type == "k_means":
        out = prior.mean(dim=2)
        for r in range(num_iterations):
            logit = (prior * F.normalize(out, dim=-1).unsqueeze(2)).sum(dim=-1)
            c = F.softmax(logit, dim=1)
            out = (c.unsqueeze(-1) * prior).sum(dim=2)
        return out
    raise ValueError(routing_type)


class CapsuleLinear(nn.Module):
    """Stand-in with the constructor signature the reference uses."""
    ROUTING_TYPE = "dynamic"
    NUM_ITERATIONS = 3

    def __init__(self, out_capsules, in_length, out_length, in_capsules=None, share_weight=True,
                 routing_type=None, num_iterations=None, **kwargs):
        super().__init__()
        if in_capsules is not None or not share_weight:
            raise ValueError("only the shared-weight mode used by the reference is restated")
        self.out_capsules, self.in_length, self.out_length = out_capsules, in_length, out_length
        self.routing_type = routing_type or self.ROUTING_TYPE
        self.num_iterations = num_iterations or self.NUM_ITERATIONS
        self.weight = nn.Parameter(torch.empty(out_capsules, out_length, in_length))
        nn.init.xavier_uniform_(self.weight)
        self.bias = None

    def forward(self, input):
        return capsule_linear(input, self.weight, self.routing_type, self.num_iterations)
