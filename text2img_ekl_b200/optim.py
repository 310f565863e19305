"""Adam for the B200 step: torch.optim.Adam semantics (the reference's define_optimizers, cub_trainer_splitz_cap_ca.py:
199-215) executed as ONE fused kernel per network over flat buffers (include/ekl_b200.h: ekl_adam_step).

The optimiser re-homes every parameter of the network into one flat fp32 buffer (shapes, strides / channels_last
layouts and state_dict contents unchanged: each parameter becomes a view), keeps exp_avg / exp_avg_sq flat, and writes a
bf16 shadow of the updated parameters in the same pass.  ops.ConvSpec uses the shadow slice of a channels_last
stride-1 / stride-2 conv filter directly as its packed forward operand."""
import torch

from . import _lib as L
from . import ops


def _pad8(n):
    return (n + 7) // 8 * 8


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.plist = params
        dev = params[0].device
        # every parameter starts at a multiple of 8 elements: 32-byte aligned fp32 views and 16-byte aligned bf16 shadows
        # (a shadow slice is handed to TMA as a packed filter operand: tensor maps need 16-byte aligned bases)
        self.offsets, off = [], 0
        for p in params:
            self.offsets.append(off)
            off += _pad8(p.numel())
        self.n = off
        self.flat_p = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.shadow = torch.zeros(self.n, device=dev, dtype=torch.bfloat16)
        self.exp_avg = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.state_dev = torch.zeros(3, device=dev, dtype=torch.float32)
        self.flat_g = None
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                view = self._view(self.flat_p, p, o)
                view.copy_(p.data)
                p.data = view
                p._ekl_shadow = self.shadow[o:o + p.numel()]
                self.state[p] = dict(step=self.state_dev[0], exp_avg=self._view(self.exp_avg, p, o),
                                     exp_avg_sq=self._view(self.exp_avg_sq, p, o))
            self.shadow.copy_(self.flat_p)

    @staticmethod
    def _view(flat, p, off):
        g = flat[off:off + p.numel()]
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            return g.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
        return g.view(p.shape)

    def refresh_shadow(self):
        """Re-derive the bf16 shadow from the fp32 masters (after the parameters were written from outside)."""
        with torch.no_grad():
            self.shadow.copy_(self.flat_p)
        ops.mark_dirty(self.plist)

    def load_state_dict(self, state_dict):
        """Accepts the state_dict of torch.optim.Adam or of this class (same format: per-parameter `step`, `exp_avg`,
        `exp_avg_sq`, indexed in parameter order).  The moments are copied INTO the flat buffers: the views held by
        self.state, and any captured graph, keep pointing at live storage."""
        ids = [i for g in state_dict["param_groups"] for i in g["params"]]
        if len(ids) != len(self.plist):
            raise ValueError("optimizer state has %d parameters, this optimizer %d" % (len(ids), len(self.plist)))
        step = None
        with torch.no_grad():
            for idx, p in zip(ids, self.plist):
                st = state_dict["state"].get(idx)
                if st is None:
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError("optimizer state %d has shape %s, parameter %s" % (idx, tuple(st["exp_avg"].shape), tuple(p.shape)))
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                step = float(st["step"]) if step is None else step
            if step is not None:
                self.state_dev[0] = step          # the kernel derives both bias corrections from the count
        grp, src = self.param_groups[0], state_dict["param_groups"][0]
        for k in ("lr", "betas", "eps"):
            if k in src:
                grp[k] = tuple(src[k]) if k == "betas" else src[k]

    def make_flat_grads(self):
        """Gradient buffer with the parameters' offsets; every p.grad becomes a view with the parameter's layout."""
        self.flat_g = torch.zeros(self.n, device=self.flat_p.device, dtype=torch.float32)
        for p, o in zip(self.plist, self.offsets):
            p.grad = self._view(self.flat_g, p, o)
        return self.flat_g

    @torch.no_grad()
    def step(self, closure=None, grads_bf16=None):
        """grads_bf16: optional bf16 buffer in the flat layout holding the gradients to apply instead of the fp32 flat
        buffer (the rank-averaged staging buffer of parallel.GradReducer)."""
        if self.flat_p.device.type != "cuda":
            raise L.EklError("FlatAdam.step drives the CUDA kernel library (ekl_adam_step): there is no CPU path; "
                             "use torch.optim.Adam for CPU tensors")
        if self.flat_g is None or any(p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o
                                      for p, o in zip(self.plist, self.offsets)):
            # gradients live elsewhere (stand-alone use): gather them into the flat layout
            old = [p.grad for p in self.plist]
            self.make_flat_grads()
            for p, g in zip(self.plist, old):
                if g is not None:
                    p.grad.copy_(g)
        self.tick()
        self.apply_slice(0, self.n, grads_bf16)
        self.finish_step()

    # ---- the same step in pieces (engine.TailUpdate: the optimiser overlaps the backward pass)
    @torch.no_grad()
    def tick(self):
        """Advance the step count / bias corrections once (device side, capturable); apply_slice() calls follow."""
        grp = self.param_groups[0]
        b1, b2 = grp["betas"]
        L.check(L.lib().ekl_adam_tick(L.ptr(self.state_dev), float(b1), float(b2), L.stream()))
        ops._count()

    @torch.no_grad()
    def apply_slice(self, lo, hi, grads_bf16=None):
        """Adam update of the flat slice [lo, hi) (multiples of 4) from the fp32 flat gradients, or from `grads_bf16`, a
        bf16 buffer in the flat layout (the rank-averaged staging buffer of parallel.GradReducer)."""
        if hi <= lo:
            return
        grp = self.param_groups[0]
        b1, b2 = grp["betas"]
        if grads_bf16 is not None:
            assert grads_bf16.dtype == torch.bfloat16 and grads_bf16.numel() == self.n
            g, g16 = grads_bf16, 1
        else:
            g, g16 = self.flat_g, 0
        L.check(L.lib().ekl_adam_apply(L.ptr(self.flat_p[lo:hi]), L.ptr(g[lo:hi]), g16, L.ptr(self.exp_avg[lo:hi]),
                                       L.ptr(self.exp_avg_sq[lo:hi]), L.ptr(self.shadow[lo:hi]), hi - lo, L.ptr(self.state_dev),
                                       float(grp["lr"]), float(b1), float(b2), float(grp["eps"]), L.stream()))
        ops._count()
        ops._acct("adam", nbytes=30.0 * (hi - lo))       # 16 B read (p, g, m, v) + 14 B written (p, m, v, bf16 shadow) per parameter

    def finish_step(self):
        ops.mark_dirty(self.plist)          # the packed data-gradient filter operands are stale now
