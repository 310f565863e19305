"""One process per GPU: gradient averaging over NCCL (NVLink 5 / NVSwitch) replaces the reference's single-process
nn.DataParallel (cub_trainer_splitz_cap_ca.py:139,163).  Pure data parallelism: per-rank batches, per-replica
BatchNorm statistics (no SyncBN -- DataParallel semantics), the mean over ranks of each network's flat gradient
buffer per optimiser update (GradReducer: bf16 payload, sliced from the tail of the buffer so that it overlaps that
network's backward); weights stay bit-identical across ranks after an initial broadcast from rank 0 because every rank
applies the same averaged gradients with the same optimiser state.
"""
import os

import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return world()
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend, init_method="env://")
    return world()


def broadcast_params(nets, src=0):
    """Make every replica start from rank `src`'s weights and buffers (DataParallel replicates GPU0's)."""
    if world()[1] == 1:
        return
    for net in nets:
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, src)


def make_allreduce():
    """Returns callable(flat_grad) averaging it over all ranks in place, or None for a single process."""
    rank, ws = world()
    if ws == 1:
        return None

    def allreduce(flat):
        if dist.get_backend() == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(ws)
    return allreduce


def shard_range(n_units, rank=None, ws=None):
    """Contiguous [lo, hi) share of n_units for this rank (used by the data side: each rank draws its own batches)."""
    r, w = world()
    rank = r if rank is None else rank
    ws = w if ws is None else ws
    per, rem = divmod(n_units, ws)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def grad_comm_mode():
    """Payload type of the gradient all-reduce: 'bf16' (default: half the NVLink bytes; the optimiser reads the averaged
    bf16 gradients directly) or 'fp32' (EKL_GRAD_COMM=fp32: the flat fp32 buffer is reduced in place)."""
    mode = os.environ.get("EKL_GRAD_COMM", "bf16")
    if mode not in ("bf16", "fp32"):
        raise ValueError("EKL_GRAD_COMM must be bf16 or fp32, got %r" % mode)
    return mode


class GradReducer:
    """Gradient average of ONE network over the ranks, overlapped with that network's backward.

    Parameters sit in the flat gradient buffer in registration order = forward order, so the deepest layers -- whose
    gradients backward produces FIRST, and which hold most of a discriminator's weights (the 1024->2048 4x4 and
    2048->1024 3x3 filters are 52 M of JOINT_D_NET256's 73 M parameters) -- form its tail.  The network's forward marks
    block boundaries (ops.grad_mark); when the gradient of a marked activation arrives, everything registered after that
    block is final: the slice [end of that block, start of what is already in flight) is rounded to bf16 into a staging
    buffer and all-reduced on a side stream while the rest of backward keeps the main stream busy.  finish() sends the
    remaining head slice, joins the side stream and returns the buffer the optimiser must read (the bf16 staging buffer,
    or the fp32 flat buffer in fp32 mode).  A network without marks (the generator: its largest filters are the FIRST
    layers, final only when backward ends) is reduced by finish() alone.

    Collectives are issued in the same order on every rank (same program, same marks) on one communicator; inside a
    captured step they become graph nodes on the side stream."""

    def __init__(self, flat, net=None, params=None, offsets=None, mode=None, min_bytes=4 << 20):
        self.flat, self.total = flat, flat.numel()
        self.mode = mode or grad_comm_mode()
        self.ws = dist.get_world_size()
        self.avg = dist.get_backend() == "nccl"
        self.stage = torch.zeros(self.total, dtype=torch.bfloat16, device=flat.device) if self.mode == "bf16" else None
        self.min_elems = max(4, min_bytes // 4)
        self.end_of = {}
        if net is not None and params is not None:
            end_of_param = {id(p): o + p.numel() for p, o in zip(params, offsets)}
            for m in net.modules():
                ends = [end_of_param[id(p)] for p in m.parameters() if id(p) in end_of_param]
                if ends:
                    self.end_of[id(m)] = (max(ends) + 3) // 4 * 4 if max(ends) < self.total else self.total
        self.side = torch.cuda.Stream() if flat.is_cuda else None
        self.active, self.lo, self.slices = False, self.total, []
        self.after_slice = None          # optional callback(lo, hi), run on the side stream behind the slice's all-reduce

    def begin(self):
        """Call right before backward of this network's own update (marks fired at any other time are ignored)."""
        self.active, self.lo, self.slices = True, self.total, []

    def on_mark(self, after_module):
        if not self.active:
            return
        lo = self.end_of.get(id(after_module))
        if lo is None or self.lo - lo < self.min_elems:
            return
        self._send(lo, self.lo)
        self.lo = lo

    def _reduce(self, lo, hi):
        src = self.flat[lo:hi]
        if self.stage is not None:
            dst = self.stage[lo:hi]
            if src.is_cuda:
                from . import _lib as L
                from . import ops
                L.check(L.lib().ekl_cast_bf16(L.ptr(src), L.ptr(dst), hi - lo, L.stream()))
                ops._count()
            else:
                dst.copy_(src)              # host tensors: the gloo tests of this class
            src = dst
        if self.avg:
            dist.all_reduce(src, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(src, op=dist.ReduceOp.SUM)
            src.div_(self.ws)

    def _send(self, lo, hi):
        self.slices.append((lo, hi))
        if self.side is None:
            self._reduce(lo, hi)
            if self.after_slice is not None:
                self.after_slice(lo, hi)
            return
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)              # the slice's gradients are complete on the issuing stream
        with torch.cuda.stream(self.side):
            self._reduce(lo, hi)
            if self.after_slice is not None:
                self.after_slice(lo, hi)        # e.g. the optimiser update of the slice (engine.TailUpdate)

    def finish(self):
        """Reduce what is left (the head of the buffer), wait for every slice; returns (gradient buffer, is_bf16): the
        whole buffer now holds the mean over ranks, exactly as one all-reduce of the flat buffer would."""
        self.active = False
        if self.lo > 0:
            self._send(0, self.lo)
            self.lo = 0
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)
        return (self.stage, True) if self.stage is not None else (self.flat, False)


def make_reducer(flat, net=None, params=None, offsets=None):
    """GradReducer for a network's flat gradient buffer, or None for a single process."""
    return GradReducer(flat, net, params, offsets) if world()[1] > 1 else None
