"""One process per GPU: gradient averaging over NCCL (NVLink 5 / NVSwitch) replaces the reference's single-process
nn.DataParallel (cub_trainer_splitz_cap_ca.py:139,163).  Pure data parallelism: per-rank batches, per-replica
BatchNorm statistics (no SyncBN -- DataParallel semantics), one all-reduce(mean) of each network's flat gradient
buffer per optimiser update; weights stay bit-identical across ranks after an initial broadcast from rank 0.
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


class TailAllreduce:
    """Gradient all-reduce of one network, started from the END of its flat gradient buffer while backward is still
    running (opt-in: EKL_BUCKET_AR=1, N > 1).

    Parameters sit in the flat buffer in registration order = forward order, so the deepest layers -- whose gradients
    backward produces FIRST, and which hold most of a discriminator's weights (the 1024->2048 4x4 and 2048->1024 3x3
    filters are 52 M of JOINT_D_NET256's 73 M parameters) -- form its tail.  The network's forward marks block boundaries
    (ops.grad_mark); when the gradient of a marked activation arrives, everything registered after that block is final:
    the slice [end of that block, start of what is already in flight) goes out as an asynchronous all-reduce that
    overlaps the rest of backward.  finish() sends the remaining head slice and makes the current stream wait for all
    of them.  Collectives are issued in the same order on every rank (same graph, same marks)."""

    def __init__(self, net, params, offsets, total, flat, min_bytes=4 << 20):
        self.flat, self.total = flat, total
        self.min_elems = max(1, min_bytes // 4)
        end_of_param = {id(p): o + p.numel() for p, o in zip(params, offsets)}
        self.end_of = {}
        for m in net.modules():
            ends = [end_of_param[id(p)] for p in m.parameters() if id(p) in end_of_param]
            if ends:
                self.end_of[id(m)] = (max(ends) + 3) // 4 * 4 if max(ends) < total else total
        self.ws = dist.get_world_size()
        self.avg = dist.get_backend() == "nccl"
        self.active, self.lo, self.works = False, total, []

    def begin(self):
        """Call right before backward of this network's own update (marks fired at any other time are ignored)."""
        self.active, self.lo, self.works = True, self.total, []

    def on_mark(self, after_module):
        if not self.active:
            return
        lo = self.end_of.get(id(after_module))
        if lo is None or self.lo - lo < self.min_elems:
            return
        self._launch(lo, self.lo)
        self.lo = lo

    def _launch(self, lo, hi):
        t = self.flat[lo:hi]
        w = dist.all_reduce(t, op=dist.ReduceOp.AVG if self.avg else dist.ReduceOp.SUM, async_op=True)
        self.works.append((w, lo, hi))

    def finish(self):
        """Reduce what is left (the head of the buffer) and wait for every slice: afterwards the whole buffer holds the
        mean over ranks, exactly as one all-reduce of the flat buffer would."""
        self.active = False
        if self.lo > 0:
            self._launch(0, self.lo)
            self.lo = 0
        for w, lo, hi in self.works:
            w.wait()
            if not self.avg:
                self.flat[lo:hi].div_(self.ws)
        n = len(self.works)
        self.works = []
        return n
