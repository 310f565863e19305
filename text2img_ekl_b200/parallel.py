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
