"""Multi-GPU plumbing: one process per GPU, envs block-partitioned across ranks, no collective inside a step.

The only reductions on this path: (1) the four episode-statistics sums per ``runner.run()`` (done
inside ``CudaVecRunner._finish_run``) and (2) a learner-side flat gradient all-reduce, provided here for callers
that keep a replica of the agent on every rank.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import torch as th
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if th.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            th.cuda.set_device(local_rank)
            kw["device_id"] = th.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def env_partition(total_envs, rank, world):
    """Contiguous block of environments owned by ``rank`` (the first ``total % world`` ranks get one extra)."""
    base, extra = divmod(total_envs, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def all_reduce_gradients(parameters, average=True):
    """One flat-buffer all-reduce of every .grad (the agent has ~40 k parameters: latency-bound, so one bucket)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = th.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def broadcast_parameters(module, src=0):
    """Replicate the agent weights of rank ``src`` (after load / checkpoint restore)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
