"""Task-sharded data parallelism for meta-training (the one strategy the build adds).

MAML tasks are independent given theta, so task t lives on rank ``t mod world``; the only
exchange is ONE all-reduce (sum) per meta-step of the flat buffer [meta-gradient | meta-loss],
after which every rank applies the identical fused clip+AdamW (deterministic replicas, no
broadcast).  One process per GPU, ``torch.distributed`` over NCCL; the same code runs over gloo
on CPU tensors for the world_size-2 tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_tasks(num_tasks, rank, world, require_even=False):
    if require_even and num_tasks % world != 0:
        raise ValueError(f"{num_tasks} tasks do not divide evenly over {world} ranks")
    return list(range(rank, num_tasks, world))


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def allreduce_meta(buf, group=None):
    """Sum the packed [grad | loss] buffer over ranks in place (one collective)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
