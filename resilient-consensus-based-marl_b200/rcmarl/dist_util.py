"""Data-parallel plumbing: one process per GPU, environments sharded, parameters replicated,
ONE all-reduce of the packed UNSCALED sums per optimisation step (the division by the global batch size happens
after the reduce, inside the apply kernels)."""
import os

import torch


def shard_envs(n_total, rank, world):
    """Contiguous shard of the environment batch owned by `rank`: (first global env index, count)."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_sums(t, world, group=None):
    """In-place SUM all-reduce of a packed sums tensor (NCCL on GPUs, gloo in the CPU tests)."""
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def init_from_env(backend=None):
    """torchrun contract: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT.  Returns (rank, world, local_rank)."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank
