"""Rank helpers (det3d/torchie/trainer/utils.py:22-43): ``get_dist_info`` and ``@master_only``."""
import functools


def get_dist_info():
    """``(rank, world_size)`` of the default torch.distributed group, ``(0, 1)`` when not initialised."""
    import sys
    dist = sys.modules.get("torch.distributed")
    if dist is None:         # nobody imported it, so no process group can exist: a single-process run (the CLI's one-shot
        return 0, 1          # selection) does not pay the import of torch for this question
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def master_only(func):
    """Run ``func`` on rank 0 only (other ranks return None), base_selector.py:57."""

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        rank, _ = get_dist_info()
        if rank == 0:
            return func(*args, **kwargs)
        return None

    return wrapper


def _group_device():
    """Device on which small control tensors live for the default group's backend (NCCL needs CUDA tensors)."""
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def agree_from_rank0(values):
    """Rank 0's tuple of integers on every rank (``values`` unchanged when not distributed).  The sharded engines do
    the greedy bookkeeping redundantly on every rank, so anything drawn from a per-process RNG (the first frame of an
    empty buffer, base_selector.py / spatial_temporal_selector.py:168) must be rank 0's draw everywhere."""
    rank, world = get_dist_info()
    if world <= 1:
        return tuple(values)
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=_group_device())
    dist.broadcast(t, src=0)
    return tuple(int(v) for v in t.tolist())


def all_ranks_max(value: int) -> int:
    """Largest ``value`` over the ranks: lets every rank take the same branch (retry / raise) after a collective call."""
    rank, world = get_dist_info()
    if world <= 1:
        return int(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(value)], dtype=torch.int64, device=_group_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def check_same_on_all_ranks(values, what: str):
    """Raise on every rank when the ranks do not hold the same integers (cheap guard before a collective kernel)."""
    rank, world = get_dist_info()
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=_group_device())
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        raise RuntimeError(f"ranks disagree on {what}: min {lo.tolist()} max {hi.tolist()} (this rank {t.tolist()})")
