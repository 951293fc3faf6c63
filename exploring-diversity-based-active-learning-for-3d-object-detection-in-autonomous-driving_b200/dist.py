"""Rank helpers (det3d/torchie/trainer/utils.py:22-43): ``get_dist_info`` and ``@master_only``."""
import functools


def get_dist_info():
    """``(rank, world_size)`` of the default torch.distributed group, ``(0, 1)`` when not initialised."""
    try:
        import torch.distributed as dist
    except ImportError:  # pragma: no cover - torch is part of the image
        return 0, 1
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
