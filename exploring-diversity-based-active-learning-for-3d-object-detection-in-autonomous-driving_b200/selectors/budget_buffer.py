"""Cumulative-budget bookkeeping of the selection buffer.

The reference keeps one JSON object per experiment, ``{"0": [], "<cumulative budget>": [frame index, ...]}``
(tools/active_select.py:68-70 writes the initial ``{"0": []}``); every ``active_select`` invocation reads it,
selects under ``budget`` MORE cost units and adds one key (det3d/selectors/base_selector.py:34,40,47-49,60-61).
The helpers here are pure functions over that dict so that the host tests can pin them without a selector.
"""
import os
from typing import Dict, List, Sequence


def latest_key(buffer: Dict[str, List[int]]) -> str:
    """Key of the most recent round: the numerically largest one (base_selector.py:47-49)."""
    return str(max(map(int, buffer)))


def next_key(buffer: Dict[str, List[int]], budget: int) -> str:
    """Key this round will be stored under: previous cumulative budget + ``budget`` (base_selector.py:40)."""
    return str(budget + int(latest_key(buffer)))


def subset_path(infos_file: str, key: str) -> str:
    """``<infos>_<key><ext>`` -- same ``str.replace`` rule as base_selector.py:69-70 (every occurrence of the
    extension inside the path is rewritten, which the consumers rely on for ``.../x.pkl`` only)."""
    ext = os.path.splitext(infos_file)[-1]
    return infos_file.replace(ext, f"_{key}{ext}")


def annotation_cost(infos: Sequence[dict], frames: Sequence[int], cost_f: float, cost_b: float):
    """``cost_f * #frames`` then ``+ #boxes * cost_b`` frame by frame, python floats in this exact order
    (base_selector.py:78-86): the greedy loop continues this running sum, so the order is part of parity."""
    total = 0
    total += cost_f * len(frames)
    for i in frames:
        total += infos[i]["gt_names"].shape[0] * cost_b
    return total


def settle_allocator(n_frames: int) -> None:
    """Finish, right after the unpickle that caused it, the allocator work the unpickle leaves behind.

    Unpickling the infos list frees one temporary per frame; glibc parks them in its unsorted bin and files at most 10,000
    of them per ``malloc`` call that has to look there, so each of the next ``n_frames / 10,000`` medium-sized allocations --
    wherever they happen -- costs 2.5-3 ms (45 ms after a 225,040-frame pool, measured; ``malloc_trim`` does not help).  A few
    throw-away allocations make the allocator do that filing here, in the constructor, next to the 3 s load it belongs to,
    instead of inside the first calls of ``select_samples``.  glibc only; anything else is left alone."""
    if n_frames < 10000:
        return
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.malloc.restype = ctypes.c_void_p
        libc.malloc.argtypes = [ctypes.c_size_t]
        libc.free.argtypes = [ctypes.c_void_p]
        for _ in range(n_frames // 5000 + 8):
            libc.free(libc.malloc(72000))
    except (OSError, AttributeError):
        pass
