"""Host-side extraction of the per-frame arrays the device needs from the infos list.

Replaces the per-frame python loops of det3d/selectors/spatial_temporal_selector.py:78-91 (ego
location), :114-129 (consecutive logfile runs), temporal_selector.py:50-55 (logfile-name groups) and
euclidean_spatial_selector.py:57-60,75-76 (city per frame).  String parsing stays on the host; the
numeric part is vectorised but keeps the reference's own numpy expressions, so the float64 ego
locations are bit-identical to the reference loop (checked on a sample at run time and in tests).
"""
import numpy as np


def logfile_names(infos):
    """``cam_front_path`` basename up to the first "__" (spatial_temporal_selector.py:79).

    Consecutive frames mostly share their directory and logfile, so a frame whose path starts with the previous
    frame's ``<dir>/<logfile>__`` and has no further "/" is given the previous name without re-parsing -- provably
    the same string as ``path.split("/")[-1].split("__")[0]`` (the name has no "/" and no "__" by construction).
    """
    names, prefix, name, plen = [], None, None, 0
    append = names.append
    for info in infos:
        p = info["cam_front_path"]
        if prefix is not None and p.startswith(prefix) and p.find("/", plen) < 0:
            append(name)
            continue
        base = p.split("/")[-1]
        name = base.split("__")[0]
        if base.startswith(name + "__"):
            prefix = p[: len(p) - len(base)] + name + "__"
            plen = len(prefix)
        else:
            prefix = None
        append(name)
    return names


def _location_loop(cals):
    return np.stack([(-(cal[:3, 3].T @ cal[:3, :3]))[:2] for cal in cals])


def ego_locations(infos, check: int = 64) -> np.ndarray:
    """float64 [N,2] ego x,y: ``-(cal[:3,3].T @ cal[:3,:3])[:2]`` with ``cal = car_from_global`` (:83-91).

    The batched ``np.matmul`` goes through the same vector-matrix inner loop as the reference's
    per-frame ``@``; ``check`` frames (spread over the pool) are recomputed with the reference's
    literal expression and must match bit-for-bit, otherwise the literal loop is used for all.
    """
    cals = [info["car_from_global"] for info in infos]
    try:
        stack = np.concatenate(cals).reshape(-1, 4, 4).astype(np.float64, copy=False)   # [N,4,4]; 3x faster than np.stack
    except ValueError:
        stack = np.stack(cals).astype(np.float64, copy=False)
    loc = (-np.matmul(stack[:, None, :3, 3], stack[:, :3, :3]))[:, 0, :2]  # [N,1,3]@[N,3,3]
    loc = np.ascontiguousarray(loc)
    n = len(cals)
    probe = np.unique(np.linspace(0, n - 1, num=min(check, n)).astype(np.int64))
    if not np.array_equal(_location_loop([cals[i] for i in probe]).view(np.uint64), loc[probe].view(np.uint64)):
        loc = _location_loop(cals)
    return loc


def consecutive_runs(names):
    """Run id per frame and the reference's ``max_temporal_distance`` (:114-129).

    A run is a maximal range of consecutive indices sharing one logfile string; the maximum is over
    the LENGTHS of all runs except the last (the reference updates it only when a run ends).
    """
    n = len(names)
    arr = np.asarray(names, dtype=object)
    change = np.ones(n, dtype=bool)
    change[1:] = arr[1:] != arr[:-1]
    run_id = (np.cumsum(change) - 1).astype(np.int32)
    starts = np.flatnonzero(change)
    lengths = np.diff(np.append(starts, n))
    max_td = int(lengths[:-1].max()) if len(lengths) > 1 else 0
    return run_id, max_td


def name_groups(names) -> np.ndarray:
    """Group id per frame by logfile NAME, first-appearance order (temporal_selector.py:50-55)."""
    ids = {}
    return np.fromiter((ids.setdefault(nm, len(ids)) for nm in names), dtype=np.int32, count=len(names))


def city_groups(names, logs) -> np.ndarray:
    """City id per frame: ``location.split("-")[-1]`` of the frame's log (euclidean_spatial_selector.py:57-60,75-76).

    Raises KeyError when a frame's logfile is missing from ``logs`` (as the reference does at :76).
    """
    log_to_loc = {l["logfile"]: l["location"].split("-")[-1] for l in logs}
    ids = {}
    return np.fromiter((ids.setdefault(log_to_loc[nm], len(ids)) for nm in names), dtype=np.int32, count=len(names))


def box_counts(infos) -> np.ndarray:
    """``info["gt_names"].shape[0]`` per frame (base_selector.py:85)."""
    return np.fromiter((len(info["gt_names"]) for info in infos), dtype=np.int32, count=len(infos))


def extract_spatial_temporal(infos, known_logfiles=None):
    """Everything SpatialTemporalSelector needs from the infos list: ``(locations, run_id, max_td, n_boxes)``.

    Under torch.distributed (more than one rank) the per-frame python work -- string parsing, stacking the 4x4
    poses, counting boxes: O(pool) and, at tens of microseconds per frame, most of the end-to-end time -- is
    done by every rank for ITS slice of the pool only; the slices (a few hundred KB) are exchanged with one
    ``all_gather_object`` and every rank assembles the same full arrays.  Run ids are consecutive-run ids of
    the whole list, so the boundaries between slices are merged from the neighbouring names.
    ``known_logfiles``: the reference raises ``KeyError`` for a frame whose logfile is not in ``logs_file``
    (spatial_temporal_selector.py:80); here every rank raises it when any rank finds one.
    """
    from .dist import get_dist_info
    rank, world = get_dist_info()
    n = len(infos)
    if world <= 1 or n < 4 * world:
        names = logfile_names(infos)
        if known_logfiles is not None:
            for nm in names:
                if nm not in known_logfiles:
                    raise KeyError(nm)
        run_id, max_td = consecutive_runs(names)
        return ego_locations(infos), run_id, max_td, box_counts(infos)
    import torch.distributed as dist
    lo, hi = rank * n // world, (rank + 1) * n // world
    part = infos[lo:hi]
    names = logfile_names(part)
    missing = None
    if known_logfiles is not None:
        missing = next((nm for nm in names if nm not in known_logfiles), None)
    arr = np.asarray(names, dtype=object)
    change = np.ones(len(names), dtype=bool)            # change[0] is fixed up against the previous slice's last name
    change[1:] = arr[1:] != arr[:-1]
    mine = {"lo": lo, "missing": missing, "first": names[0], "last": names[-1], "change": np.packbits(change),
            "n": len(names), "loc": ego_locations(part), "boxes": box_counts(part)}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    for p in parts:
        if p["missing"] is not None:
            raise KeyError(p["missing"])
    change = np.concatenate([np.unpackbits(p["change"])[: p["n"]].astype(bool) for p in parts])
    for prev, cur in zip(parts, parts[1:]):
        change[cur["lo"]] = prev["last"] != cur["first"]
    run_id = (np.cumsum(change) - 1).astype(np.int32)
    starts = np.flatnonzero(change)
    lengths = np.diff(np.append(starts, n))
    max_td = int(lengths[:-1].max()) if len(lengths) > 1 else 0
    loc = np.ascontiguousarray(np.concatenate([p["loc"] for p in parts]))
    boxes = np.concatenate([p["boxes"] for p in parts]).astype(np.int32)
    return loc, run_id, max_td, boxes
