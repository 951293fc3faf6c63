"""Host-side extraction of the per-frame arrays the device needs from the infos list.

Replaces the per-frame python loops of det3d/selectors/spatial_temporal_selector.py:78-91 (ego
location), :114-129 (consecutive logfile runs), temporal_selector.py:50-55 (logfile-name groups) and
euclidean_spatial_selector.py:57-60,75-76 (city per frame).  String parsing stays on the host; the
numeric part is vectorised but keeps the reference's own numpy expressions, so the float64 ego
locations are bit-identical to the reference loop (checked on a sample at run time and in tests).
"""
import importlib.machinery
import importlib.util
import os

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


LAST_EGO_CHECK = {}    # outcome of the last all-frames device check of the ego locations (tests, bench.py)


def _device_check(device, pose9, loc, fused_modes, frames) -> bool:
    """True when the device's literal-order recomputation of EVERY frame (csrc/ego_check.cu) reproduces ``loc`` bit for
    bit under one of ``fused_modes``; ``ADSEL_EGO_CHECK=probe`` keeps the probe frames as the only evidence."""
    if device is None or os.environ.get("ADSEL_EGO_CHECK", "all") == "probe":
        return True
    from . import native
    for fused in fused_modes:
        n_bad, first = native.check_ego_locations(device, pose9, loc, fused)
        LAST_EGO_CHECK.clear()
        LAST_EGO_CHECK.update(frames=frames, n_bad=n_bad, first_bad=first, fused=bool(fused))
        if n_bad == 0:
            return True
    import warnings
    warnings.warn(f"ego locations: {LAST_EGO_CHECK['n_bad']} of {frames} frames (first: {LAST_EGO_CHECK['first_bad']}) differ from the "
                  "literal order recomputed on the device; falling back to the reference's per-frame expression")
    return False


def ego_locations(infos, check: int = 64, device=None) -> np.ndarray:
    """float64 [N,2] ego x,y: ``-(cal[:3,3].T @ cal[:3,:3])[:2]`` with ``cal = car_from_global`` (:83-91).

    The batched ``np.matmul`` goes through the same vector-matrix inner loop as the reference's
    per-frame ``@``; ``check`` frames (spread over the pool) are recomputed with the reference's
    literal expression and must match bit-for-bit, otherwise the literal loop is used for all.
    With ``device`` (a CUDA device index) every frame is additionally redone on the GPU in the literal
    order (fma chain or separately rounded products, whichever this host's numpy uses).
    """
    cals = [info["car_from_global"] for info in infos]
    # numpy's vector-matrix product rounds differently for Fortran-ordered / strided operands (another BLAS kernel), so
    # the batched product below stands for the per-frame one only when every matrix has the layout of the stacked copy
    uniform = all(isinstance(c, np.ndarray) and c.shape == (4, 4) and c.dtype == np.float64 and c.flags.c_contiguous
                  for c in cals)
    if not uniform:
        return _location_loop(cals)
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
    elif device is not None and not _device_check(device, stack[:, :3, [0, 1, 3]].reshape(-1, 9), loc, (True, False), n):
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


_HOSTEXT = None


def _hostext():
    """The optional CPython helper (csrc/hostext.c), loaded from the in-tree build; ``False`` when unavailable."""
    global _HOSTEXT
    if _HOSTEXT is None:
        _HOSTEXT = False
        if os.environ.get("ADSEL_HOSTEXT", "1") != "0":
            try:
                from . import build as _build
                path = _build.build_hostext()
                if path and os.path.exists(path):
                    loader = importlib.machinery.ExtensionFileLoader("_adsel_hostext", path)
                    spec = importlib.util.spec_from_loader("_adsel_hostext", loader)
                    mod = importlib.util.module_from_spec(spec)
                    loader.exec_module(mod)
                    _HOSTEXT = mod
            except Exception:          # host-side convenience only: the python implementation below is equivalent
                _HOSTEXT = False
    return _HOSTEXT


def _change_flags(names) -> np.ndarray:
    """bool [m]: names[i] != names[i-1] (first entry True) -- the consecutive-run boundaries (:114-129)."""
    arr = np.asarray(names, dtype=object)
    change = np.ones(len(names), dtype=bool)
    change[1:] = arr[1:] != arr[:-1]
    return change


def runs_from_change(change):
    """``(run_id int32 [m], max_temporal_distance)`` from run-boundary flags, as ``consecutive_runs`` defines them."""
    run_id = (np.cumsum(change) - 1).astype(np.int32)
    starts = np.flatnonzero(change)
    lengths = np.diff(np.append(starts, len(change)))
    max_td = int(lengths[:-1].max()) if len(lengths) > 1 else 0
    return run_id, max_td


def extract_slice(infos, lo=0, hi=None, check: int = 64, with_change: bool = False, device=None):
    """``(names, locations, n_boxes)`` of ``infos[lo:hi]`` -- logfile_names / ego_locations / box_counts in one pass
    over the list, in C when the helper is built.  The helper's ego locations are compared bit-for-bit with the
    reference's literal expression on ``check`` probe frames (numpy's vector-matrix product uses fused
    multiply-adds on FMA hosts and separately rounded products elsewhere: both orders are tried) and the python
    implementation is used if neither matches or the infos are not the plain dicts / float64 matrices expected.
    With ``device`` (a CUDA device index) the pinned order is then verified on EVERY frame by one fp64 kernel
    (csrc/ego_check.cu); a mismatch anywhere sends the whole slice through the reference's literal per-frame loop."""
    hi = len(infos) if hi is None else hi
    ext = _hostext()
    if ext and isinstance(infos, list) and hi > lo:
        probe = np.unique(np.linspace(lo, hi - 1, num=min(check, hi - lo)).astype(np.int64))
        try:
            want = _location_loop([infos[i]["car_from_global"] for i in probe]).view(np.uint64)
            for fused in (True, False):
                loc_b, box_b, names, chg_b, pose_b = ext.extract(infos, lo, hi, fused)
                loc = np.frombuffer(loc_b, dtype=np.float64).reshape(-1, 2)
                if np.array_equal(loc[probe - lo].view(np.uint64), want):
                    if not _device_check(device, np.frombuffer(pose_b, dtype=np.float64), loc, (fused,), hi - lo):
                        loc = _location_loop([infos[i]["car_from_global"] for i in range(lo, hi)])
                    out = (names, loc, np.frombuffer(box_b, dtype=np.int32))
                    return out + (np.frombuffer(chg_b, dtype=np.uint8).astype(bool),) if with_change else out
        except (TypeError, KeyError, ValueError):
            pass
    part = infos[lo:hi]
    names = logfile_names(part)
    out = (names, ego_locations(part, check, device), box_counts(part))
    return out + (_change_flags(names),) if with_change else out


def extract_spatial_temporal(infos, known_logfiles=None, device=None):
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
        names, loc, boxes, change = extract_slice(infos, with_change=True, device=device)
        if known_logfiles is not None:
            for i in np.flatnonzero(change):            # every distinct consecutive name, in order (KeyError like :80)
                if names[i] not in known_logfiles:
                    raise KeyError(names[i])
        run_id, max_td = runs_from_change(change)
        return loc, run_id, max_td, boxes
    import time
    import torch.distributed as dist
    t0 = time.perf_counter()
    lo, hi = rank * n // world, (rank + 1) * n // world
    names, loc_mine, boxes_mine, change = extract_slice(infos, lo, hi, with_change=True, device=device)
    missing = None                                      # change[0] is fixed up against the previous slice's last name
    if known_logfiles is not None:
        missing = next((names[i] for i in np.flatnonzero(change) if names[i] not in known_logfiles), None)
    mine = {"lo": lo, "missing": missing, "first": names[0], "last": names[-1], "change": np.packbits(change),
            "n": len(names), "loc": loc_mine, "boxes": boxes_mine}
    t1 = time.perf_counter()
    parts = _all_gather_slices(dist, world, mine, -(-n // world))
    t2 = time.perf_counter()
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
    LAST_TIMING.update({"parse_own_slice_ms": (t1 - t0) * 1e3, "all_gather_ms": (t2 - t1) * 1e3,
                        "assemble_ms": (time.perf_counter() - t2) * 1e3})
    return loc, run_id, max_td, boxes


LAST_TIMING = {}   # host-side timing of the last sharded extraction (reported by bench.py)
_TEXT = 256   # bytes reserved per string field of a gathered slice header


def _all_gather_slices(dist, world, mine, cap):
    """All-gather of the per-rank slice records as ONE fixed-size byte tensor per rank (a single NCCL / gloo
    collective; ``all_gather_object`` costs tens of milliseconds in pickling and staging for the same data).
    Layout: int64 lo, n | three length-prefixed utf-8 fields (missing, first, last) | loc f64[cap,2] |
    boxes i32[cap] | packed change bits u8[ceil(cap/8)]."""
    import torch
    nbits = -(-cap // 8)
    size = 16 + 3 * (8 + _TEXT) + cap * 16 + cap * 4 + nbits
    buf = np.zeros(size, dtype=np.uint8)
    buf[:16] = np.array([mine["lo"], mine["n"]], dtype=np.int64).view(np.uint8)
    off = 16
    for key in ("missing", "first", "last"):
        raw = b"" if mine[key] is None else mine[key].encode("utf-8")
        if len(raw) > _TEXT:
            raise ValueError(f"logfile name longer than {_TEXT} bytes: {mine[key][:40]}...")
        flag = -1 if mine[key] is None else len(raw)
        buf[off:off + 8] = np.array([flag], dtype=np.int64).view(np.uint8)
        buf[off + 8:off + 8 + len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        off += 8 + _TEXT
    k = mine["n"]
    buf[off:off + k * 16] = np.ascontiguousarray(mine["loc"], dtype=np.float64).view(np.uint8).reshape(-1)
    off += cap * 16
    buf[off:off + k * 4] = np.ascontiguousarray(mine["boxes"], dtype=np.int32).view(np.uint8)
    off += cap * 4
    buf[off:off + len(mine["change"])] = mine["change"]
    if dist.get_backend() == "nccl":
        t = torch.from_numpy(buf).cuda()
        out = torch.empty(world * size, dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t)
        allb = out.cpu().numpy().reshape(world, size)
    else:
        outs = [torch.empty(size, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(outs, torch.from_numpy(buf))
        allb = np.stack([o.numpy() for o in outs])
    parts = []
    for r in range(world):
        b = allb[r]
        lo, k = (int(v) for v in b[:16].view(np.int64))
        rec = {"lo": lo, "n": k}
        off = 16
        for key in ("missing", "first", "last"):
            flag = int(b[off:off + 8].view(np.int64)[0])
            rec[key] = None if flag < 0 else bytes(b[off + 8:off + 8 + flag]).decode("utf-8")
            off += 8 + _TEXT
        rec["loc"] = b[off:off + k * 16].view(np.float64).reshape(k, 2)
        off += cap * 16
        rec["boxes"] = b[off:off + k * 4].view(np.int32)
        off += cap * 4
        rec["change"] = b[off:off + nbits]
        parts.append(rec)
    return parts
