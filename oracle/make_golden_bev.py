"""oracle/make_golden_bev.py -- TEST INFRASTRUCTURE ONLY.

Pins the BINNING of the BEV density descriptor to the reference's own BEV map: runs the UNMODIFIED numba function
``points_to_bev`` of det3d/ops/point_cloud/bev_ops.py (loaded by file path; it needs numpy + numba only) on
deterministic synthetic points and records its point-count channel ``bev_map[-1]`` (:53) ->
tests/golden/bev_counts.json.  ``max_voxels`` is set above the number of points so that the order-dependent early
``break`` (:48-49) never fires; non-finite coordinates are not fed (the reference would index out of bounds).

No selector of the reference consumes this map, so the DESCRIPTOR (levels, distances) stays unpinned by the
reference; what this golden pins is the float32 cell arithmetic (subtract, true division, floor, half-open
bounds, y-major layout) that oracle/bev_oracle.py and the CUDA histogram restate.

    python oracle/make_golden_bev.py
"""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import adsel_b200  # noqa: E402,F401
from adsel_b200 import synthetic  # noqa: E402
from reference_loader import REFERENCE_ROOT  # noqa: E402

NUSC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)   # examples/active/cbgs_spatial_temporal.py:280
CASES = {   # case id -> (voxel size, frames, points per frame)
    "grid128": ((0.8, 0.8, 8.0), 4, 3000),
    "grid256": ((0.4, 0.4, 8.0), 2, 5000),
    "grid512_z2": ((0.2, 0.2, 4.0), 1, 4000),      # two height slices: the count channel sums over z
}


def edge_points():
    """Points on and next to every face of the range, in float32 (the cases where subtract/divide/floor matter)."""
    lo, hi = np.float32(-51.2), np.float32(51.2)
    vals = [lo, np.nextafter(lo, np.float32(0)), np.nextafter(lo, np.float32(-100)), hi, np.nextafter(hi, np.float32(0)),
            np.nextafter(hi, np.float32(100)), np.float32(0.0), np.float32(-0.0), np.float32(0.8), np.float32(-0.8),
            np.nextafter(np.float32(0.8), np.float32(0)), np.float32(50.4), np.nextafter(np.float32(50.4), np.float32(0)),
            np.float32(1e-30), np.float32(-1e-30), np.float32(25.6), np.float32(-25.6)]
    zs = [np.float32(-5.0), np.nextafter(np.float32(-5.0), np.float32(-10)), np.float32(3.0),
          np.nextafter(np.float32(3.0), np.float32(0)), np.float32(-1.0), np.nextafter(np.float32(-1.0), np.float32(0))]
    pts = [(x, y, z, np.float32(1.0)) for x in vals for y in vals[:8] for z in zs[:3]]
    pts += [(x, np.float32(3.3), z, np.float32(1.0)) for x in vals for z in zs]
    return np.array(pts, dtype=np.float32)


def case_points(case, frame, n_points):
    """float32 [m,4] x,y,z,intensity of one frame: the package's synthetic LiDAR plus the edge points."""
    pts = synthetic.synthetic_lidar_points(1000 * (sorted(CASES).index(case) + 1) + frame, n_points)[:, :4]
    return np.ascontiguousarray(np.concatenate([pts, edge_points()]), dtype=np.float32)


def main():
    path = os.path.join(REFERENCE_ROOT, "det3d", "ops", "point_cloud", "bev_ops.py")
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_bev_ops", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import numba
    golden = {"numpy": np.__version__, "numba": numba.__version__, "range": list(NUSC_RANGE),
              "source": "det3d/ops/point_cloud/bev_ops.py:points_to_bev, bev_map[-1]", "cases": {}}
    for case, (voxel, frames, n_points) in CASES.items():
        out = []
        for f in range(frames):
            pts = case_points(case, f, n_points)
            bev = mod.points_to_bev(pts, list(voxel), list(NUSC_RANGE), with_reflectivity=False, max_voxels=10 ** 7)
            count = bev[-1]                                   # [H = ny, W = nx] float32 point counts
            assert count.dtype == np.float32 and float(count.max()) < 2 ** 24
            cells = np.flatnonzero(count.reshape(-1))
            out.append({"n_points": int(len(pts)), "shape": list(count.shape), "inside": int(count.sum()),
                        "cells": cells.tolist(), "counts": count.reshape(-1)[cells].astype(np.int64).tolist()})
            print(case, f, count.shape, "inside", int(count.sum()), "of", len(pts), "cells", len(cells))
        golden["cases"][case] = {"voxel": list(voxel), "n_points": n_points, "frames": out}
    out_file = os.path.join(ROOT, "tests", "golden", "bev_counts.json")
    with open(out_file, "w") as f:
        json.dump(golden, f)
    print("wrote", out_file, os.path.getsize(out_file), "bytes")


if __name__ == "__main__":
    main()
