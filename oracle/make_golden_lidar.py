"""oracle/make_golden_lidar.py -- TEST INFRASTRUCTURE ONLY.

Pins the multi-sweep LiDAR loader in front of the BEV descriptor (``load_multisweep_points``) to the reference's
own loading code.  det3d/datasets/pipelines/loading.py cannot be imported (pycocotools, det3d.core), but the three
functions the nuScenes branch uses -- ``read_file`` (:17-31), ``remove_close`` (:34-43), ``read_sweep`` (:46-63) --
are self-contained numpy; they are taken UNMODIFIED out of the reference file with ``ast`` and executed, and the
nuScenes branch of ``LoadPointCloudFromFile.__call__`` (:98-126) is driven exactly as written there (key frame,
zero time lag, ``np.random.choice`` over the sweeps, ``np.hstack([points, times])``).

The reference draws the sweeps in random order; a histogram does not depend on point order, so the golden stores,
per frame, the number of points and a SHA-256 over the rows sorted lexicographically as raw float32 bit patterns.

    python oracle/make_golden_lidar.py          # -> tests/golden/lidar_multisweep.json
"""
import ast
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import adsel_b200  # noqa: E402,F401
from adsel_b200 import synthetic  # noqa: E402
from reference_loader import REFERENCE_ROOT  # noqa: E402

SPEC = dict(pool="P300", frames=6, n_points=1500, file_sweeps=4, seed=7)   # 4 files per frame: key frame + 3 sweeps


def rows_digest(points):
    """(count, sha256) of float32 [P,5] rows, independent of row order (sorted by their bit patterns)."""
    bits = np.ascontiguousarray(points, dtype=np.float32).view(np.uint32).reshape(-1, 5)
    order = np.lexsort(bits.T[::-1])
    return int(len(bits)), hashlib.sha256(bits[order].tobytes()).hexdigest()


def write_case(out_dir):
    """The synthetic pool's first frames with LiDAR files; sweeps get a rotation + translation and a time lag, and
    a cluster of returns inside the 1 m box that ``remove_close`` must drop (sweeps only, not the key frame)."""
    infos, _ = synthetic.generate_pool(getattr(synthetic, SPEC["pool"]))
    infos = infos[: SPEC["frames"]]
    synthetic.write_lidar_files(infos, out_dir, n_points=SPEC["n_points"], nsweeps=SPEC["file_sweeps"], seed=SPEC["seed"])
    for i, info in enumerate(infos):
        info["lidar_path"] = os.path.join(out_dir, info["lidar_path"])
        for s, sweep in enumerate(info["sweeps"]):
            sweep["lidar_path"] = os.path.join(out_dir, sweep["lidar_path"])
            a = 0.01 * (i + 1) * (s + 1)
            tm = np.eye(4)
            tm[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
            tm[:3, 3] = [0.3 * (s + 1), -0.2 * (s + 1), 0.01 * i]
            sweep["transform_matrix"] = tm if (i, s) != (1, 0) else None      # the reference allows None (:54)
            pts = np.fromfile(sweep["lidar_path"], dtype=np.float32).reshape(-1, 5)
            pts[:40, :2] = np.linspace(-1.2, 1.2, 80, dtype=np.float32).reshape(40, 2)   # around the 1 m box, both sides
            pts[40, :2] = [1.0, 0.0]          # |x| == radius is kept (strict <)
            pts[41, :2] = [-0.999999, 0.999999]
            pts.tofile(sweep["lidar_path"])
        if i == 2:                            # a trailing partial record: read_file drops it (:24-26)
            with open(info["lidar_path"], "ab") as f:
                f.write(np.array([1.0, 2.0], dtype=np.float32).tobytes())
    return infos


def reference_functions():
    path = os.path.join(REFERENCE_ROOT, "det3d", "datasets", "pipelines", "loading.py")
    tree = ast.parse(open(path).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("read_file", "remove_close", "read_sweep")]
    assert len(wanted) == 3
    ns = {"np": np}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), ns)
    return ns


def reference_combined(ns, info, nsweeps):
    """The nuScenes branch of LoadPointCloudFromFile.__call__ (loading.py:98-126), statement for statement."""
    points = ns["read_file"](str(info["lidar_path"]))
    sweep_points_list = [points]
    sweep_times_list = [np.zeros((points.shape[0], 1))]
    assert (nsweeps - 1) <= len(info["sweeps"])
    for i in np.random.choice(len(info["sweeps"]), nsweeps - 1, replace=False):
        points_sweep, times_sweep = ns["read_sweep"](info["sweeps"][i])
        sweep_points_list.append(points_sweep)
        sweep_times_list.append(times_sweep)
    points = np.concatenate(sweep_points_list, axis=0)
    times = np.concatenate(sweep_times_list, axis=0).astype(points.dtype)
    return np.hstack([points, times])


def main():
    ns = reference_functions()
    golden = {"numpy": np.__version__, "spec": SPEC, "source": "det3d/datasets/pipelines/loading.py:17-63,98-126",
              "frames": []}
    with tempfile.TemporaryDirectory() as wd:
        infos = write_case(wd)
        np.random.seed(0)
        for i, info in enumerate(infos):
            combined = reference_combined(ns, info, SPEC["file_sweeps"])      # all sweeps: the choice is a permutation
            assert combined.dtype == np.float32 and combined.shape[1] == 5
            n, digest = rows_digest(combined)
            key_only = rows_digest(reference_combined(ns, info, 1))
            golden["frames"].append({"points": n, "sha256": digest, "key_frame_points": key_only[0],
                                     "key_frame_sha256": key_only[1]})
            print(i, n, digest[:16], key_only[0])
    out_file = os.path.join(ROOT, "tests", "golden", "lidar_multisweep.json")
    with open(out_file, "w") as f:
        json.dump(golden, f)
    print("wrote", out_file)


if __name__ == "__main__":
    main()
