"""Synthetic nuScenes-trainval-shaped pools.

Writes exactly the on-disk inputs the reference selector reads
(det3d/selectors/base_selector.py:31-40, spatial_temporal_selector.py:65-91,110-134):

* an infos pkl: ``list[dict]`` in capture order with ``cam_front_path`` (logfile name up to the
  first ``"__"``), ``car_from_global`` (float64 4x4, inverse ego pose, as produced by
  det3d/datasets/nuscenes/nusc_common.py:404-417) and ``gt_names`` (ndarray of str, one per box);
* ``log.json``: list of ``{"logfile", "location", ...}`` records;
* a buffer JSON ``{"0": []}`` (tools/active_select.py:68-71).

There is no dataset in this environment, so every test, the bench and smoke() use these pools.
"""
import json
import os
import pickle
from dataclasses import dataclass

import numpy as np

LOCATIONS = ("boston-seaport", "singapore-onenorth", "singapore-queenstown", "singapore-hollandvillage")
CLASS_NAMES = np.array(
    ["car", "truck", "construction_vehicle", "bus", "trailer", "barrier", "motorcycle", "bicycle",
     "pedestrian", "traffic_cone"]
)


@dataclass(frozen=True)
class PoolSpec:
    """Shape of one synthetic pool: ``n_long`` scenes of ``long_len`` frames then the rest of ``short_len``."""
    name: str
    n_scenes: int
    n_long: int
    long_len: int
    short_len: int
    scenes_per_log: int = 12
    seed: int = 20221107

    @property
    def n_frames(self) -> int:
        return self.n_long * self.long_len + (self.n_scenes - self.n_long) * self.short_len


# Pools named by BASELINE.json's configs.
P28K = PoolSpec("P28K", 700, 130, 41, 40)        # 28,130 frames, 59 logs
P850 = PoolSpec("P850", 850, 80, 34, 33)         # 28,130 frames, 71 logs (temporal-only config)
P281K = PoolSpec("P281K", 7000, 1300, 41, 40)    # 281,300 frames, 584 logs
# Small pools for unit tests (same generator, seconds on the CPU oracle).
P2K = PoolSpec("P2K", 52, 19, 41, 40)            # 2,099 frames
P8K = PoolSpec("P8K", 200, 0, 41, 40)            # 8,000 frames
P300 = PoolSpec("P300", 12, 0, 41, 25, scenes_per_log=3)   # 300 frames


def _logfile_name(rng: np.random.Generator, k: int) -> str:
    vehicle = "n008" if rng.random() < 0.45 else "n015"
    tz = "-0400" if vehicle == "n008" else "+0800"
    month, day = 5 + (k % 5), 1 + int(rng.integers(0, 28))
    hh, mm, ss = int(rng.integers(8, 19)), int(rng.integers(0, 60)), int(rng.integers(0, 60))
    # k keeps names unique even when the random clock fields collide
    return f"{vehicle}-2018-{month:02d}-{day:02d}-{hh:02d}-{mm:02d}-{ss:02d}{tz}-k{k:04d}"


def generate_pool(spec: PoolSpec, box_m: float = 1600.0):
    """Return ``(infos, logs)`` for ``spec``.

    Ego track per log: start ~U([200, 200+box_m]^2) m, heading ~U(0, 2pi); every 0.5 s key-frame
    ``v <- clip(v + N(0, 0.6), 0.05, 15)`` m/s and ``heading += N(0, 0.03)``.  The speed floor keeps
    every kNN edge above scipy's 1e-8 "no edge" threshold and avoids duplicate positions.
    Consecutive scenes of one log are separated by a 0-60 s untracked gap of the same drive.
    """
    rng = np.random.default_rng(spec.seed)
    n_logs = -(-spec.n_scenes // spec.scenes_per_log)
    logs, infos = [], []
    scene = 0
    for k in range(n_logs):
        logfile = _logfile_name(rng, k)
        vehicle = logfile.split("-")[0]
        location = LOCATIONS[0] if vehicle == "n008" else LOCATIONS[1 + int(rng.integers(0, 3))]
        logs.append({"token": f"log{k:04d}", "logfile": logfile, "vehicle": vehicle,
                     "date_captured": logfile[5:15], "location": location})
        pos = rng.uniform(200.0, 200.0 + box_m, size=2)
        heading = rng.uniform(0.0, 2.0 * np.pi)
        speed = rng.uniform(1.0, 10.0)
        t_us = 1526915243000000 + k * 86_400_000_000
        box_rate = rng.uniform(12.0, 15.0)
        for _ in range(min(spec.scenes_per_log, spec.n_scenes - scene)):
            n_frames = spec.long_len if scene < spec.n_long else spec.short_len
            for f in range(n_frames):
                c, s = np.cos(heading), np.sin(heading)
                rot = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
                trans = np.array([pos[0], pos[1], 0.0])
                car_from_global = np.eye(4)
                car_from_global[:3, :3] = rot.T
                car_from_global[:3, 3] = rot.T.dot(-trans)
                n_box = int(min(150, rng.poisson(rng.gamma(2.0, box_rate))))
                infos.append({
                    "lidar_path": f"samples/LIDAR_TOP/{logfile}__LIDAR_TOP__{t_us}.pcd.bin",
                    "cam_front_path": f"samples/CAM_FRONT/{logfile}__CAM_FRONT__{t_us}.jpg",
                    "token": f"s{scene:05d}f{f:02d}",
                    "timestamp": t_us * 1e-6,
                    "car_from_global": car_from_global,
                    "gt_names": CLASS_NAMES[rng.integers(0, len(CLASS_NAMES), size=n_box)],
                    "sweeps": [],
                })
                speed = float(np.clip(speed + rng.normal(0.0, 0.6), 0.05, 15.0))
                heading += rng.normal(0.0, 0.03)
                pos = pos + 0.5 * speed * np.array([np.cos(heading), np.sin(heading)])
                t_us += 500_000
            gap_s = rng.uniform(0.0, 60.0)
            pos = pos + gap_s * speed * np.array([np.cos(heading), np.sin(heading)])
            heading += rng.normal(0.0, 0.3)
            t_us += int(gap_s * 1e6)
            scene += 1
    assert len(infos) == spec.n_frames, (len(infos), spec.n_frames)
    return infos, logs


def write_pool(spec: PoolSpec, out_dir: str, box_m: float = 1600.0) -> dict:
    """Generate ``spec`` and write infos pkl / log.json / empty buffer JSON under ``out_dir``."""
    os.makedirs(out_dir, exist_ok=True)
    infos, logs = generate_pool(spec, box_m=box_m)
    paths = {
        "infos": os.path.join(out_dir, "infos_train_10sweeps_withvelo.pkl"),
        "logs": os.path.join(out_dir, "log.json"),
        "buffer": os.path.join(out_dir, "buffer.json"),
        "distance_store": os.path.join(out_dir, "dijkstra_distance_map.npy"),
    }
    with open(paths["infos"], "wb") as f:
        pickle.dump(infos, f, protocol=2)
    with open(paths["logs"], "w") as f:
        json.dump(logs, f)
    with open(paths["buffer"], "w") as f:
        json.dump({"0": []}, f, indent=4)
    return paths


def synthetic_lidar_points(index: int, n_points: int = 3000, seed: int = 7) -> np.ndarray:
    """Deterministic multi-sweep-shaped points float32 [n_points,5] for frame ``index`` (tests)."""
    rng = np.random.default_rng(seed * 1_000_003 + index)
    ang = rng.uniform(0.0, 2.0 * np.pi, n_points)
    rad = 1.5 + 68.5 * rng.random(n_points) ** 2
    centre = 6.0 * np.array([np.cos(index * 0.7), np.sin(index * 1.3)])          # frame-dependent scene layout
    pts = np.empty((n_points, 5), dtype=np.float32)
    pts[:, 0] = rad * np.cos(ang) + centre[0] * (rng.random(n_points) < 0.3)
    pts[:, 1] = rad * np.sin(ang) + centre[1] * (rng.random(n_points) < 0.3)
    pts[:, 2] = -1.84 + rad * rng.uniform(-0.5, 0.2, n_points)
    pts[:, 3] = rng.uniform(0, 255, n_points)
    pts[:, 4] = 0.05 * rng.integers(0, 10, n_points)
    return pts


def write_lidar_files(infos, out_dir: str, n_points: int = 1500, nsweeps: int = 3, seed: int = 7) -> None:
    """Write nuScenes-style ``.pcd.bin`` files (5 float32 per point) and fill ``lidar_path`` / ``sweeps``."""
    os.makedirs(os.path.join(out_dir, "samples"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "sweeps"), exist_ok=True)
    for i, info in enumerate(infos):
        pts = synthetic_lidar_points(i, n_points * nsweeps, seed)
        info["lidar_path"] = f"samples/f{i:06d}.pcd.bin"
        pts[:n_points].tofile(os.path.join(out_dir, info["lidar_path"]))
        info["sweeps"] = []
        for s in range(1, nsweeps):
            path = f"sweeps/f{i:06d}_s{s}.pcd.bin"
            pts[s * n_points:(s + 1) * n_points].tofile(os.path.join(out_dir, path))
            tm = np.eye(4)
            tm[:2, 3] = 0.3 * s, -0.2 * s
            info["sweeps"].append({"lidar_path": path, "transform_matrix": tm, "time_lag": 0.05 * s})


def synthetic_features(n: int, dim: int = 512, seed: int = 11) -> np.ndarray:
    """float32 [n, dim] stand-in for the mean-pooled FPN features the reference's feature selectors store with
    ``torch.save`` (feature_selector.py:66-78): non-negative, post-ReLU-like, a few latent "scene types" plus
    per-frame noise so that distances are structured rather than uniform."""
    rng = np.random.default_rng(seed)
    centres = np.abs(rng.normal(0.0, 1.0, size=(24, dim))).astype(np.float32)
    mix = rng.dirichlet(np.full(24, 0.3), size=n).astype(np.float32)
    feats = mix @ centres + np.abs(rng.normal(0.0, 0.05, size=(n, dim))).astype(np.float32)
    return np.ascontiguousarray(feats, dtype=np.float32)
