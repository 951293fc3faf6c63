"""``BevSpatialSelector`` -- BEV-density diversity (BASELINE.json config 2; north_star kernels 1 and 3).

EXTENSION: the reference has no LiDAR-descriptor selector.  The engine is the
reference's one-term greedy loop (det3d/selectors/spatial_selector.py:101-138: raw distances, no
normalisation, cost-budget stop, assertion against buffer and selection); the distance term is
the squared Euclidean distance between per-frame BEV density-level descriptors:

  points   multi-sweep LiDAR, float32 [P,5] = x,y,z,intensity,dt, loaded like
           det3d/datasets/pipelines/loading.py:17-63,98-126 (key-frame + sweeps transformed by
           ``transform_matrix``, returns closer than 1 m in x and y removed from sweeps)
  counts   bev_ops.py:27-53 binning over ``point_range`` / ``voxel_size`` (csrc/bev.cu)
  level    bit_length(count), 0..31
  d2(i,j)  |level_i - level_j|^2 as exact integers, Gram tiles on tcgen05 (csrc/pairdist.cu)
"""
import logging
import os
from typing import Callable, Optional

import numpy as np

from .. import pool, sharded
from .base_selector import BaseSelector
from .registry import SELECTORS


def read_lidar_file(path, num_point_feature=4):
    """loading.py:17-31: float32 records of 5 values, first ``num_point_feature`` kept."""
    points = np.fromfile(path, dtype=np.float32)
    points = points[: points.shape[0] - points.shape[0] % 5]
    return points.reshape(-1, 5)[:, :num_point_feature]


def load_multisweep_points(info, data_root="", nsweeps=10):
    """Key-frame plus the first ``nsweeps-1`` sweeps as [P,5] (loading.py:34-63,98-126).

    The reference draws the sweeps with ``np.random.choice(..., replace=False)``; a BEV histogram
    does not depend on point order, so the first ``nsweeps-1`` are taken in order.
    """
    key = read_lidar_file(os.path.join(data_root, str(info["lidar_path"])))
    chunks = [np.hstack([key, np.zeros((key.shape[0], 1), dtype=np.float32)])]
    assert (nsweeps - 1) <= len(info["sweeps"]), \
        "nsweeps {} should not greater than list length {}.".format(nsweeps, len(info["sweeps"]))
    for sweep in info["sweeps"][: nsweeps - 1]:
        pts = read_lidar_file(os.path.join(data_root, str(sweep["lidar_path"]))).T          # [4,P]
        keep = ~((np.abs(pts[0]) < 1.0) & (np.abs(pts[1]) < 1.0))                            # remove_close, :34-43
        pts = pts[:, keep]
        if sweep["transform_matrix"] is not None:
            pts[:3, :] = sweep["transform_matrix"].dot(np.vstack((pts[:3, :], np.ones(pts.shape[1]))))[:3, :]
        t = np.full((1, pts.shape[1]), sweep["time_lag"], dtype=np.float32)
        chunks.append(np.vstack([pts, t]).T.astype(np.float32))
    return np.concatenate(chunks, axis=0)


@SELECTORS.register_module
class BevSpatialSelector(BaseSelector):
    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 detector=None, dataloader=None, logger: Optional[logging.Logger] = None, pred: bool = False,
                 data_root: str = "", nsweeps: int = 10,
                 point_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0),     # cbgs_spatial_temporal.py:280
                 voxel_size=(0.8, 0.8, 8.0), batch_frames: int = 1024,
                 points_fn: Optional[Callable] = None, cost_b: float = 0.04, cost_f: float = 0.12,
                 io_threads: int = 8) -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)
        self.data_root, self.nsweeps = data_root, nsweeps
        self.point_range, self.voxel_size = tuple(point_range), tuple(voxel_size)
        self.batch_frames = int(batch_frames)
        self.points_fn = points_fn
        self.io_threads = max(1, int(io_threads))
        self._io_pool = None

    def _points(self, index):
        info = self.infos_origin[index]
        if self.points_fn is not None:
            return np.ascontiguousarray(self.points_fn(info, index), dtype=np.float32)
        return load_multisweep_points(info, self.data_root, self.nsweeps)

    def load_frames(self, lo, hi):
        """``(points float32 [P,5], frame_off int64 [hi-lo+1])`` of frames [lo, hi), in frame order.  Frames read
        from files are loaded by ``io_threads`` threads (np.fromfile and the sweep transforms release the GIL); a
        user ``points_fn`` is called sequentially, it need not be thread-safe.  An exception of any frame propagates."""
        if self.points_fn is None and self.io_threads > 1 and hi - lo > 1:
            if self._io_pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._io_pool = ThreadPoolExecutor(max_workers=self.io_threads, thread_name_prefix="adsel-lidar")
            frames = list(self._io_pool.map(self._points, range(lo, hi)))
        else:
            frames = [self._points(i) for i in range(lo, hi)]
        off = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.int64)
        points = np.concatenate(frames, axis=0) if frames else np.empty((0, 5), dtype=np.float32)
        return points, off

    def select_samples(self, **kwargs) -> None:
        import random
        n = len(self.infos_origin)
        handle = sharded.make_engine(self._device_index(kwargs))   # one GPU, or this rank's slice of the frames

        def load_batch(h, lo, hi):                                 # stream the pool through HBM
            h.bev_set_points(*self.load_frames(lo, hi))

        try:
            handle.dense_build(n, self.batch_frames, load_batch, self.point_range, self.voxel_size)
            sampled = list(self.buffer[self.get_max_key()])
            first_index = random.choice(range(n)) if len(sampled) == 0 else -1   # spatial_selector.py:111
            if getattr(handle, "world", 1) > 1:     # sharded engine: every rank starts from rank 0's draw and the same buffer
                from .. import dist as _d
                (first_index,) = _d.agree_from_rank0((first_index,))
                _d.check_same_on_all_ranks((len(sampled), sum(sampled) % (1 << 31), int(self.current_budget)),
                                           "the buffer they resume from")
            selected, cost = handle.dense_select(pool.box_counts(self.infos_origin), cost_b=self.cost_b,
                                                 cost_f=self.cost_f, cost_start=self.get_cost_amount(),
                                                 budget=int(self.current_budget), forbid_sampled=True,
                                                 sampled=sampled, first_index=first_index)
            self.last_stats = {"stage_ms": handle.stage_ms(), "counters": handle.counters(), "cost": cost}
            self.selected_index[self.current_budget] = sampled + selected
        finally:
            handle.close()
            if self._io_pool is not None:
                self._io_pool.shutdown(wait=False)
                self._io_pool = None
