"""``SpatialTemporalSelector`` -- the hot path (det3d/selectors/spatial_temporal_selector.py:16-193).

Same registry name, constructor keywords, assertions and results as the reference; the work
(kNN graph, geodesic rows, temporal term, normalise, aggregate, greedy loop, cost stop) runs in the
sm_100a library through ``native.Handle``.
"""
import logging
import os
from typing import Optional

from .. import pool, sharded
from ..fileio import load
from .base_selector import BaseSelector
from .registry import SELECTORS


@SELECTORS.register_module
class SpatialTemporalSelector(BaseSelector):
    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 detector=None, dataloader=None, logger: Optional[logging.Logger] = None, pred: bool = False,
                 k: int = 8,
                 logs_file: str = "/home/st2000/data/Datasets/nuScenes/train/v1.0-trainval/log.json",
                 normalize: str = "exp",
                 distance_store_file: str = "/home/st2000/data/buffers/dijkstra_distance_map.npy",
                 cost_b: float = 0.04, cost_f: float = 0.12, lambda_t: float = 1, aggregate: str = "sum") -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)
        self.logs_file = logs_file
        assert normalize in ["linear", "exp"]
        self.normalize = normalize
        self.k = k
        self.lambda_t = lambda_t
        # Accepted for drop-in compatibility.  The reference caches a dense N x N float64 geodesic map
        # there (6.33 GB at N=28,130, :60-63,106); here the rows are rebuilt on the GPU in milliseconds,
        # so the file is neither read nor written.
        self.distance_store_file = distance_store_file
        assert aggregate in ["sum", "min", "max"]
        self.aggregate = aggregate
        self.logger.info(f"lambda_t: {self.lambda_t}")

    def select_samples(self, **kwargs) -> None:
        if self.distance_store_file and os.path.exists(self.distance_store_file):
            self.logger.warning(f"{self.distance_store_file} exists but is ignored: the geodesic rows are rebuilt on the GPU "
                                "(the reference would load this cached N x N map instead of recomputing it)")
        import time
        t0 = time.perf_counter()
        infos = self.infos_origin
        known = {l["logfile"] for l in load(self.logs_file)}           # the reference raises KeyError at :80
        t_logs = time.perf_counter()
        # :83-91 ego locations, :114-129 consecutive logfile runs, base_selector.py:85 box counts
        locations, run_id, max_td, n_boxes = pool.extract_spatial_temporal(infos, known, device=self._device_index(kwargs))
        t1 = time.perf_counter()
        handle = sharded.make_engine(self._device_index(kwargs))   # one GPU, or this rank's shard
        try:
            handle.set_pool(locations, run_id, n_boxes)
            handle.build_knn_graph(self.k)                # :93-101 (+ scipy's dense->CSR rules)
            max_sd = handle.geodesic_max_finite() if self.normalize == "linear" else 0.0   # :139
            t2 = time.perf_counter()
            self._run_greedy(handle, spatial="geodesic", temporal=True, normalize=self.normalize,
                             aggregate=self.aggregate, lambda_t=self.lambda_t, max_spatial=max_sd,
                             max_temporal=max_td)
            t3 = time.perf_counter()
            self.last_stats["host_ms"] = {"extract": (t1 - t0) * 1e3, "logs_file": (t_logs - t0) * 1e3, "pool+graph": (t2 - t1) * 1e3,
                                          "lists+greedy+readback": (t3 - t2) * 1e3, **pool.LAST_TIMING}
        finally:
            handle.close()
