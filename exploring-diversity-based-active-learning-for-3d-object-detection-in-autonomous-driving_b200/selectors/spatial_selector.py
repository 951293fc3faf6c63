"""``SpatialSelector``: geodesic term only, raw distances (det3d/selectors/spatial_selector.py:16-138)."""
import logging
import os
from typing import Optional

from .. import pool, sharded
from ..fileio import load
from .base_selector import BaseSelector
from .registry import SELECTORS


@SELECTORS.register_module
class SpatialSelector(BaseSelector):
    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 detector=None, dataloader=None, logger: Optional[logging.Logger] = None, pred: bool = False,
                 k: int = 8,
                 logs_file: str = "/home/st2000/data/Datasets/nuScenes/train/v1.0-trainval/log.json",
                 distance_store_file: str = "/home/st2000/data/buffers/dijkstra_distance_map.npy",
                 cost_b: float = 0.04, cost_f: float = 0.12) -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)
        self.logs_file = logs_file
        self.k = k
        self.distance_store_file = distance_store_file   # accepted, unused (see SpatialTemporalSelector)

    def select_samples(self, **kwargs) -> None:
        if self.distance_store_file and os.path.exists(self.distance_store_file):
            self.logger.warning(f"{self.distance_store_file} exists but is ignored: the geodesic rows are rebuilt on the GPU "
                                "(the reference would load this cached N x N map instead of recomputing it)")
        infos = self.infos_origin
        names, locations, n_boxes = pool.extract_slice(infos, device=self._device_index(kwargs))
        known = {l["logfile"] for l in load(self.logs_file)}
        for nm in names:                                  # KeyError like spatial_selector.py:71
            if nm not in known:
                raise KeyError(nm)
        run_id, _ = pool.consecutive_runs(names)
        handle = sharded.make_engine(self._device_index(kwargs))   # one GPU, or this rank's shard
        try:
            handle.set_pool(locations, run_id, n_boxes)
            handle.build_knn_graph(self.k)
            self._run_greedy(handle, spatial="geodesic", temporal=False, normalize="raw", forbid_sampled=True)
        finally:
            handle.close()
