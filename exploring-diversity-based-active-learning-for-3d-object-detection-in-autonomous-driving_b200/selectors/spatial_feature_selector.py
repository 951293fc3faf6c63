"""``SpatialFeatureSelector`` / ``SpatialTemporalFeatureSelector`` -- detector-feature diversity added to the geodesic (and
temporal) terms (det3d/selectors/spatial_feature_selector.py:18-237, spatial_temporal_feature_selector.py:18-258).

Same registry names, constructor keywords, assertions and results as the reference.  The kNN graph, the geodesic rows,
the N x N float32 feature map (csrc/featdist.cu: the reference's map bit for bit) and the greedy loop (csrc/fused.cu)
run on the GPU.  Producing the features needs the trained detector (``pred=True`` -> ``buffer_pred``), which is out of
scope: they are read from ``buffer_path`` as the reference does with ``pred=False``.
"""
import logging
import os
import random
from typing import Optional

from .. import pool, sharded
from ..fileio import load
from .base_selector import BaseSelector
from .feature_selector import load_features
from .registry import SELECTORS


class _FusedFeatureBase(BaseSelector):
    temporal = False

    def _select(self, kwargs, **loop_kw) -> None:
        if self.pred:
            raise NotImplementedError("pred=True runs the trained detector over the pool, which is outside this package: "
                                      "store the features and pass pred=False")
        if self.distance_store_file and os.path.exists(self.distance_store_file):
            self.logger.warning(f"{self.distance_store_file} exists but is ignored: the geodesic rows are rebuilt on the GPU")
        infos = self.infos_origin
        feats = load_features(self.buffer_path)
        assert feats.shape[0] == len(infos), f"{self.buffer_path} holds {feats.shape[0]} feature rows for {len(infos)} frames"
        names, locations, n_boxes = pool.extract_slice(infos, device=self._device_index(kwargs))
        known = {l["logfile"] for l in load(self.logs_file)}
        for nm in names:                                  # KeyError like spatial_feature_selector.py:152
            if nm not in known:
                raise KeyError(nm)
        run_id, max_td = pool.consecutive_runs(names)
        handle = sharded.make_engine(self._device_index(kwargs))
        if not hasattr(handle, "select_fused_feature"):
            raise NotImplementedError(f"{type(self).__name__} runs on one GPU (set ADSEL_SHARD=0 under torch.distributed)")
        try:
            handle.set_pool(locations, run_id, n_boxes)
            handle.build_knn_graph(self.k)
            handle.feature_distance_build(feats, self.p)
            sampled = list(self.buffer[self.get_max_key()])
            first_index = random.choice(range(len(infos))) if len(sampled) == 0 else -1
            try:
                selected, cost = handle.select_fused_feature(
                    temporal=self.temporal, cost_b=self.cost_b, cost_f=self.cost_f, cost_start=self.get_cost_amount(),
                    budget=int(self.current_budget), sampled=sampled, first_index=first_index, max_temporal=max_td, **loop_kw)
            except AssertionError as e:
                idx = int(str(e).split()[1])
                if loop_kw.get("forbid_sampled") and idx in sampled:      # spatial_feature_selector.py:222
                    raise AssertionError(f"id: {idx} not in the left index list") from None
                raise
            self.last_stats = {"stage_ms": handle.stage_ms(), "counters": handle.counters(), "cost": cost}
            self.selected_index[self.current_budget] = sampled + selected
        finally:
            handle.close()


@SELECTORS.register_module
class SpatialFeatureSelector(_FusedFeatureBase):
    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 buffer_path: str = "", detector=None, dataloader=None, logger: Optional[logging.Logger] = None,
                 pred: bool = False, k: int = 8, p: int = 2,
                 logs_file: str = "/home/st2000/data/Datasets/nuScenes/train/train/v1.0-trainval/log.json",
                 distance_store_file: str = "/home/st2000/data/buffers/dijkstra_distance_map.npy",
                 cost_b: float = 0.04, cost_f: float = 0.12, aggregate: str = "sum") -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)
        self.logs_file, self.k, self.distance_store_file, self.buffer_path = logs_file, k, distance_store_file, buffer_path
        assert p in [1, 2]
        self.p = p
        assert aggregate in ["sum", "min", "max"]
        self.aggregate = aggregate

    def select_samples(self, **kwargs) -> None:
        # seeds from the SPATIAL rows only (:201-209); asserts against the buffer as well as the new picks (:222-223)
        self._select(kwargs, aggregate=self.aggregate, seed_spatial_only=True, forbid_sampled=True)


@SELECTORS.register_module
class SpatialTemporalFeatureSelector(_FusedFeatureBase):
    temporal = True

    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 buffer_path: str = "/home/st2000/data/buffers/feature_pred.pt", detector=None, dataloader=None,
                 logger: Optional[logging.Logger] = None, pred: bool = True, k: int = 8, p: int = 2,
                 logs_file: str = "/home/st2000/data/Datasets/nuScenes/train/v1.0-trainval/log.json",
                 distance_store_file: str = "/home/st2000/data/buffers/dijkstra_distance_map.npy",
                 cost_b: float = 0.04, cost_f: float = 0.12, lambda_f: float = 1, lambda_t: float = 1) -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)
        self.buffer_path, self.logs_file, self.k, self.distance_store_file = buffer_path, logs_file, k, distance_store_file
        assert p in [1, 2]
        self.p = p
        self.lambda_f, self.lambda_t = lambda_f, lambda_t
        self.logger.info(f"lambda_f: {self.lambda_f}")
        self.logger.info(f"lambda_t: {self.lambda_t}")

    def select_samples(self, **kwargs) -> None:
        # seeds from the aggregated map (:225-234); asserts against the new picks only (:245)
        self._select(kwargs, aggregate="sum", lambda_t=self.lambda_t, lambda_f=self.lambda_f, seed_spatial_only=False,
                     forbid_sampled=False)
