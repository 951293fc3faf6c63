"""``FeatureSelector`` -- detector-feature diversity (det3d/selectors/feature_selector.py:16-172).

Same registry name, constructor keywords and result layout as the reference.  The reference fills an N x N
float32 distance map row by row on the CPU (:87-109; ``p=2`` is ``sqrt((a-b)**2).sum(1)``, i.e. also an L1
distance) and runs the usual greedy loop on it (:143-172); here the map is one CUDA kernel (csrc/featdist.cu) and
the loop is the dense greedy kernel.  Producing the features needs the trained detector (``pred=True`` ->
``buffer_pred``, :51-85), which is out of scope: the features are read from ``buffer_path`` like the
reference does with ``pred=False`` (``torch.load``; ``.npy`` is accepted as well).
"""
import logging
import os
from typing import Optional

import numpy as np

from .. import pool, sharded
from .base_selector import BaseSelector
from .registry import SELECTORS


def load_features(path) -> np.ndarray:
    """float32 [N, C] features stored by the reference's ``buffer_pred`` (``torch.save`` of a CPU tensor)."""
    if str(path).endswith(".npy"):
        return np.ascontiguousarray(np.load(path), dtype=np.float32)
    import torch
    feats = torch.load(path, map_location="cpu")
    return np.ascontiguousarray(feats.detach().to(torch.float32).numpy())


@SELECTORS.register_module
class FeatureSelector(BaseSelector):
    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 buffer_path: str = "/home/st2000/data/buffers/feature_pred.pt", p: int = 2, detector=None,
                 dataloader=None, logger: Optional[logging.Logger] = None,
                 distance_store_file: str = "/home/st2000/data/buffers/feature_distance_map.npy", pred: bool = True,
                 cost_b: float = 0.04, cost_f: float = 0.12) -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)
        self.buffer_path = buffer_path
        # accepted for drop-in compatibility; the 4 N^2-byte map is rebuilt on the GPU in milliseconds instead
        self.distance_store_file = distance_store_file
        assert p in [1, 2]
        self.p = p

    def select_samples(self, **kwargs) -> None:
        if self.pred:
            raise NotImplementedError("pred=True runs the trained detector over the pool (feature_selector.py:51-85), "
                                      "which is outside this package: store the features and pass pred=False")
        feats = load_features(self.buffer_path)
        n = len(self.infos_origin)
        assert feats.shape[0] == n, f"{self.buffer_path} holds {feats.shape[0]} feature rows for {n} frames"
        sampled = list(self.buffer[self.get_max_key()])
        if len(sampled) == 0:
            # the reference stacks the rows of the frames already in the buffer (:143-147): an empty buffer fails there
            raise RuntimeError("stack expects a non-empty TensorList")
        handle = sharded.make_engine(self._device_index(kwargs))
        if not hasattr(handle, "feature_distance_build"):
            raise NotImplementedError("FeatureSelector runs on one GPU (set ADSEL_SHARD=0 under torch.distributed)")
        try:
            handle.feature_distance_build(feats, self.p)                       # :87-109
            selected, cost = handle.select_dense(pool.box_counts(self.infos_origin), cost_b=self.cost_b,
                                                 cost_f=self.cost_f, cost_start=self.get_cost_amount(),
                                                 budget=int(self.current_budget), forbid_sampled=False,
                                                 sampled=sampled, first_index=-1)
            self.last_stats = {"stage_ms": handle.stage_ms(), "counters": handle.counters(), "cost": cost}
            self.selected_index[self.current_budget] = selected + sampled      # :172 (new frames first)
        finally:
            handle.close()
