"""``TemporalSelector``: frame-index gaps inside a logfile-NAME group, margin 1e6 outside, raw
(det3d/selectors/temporal_selector.py:16-104)."""
import logging
from typing import Optional

import numpy as np

from .. import pool, sharded
from .base_selector import BaseSelector
from .registry import SELECTORS


@SELECTORS.register_module
class TemporalSelector(BaseSelector):
    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 detector=None, dataloader=None, logger: Optional[logging.Logger] = None, pred: bool = False,
                 cost_b: float = 0.04, cost_f: float = 0.12) -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)

    def select_samples(self, **kwargs) -> None:
        infos = self.infos_origin
        groups = pool.name_groups(pool.logfile_names(infos))     # :50-55
        handle = sharded.make_engine(self._device_index(kwargs))   # one GPU, or this rank's shard
        try:
            handle.set_pool(np.zeros((len(infos), 2)), groups, pool.box_counts(infos))   # no ego pose in this selector
            self._run_greedy(handle, spatial="none", temporal=True, normalize="raw", forbid_sampled=True)
        finally:
            handle.close()
