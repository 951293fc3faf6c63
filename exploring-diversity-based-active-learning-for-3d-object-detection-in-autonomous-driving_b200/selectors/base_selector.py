"""``BaseSelector``: buffer JSON, cumulative-budget bookkeeping, cost model, dump
(det3d/selectors/base_selector.py:12-86), plus the shared device-side greedy driver.
"""
import logging
import os
import random
from typing import Optional

from ..dist import master_only
from ..fileio import dump, load
from . import budget_buffer as bb
from .registry import SELECTORS


@SELECTORS.register_module
class BaseSelector(object):
    """Same constructor, attributes and methods as the reference class.

    ``detector`` / ``dataloader`` are accepted and stored for drop-in compatibility
    (tools/active_select.py:152-157 always passes them) but never used by the data-only selectors.
    """

    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None,
                 infos_origin="", detector=None, dataloader=None, logger: Optional[logging.Logger] = None,
                 pred: bool = False, cost_b: float = 0.04, cost_f: float = 0.12) -> None:
        super().__init__()
        # files: the buffer is re-written in place unless a separate dump name is given (:33)
        self.buffer_file = buffer_file
        self.dump_file_name = buffer_file if dump_file_name is None else dump_file_name
        self.infos_file = infos_origin
        self.buffer = load(buffer_file)                 # {"0": [], "<cumulative budget>": [idx, ...]}
        self.infos_origin = load(infos_origin)          # list[dict], whole pool in RAM (:39)
        bb.settle_allocator(len(self.infos_origin))
        # budget: ``budget`` more cost units on top of the latest round (:40)
        self.budget = budget
        self.cost_b, self.cost_f = cost_b, cost_f
        self.current_budget = bb.next_key(self.buffer, budget)
        self.selected_index = {}
        # carried for the caller, untouched by the data-only selectors
        self.detector, self.dataloader, self.pred = detector, dataloader, pred
        self.logger = logger if logger is not None else logging.getLogger(__file__)
        self.last_stats = {}

    # -- reference API ------------------------------------------------------------------------
    def get_max_key(self):
        return bb.latest_key(self.buffer)

    def get_selected_samples(self):
        return self.selected_index

    def get_cost_amount(self):
        """Annotation cost of the frames already in the buffer (:78-86)."""
        return bb.annotation_cost(self.infos_origin, self.buffer[self.get_max_key()], self.cost_f, self.cost_b)

    def select_samples(self, **kwargs) -> None:
        return

    @master_only
    def dump_file(self) -> None:
        """Write the updated buffer JSON and the infos subset ``<infos>_<budget>.<ext>`` (:57-73)."""
        self.buffer.update(self.selected_index)
        dump(self.buffer, self.dump_file_name)
        self.logger.info(f"update the buffer, and save as {self.dump_file_name}")
        key = str(self.current_budget)
        out_path = bb.subset_path(self.infos_file, key)
        dump([self.infos_origin[i] for i in self.buffer[key]], out_path)
        self.logger.info(f"sample the {key} infos and save as {out_path}")

    # -- shared by the diversity selectors ---------------------------------------------------
    @staticmethod
    def _device_index(kwargs) -> int:
        if kwargs.get("local_rank") is not None:
            return int(kwargs["local_rank"])
        return int(os.environ.get("LOCAL_RANK", "0"))

    def _run_greedy(self, handle, **select_kwargs) -> None:
        """Seed / resume exactly like the reference (:157-170) and run the device greedy loop (:172-193)."""
        sampled = list(self.buffer[self.get_max_key()])
        first_index = -1
        if len(sampled) == 0:
            # consumes the caller-seeded python RNG exactly like the reference (:168)
            first_index = random.choice(range(len(self.infos_origin)))
        if getattr(handle, "world", 1) > 1:
            # The ranks of a sharded engine do the bookkeeping redundantly: they must start from rank 0's draw and the
            # same buffer (in the reference an unsynchronised per-rank RNG is harmless: full replicas, rank 0 dumps).
            from .. import dist as _d
            (first_index,) = _d.agree_from_rank0((first_index,))
            _d.check_same_on_all_ranks((len(sampled), sum(sampled) % (1 << 31), int(self.current_budget)),
                                       "the buffer they resume from")
        selected, cost = handle.select(cost_b=self.cost_b, cost_f=self.cost_f, cost_start=self.get_cost_amount(),
                                       budget=int(self.current_budget), sampled=sampled, first_index=first_index,
                                       **select_kwargs)
        self.last_stats = {"stage_ms": handle.stage_ms(), "counters": handle.counters(), "cost": cost}
        self.selected_index[self.current_budget] = sampled + selected
