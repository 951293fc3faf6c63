"""``RandomSelector``: the uniform-random baseline the diversity selectors are compared with
(det3d/selectors/random_selector.py:12-64).

Host-only: there is no distance map and nothing to put on the GPU.  What has to be reproduced exactly is
the consumption of Python's ``random`` stream (one ``random.choice`` over the not-yet-labelled indices per
pick, in ascending-index order, :44-47,56,63), the two-step cost accumulation (``+= cost_f`` and then
``+= boxes * cost_b``, :57-58 -- a different rounding order from the diversity selectors) and the result
order ``newly selected + previously sampled`` (:66).
"""
import bisect
import logging
import random
from typing import Optional

from .base_selector import BaseSelector
from .registry import SELECTORS


@SELECTORS.register_module
class RandomSelector(BaseSelector):
    def __init__(self, budget: int, buffer_file: str, dump_file_name: Optional[str] = None, infos_origin="",
                 detector=None, dataloader=None, logger: Optional[logging.Logger] = None, pred: bool = False,
                 cost_b: float = 0.04, cost_f: float = 0.12) -> None:
        super().__init__(budget, buffer_file, dump_file_name, infos_origin=infos_origin, detector=detector,
                         dataloader=dataloader, logger=logger, pred=pred, cost_b=cost_b, cost_f=cost_f)

    def select_samples(self, **kwargs) -> None:
        sampled = self.buffer[self.get_max_key()]
        taken = set(sampled)
        remaining = [i for i in range(len(self.infos_origin)) if i not in taken]    # stays sorted
        self.logger.info(f"have selected {self.get_max_key()} examples, need to sample {self.budget} examples "
                         f"from {len(remaining)} examples")
        limit = int(self.current_budget)
        cost = self.get_cost_amount()
        picked = []
        while True:
            frame = random.choice(remaining)         # IndexError on an exhausted pool, as in the reference
            cost += self.cost_f
            cost += self.infos_origin[frame]["gt_names"].shape[0] * self.cost_b
            if cost > limit:
                break                                # the overflowing frame is dropped but its draw is consumed
            picked.append(frame)
            del remaining[bisect.bisect_left(remaining, frame)]
        self.last_stats = {"cost": cost}
        self.selected_index[self.current_budget] = picked + sampled
