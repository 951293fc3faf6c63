#!/usr/bin/env python
"""Data-only replacement of the reference CLI (tools/active_select.py:22-163).

Same flags and the same two-invocation protocol: the first run (no buffer file yet) writes the
empty buffer ``{"0": []}`` with ``indent=4`` and exits (:68-71,111-114); later runs build the
selector from ``cfg.selector``, call ``select_samples(local_rank=...)`` and ``dump_file()``
(:152-163).  The reference also builds a detector, a dataset and a dataloader it does not need for
the data-only selectors (:119-148); that part is skipped.  Seeds as in :76-80.

The config is a python file that defines ``selector = dict(type=..., budget=..., buffer_file=...,
infos_origin=..., logs_file=...)`` (the selector block of examples/active/cbgs_spatial_temporal.py:361-367;
the reference's full configs cannot be imported without the detector stack), or a .json/.yaml file
with a top-level ``selector`` mapping.
"""
import argparse
import logging
import os
import random
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import adsel_b200  # noqa: E402,F401
from adsel_b200 import fileio  # noqa: E402
from adsel_b200.selectors import build_selector  # noqa: E402


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Active selection (data-only selectors, B200)")
    p.add_argument("--config", required=True, help="config file with a `selector` dict")
    p.add_argument("--work_dir", help="logged only; the buffer lives at cfg.selector.buffer_file")
    p.add_argument("--checkpoint", default=None, help="accepted for compatibility, unused (no detector is built)")
    p.add_argument("--txt_result", type=bool, default=False, help="accepted for compatibility, unused")
    p.add_argument("--gpus", type=int, default=2, help="accepted for compatibility; ranks come from the launcher")
    p.add_argument("--budget", type=int, default=2400, help="annotation-cost budget of this round (reference default)")
    p.add_argument("--pred", action="store_true")
    p.add_argument("--launcher", choices=["none", "pytorch", "slurm", "mpi"], default="pytorch")
    # torch.distributed.launch passes --local_rank (the reference's default is 0); torchrun only sets LOCAL_RANK
    p.add_argument("--local_rank", type=int, default=int(os.environ.get("LOCAL_RANK", 0)))
    args = p.parse_args(argv)
    if "LOCAL_RANK" not in os.environ:                  # :60-61
        os.environ["LOCAL_RANK"] = str(args.local_rank)
    return args


def load_selector_cfg(path):
    if path.endswith(".py"):
        cfg = runpy.run_path(path)
    else:
        cfg = fileio.load(path)
    if "selector" not in cfg:
        raise KeyError(f"{path} defines no `selector`")
    return dict(cfg["selector"])


def main(argv=None):
    random.seed(3407)
    np.random.seed(3407)
    args = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    logger = logging.getLogger("active_select")
    selector_cfg = load_selector_cfg(args.config)
    if args.budget is not None:
        selector_cfg["budget"] = args.budget            # always overrides the config (:87-88)
    distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if distributed:
        import torch
        torch.cuda.set_device(args.local_rank)
        torch.distributed.init_process_group(backend="nccl", init_method="env://")
    logger.info("Distributed testing: {}".format(distributed))
    logger.info(f"work dir: {args.work_dir}")
    if not os.path.exists(selector_cfg["buffer_file"]):
        fileio.dump({"0": []}, selector_cfg["buffer_file"], indent=4)
        logger.info(f"init a empty buffer, and save as {selector_cfg['buffer_file']}")
        return 0
    selector_cfg.update({"detector": None, "dataloader": None, "logger": logger, "pred": args.pred})
    selector = build_selector(selector_cfg)
    logger.info("begin selection")
    selector.select_samples(local_rank=args.local_rank)
    selector.dump_file()
    return 0


if __name__ == "__main__":
    sys.exit(main())
