#!/usr/bin/env python
"""Data-only replacement of the reference CLI (tools/active_select.py:22-163).

Same flags and the same two-invocation protocol: the first run (no buffer file yet) writes the
empty buffer ``{"0": []}`` with ``indent=4`` and exits (:68-71,111-114); later runs build the
selector from ``cfg.selector``, call ``select_samples(local_rank=...)`` and ``dump_file()``
(:152-163).  The reference also builds a detector, a dataset and a dataloader it does not need for
the data-only selectors (:119-148); that part is skipped.  Seeds as in :76-80.

The config is a python file that defines ``selector = dict(type=..., budget=..., buffer_file=...,
infos_origin=..., logs_file=...)`` -- the reference's own config files work unmodified
(examples/active/cbgs_spatial_temporal.py:361-367 is such a block; the two detector-stack helpers those files
import are replaced by placeholders when ``det3d`` is not installed, see ``run_config_file``) -- or a
.json/.yaml file with a top-level ``selector`` mapping.
"""
import argparse
import importlib.abc
import importlib.machinery
import logging
import os
import random
import runpy
import sys
import types

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


class _PlaceholderModule(types.ModuleType):
    """Stands in for ``det3d`` and its sub-modules while a reference config file is executed outside the reference
    tree: every attribute is a function that returns its first argument (``build_box_coder(cfg) -> cfg``,
    ``get_downsample_factor(model) -> model``), which is all the detector half of such a config needs to evaluate."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: a[0] if a else None


class _PlaceholderFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == "det3d" or fullname.startswith("det3d."):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _PlaceholderModule(spec.name)

    def exec_module(self, module):
        pass


def run_config_file(path):
    """Execute a python config like ``Config.fromfile`` does (det3d/torchie/utils/config.py).  The reference's own
    configs (examples/active/cbgs_*.py) import two helpers of the detector stack at module level
    (``det3d.builder.build_box_coder``, ``det3d.utils.config_tool.get_downsample_factor``); when ``det3d`` is not
    importable they are replaced by placeholders for the duration of the execution -- only ``selector`` is read."""
    try:
        return runpy.run_path(path)
    except ModuleNotFoundError as e:
        if not (e.name or "").split(".")[0] == "det3d":
            raise
    finder = _PlaceholderFinder()
    sys.meta_path.insert(0, finder)
    try:
        return runpy.run_path(path)
    finally:
        sys.meta_path.remove(finder)
        for name in [m for m in sys.modules if m == "det3d" or m.startswith("det3d.")]:
            if isinstance(sys.modules[name], _PlaceholderModule):
                del sys.modules[name]


def load_selector_cfg(path):
    if path.endswith(".py"):
        cfg = run_config_file(path)
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
