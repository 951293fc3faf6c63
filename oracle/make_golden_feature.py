"""oracle/make_golden_feature.py -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference ``FeatureSelector`` (det3d/selectors/feature_selector.py, ``pred=False``: features
loaded with ``torch.load``) on synthetic features and records what it selects -> tests/golden/<pool>_feature.json.
The reference needs a non-empty buffer (it stacks the rows of the frames already selected, :143-148), so the
first round of every case is a fixed pseudo-random buffer.

    python oracle/make_golden_feature.py --pool P2K
"""
import argparse
import contextlib
import io
import json
import logging
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import adsel_b200  # noqa: E402,F401
from adsel_b200 import synthetic  # noqa: E402
from reference_loader import load_reference_selectors  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool", default="P2K")
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--fused", action="store_true",
                    help="SpatialFeatureSelector / SpatialTemporalFeatureSelector -> <pool>_fused_feature.json")
    args = ap.parse_args()
    if args.fused:
        return main_fused(args)
    logging.basicConfig(level=logging.WARNING)
    os.environ["TQDM_DISABLE"] = "1"
    spec = getattr(synthetic, args.pool)
    cls = load_reference_selectors()["FeatureSelector"]
    feats = synthetic.synthetic_features(spec.n_frames, args.dim)
    golden = {"pool": spec.name, "n_frames": spec.n_frames, "dim": args.dim, "feature_seed": 11, "torch": torch.__version__,
              "cases": {}}
    with tempfile.TemporaryDirectory() as wd:
        paths = synthetic.write_pool(spec, wd)
        feat_path = os.path.join(wd, "feature_pred.pt")
        torch.save(torch.from_numpy(feats), feat_path)
        first = sorted(np.random.default_rng(5).choice(spec.n_frames, size=40, replace=False).tolist())
        for cid, p, rounds in (("feature_p2", 2, [100]), ("feature_p1", 1, [100]), ("feature_p2_rounds", 2, [60, 60])):
            buffer_file = os.path.join(wd, f"buffer_{cid}.json")
            json.dump({"0": [], "50": first}, open(buffer_file, "w"))
            out = {}
            for budget in rounds:
                sel = cls(budget=budget, buffer_file=buffer_file, infos_origin=paths["infos"], buffer_path=feat_path, p=p,
                          distance_store_file=os.path.join(wd, f"dist_{cid}.npy"), pred=False, logger=logging.getLogger("ref"))
                with contextlib.redirect_stdout(io.StringIO()):
                    sel.select_samples(local_rank=0)
                out.update({k: [int(i) for i in v] for k, v in sel.get_selected_samples().items()})
                sel.buffer.update(sel.selected_index)
                json.dump(sel.buffer, open(buffer_file, "w"))
            golden["cases"][cid] = {"selector": "FeatureSelector", "p": p, "initial_buffer": {"0": [], "50": first},
                                    "rounds": rounds, "selected_index": out}
            print(cid, {k: len(v) for k, v in out.items()}, flush=True)
    out_file = os.path.join(ROOT, "tests", "golden", f"{spec.name}_feature.json")
    json.dump(golden, open(out_file, "w"))
    print("wrote", out_file, os.path.getsize(out_file), "bytes")


FUSED_CASES = [   # (case id, selector, ctor kwargs, initial buffer kind, chained budgets)
    ("sf_sum", "SpatialFeatureSelector", {}, "empty", [300]),
    ("sf_min", "SpatialFeatureSelector", {"aggregate": "min"}, "empty", [300]),
    ("sf_max", "SpatialFeatureSelector", {"aggregate": "max"}, "empty", [300]),
    ("sf_p1_k5", "SpatialFeatureSelector", {"p": 1, "k": 5}, "empty", [300]),
    ("sf_rounds", "SpatialFeatureSelector", {}, "empty", [150, 150]),      # 2nd round: seeded from SPATIAL rows only (:201-204)
    ("sf_resume", "SpatialFeatureSelector", {"aggregate": "max"}, "seeded", [200]),
    ("stf", "SpatialTemporalFeatureSelector", {}, "empty", [300]),
    ("stf_lambdas", "SpatialTemporalFeatureSelector", {"lambda_t": 0.5, "lambda_f": 2.0}, "empty", [300]),
    ("stf_p1", "SpatialTemporalFeatureSelector", {"p": 1}, "empty", [300]),
    ("stf_rounds", "SpatialTemporalFeatureSelector", {"lambda_f": 0.25}, "empty", [150, 150]),
    ("stf_resume", "SpatialTemporalFeatureSelector", {}, "seeded", [200]),
]


def main_fused(args):
    """The two selectors that add the feature term to the spatial (/temporal) terms
    (det3d/selectors/spatial_feature_selector.py, spatial_temporal_feature_selector.py), unmodified, pred=False."""
    import random
    logging.basicConfig(level=logging.WARNING)
    os.environ["TQDM_DISABLE"] = "1"
    spec = getattr(synthetic, args.pool)
    classes = load_reference_selectors()
    feats = synthetic.synthetic_features(spec.n_frames, args.dim)
    golden = {"pool": spec.name, "n_frames": spec.n_frames, "dim": args.dim, "feature_seed": 11, "torch": torch.__version__,
              "numpy": np.__version__, "python_random_seed": 3407, "cases": {}}
    with tempfile.TemporaryDirectory() as wd:
        paths = synthetic.write_pool(spec, wd)
        feat_path = os.path.join(wd, "feature_pred.pt")
        torch.save(torch.from_numpy(feats), feat_path)
        seeded = sorted(np.random.default_rng(5).choice(spec.n_frames, size=40, replace=False).tolist())
        for cid, sel_name, kw, start, rounds in FUSED_CASES:
            initial = {"0": []} if start == "empty" else {"0": [], "50": seeded}
            buffer_file = os.path.join(wd, f"buffer_{cid}.json")
            json.dump(initial, open(buffer_file, "w"))
            store = os.path.join(wd, f"geodesic_k{kw.get('k', 8)}.npy")      # the geodesic map only depends on k
            out = {}
            for budget in rounds:
                random.seed(3407)                                             # tools/active_select.py:80
                sel = classes[sel_name](budget=budget, buffer_file=buffer_file, infos_origin=paths["infos"],
                                        buffer_path=feat_path, logs_file=paths["logs"], distance_store_file=store,
                                        pred=False, logger=logging.getLogger("ref"), **kw)
                with contextlib.redirect_stdout(io.StringIO()):
                    sel.select_samples(local_rank=0)
                out.update({k: [int(i) for i in v] for k, v in sel.get_selected_samples().items()})
                sel.buffer.update(sel.selected_index)
                json.dump(sel.buffer, open(buffer_file, "w"))
            golden["cases"][cid] = {"selector": sel_name, "kwargs": kw, "initial_buffer": initial, "rounds": rounds,
                                    "selected_index": out}
            print(cid, {k: len(v) for k, v in out.items()}, flush=True)
    out_file = os.path.join(ROOT, "tests", "golden", f"{spec.name}_fused_feature.json")
    json.dump(golden, open(out_file, "w"))
    print("wrote", out_file, os.path.getsize(out_file), "bytes")


if __name__ == "__main__":
    main()
