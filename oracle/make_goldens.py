"""oracle/make_goldens.py -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference selectors (oracle/reference_loader.py) on the synthetic pools of
``adsel_b200.synthetic`` and records what they select as small JSON fixtures under tests/golden/.
Run in this container (the reference tree is not on the GPU box):

    python oracle/make_goldens.py --pool P2K            # seconds
    python oracle/make_goldens.py --pool P28K --only st # ~8 min, ~32 GB RSS

Seeding follows tools/active_select.py:76-80: ``random.seed(3407)`` immediately before each
selector is constructed, so the first pick of a fresh buffer is ``random.choice(range(N))``.
Each case starts from a fresh ``{"0": []}`` buffer unless ``rounds`` chains several budgets (the
second round then resumes from the first round's buffer).
"""
import argparse
import contextlib
import io
import json
import logging
import os
import random
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import adsel_b200  # noqa: E402,F401
from adsel_b200 import synthetic  # noqa: E402
from reference_loader import load_reference_selectors  # noqa: E402

# (case id, selector, ctor kwargs, budgets chained as successive rounds)
ST_VARIANTS = [
    ("st_exp_sum", {}),
    ("st_exp_min", {"aggregate": "min"}),
    ("st_exp_max", {"aggregate": "max"}),
    ("st_linear_sum", {"normalize": "linear"}),
    ("st_linear_min", {"normalize": "linear", "aggregate": "min"}),
    ("st_linear_max", {"normalize": "linear", "aggregate": "max"}),
    ("st_exp_sum_l05", {"lambda_t": 0.5}),
    ("st_exp_sum_cost", {"cost_b": 0.1, "cost_f": 0.3}),
]


def run_case(classes, sel_name, paths, kwargs, rounds, workdir, keep_store=None):
    buffer_file = os.path.join(workdir, f"buffer_{time.time_ns()}.json")
    with open(buffer_file, "w") as f:
        json.dump({"0": []}, f, indent=4)
    out, times = {}, []
    for budget in rounds:
        kw = dict(budget=budget, buffer_file=buffer_file, infos_origin=paths["infos"],
                  logger=logging.getLogger("ref"))
        if sel_name in ("SpatialTemporalSelector", "SpatialSelector", "EuSpatialSelector"):
            kw["logs_file"] = paths["logs"]
            kw["distance_store_file"] = keep_store or os.path.join(workdir, f"store_{sel_name}_{time.time_ns()}.npy")
        kw.update(kwargs)
        random.seed(3407)
        sel = classes[sel_name](**kw)
        t0 = time.perf_counter()
        try:
            with contextlib.redirect_stdout(io.StringIO()):   # "cost amount: ..." progress lines
                sel.select_samples(local_rank=0)
        except AssertionError as e:                           # pool exhausted (reference :182)
            out[sel.current_budget] = {"raises": "AssertionError", "message": str(e)}
            times.append(time.perf_counter() - t0)
            break
        times.append(time.perf_counter() - t0)
        out.update({k: [int(i) for i in v] for k, v in sel.get_selected_samples().items()})
        sel.buffer.update(sel.selected_index)
        with open(buffer_file, "w") as f:
            json.dump(sel.buffer, f)
        if keep_store is None and "distance_store_file" in kw and os.path.exists(kw["distance_store_file"]):
            os.remove(kw["distance_store_file"])
    return out, times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pool", default="P2K")
    ap.add_argument("--only", default="all", help="all | st (SpatialTemporal exp/sum only) | st_agg | st_linear | temporal (TemporalSelector only) | random | spatial (SpatialSelector at budget 1200)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--budget", type=int, default=600, help="base budget B; chained cases use B/2,B/2,B")
    args = ap.parse_args()
    logging.basicConfig(level=logging.WARNING)
    os.environ["TQDM_DISABLE"] = "1"
    B, H = args.budget, args.budget // 2
    spec = getattr(synthetic, args.pool)
    classes = load_reference_selectors()
    os.makedirs(args.out, exist_ok=True)
    golden = {"pool": spec.name, "n_frames": spec.n_frames, "seed": spec.seed,
              "numpy": np.__version__, "python_random_seed": 3407, "cases": {}}
    import scipy
    golden["scipy"] = scipy.__version__
    with tempfile.TemporaryDirectory() as wd:
        paths = synthetic.write_pool(spec, wd)
        store = os.path.join(wd, "geodesic.npy")   # shared: the geodesic map is identical across ST cases
        cases = []
        if args.only == "temporal":     # BASELINE.json config 4: temporal-only selector, 850 scenes, budget 1200
            cases += [("temporal_1200", "TemporalSelector", {}, [1200]),
                      ("temporal_rounds", "TemporalSelector", {}, [600, 600])]
        elif args.only == "spatial":    # BASELINE config 2's engine at full size: SpatialSelector, budget 1200 (~6 min at P28K)
            cases += [("spatial_1200", "SpatialSelector", {}, [1200])]
        elif args.only == "st_linear":  # far = inf rows at full size: the lazy-row path (one SSSP per selected frame)
            cases += [("st_linear_sum_1200", "SpatialTemporalSelector", {"normalize": "linear"}, [1200]),
                      ("st_linear_min_600", "SpatialTemporalSelector", {"normalize": "linear", "aggregate": "min"}, [600]),
                      ("st_linear_sum_rounds", "SpatialTemporalSelector", {"normalize": "linear"}, [300, 300])]
        elif args.only == "random":     # the random baseline every experiment of the paper is compared with
            cases += [("random", "RandomSelector", {}, [B]),
                      ("random_rounds", "RandomSelector", {}, [H, H, B]),
                      ("random_cost", "RandomSelector", {"cost_b": 0.1, "cost_f": 0.3}, [B])]
        elif args.only == "st_agg":     # the other aggregates / weights of the exp-normalised rows at full size (~5 min at P28K)
            cases += [("st_exp_min_1200", "SpatialTemporalSelector", {"aggregate": "min"}, [1200]),
                      ("st_exp_max_1200", "SpatialTemporalSelector", {"aggregate": "max"}, [1200]),
                      ("st_exp_sum_l05_cost_1200", "SpatialTemporalSelector", {"lambda_t": 0.5, "cost_b": 0.1, "cost_f": 0.3}, [1200]),
                      ("st_exp_max_rounds", "SpatialTemporalSelector", {"aggregate": "max"}, [300, 300])]
        elif args.only == "st":
            cases += [("st_exp_sum", "SpatialTemporalSelector", {}, [600]),
                      ("st_exp_sum_1200", "SpatialTemporalSelector", {}, [1200]),
                      ("st_exp_sum_4800", "SpatialTemporalSelector", {}, [4800]),
                      ("st_exp_sum_rounds", "SpatialTemporalSelector", {}, [600, 600])]  # noqa
        else:
            for cid, kw in ST_VARIANTS:
                cases.append((cid, "SpatialTemporalSelector", kw, [B]))
            cases += [("st_exp_sum_2x", "SpatialTemporalSelector", {}, [2 * B]),
                      ("st_exp_sum_rounds", "SpatialTemporalSelector", {}, [H, H, B]),
                      ("st_linear_sum_rounds", "SpatialTemporalSelector", {"normalize": "linear"}, [H, H]),
                      ("st_exhaust", "SpatialTemporalSelector", {}, [1000 * B]),
                      ("spatial", "SpatialSelector", {}, [B]),
                      ("spatial_rounds", "SpatialSelector", {}, [H, H]),
                      ("temporal", "TemporalSelector", {}, [B]),
                      ("temporal_rounds", "TemporalSelector", {}, [H, H]),
                      ("euclid", "EuSpatialSelector", {}, [B]),
                      ("euclid_rounds", "EuSpatialSelector", {}, [H, H])]
        for cid, sel_name, kw, rounds in cases:
            keep = store if sel_name in ("SpatialTemporalSelector", "SpatialSelector") else None
            sel_out, times = run_case(classes, sel_name, paths, kw, rounds, wd, keep_store=keep)
            golden["cases"][cid] = {"selector": sel_name, "kwargs": kw, "rounds": rounds,
                                    "selected_index": sel_out, "seconds": [round(t, 3) for t in times]}
            print(cid, {k: (len(v) if isinstance(v, list) else v["raises"]) for k, v in sel_out.items()},
                  [round(t, 2) for t in times], flush=True)
    golden["base_budget"] = B
    out_file = os.path.join(args.out, f"{spec.name}_{args.only}.json")
    with open(out_file, "w") as f:
        json.dump(golden, f)
    print("wrote", out_file, os.path.getsize(out_file), "bytes")


if __name__ == "__main__":
    main()
