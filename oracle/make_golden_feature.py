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
    args = ap.parse_args()
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


if __name__ == "__main__":
    main()
