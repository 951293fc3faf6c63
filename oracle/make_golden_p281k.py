"""oracle/make_golden_p281k.py -- TEST INFRASTRUCTURE ONLY.

BASELINE.json config 5: the 281,300-frame pool (10x trainval), budget 4800.  The dense reference cannot
run it (one N x N float64 map = 633 GB, it holds four), so the fixture is produced by the STREAMING
restatement (oracle/restatement.py: one scipy Dijkstra row per selected frame), which tests/test_oracle.py
pins index-for-index against the unmodified reference on every smaller pool including the 28,130-frame one
(SURVEY.md §8c).  Writes tests/golden/P281K_st_restatement.json (about 25 KB).

    python oracle/make_golden_p281k.py            # ~5-10 min on one core
"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import adsel_b200  # noqa: E402,F401
from adsel_b200 import synthetic  # noqa: E402
import restatement as R  # noqa: E402


def main():
    spec = synthetic.P281K
    t0 = time.time()
    infos, logs = synthetic.generate_pool(spec)
    print(f"pool generated: {len(infos)} frames in {time.time() - t0:.1f} s", flush=True)
    random.seed(3407)                                   # tools/active_select.py:80
    first = random.choice(range(len(infos)))
    t0 = time.time()
    key, sel = R.select("spatial_temporal", infos, logs, {"0": []}, 4800, first)
    dt = time.time() - t0
    print(f"selected {len(sel)} frames in {dt:.1f} s", flush=True)
    out = {"pool": spec.name, "n_frames": spec.n_frames, "seed": spec.seed, "first_index": first,
           "generator": "oracle/make_golden_p281k.py (streaming restatement; the dense reference needs 4 x 633 GB)",
           "selector": "SpatialTemporalSelector", "kwargs": {}, "budget": 4800, "seconds": dt,
           "selected_index": {key: sel}}
    path = os.path.join(ROOT, "tests", "golden", "P281K_st_restatement.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path)


if __name__ == "__main__":
    main()
