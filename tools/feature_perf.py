#!/usr/bin/env python
"""Feature-distance map timing (torch-order kernel vs index-order kernel) -- development aid."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adsel_b200  # noqa: E402,F401
from adsel_b200 import native, synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28130
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 512
feats = synthetic.synthetic_features(n, dim, 11)
h = native.Handle(0)
for var in ("", "1"):
    if var:
        os.environ["ADSEL_FEATDIST_INDEX_ORDER"] = var
    else:
        os.environ.pop("ADSEL_FEATDIST_INDEX_ORDER", None)
    ts = []
    for _ in range(3):
        h.feature_distance_build(feats, 2)
        ts.append(h.stage_ms()["pairdist"])
    print(("index order" if var else "torch order"), f"n={n} dim={dim}: {min(ts):.2f} ms", flush=True)
h.close()
