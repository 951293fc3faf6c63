#!/usr/bin/env python
"""Dense greedy loop timing (single CTA vs thread-block cluster) on random integer descriptors (development aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adsel_b200  # noqa: E402,F401
from adsel_b200 import native  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 28130
    dim = 256
    rng = np.random.default_rng(3)
    levels = native.bf16_bits(rng.integers(0, 32, size=(n, dim)))
    boxes = rng.integers(0, 60, size=n).astype(np.int32)
    h = native.Handle(0)
    h.pairdist_set_descriptors(levels)
    h.pairdist_build()
    res = {}
    for var in ("0", "1"):
        os.environ["ADSEL_DENSE_CLUSTER"] = var
        best = None
        for _ in range(4):
            sel, cost = h.select_dense(boxes, budget=1200, first_index=17)
            t = h.stage_ms()["dense_greedy"]
            best = t if best is None else min(best, t)
        res[var] = sel
        print(f"ADSEL_DENSE_CLUSTER={var}: {best:.3f} ms, {len(sel)} frames, {best * 1e3 / len(sel):.2f} us/step", flush=True)
    print("identical:", res["0"] == res["1"])
    h.close()


if __name__ == "__main__":
    main()
