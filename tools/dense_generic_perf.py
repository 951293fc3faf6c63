"""Times the generic (> 32,768 columns) dense greedy variant (development aid, GPU box only)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import adsel_b200  # noqa: F401
from adsel_b200 import native
h = native.Handle(0)
for n in (35_328, 52_000):
    rng = np.random.default_rng(n)
    lv = rng.integers(0, 32, size=(n, 8))
    h.pairdist_set_descriptors(native.bf16_bits(lv))
    h.pairdist_build()
    nb = rng.integers(0, 60, size=n).astype(np.int32)
    for it in range(2):
        sel, _ = h.select_dense(nb, budget=1200, first_index=7)
    ms = h.stage_ms()["dense_greedy"]
    print("n", n, "steps", len(sel), "ms", ms, "us/step", ms * 1e3 / len(sel))
