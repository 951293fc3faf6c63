"""Times the tcgen05 distance tiles and the dense greedy loop (development aid, GPU box only)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import adsel_b200
from adsel_b200 import native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28130
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
rng = np.random.default_rng(0)
lv = rng.integers(0, 20, size=(n, dim), dtype=np.int64)
lv[rng.random((n, dim)) < 0.7] = 0
h = native.Handle(0)
h.pairdist_set_descriptors(native.bf16_bits(lv))
for it in range(4):
    h.pairdist_build()
    ms = h.stage_ms()["pairdist"]
    print(f"n={n} dim={dim} ms={ms:.3f} TFLOP/s={2.0*n*n*dim/ms/1e9:.1f} frac_of_1393.8={2.0*n*n*dim/ms/1e9/1393.8:.3f}")
rows = h.pairdist_get_rows(n - 64, 64)
a = lv[n - 64:].astype(np.int64)            # exact integer check of the last rows (inline: tools do not use oracle/)
ref = (a ** 2).sum(1)[:, None] + (lv ** 2).sum(1)[None, :] - 2 * (a @ lv.T)
print("last rows exact:", np.array_equal(rows.astype(np.int64), ref))
nb = rng.integers(0, 60, size=n).astype(np.int32)
sel, cost = h.select_dense(nb, budget=1200, first_index=1862)
print("dense greedy", len(sel), "frames", h.stage_ms()["dense_greedy"], "ms", h.counters())
