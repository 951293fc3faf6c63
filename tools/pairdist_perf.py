import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import adsel_b200
from adsel_b200 import native
import bev_oracle as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28130
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
rng = np.random.default_rng(0)
lv = rng.integers(0, 20, size=(n, dim), dtype=np.int64)
lv[rng.random((n, dim)) < 0.7] = 0
h = native.Handle(0)
h.pairdist_set_descriptors(B.bf16_bits(lv))
for it in range(4):
    h.pairdist_build()
    ms = h.stage_ms()["pairdist"]
    print(f"n={n} dim={dim} ms={ms:.3f} TFLOP/s={2.0*n*n*dim/ms/1e9:.1f} frac_of_1393.8={2.0*n*n*dim/ms/1e9/1393.8:.3f}")
rows = h.pairdist_get_rows(n - 64, 64)
ref = B.sq_distance_matrix(lv[n - 64:], lv)
print("last rows exact:", np.array_equal(rows.astype(np.int64), ref))
nb = rng.integers(0, 60, size=n).astype(np.int32)
sel, cost = h.select_dense(nb, budget=1200, first_index=1862)
print("dense greedy", len(sel), "frames", h.stage_ms()["dense_greedy"], "ms", h.counters())
