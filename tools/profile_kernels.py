"""Small driver for ncu captures of the two roofline kernels (bev histogram, tcgen05 distance tiles)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import adsel_b200
from adsel_b200 import native
F = int(sys.argv[1]) if len(sys.argv) > 1 else 768
N = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
h = native.Handle(0)
sizes = 300_000 + (np.arange(F) * 977) % 4001 - 2000
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
h.bev_generate(off, 10, 1)
for it in range(3):
    h.bev_build(aggregate_atomics=False)
print("bev ms", h.stage_ms()["bev_histogram"], "GB/s", off[-1] * 20 / h.stage_ms()["bev_histogram"] / 1e6)
rng = np.random.default_rng(0)
lv = rng.integers(0, 20, size=(N, 16384), dtype=np.int64)
lv[rng.random((N, 16384)) < 0.7] = 0
h.pairdist_set_descriptors(native.bf16_bits(lv))
for it in range(3):
    h.pairdist_build()
ms = h.stage_ms()["pairdist"]
print("pairdist ms", ms, "TFLOP/s", 2.0 * N * N * 16384 / ms / 1e9)
h.close()
