"""Times the feature-distance map kernel (development aid, GPU box only)."""
import sys
sys.path.insert(0, "/root/repo")
import adsel_b200  # noqa: F401
from adsel_b200 import native, synthetic
h = native.Handle(0)
for n in (8000, 28130):
    f = synthetic.synthetic_features(n, 512)
    for it in range(3):
        h.feature_distance_build(f, 2)
    ms = h.stage_ms()["pairdist"]
    print("feature map n", n, "ms", ms, "T pair-terms/s", n * n * 512 / ms / 1e9, "FP32 TFLOP/s (2 adds per term)", 2 * n * n * 512 / ms / 1e9)
