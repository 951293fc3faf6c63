"""GPU: the device float32 exp (csrc/common.cuh::np_exp_f32) is numpy's float32 exp bit for bit -- the building block
of the fused feature selectors' normalisation (spatial_feature_selector.py:190).  Checked through the C ABI
(adsel_exp_f32) against oracle/npexp.c::oracle_npexpf, which is pinned to np.exp on all 2^32 float32 inputs."""
import numpy as np
import pytest

import restatement as R
from adsel_b200 import native

pytestmark = pytest.mark.gpu


def test_device_exp_f32_is_numpys_float32_exp_bit_for_bit():
    h = native.Handle(0)
    try:
        def check(x):
            got, want = h.exp_f32(x), R.npexp_f32(x)
            same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), (x[~same][:5], got[~same][:5], want[~same][:5])

        # every 64th float32 bit pattern over the whole range (both signs, denormals, infinities, NaNs)
        check(np.arange(0, 1 << 32, 64, dtype=np.uint64).astype(np.uint32).view(np.float32))
        rng = np.random.default_rng(3)
        check((-rng.uniform(0.0, 104.0, 4_000_000)).astype(np.float32))          # -feature distance: the selectors' range
        check(rng.uniform(-1e-3, 1e-3, 200_000).astype(np.float32))
        # every input whose result is denormal or at the underflow / overflow edge
        lo, hi = np.float32(-104.0).view(np.uint32), np.float32(-87.0).view(np.uint32)
        check(np.arange(int(hi), int(lo) + 1, dtype=np.uint32).view(np.float32))
        check(np.arange(int(np.float32(88.0).view(np.uint32)), int(np.float32(89.0).view(np.uint32)), dtype=np.uint32).view(np.float32))
        d = np.abs(rng.normal(0, 20, 100_000)).astype(np.float32)
        assert np.array_equal((np.float32(1) - h.exp_f32(-d)).view(np.uint32), R.exp_normalize_f32(d).view(np.uint32))
    finally:
        h.close()
