"""CPU: oracle/npexp.c is pinned bit-for-bit against numpy's own float64 exp."""
import numpy as np
import pytest

import restatement as R


def _numpy_uses_svml():
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
    except Exception:  # pragma: no cover
        return False
    return bool(feats.get("AVX512_SKX"))


@pytest.mark.skipif(not _numpy_uses_svml(), reason="numpy dispatches float64 exp to SVML only on AVX512_SKX hosts")
def test_npexp_bit_exact_against_numpy():
    rng = np.random.default_rng(7)
    for lo, hi, n in [(-40.0, 0.0, 2_000_000), (-707.0, -600.0, 100_000), (-1.0, 1.0, 500_000),
                      (0.0, 707.0, 200_000), (-1e-15, 1e-15, 50_000), (-1e-3, 0.0, 200_000)]:
        x = rng.uniform(lo, hi, n)
        assert np.array_equal(R.npexp(x).view(np.uint64), np.exp(x).view(np.uint64)), (lo, hi)
    x = -np.arange(0, 700, dtype=np.float64)                      # the temporal term's integer gaps
    assert np.array_equal(R.npexp(x).view(np.uint64), np.exp(x).view(np.uint64))
    x = -np.arange(0, 640) / 16.0 * np.log(2.0)                   # table-index boundaries
    assert np.array_equal(R.npexp(x).view(np.uint64), np.exp(x).view(np.uint64))


def test_exp_normalize_saturates_to_one_beyond_38():
    # the truncation radius of the geodesic search relies on this
    d = np.concatenate([np.linspace(38.0, 60.0, 10001), [1e3, 1e6, np.inf]])
    assert np.all(R.exp_normalize(d) == 1.0)
    assert R.exp_normalize(np.array([0.0]))[0] == 0.0
    assert R.exp_normalize(np.array([37.0]))[0] < 1.0


def _numpy_has_simd_float_exp():
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as feats
    except Exception:  # pragma: no cover
        return False
    return bool(feats.get("AVX512F") or (feats.get("AVX2") and feats.get("FMA3")))


@pytest.mark.skipif(not _numpy_has_simd_float_exp(), reason="numpy's float32 exp kernel needs AVX512F or AVX2+FMA3")
def test_npexp_f32_bit_exact_against_numpy():
    """float32 exp (the feature term's normalisation, spatial_feature_selector.py:190).  The whole float32 range
    (2^32 inputs, ~2 min) was compared in the build container; here: every 2^8-th bit pattern plus dense samples of
    the ranges the selectors use and the overflow / underflow / denormal edges.  ADSEL_EXHAUSTIVE=1 runs it all."""
    import os
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning)

    def check(x):
        a, b = np.exp(x), R.npexp_f32(x)
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))     # NaN payloads aside
        assert same.all(), x[~same][:5]

    stride = 1 if os.environ.get("ADSEL_EXHAUSTIVE") else 256
    for c in range(0, 256):
        u = (np.arange(0, 1 << 24, stride, dtype=np.uint64) + (c << 24)).astype(np.uint32)
        check(u.view(np.float32))
    rng = np.random.default_rng(3)
    check((-rng.uniform(0.0, 104.0, 3_000_000)).astype(np.float32))       # -feature distance
    check(rng.uniform(-1e-3, 1e-3, 200_000).astype(np.float32))
    edges = np.array([0.0, -0.0, 88.72283, 88.72284, 88.7229, -87.33654, -87.33655, -103.97208, -103.9721, -104.0,
                      np.inf, -np.inf, np.nan], dtype=np.float32)
    check(edges)
    d = np.abs(rng.normal(0, 20, 100_000)).astype(np.float32)
    assert np.array_equal(R.exp_normalize_f32(d).view(np.uint32), (1 - np.exp(-d)).view(np.uint32))
    assert R.exp_normalize_f32(d).dtype == np.float32
