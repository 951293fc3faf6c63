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
