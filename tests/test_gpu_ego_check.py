"""GPU: the all-frames ego-location check (csrc/ego_check.cu) -- every frame's ``-(cal[:3,3].T @ cal[:3,:3])[:2]``
(spatial_temporal_selector.py:83-91) redone on the device in the literal order and compared bit-for-bit with what the
host extraction produced."""
import numpy as np
import pytest

import restatement as R
from adsel_b200 import native, pool as P

pytestmark = pytest.mark.gpu


def _pose9(infos):
    return np.stack([i["car_from_global"][:3][:, [0, 1, 3]].reshape(9) for i in infos])


def test_every_frame_of_a_pool_is_verified_on_the_device(pools):
    infos = pools("P2K")["infos"]
    P.LAST_EGO_CHECK.clear()
    names, loc, boxes = P.extract_slice(infos, device=0)
    assert P.LAST_EGO_CHECK["frames"] == len(infos) and P.LAST_EGO_CHECK["n_bad"] == 0 and P.LAST_EGO_CHECK["first_bad"] == -1
    want = R.ego_locations(infos)                       # the reference's literal per-frame expression
    assert np.array_equal(loc.view(np.uint64), want.view(np.uint64))
    fused = P.LAST_EGO_CHECK["fused"]
    assert native.check_ego_locations(0, _pose9(infos), want, fused) == (0, -1)
    # a slice checks only its frames
    P.extract_slice(infos, 700, 1501, device=0)
    assert P.LAST_EGO_CHECK["frames"] == 801 and P.LAST_EGO_CHECK["n_bad"] == 0
    # one ulp off in two frames, a flipped sign of zero in a third: counted, and the first is named
    off = want.copy()
    off[17, 1] = np.nextafter(off[17, 1], np.inf)
    off[1400, 0] = np.nextafter(off[1400, 0], -np.inf)
    assert native.check_ego_locations(0, _pose9(infos), off, fused) == (2, 17)
    # the other rounding order is a different function of general matrices (this is what the probe frames pin)
    rng = np.random.default_rng(5)
    mats = rng.normal(size=(4096, 4, 4))
    gen = [{"car_from_global": m} for m in mats]
    lit = R.ego_locations(gen)
    assert native.check_ego_locations(0, _pose9(gen), lit, fused) == (0, -1)
    n_bad, first = native.check_ego_locations(0, _pose9(gen), lit, not fused)
    assert n_bad > 100 and first >= 0


def test_special_values_and_the_python_path():
    """NaN matches NaN whatever its payload, -0.0 is not +0.0, infinities are compared like any number; the batched numpy
    path (no C helper) is checked under both rounding orders."""
    pose = np.zeros((6, 9))
    pose[1, [2, 5, 8]] = [1.0, 2.0, 3.0]                 # translation only: both coordinates are -(0) = -0.0
    pose[2] = [np.nan, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]
    pose[3] = [np.inf, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]
    pose[4] = [1e308, 1.0, 10.0, 1e308, 1.0, 10.0, 1.0, 1.0, 1.0]     # overflow in the products
    pose[5] = [3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0, 11.0]
    mats = np.zeros((6, 4, 4))
    mats[:, :3, 0], mats[:, :3, 1], mats[:, :3, 3] = pose[:, 0::3], pose[:, 1::3], pose[:, 2::3]
    with np.errstate(all="ignore"):
        lit = np.stack([(-(m[:3, 3].T @ m[:3, :3]))[:2] for m in mats])
    for fused in (True, False):
        assert native.check_ego_locations(0, pose, lit, fused) == (0, -1), fused      # exact small integers / specials
    assert np.signbit(lit[1]).all()
    flipped = lit.copy()
    flipped[1, 0] = 0.0
    assert native.check_ego_locations(0, pose, flipped, True) == (1, 1)
    other_nan = lit.copy()
    other_nan[2, 0] = np.float64(np.frombuffer(np.uint64(0xFFF8000000000123).tobytes(), dtype=np.float64)[0])
    assert native.check_ego_locations(0, pose, other_nan, True) == (0, -1)
    assert native.check_ego_locations(0, np.zeros((0, 9)), np.zeros((0, 2)), True) == (0, -1)
    with pytest.raises(native.NativeError):
        native.check_ego_locations(99, pose, lit, True)
    infos = [{"car_from_global": m} for m in np.random.default_rng(1).normal(size=(500, 4, 4))]
    P.LAST_EGO_CHECK.clear()
    loc = P.ego_locations(infos, device=0)
    assert P.LAST_EGO_CHECK["frames"] == 500 and P.LAST_EGO_CHECK["n_bad"] == 0
    assert np.array_equal(loc.view(np.uint64), R.ego_locations(infos).view(np.uint64))


def test_selector_runs_the_check(pools, tmp_path):
    import json
    import random
    from adsel_b200.selectors import build_selector
    pl = pools("P850")
    buf = tmp_path / "buffer.json"
    buf.write_text(json.dumps({"0": []}))
    P.LAST_EGO_CHECK.clear()
    random.seed(3407)
    sel = build_selector(dict(type="SpatialTemporalSelector", budget=20, buffer_file=str(buf), infos_origin=pl["paths"]["infos"],
                              logs_file=pl["paths"]["logs"], distance_store_file=str(tmp_path / "d.npy")))
    sel.select_samples()
    assert P.LAST_EGO_CHECK["frames"] == len(pl["infos"]) and P.LAST_EGO_CHECK["n_bad"] == 0
