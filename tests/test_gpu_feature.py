"""GPU: feature-distance map (csrc/featdist.cu), FeatureSelector and the fused feature selectors (csrc/fused.cu) against
the reference.

The map is float32.  For feature dimensions that are multiples of 32 (the reference's 512) the kernel follows torch's
vectorised cascade sum and the map equals the reference's bit for bit; other dimensions are summed in index order and
agree to 1e-5 relative (north_star's fp32 tolerance)."""
import json
import pickle

import numpy as np
import pytest

import restatement as R
from adsel_b200 import native, synthetic
from adsel_b200.selectors import build_selector
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,dim,p", [(1, 8, 1), (130, 96, 2), (700, 512, 2), (513, 33, 1)])
def test_feature_distance_rows_within_fp32_tolerance(n, dim, p):
    feats = synthetic.synthetic_features(n, dim, seed=n + dim)
    feats[n // 2] = feats[0]                                  # an exact duplicate: distance exactly 0
    h = native.Handle(0)
    try:
        h.feature_distance_build(feats, p)
        got = h.feature_distance_rows(0, n)
        ref = np.stack([R.feature_row(feats, i, p) for i in range(n)])
        assert np.allclose(got, ref, rtol=1e-5, atol=0.0)   # tolerance stated in BASELINE.json north_star for fp32
        assert np.all(np.diag(got) == 0.0) and got[0, n // 2] == 0.0
        assert np.array_equal(got, got.T)                    # |a-b| == |b-a| term by term, same summation order
    finally:
        h.close()


@pytest.mark.parametrize("n,dim,p", [(130, 96, 2), (300, 512, 1), (300, 512, 2), (257, 32, 1), (150, 768, 2), (90, 1024, 1)])
def test_feature_distance_rows_bit_for_bit_in_torchs_summation_order(n, dim, p):
    feats = synthetic.synthetic_features(n, dim, seed=n + dim)
    feats[n // 2] = feats[0]
    h = native.Handle(0)
    try:
        h.feature_distance_build(feats, p)
        got = h.feature_distance_rows(0, n)
        ref = np.stack([R.feature_row(feats, i, p) for i in range(n)])       # torch itself
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        with pytest.raises(native.NativeError):
            bad = feats.copy()
            bad[3, 5] = np.nan
            h.feature_distance_build(bad, p)
    finally:
        h.close()


def test_fused_feature_selectors_drop_in_select_like_the_reference(pools, tmp_path):
    """SpatialFeatureSelector / SpatialTemporalFeatureSelector through the registry, 11 cases recorded from the UNMODIFIED
    classes: sum / min / max, p = 1, other k and lambdas, chained rounds, a resumed buffer."""
    import random
    import torch
    golden = load_golden("P2K_fused_feature.json")
    pl = pools("P2K")
    feats = synthetic.synthetic_features(golden["n_frames"], golden["dim"], golden["feature_seed"])
    feat_path = str(tmp_path / "feature_pred.pt")
    torch.save(torch.from_numpy(feats), feat_path)
    for cid, case in golden["cases"].items():
        buffer_file = str(tmp_path / f"buffer_{cid}.json")
        json.dump(case["initial_buffer"], open(buffer_file, "w"))
        res = {}
        for budget in case["rounds"]:
            random.seed(3407)
            sel = build_selector(dict(type=case["selector"], budget=budget, buffer_file=buffer_file,
                                      infos_origin=pl["paths"]["infos"], buffer_path=feat_path, logs_file=pl["paths"]["logs"],
                                      distance_store_file=str(tmp_path / "unused.npy"), pred=False, **case["kwargs"]))
            sel.select_samples(local_rank=0)
            sel.dump_file()
            res.update(sel.get_selected_samples())
        assert res == case["selected_index"], cid
    assert len(golden["cases"]) == 11


def test_p2_extreme_magnitudes_in_torchs_order_too():
    """The torch-order kernel skips the IEEE square root of p=2 only where sqrt(fl(d*d)) == |d| is provable; rows / tiles with
    values outside that range take the literal expression (dimension 64: the bit-exact kernel)."""
    rng = np.random.default_rng(8)
    feats = rng.standard_normal((200, 64)).astype(np.float32)
    feats[3, 5] = 1e-30
    feats[150, 9] = 3e19
    feats[77] = 0.0
    h = native.Handle(0)
    try:
        for p in (1, 2):
            h.feature_distance_build(feats, p)
            got = h.feature_distance_rows(0, len(feats))
            ref = np.stack([R.feature_row(feats, i, p) for i in range(len(feats))])
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), p
    finally:
        h.close()


def test_p2_extreme_magnitudes_follow_the_literal_expression():
    """sqrt(fl(d*d)) differs from |d| only when d*d underflows / overflows float32: those terms are evaluated literally."""
    feats = np.zeros((4, 8), dtype=np.float32)
    feats[1, 0] = 1e-30          # d*d underflows to 0 -> term 0, although |d| > 0
    feats[2, 1] = 3e19           # d*d overflows to inf -> term inf
    feats[3, 2] = 2.5
    h = native.Handle(0)
    try:
        h.feature_distance_build(feats, 2)
        got = h.feature_distance_rows(0, 4)
        ref = np.stack([R.feature_row(feats, i, 2) for i in range(4)])
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        h.feature_distance_build(feats, 1)
        ref1 = np.stack([R.feature_row(feats, i, 1) for i in range(4)])
        assert np.array_equal(h.feature_distance_rows(0, 4).view(np.uint32), ref1.view(np.uint32))
    finally:
        h.close()


def test_feature_selector_drop_in_selects_like_the_reference(pools, tmp_path):
    import torch
    golden = load_golden("P2K_feature.json")
    pl = pools("P2K")
    feats = synthetic.synthetic_features(golden["n_frames"], golden["dim"], golden["feature_seed"])
    feat_path = str(tmp_path / "feature_pred.pt")
    torch.save(torch.from_numpy(feats), feat_path)
    for cid, case in golden["cases"].items():
        buffer_file = str(tmp_path / f"buffer_{cid}.json")
        json.dump(case["initial_buffer"], open(buffer_file, "w"))
        res = {}
        for budget in case["rounds"]:
            sel = build_selector(dict(type="FeatureSelector", budget=budget, buffer_file=buffer_file,
                                      infos_origin=pl["paths"]["infos"], buffer_path=feat_path, p=case["p"],
                                      distance_store_file=str(tmp_path / "unused.npy"), pred=False))
            sel.select_samples(local_rank=0)
            sel.dump_file()
            res.update(sel.get_selected_samples())
        assert res == case["selected_index"], cid
        last = max(res, key=int)
        sub = pickle.load(open(pl["paths"]["infos"].replace(".pkl", f"_{last}.pkl"), "rb"))
        assert [s["token"] for s in sub] == [pl["infos"][i]["token"] for i in res[last]]
    with pytest.raises(NotImplementedError):
        build_selector(dict(type="FeatureSelector", budget=10, buffer_file=buffer_file, infos_origin=pl["paths"]["infos"],
                            buffer_path=feat_path, pred=True)).select_samples(local_rank=0)
