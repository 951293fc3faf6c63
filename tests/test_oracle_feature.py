"""CPU: the feature-distance restatement (oracle/restatement.py, the reference's own torch expressions) against the
golden produced by the UNMODIFIED reference FeatureSelector (oracle/make_golden_feature.py)."""
import numpy as np

import restatement as R
from adsel_b200 import synthetic
from conftest import load_golden


def test_feature_restatement_reproduces_the_reference(pools):
    golden = load_golden("P2K_feature.json")
    infos = pools("P2K")["infos"]
    feats = synthetic.synthetic_features(golden["n_frames"], golden["dim"], golden["feature_seed"])
    n_boxes = np.array([i["gt_names"].shape[0] for i in infos])
    for cid, case in golden["cases"].items():
        buffer = {k: list(v) for k, v in case["initial_buffer"].items()}
        for budget in case["rounds"]:
            key, sel = R.feature_select(feats, n_boxes, buffer, budget, p=case["p"])
            assert sel == case["selected_index"][key], cid
            buffer[key] = sel


def test_p2_is_an_l1_distance_like_p1():
    """feature_selector.py:99: ``sqrt((a-b)**2).sum(1)`` -- the square root is taken per element."""
    feats = synthetic.synthetic_features(64, 96, 3)
    r1, r2 = R.feature_row(feats, 5, 1), R.feature_row(feats, 5, 2)
    assert np.allclose(r1, r2, rtol=1e-6)
    assert not np.allclose(r2, np.sqrt(((feats - feats[5]) ** 2).sum(1)), rtol=1e-3)


def test_fused_feature_restatement_reproduces_the_reference(pools):
    """SpatialFeatureSelector / SpatialTemporalFeatureSelector (feature term added to the geodesic / temporal terms):
    the streaming restatement against what the UNMODIFIED classes select (oracle/make_golden_feature.py --fused),
    including the spatial-only seeding of the spatial+feature selector and the float32 feature normalisation."""
    from conftest import first_index_for
    golden = load_golden("P2K_fused_feature.json")
    infos = pools("P2K")["infos"]
    feats = synthetic.synthetic_features(golden["n_frames"], golden["dim"], golden["feature_seed"])
    first = first_index_for(len(infos))
    kinds = {"SpatialFeatureSelector": "spatial_feature", "SpatialTemporalFeatureSelector": "spatial_temporal_feature"}
    for cid, case in golden["cases"].items():
        buffer = {k: list(v) for k, v in case["initial_buffer"].items()}
        for budget in case["rounds"]:
            key, sel = R.fused_feature_select(kinds[case["selector"]], infos, feats, buffer, budget, first, **case["kwargs"])
            assert sel == case["selected_index"][key], cid
            buffer[key] = sel
    assert len(golden["cases"]) == 11


def test_torch_sum_restatement_is_torchs_sum_bit_for_bit():
    """The summation order csrc/featdist.cu follows (oracle/restatement.py::torch_sum_f32) IS torch's CPU ``.sum(1)`` of a
    contiguous float32 matrix: lengths below / at / above one cascade block, multiples of 32 and ragged tails."""
    import torch
    rng = np.random.default_rng(5)
    for dim in (16, 24, 32, 100, 512, 520, 768, 1000, 1024, 2048, 8192):
        feats = rng.standard_normal((12, dim)).astype(np.float32)
        t = torch.from_numpy(feats)
        for p in (1, 2):
            want = (torch.abs(t - t[3]).sum(1) if p == 1 else torch.sqrt((t - t[3]) ** 2).sum(1)).numpy()
            terms = np.abs(feats - feats[3]) if p == 1 else np.sqrt((feats - feats[3]) ** 2)
            got = np.array([R.torch_sum_f32(row) for row in terms], dtype=np.float32)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dim, p)
