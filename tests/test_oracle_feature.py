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
