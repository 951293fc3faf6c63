"""CPU: the oracle restatement reproduces the unmodified reference (goldens) index for index."""
import numpy as np
import pytest

import restatement as R
from conftest import KIND, first_index_for, load_golden


def run_oracle_case(pool, case):
    infos, logs = pool["infos"], pool["logs"]
    kw = dict(case["kwargs"])
    cb, cf = kw.pop("cost_b", 0.04), kw.pop("cost_f", 0.12)
    buffer, res = {"0": []}, {}
    for budget in case["rounds"]:
        first = first_index_for(len(infos))
        try:
            key, sel = R.select(KIND[case["selector"]], infos, logs, buffer, budget, first, cost_b=cb, cost_f=cf, **kw)
        except AssertionError as e:
            key = str(budget + max(int(k) for k in buffer))
            res[key] = {"raises": "AssertionError", "message": str(e)}
            break
        buffer[key] = sel
        res[key] = sel
    return res


@pytest.mark.parametrize("pool_name", ["P300", "P2K"])
def test_oracle_matches_reference_goldens(pools, pool_name):
    golden = load_golden(f"{pool_name}_all.json")
    pool = pools(pool_name)
    assert golden["n_frames"] == len(pool["infos"])
    for cid, case in golden["cases"].items():
        assert run_oracle_case(pool, case) == case["selected_index"], cid


def test_oracle_matches_reference_goldens_p8k_exp(pools):
    golden = load_golden("P8K_all.json")
    pool = pools("P8K")
    for cid in ("st_exp_sum", "st_exp_sum_rounds", "temporal"):
        case = golden["cases"][cid]
        assert run_oracle_case(pool, case) == case["selected_index"], cid


def test_first_index_is_the_references_random_choice():
    # random.seed(3407); random.choice(range(28130)) == 1862, range(281300) -> 29798
    assert first_index_for(28130) == 1862
    assert first_index_for(281300) == 29798


def test_null_edge_rule_drops_tiny_weights():
    # scipy's dense->CSR conversion masks |w| <= 1e-8 (np.ma.masked_values atol): duplicates lose their edge
    loc = np.array([[0.0, 0.0], [5e-9, 0.0], [1.0, 0.0], [2.0, 0.0], [3.5, 0.0]])
    csr = R.knn_graph_csr(loc, k=2)
    dense = csr.toarray()
    assert dense[0, 1] == 0.0 and dense[1, 0] == 0.0
    assert dense[0, 2] == 1.0 and dense[2, 0] == 1.0
    rows = R.geodesic_rows(csr, [0])
    assert rows[0, 1] == 1.0 + (1.0 - 5e-9)      # reached through vertex 2, not directly


def test_consecutive_runs_ignores_last_run_for_max():
    infos = [{"cam_front_path": f"x/{lf}__CAM_FRONT__{i}.jpg"} for i, lf in enumerate("aabbbaacccc")]
    run_id, max_td = R.consecutive_runs(infos)
    assert run_id.tolist() == [0, 0, 1, 1, 1, 2, 2, 3, 3, 3, 3]
    assert max_td == 3      # the final run of 4 is never counted (reference :127-129)


def test_restatement_reproduces_the_p850_temporal_golden(pools):
    """BASELINE.json config 4 (TemporalSelector, 850 scenes, budget 1200): oracle port vs the unmodified reference."""
    golden = load_golden("P850_temporal.json")
    pl = pools("P850")
    case = golden["cases"]["temporal_1200"]
    key, sel = R.select("temporal", pl["infos"], pl["logs"], {"0": []}, 1200, first_index_for(len(pl["infos"])))
    assert sel == case["selected_index"][key]


def test_restatement_reproduces_the_p28k_spatial_golden(pools):
    """BASELINE.json config 2's engine at full size: the unmodified SpatialSelector on the 28,130-frame pool at budget
    1200 (oracle/make_goldens.py --pool P28K --only spatial) vs the streaming restatement."""
    golden = load_golden("P28K_spatial.json")
    pl = pools("P28K")
    case = golden["cases"]["spatial_1200"]
    key, sel = R.select("spatial", pl["infos"], pl["logs"], {"0": []}, 1200, first_index_for(len(pl["infos"])))
    assert key == "1200" and sel == case["selected_index"][key]


def test_restatement_reproduces_a_full_size_min_aggregate_golden(pools):
    """28,130 frames, exp normalisation, aggregate="min", budget 1200 (oracle/make_goldens.py --pool P28K --only st_agg):
    the streaming restatement against what the unmodified reference selected."""
    golden = load_golden("P28K_st_agg.json")
    pl = pools("P28K")
    case = golden["cases"]["st_exp_min_1200"]
    key, sel = R.select("spatial_temporal", pl["infos"], pl["logs"], {"0": []}, 1200, first_index_for(len(pl["infos"])),
                        **case["kwargs"])
    assert key == "1200" and sel == case["selected_index"][key]
