"""GPU: rows without a saturation radius (normalize="linear", SpatialSelector's raw distances) through the lazy loop
(csrc/lazy.cu: one truncation-free search per selected frame) -- against the reference's goldens."""
import pytest

from conftest import load_golden
from test_gpu_parity import _run_case_through_selectors

pytestmark = pytest.mark.gpu

FAR_INF = ("st_linear_sum", "st_linear_min", "st_linear_max", "st_linear_sum_rounds", "spatial", "spatial_rounds")


@pytest.mark.parametrize("pool_name", ["P300", "P2K", "P8K"])
def test_lazy_rows_select_what_the_reference_selects(pools, tmp_path, monkeypatch, pool_name):
    """ADSEL_LAZY=1 forces the lazy loop on the small pools whose goldens the list path already reproduces: every
    far = inf case (three aggregates, chained rounds with resumed buffers, SpatialSelector with forbid_sampled)."""
    monkeypatch.setenv("ADSEL_LAZY", "1")
    golden = load_golden(f"{pool_name}_all.json")
    pl = pools(pool_name)
    ran = 0
    for cid in FAR_INF:
        case = golden["cases"][cid]
        res, _ = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], f"{pool_name}/{cid}"
        ran += 1
    assert ran == len(FAR_INF)


def test_lazy_rows_raise_the_reference_assertion_on_exhaustion(pools, tmp_path, monkeypatch):
    monkeypatch.setenv("ADSEL_LAZY", "1")
    pl = pools("P300")
    case = {"selector": "SpatialTemporalSelector", "kwargs": {"normalize": "linear"}, "rounds": [10 ** 6]}
    res, _ = _run_case_through_selectors(pl, case, tmp_path)
    (only,) = res.values()
    assert only["raises"] == "AssertionError" and "has been selected" in only["message"]


def test_p28k_spatial_selector_budget_1200_identical_to_reference(pools, tmp_path):
    """BASELINE.json config 2's engine at full size: SpatialSelector on the 28,130-frame pool, budget 1200 (1,020 frames).
    Lazy rows by default at this size: the all-pairs lists would need the sum of squared component sizes."""
    golden = load_golden("P28K_spatial.json")
    pl = pools("P28K")
    for cid, case in golden["cases"].items():
        res, _ = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], cid


def test_p28k_linear_normalisation_identical_to_reference(pools, tmp_path):
    """normalize="linear" at full size (sum at budget 1200, min at 600, a chained 300+300 resume)."""
    golden = load_golden("P28K_st_linear.json")
    pl = pools("P28K")
    for cid, case in golden["cases"].items():
        res, _ = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], cid
