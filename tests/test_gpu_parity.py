"""GPU: parity of the sm_100a path against the oracle and the reference's goldens, through the C ABI."""
import json
import os
import pickle
import random

import numpy as np
import pytest

import restatement as R
from adsel_b200 import native, pool
from adsel_b200.selectors import build_selector
from conftest import KIND, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = native.Handle(0)
    yield h
    h.close()


def test_device_exp_is_numpys_exp_bit_for_bit(handle):
    rng = np.random.default_rng(11)
    xs = [rng.uniform(-40.0, 0.0, 4_000_000), rng.uniform(-707.0, 707.0, 1_000_000), rng.uniform(-1e-3, 1e-3, 500_000),
          -np.arange(0, 700, dtype=np.float64), -np.arange(0, 640) / 16.0 * np.log(2.0), np.array([0.0, -0.0, -38.0, -37.43])]
    for x in xs:
        assert np.array_equal(handle.exp_f64(x).view(np.uint64), R.npexp(x).view(np.uint64))


def _set_spatial_pool(handle, infos):
    loc = pool.ego_locations(infos)
    run_id, max_td = pool.consecutive_runs(pool.logfile_names(infos))
    handle.set_pool(loc, run_id, pool.box_counts(infos))
    return loc, run_id, max_td


@pytest.mark.parametrize("pool_name,k", [("P300", 8), ("P2K", 8), ("P2K", 4), ("P2K", 11), ("P8K", 8)])
def test_knn_matches_ckdtree(handle, pools, pool_name, k):
    infos = pools(pool_name)["infos"]
    loc, _, _ = _set_spatial_pool(handle, infos)
    handle.build_knn_graph(k)
    dist, idx = handle.get_knn()
    ref_d, ref_i = R.knn_query(loc, k)
    assert np.array_equal(dist.view(np.uint64), ref_d.view(np.uint64))      # bit-exact float64 distances
    assert np.array_equal(idx, ref_i)


@pytest.mark.parametrize("pool_name", ["P300", "P2K", "P8K"])
def test_csr_matches_scipy_conversion(handle, pools, pool_name):
    infos = pools(pool_name)["infos"]
    loc, _, _ = _set_spatial_pool(handle, infos)
    handle.build_knn_graph(8)
    row_ptr, col, w = handle.get_csr()
    ref = R.knn_graph_csr(loc, 8)
    ref.sort_indices()
    assert np.array_equal(row_ptr, ref.indptr)
    for i in range(len(loc)):
        o = np.argsort(col[row_ptr[i]:row_ptr[i + 1]])
        assert np.array_equal(col[row_ptr[i]:row_ptr[i + 1]][o], ref.indices[ref.indptr[i]:ref.indptr[i + 1]])
        assert np.array_equal(w[row_ptr[i]:row_ptr[i + 1]][o].view(np.uint64),
                              ref.data[ref.indptr[i]:ref.indptr[i + 1]].view(np.uint64))


def test_csr_drops_null_edges(handle):
    loc = np.array([[0.0, 0.0], [5e-9, 0.0], [1.0, 0.0], [2.0, 0.0], [3.5, 0.0]])
    handle.set_pool(loc, np.zeros(5, np.int32), np.ones(5, np.int32))
    handle.build_knn_graph(2)
    row_ptr, col, w = handle.get_csr()
    ref = R.knn_graph_csr(loc, 2)
    assert np.array_equal(row_ptr, ref.indptr)
    assert 1 not in col[row_ptr[0]:row_ptr[1]].tolist()
    rows = handle.geodesic_rows([0])
    assert np.array_equal(rows.view(np.uint64), R.geodesic_rows(ref, [0]).view(np.uint64))


@pytest.mark.parametrize("pool_name,m", [("P300", 300), ("P2K", 257), ("P8K", 64)])
def test_geodesic_rows_match_scipy_dijkstra_bitwise(handle, pools, pool_name, m):
    infos = pools(pool_name)["infos"]
    loc, _, _ = _set_spatial_pool(handle, infos)
    handle.build_knn_graph(8)
    rng = np.random.default_rng(5)
    src = rng.choice(len(loc), size=m, replace=False) if m < len(loc) else np.arange(len(loc))
    got = handle.geodesic_rows(src)
    ref = R.geodesic_rows(R.knn_graph_csr(loc, 8), src)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    # truncated search: exact below the radius, "far" (inf) at or beyond it
    got38 = handle.geodesic_rows(src, radius=38.0)
    inside = ref < 38.0
    assert np.array_equal(got38[inside].view(np.uint64), ref[inside].view(np.uint64))
    assert np.all(np.isinf(got38[~inside]))


def test_geodesic_max_finite(handle, pools):
    infos = pools("P2K")["infos"]
    loc, _, _ = _set_spatial_pool(handle, infos)
    handle.build_knn_graph(8)
    assert handle.geodesic_max_finite() == R.geodesic_max_finite(R.knn_graph_csr(loc, 8))


def _run_case_through_selectors(pl, case, tmp_path):
    """Drive the drop-in selector classes exactly like tools/active_select.py:152-163 does."""
    buffer_file = str(tmp_path / "buffer.json")
    json.dump({"0": []}, open(buffer_file, "w"), indent=4)
    res = {}
    for budget in case["rounds"]:
        cfg = dict(type=case["selector"], budget=budget, buffer_file=buffer_file, infos_origin=pl["paths"]["infos"],
                   detector=None, dataloader=None, logger=None, pred=False)
        if case["selector"] != "TemporalSelector":
            cfg.update(logs_file=pl["paths"]["logs"], distance_store_file=str(tmp_path / "unused.npy"))
        cfg.update(case["kwargs"])
        random.seed(3407)
        sel = build_selector(cfg)
        try:
            sel.select_samples(local_rank=0)
        except AssertionError as e:
            res[sel.current_budget] = {"raises": "AssertionError", "message": str(e)}
            break
        sel.dump_file()
        res.update(sel.get_selected_samples())
    return res, buffer_file


@pytest.mark.parametrize("pool_name", ["P300", "P2K", "P8K"])
def test_selected_index_identical_to_reference(pools, tmp_path, pool_name):
    golden = load_golden(f"{pool_name}_all.json")
    pl = pools(pool_name)
    for cid, case in golden["cases"].items():
        res, buffer_file = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], f"{pool_name}/{cid}"
        if all(isinstance(v, list) for v in res.values()):
            on_disk = json.load(open(buffer_file))
            assert on_disk == {"0": [], **res}
            last = max(res, key=int)
            sub = pickle.load(open(pl["paths"]["infos"].replace(".pkl", f"_{last}.pkl"), "rb"))
            assert [s["token"] for s in sub] == [pl["infos"][i]["token"] for i in res[last]]


def test_p28k_budgets_600_1200_4800_identical_to_reference(pools, tmp_path):
    """BASELINE.json target: identical selected-index JSON on the 28,130-frame pool."""
    golden = load_golden("P28K_st.json")
    pl = pools("P28K")
    for cid, case in golden["cases"].items():
        res, _ = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], cid


def test_p28k_other_aggregates_and_weights_identical_to_reference(pools, tmp_path):
    """Full size, the other exp-normalised variants (the one-warp kernel's `gval` table holds aggregate(s_far, tlut[d])):
    min / max aggregates, lambda_t = 0.5 with other cost weights, chained rounds -- what the unmodified reference selects
    (oracle/make_goldens.py --pool P28K --only st_agg)."""
    golden = load_golden("P28K_st_agg.json")
    pl = pools("P28K")
    for cid, case in golden["cases"].items():
        res, _ = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], cid


def test_fps_vector_matches_oracle(handle, pools):
    pl = pools("P2K")
    infos = pl["infos"]
    loc, run_id, max_td = _set_spatial_pool(handle, infos)
    handle.build_knn_graph(8)
    sel, cost = handle.select(spatial="geodesic", temporal=True, normalize="exp", aggregate="sum", lambda_t=1.0,
                              budget=300, first_index=17)
    prov = R.RowProvider("spatial_temporal", infos)
    fps = np.stack([prov.row(i) for i in sel]).min(0)
    assert np.array_equal(handle.get_fps().view(np.uint64), fps.view(np.uint64))
    n_boxes = pool.box_counts(infos)
    ref_sel = R.greedy_select(prov.row, len(infos), n_boxes, [], 17, "300")
    assert sel == ref_sel


def test_ragged_and_edge_inputs(handle):
    # one long run + singleton runs, k larger than some components, every frame its own group
    rng = np.random.default_rng(3)
    n = 97
    loc = np.cumsum(rng.normal(0, 3.0, size=(n, 2)), axis=0)
    run_id = np.concatenate([np.zeros(60), np.arange(1, n - 60 + 1)]).astype(np.int32)
    n_boxes = rng.integers(0, 5, size=n).astype(np.int32)
    handle.set_pool(loc, run_id, n_boxes)
    handle.build_knn_graph(3)
    infos = [{"cam_front_path": f"a/L{r}__CAM_FRONT__{i}.jpg", "gt_names": np.zeros(b, dtype="<U3"),
              "car_from_global": np.eye(4)} for i, (r, b) in enumerate(zip(run_id, n_boxes))]
    csr = R.knn_graph_csr(loc, 3)

    def row(i):
        s = R.exp_normalize(R.geodesic_rows(csr, [i])[0])
        t = R.exp_normalize(R.temporal_row(run_id, i))
        return s + 1 * t
    ref = R.greedy_select(row, n, n_boxes, [], 5, "9")
    got, _ = handle.select(spatial="geodesic", temporal=True, normalize="exp", budget=9, first_index=5)
    assert got == ref
    # zero-box frames and a budget that admits exactly the seed
    got, cost = handle.select(spatial="geodesic", temporal=True, normalize="exp", budget=0, first_index=5)
    assert got == [5]
    with pytest.raises(native.NativeError):
        handle.select(spatial="geodesic", temporal=True, normalize="exp", budget=5, first_index=n)


def test_cli_two_invocations_match_reference(pools, tmp_path):
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("active_select_cli", os.path.join(ROOT, "tools", "active_select.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    pl = pools("P2K")
    golden = load_golden("P2K_all.json")["cases"]["st_exp_sum_rounds"]
    buf = str(tmp_path / "cli_buffer.json")
    cfg = tmp_path / "cfg.py"
    cfg.write_text(f"selector = dict(type='SpatialTemporalSelector', budget=1, buffer_file={buf!r}, "
                   f"infos_origin={pl['paths']['infos']!r}, logs_file={pl['paths']['logs']!r})\n")
    assert cli.main(["--config", str(cfg), "--budget", "300"]) == 0      # first run: empty buffer only
    for b in golden["rounds"]:
        assert cli.main(["--config", str(cfg), "--budget", str(b)]) == 0
    assert json.load(open(buf)) == {"0": [], **golden["selected_index"]}


def test_grid_knn_equals_bruteforce_on_tie_heavy_and_degenerate_inputs(handle, monkeypatch):
    """The uniform-grid search and the brute-force scan are both exact with the same (d2, index) order:
    integer lattices (masses of exact distance ties), duplicated points, a single cluster far from a few
    outliers (almost everything in one cell) and collinear points must give identical tables."""
    rng = np.random.default_rng(12)
    lattice = np.stack(np.meshgrid(np.arange(48.0), np.arange(40.0)), -1).reshape(-1, 2)
    dup = np.repeat(rng.uniform(0, 50, size=(400, 2)), 4, axis=0)
    cluster = np.concatenate([rng.normal(0, 0.5, size=(3000, 2)), rng.uniform(-5e4, 5e4, size=(8, 2))])
    line = np.stack([np.linspace(0, 900, 2500), np.zeros(2500)], 1)
    for name, loc in (("lattice", lattice), ("dup", dup), ("cluster", cluster), ("line", line)):
        loc = np.ascontiguousarray(loc[rng.permutation(len(loc))])
        n = len(loc)
        tables = {}
        for mode in ("brute", "grid"):
            monkeypatch.setenv("ADSEL_KNN", mode)
            handle.set_pool(loc, np.zeros(n, np.int32), np.ones(n, np.int32))
            handle.build_knn_graph(8)
            tables[mode] = handle.get_knn()
        assert np.array_equal(tables["brute"][1], tables["grid"][1]), name
        assert np.array_equal(tables["brute"][0].view(np.uint64), tables["grid"][0].view(np.uint64)), name
    monkeypatch.delenv("ADSEL_KNN")


@pytest.mark.parametrize("cluster", ["1", "2", "4", "8", "16"])
def test_resident_variants_of_the_greedy_loop_select_identically(handle, pools, monkeypatch, cluster):
    """ADSEL_GREEDY_CLUSTER: fps resident in the shared memory of 1 CTA or of a 2..16-CTA thread-block cluster
    (exchange through DSMEM) against ADSEL_GREEDY_CLUSTER=0, the global-memory kernel."""
    infos = pools("P8K")["infos"]
    loc, run_id, max_td = _set_spatial_pool(handle, infos)
    handle.build_knn_graph(8)
    kw = dict(spatial="geodesic", temporal=True, normalize="exp", aggregate="sum", lambda_t=1.0, budget=600, first_index=4242)
    monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", "0")
    want, want_cost = handle.select(**kw)
    want_fps = handle.get_fps()
    want2, _ = handle.select(**{**kw, "budget": 900, "sampled": want, "first_index": -1, "cost_start": want_cost})
    monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", cluster)
    got, cost = handle.select(**kw)
    assert got == want and cost == want_cost
    assert np.array_equal(handle.get_fps().view(np.uint64), want_fps.view(np.uint64))
    got2, _ = handle.select(**{**kw, "budget": 900, "sampled": got, "first_index": -1, "cost_start": cost})
    assert got2 == want2
    # exhaustion: every frame taken, then the reference's assertion (spatial_temporal_selector.py:182)
    small = pools("P300")["infos"]
    _set_spatial_pool(handle, small)
    handle.build_knn_graph(8)
    for var in ("0", cluster):
        monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", var)
        with pytest.raises(AssertionError, match="has been selected"):
            handle.select(**{**kw, "budget": 10 ** 6, "first_index": 7})
    if cluster == "1":      # every block size of the one-CTA step: same selection, same final fps, same exhaustion
        for rb in ("128", "256", "512"):
            monkeypatch.setenv("ADSEL_FAST_BLOCK", rb)
            with pytest.raises(AssertionError, match="has been selected"):
                handle.select(**{**kw, "budget": 10 ** 6, "first_index": 7})
            _set_spatial_pool(handle, infos)
            handle.build_knn_graph(8)
            got, cost = handle.select(**kw)
            assert got == want and cost == want_cost, rb
            assert np.array_equal(handle.get_fps().view(np.uint64), want_fps.view(np.uint64)), rb
            assert handle.greedy_info()["kernel"] == "greedy_fast_kernel"
            _set_spatial_pool(handle, small)
            handle.build_knn_graph(8)
        monkeypatch.delenv("ADSEL_FAST_BLOCK")
    monkeypatch.delenv("ADSEL_GREEDY_CLUSTER")


@pytest.mark.parametrize("variant", ["0", "2", "generic", "block256", "block512"])
def test_every_golden_case_with_the_global_memory_the_cluster_and_the_generic_resident_kernel(pools, tmp_path, monkeypatch, variant):
    """The default (one resident CTA, fast step where the mode allows; 128-frame blocks at this pool size) is what the
    golden tests above run; the same 18 cases through the global-memory kernel (pools beyond 16 CTAs' shared memory, the
    multi-GPU column shards), a 2-CTA cluster, the generic one-CTA resident kernel, and the fast step with the 256- and
    512-frame blocks larger pools get (ADSEL_FAST_BLOCK)."""
    if variant == "generic":
        monkeypatch.setenv("ADSEL_GREEDY_FAST", "0")
        variant = "1"
    if variant.startswith("block"):
        monkeypatch.setenv("ADSEL_FAST_BLOCK", variant[5:])
        variant = "1"
    monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", variant)
    golden = load_golden("P2K_all.json")
    pl = pools("P2K")
    for cid, case in golden["cases"].items():
        res, _ = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], f"{variant}/{cid}"
    monkeypatch.delenv("ADSEL_GREEDY_CLUSTER")


def test_p850_temporal_selector_budget_1200_identical_to_reference(pools, tmp_path):
    """BASELINE.json config 4: temporal-only selector over the 850-scene pool (28,130 frames, 71 logs), budget 1200."""
    golden = load_golden("P850_temporal.json")
    pl = pools("P850")
    for cid, case in golden["cases"].items():
        res, _ = _run_case_through_selectors(pl, case, tmp_path)
        assert res == case["selected_index"], cid


def test_one_warp_kernel_against_the_global_memory_kernel(handle, pools, monkeypatch):
    """greedy_warp_kernel (the default for geodesic + exp rows on a pool that fits one CTA's shared memory): same
    selection, cost, final fps, resumed round and exhaustion error as ADSEL_GREEDY_CLUSTER=0 -- with the speculative
    temporal-group window (groups >= 38 frames), the bounded loop (shorter groups: the LUT is shorter than its reach),
    no temporal term at all, and rows longer than the 64 entries a slot holds."""
    infos = pools("P8K")["infos"]
    for groups in ("runs", "short", "none"):
        loc, run_id, max_td = _set_spatial_pool(handle, infos)
        if groups == "short":                      # consecutive groups of 20 frames
            handle.set_pool(loc, (np.arange(len(infos)) // 20).astype(np.int32), np.ones(len(infos), np.int32))
        handle.build_knn_graph(8)
        kw = dict(spatial="geodesic", temporal=groups != "none", normalize="exp", aggregate="sum", lambda_t=1.0, budget=700,
                  first_index=4242)
        monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", "0")
        want, want_cost = handle.select(**kw)
        assert handle.greedy_info()["kernel"] == "greedy_select_kernel"
        want_fps = handle.get_fps()
        want2, _ = handle.select(**{**kw, "budget": 1000, "sampled": want, "first_index": -1, "cost_start": want_cost})
        monkeypatch.delenv("ADSEL_GREEDY_CLUSTER")
        got, cost = handle.select(**kw)
        assert handle.greedy_info() == {"kernel": "greedy_warp_kernel", "cluster": 1}
        assert got == want and cost == want_cost, groups
        assert np.array_equal(handle.get_fps().view(np.uint64), want_fps.view(np.uint64)), groups
        got2, _ = handle.select(**{**kw, "budget": 1000, "sampled": got, "first_index": -1, "cost_start": cost})
        assert got2 == want2, groups
        for agg in ("min", "max"):
            if groups == "none":
                continue
            monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", "0")
            w, _ = handle.select(**{**kw, "aggregate": agg})
            monkeypatch.delenv("ADSEL_GREEDY_CLUSTER")
            g, _ = handle.select(**{**kw, "aggregate": agg})
            assert g == w and handle.greedy_info()["kernel"] == "greedy_warp_kernel", (groups, agg)
    # dense cluster of frames: rows far longer than a slot (the tail comes from the list pool)
    rng = np.random.default_rng(3)
    n = 3000
    loc = np.concatenate([rng.normal(scale=6.0, size=(400, 2)), rng.uniform(-400, 400, size=(n - 400, 2))])
    handle.set_pool(loc, (np.arange(n) // 45).astype(np.int32), rng.integers(0, 30, n).astype(np.int32))
    handle.build_knn_graph(8)
    kw = dict(spatial="geodesic", temporal=True, normalize="exp", budget=300, first_index=5)
    monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", "0")
    want, want_cost = handle.select(**kw)
    want_fps = handle.get_fps()
    monkeypatch.delenv("ADSEL_GREEDY_CLUSTER")
    got, cost = handle.select(**kw)
    assert handle.greedy_info()["kernel"] == "greedy_warp_kernel" and got == want and cost == want_cost
    assert np.array_equal(handle.get_fps().view(np.uint64), want_fps.view(np.uint64))
    assert handle.counters()["list_entries"] / n > 20 and np.isfinite(handle.geodesic_rows(np.array([0], np.int32), radius=38.0)).sum() > 64
    # exhaustion: every frame taken, then the reference's assertion (spatial_temporal_selector.py:182)
    small = pools("P300")["infos"]
    _set_spatial_pool(handle, small)
    handle.build_knn_graph(8)
    with pytest.raises(AssertionError, match="has been selected"):
        handle.select(spatial="geodesic", temporal=True, normalize="exp", budget=10 ** 6, first_index=7)
    assert handle.greedy_info()["kernel"] == "greedy_warp_kernel"
    # out_idx capacity
    with pytest.raises(Exception, match="capacity"):
        handle.select(spatial="geodesic", temporal=True, normalize="exp", budget=10 ** 6, first_index=7, out_cap=10)


@pytest.mark.parametrize("bpc", ["4", "2", "1"])
def test_one_warp_kernel_over_a_cluster_of_passive_memory_ctas(handle, pools, monkeypatch, bpc):
    """greedy_warp_cluster_kernel: pools beyond one CTA's shared memory keep the one-warp loop; the min-distance vector is
    split over the shared memory of a 2..16-CTA cluster and reached through distributed shared memory.
    ADSEL_WARP_CLUSTER_BPC forces small slices (4 / 2 / 1 blocks of 512 frames per CTA: clusters of 4 / 8 / 16 CTAs on the
    8,000-frame pool) so that every cluster size, the super-block level and the remote paths run on a pool the test can
    afford; compared with the global-memory kernel: selection, cost, final fps, resumed round, exhaustion."""
    infos = pools("P8K")["infos"]
    for groups in ("runs", "short", "none"):
        loc, run_id, max_td = _set_spatial_pool(handle, infos)
        if groups == "short":
            handle.set_pool(loc, (np.arange(len(infos)) // 20).astype(np.int32), np.ones(len(infos), np.int32))
        handle.build_knn_graph(8)
        kw = dict(spatial="geodesic", temporal=groups != "none", normalize="exp", aggregate="sum", lambda_t=1.0, budget=700,
                  first_index=4242)
        monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", "0")
        want, want_cost = handle.select(**kw)
        want_fps = handle.get_fps()
        want2, _ = handle.select(**{**kw, "budget": 1000, "sampled": want, "first_index": -1, "cost_start": want_cost})
        monkeypatch.delenv("ADSEL_GREEDY_CLUSTER")
        monkeypatch.setenv("ADSEL_WARP_CLUSTER_BPC", bpc)
        got, cost = handle.select(**kw)
        assert handle.greedy_info() == {"kernel": "greedy_warp_cluster_kernel", "cluster": {"4": 4, "2": 8, "1": 16}[bpc]}
        assert got == want and cost == want_cost, groups
        assert np.array_equal(handle.get_fps().view(np.uint64), want_fps.view(np.uint64)), groups
        got2, _ = handle.select(**{**kw, "budget": 1000, "sampled": got, "first_index": -1, "cost_start": cost})
        assert got2 == want2, groups
        monkeypatch.delenv("ADSEL_WARP_CLUSTER_BPC")
    # exhaustion on a pool of 5 blocks (2,099 frames): every frame taken, then the reference's assertion (:182)
    small = pools("P2K")["infos"]
    _set_spatial_pool(handle, small)
    handle.build_knn_graph(8)
    monkeypatch.setenv("ADSEL_WARP_CLUSTER_BPC", "1" if bpc == "1" else "2")
    with pytest.raises(AssertionError, match="has been selected"):
        handle.select(spatial="geodesic", temporal=True, normalize="exp", budget=10 ** 6, first_index=7)
    assert handle.greedy_info()["kernel"] == "greedy_warp_cluster_kernel"
    with pytest.raises(Exception, match="capacity"):
        handle.select(spatial="geodesic", temporal=True, normalize="exp", budget=10 ** 6, first_index=7, out_cap=10)
    monkeypatch.delenv("ADSEL_WARP_CLUSTER_BPC")


@pytest.mark.parametrize("world", [2, 8])
def test_cluster_kernel_at_the_sizes_of_the_weak_scaled_pools(monkeypatch, world):
    """greedy_warp_cluster_kernel at full size -- bench.py's weak-scaled pools of 56,260 (2 CTAs) and 225,040 frames (8 CTAs:
    54 blocks in CTA 0, 56 in each passive CTA) -- against the global-memory kernel: selection, cost, final fps, resumed round."""
    from adsel_b200 import native, synthetic
    base = synthetic.P28K
    spec = synthetic.PoolSpec(f"P28Kx{world}", base.n_scenes * world, base.n_long * world, base.long_len, base.short_len,
                              base.scenes_per_log, base.seed)
    infos, _ = synthetic.generate_pool(spec, box_m=1600.0 * world ** 0.5)
    h = native.Handle(0)
    try:
        _set_spatial_pool(h, infos)
        h.build_knn_graph(8)
        kw = dict(spatial="geodesic", temporal=True, normalize="exp", budget=2400, first_index=len(infos) // 3)
        monkeypatch.setenv("ADSEL_GREEDY_CLUSTER", "0")
        want, want_cost = h.select(**kw)
        want_fps = h.get_fps()
        want2, _ = h.select(**{**kw, "budget": 2700, "sampled": want, "first_index": -1, "cost_start": want_cost})
        monkeypatch.delenv("ADSEL_GREEDY_CLUSTER")
        got, cost = h.select(**kw)
        assert h.greedy_info() == {"kernel": "greedy_warp_cluster_kernel", "cluster": world}
        assert got == want and cost == want_cost
        assert np.array_equal(h.get_fps().view(np.uint64), want_fps.view(np.uint64))
        got2, _ = h.select(**{**kw, "budget": 2700, "sampled": got, "first_index": -1, "cost_start": cost})
        assert got2 == want2 and len(got2) > 100
    finally:
        h.close()
