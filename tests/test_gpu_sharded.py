"""GPU: the pool-sharded path -- kNN rows written into every rank's table, column-sharded
neighbour lists and one mailbox exchange per greedy step -- must select exactly what the single-GPU path
and the reference select.  ``LocalWorld`` drives all ranks from this process; with one visible GPU the
ranks are CTAs of the same device (the exchange protocol is the same code), with several GPUs the last
test also spreads them over devices (peer access instead of CUDA IPC)."""
import random

import numpy as np
import pytest

import restatement as R
from adsel_b200 import native, pool, sharded
from conftest import KIND, load_golden

pytestmark = pytest.mark.gpu


def _arrays(infos):
    loc = pool.ego_locations(infos)
    run_id, max_td = pool.consecutive_runs(pool.logfile_names(infos))
    return loc, run_id, max_td, pool.box_counts(infos)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharded_knn_table_and_csr_equal_single_gpu(pools, world):
    loc, run_id, _, n_boxes = _arrays(pools("P2K")["infos"])
    w = sharded.LocalWorld(world, timeout=10.0)
    try:
        w.set_pool(loc, run_id, n_boxes)
        w.build_knn_graph(8)
        ref_d, ref_i = R.knn_query(loc, 8)
        ranges = w.shard_ranges()
        assert ranges[0][0] == 0 and ranges[-1][1] == len(loc)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert ranges == [sharded.shard_bounds(len(loc), world, r) for r in range(world)]
        for h in w.handles:                       # every rank holds the complete table, written by all ranks
            dist, idx = h.get_knn()
            assert np.array_equal(dist.view(np.uint64), ref_d.view(np.uint64))
            assert np.array_equal(idx, ref_i)
            row_ptr, col, wgt = h.get_csr()
            ref = R.knn_graph_csr(loc, 8)
            assert np.array_equal(row_ptr, ref.indptr)
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_selection_identical_to_single_gpu_and_oracle(pools, world):
    infos = pools("P2K")["infos"]
    loc, run_id, max_td, n_boxes = _arrays(infos)
    single = native.Handle(0)
    single.set_pool(loc, run_id, n_boxes)
    single.build_knn_graph(8)
    kw = dict(spatial="geodesic", temporal=True, normalize="exp", aggregate="sum", lambda_t=1.0, budget=300, first_index=17)
    want, want_cost = single.select(**kw)
    want_fps = single.get_fps()
    single.close()
    prov = R.RowProvider("spatial_temporal", infos)
    assert want == R.greedy_select(prov.row, len(infos), n_boxes, [], 17, "300")
    w = sharded.LocalWorld(world, timeout=10.0)
    try:
        w.set_pool(loc, run_id, n_boxes)
        w.build_knn_graph(8)
        got, cost = w.select(**kw)
        assert got == want and cost == want_cost
        assert np.array_equal(w.get_fps().view(np.uint64), want_fps.view(np.uint64))
        # a second round resumes from the first (sampled rows applied on every rank's slice)
        got2, _ = w.select(**{**kw, "budget": 500, "sampled": got, "first_index": -1, "cost_start": cost})
        single = native.Handle(0)
        single.set_pool(loc, run_id, n_boxes)
        single.build_knn_graph(8)
        ref2, _ = single.select(**{**kw, "budget": 500, "sampled": want, "first_index": -1, "cost_start": want_cost})
        single.close()
        assert got2 == ref2 and len(got2) > 0
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_goldens_all_variants(pools, world):
    """Every normalize x aggregate variant and the one-term siblings of the reference, sharded."""
    golden = load_golden("P300_all.json")
    pl = pools("P300")
    infos, logs = pl["infos"], pl["logs"]
    names = pool.logfile_names(infos)
    loc = pool.ego_locations(infos)
    n_boxes = pool.box_counts(infos)
    run_id, max_td = pool.consecutive_runs(names)
    checked = 0
    for cid, case in golden["cases"].items():
        if len(case["rounds"]) != 1 or not all(isinstance(v, list) for v in case["selected_index"].values()):
            continue
        kind = KIND[case["selector"]]
        kwargs = dict(case["kwargs"])
        cost_b, cost_f = kwargs.pop("cost_b", 0.04), kwargs.pop("cost_f", 0.12)
        w = sharded.LocalWorld(world, timeout=10.0)
        try:
            if kind == "spatial_temporal":
                w.set_pool(loc, run_id, n_boxes)
                w.build_knn_graph(kwargs.get("k", 8))
                normalize = kwargs.get("normalize", "exp")
                max_sd = w.geodesic_max_finite() if normalize == "linear" else 0.0
                sel_kw = dict(spatial="geodesic", temporal=True, normalize=normalize, aggregate=kwargs.get("aggregate", "sum"),
                              lambda_t=kwargs.get("lambda_t", 1), max_spatial=max_sd, max_temporal=max_td)
            elif kind == "spatial":
                w.set_pool(loc, run_id, n_boxes)
                w.build_knn_graph(kwargs.get("k", 8))
                sel_kw = dict(spatial="geodesic", temporal=False, normalize="raw", forbid_sampled=True)
            elif kind == "temporal":
                w.set_pool(loc, pool.name_groups(names), n_boxes)
                sel_kw = dict(spatial="none", temporal=True, normalize="raw", forbid_sampled=True)
            else:
                w.set_pool(loc, pool.city_groups(names, logs), n_boxes)
                sel_kw = dict(spatial="euclid", temporal=False, normalize="raw", forbid_sampled=True)
            budget = case["rounds"][0]
            random.seed(3407)
            first = random.choice(range(len(infos)))
            got, _ = w.select(cost_b=cost_b, cost_f=cost_f, cost_start=0, budget=budget, first_index=first, **sel_kw)
            assert got == case["selected_index"][str(budget)], f"{cid} world={world}"
            checked += 1
        finally:
            w.close()
    assert checked >= 8


def test_sharded_p28k_budget_4800_identical_to_reference(pools):
    golden = load_golden("P28K_st.json")["cases"]["st_exp_sum_4800"]["selected_index"]["4800"]
    infos = pools("P28K")["infos"]
    loc, run_id, max_td, n_boxes = _arrays(infos)
    random.seed(3407)
    first = random.choice(range(len(infos)))
    w = sharded.LocalWorld(8, timeout=20.0)
    try:
        w.set_pool(loc, run_id, n_boxes)
        w.build_knn_graph(8)
        got, _ = w.select(spatial="geodesic", temporal=True, normalize="exp", aggregate="sum", lambda_t=1.0, budget=4800,
                          first_index=first)
        assert got == golden
    finally:
        w.close()


def test_sharded_list_pool_overflow_is_retried_collectively(pools):
    """A rank whose neighbour-list pool is too small aborts every rank in the same exchange; the pool is
    grown and the second attempt gives the normal result."""
    infos = pools("P2K")["infos"]
    loc, run_id, max_td, n_boxes = _arrays(infos)
    # 60 near-coincident frames: every one of them has all the others within the 38 m radius
    loc = loc.copy()
    loc[:] = loc[0] + np.random.default_rng(0).uniform(0.0, 30.0, size=loc.shape)
    single = native.Handle(0)
    single.set_pool(loc, run_id, n_boxes)
    single.build_knn_graph(8)
    kw = dict(spatial="geodesic", temporal=True, normalize="exp", budget=40, first_index=3)
    want, _ = single.select(**kw)          # ~n*n list entries > the default pool of 64 n: grown inside
    assert single.counters()["list_entries"] > 64 * len(loc)
    single.close()
    w = sharded.LocalWorld(2, timeout=10.0)
    try:
        w.set_pool(loc, run_id, n_boxes)
        w.build_knn_graph(8)
        got, _ = w.select(**kw)
        assert got == want
    finally:
        w.close()


def test_sharded_over_two_devices_when_available(pools):
    if native.device_count() < 2:
        pytest.skip("needs two GPUs")
    infos = pools("P8K")["infos"]
    loc, run_id, max_td, n_boxes = _arrays(infos)
    golden = load_golden("P8K_all.json")["cases"]["st_exp_sum"]
    budget = golden["rounds"][0]
    random.seed(3407)
    first = random.choice(range(len(infos)))
    w = sharded.LocalWorld(2, devices=[0, 1], timeout=20.0)
    try:
        w.set_pool(loc, run_id, n_boxes)
        w.build_knn_graph(8)
        got, _ = w.select(spatial="geodesic", temporal=True, normalize="exp", budget=budget, first_index=first)
        assert got == golden["selected_index"][str(budget)]
    finally:
        w.close()


# ---- dense (BEV descriptor) path sharded -------------------------------------------------------------
import bev_oracle as B  # noqa: E402


@pytest.mark.parametrize("world,n,dim", [(2, 700, 1024), (3, 1000, 256), (8, 300, 64)])
def test_sharded_distance_columns_and_dense_selection(world, n, dim):
    rng = np.random.default_rng(n + dim)
    lv = rng.integers(0, 32, size=(n, dim))
    lv[rng.random((n, dim)) < 0.85] = 0
    lv[100:140] = lv[99]                      # exact ties across (and inside) column shards
    n_boxes = rng.integers(0, 60, size=n).astype(np.int32)
    d2 = B.sq_distance_matrix(lv, lv)
    ref = R.greedy_select(lambda i: d2[i].astype(np.float64), n, n_boxes, [], 5, "150", forbid_sampled=True)
    w = sharded.LocalWorld(world, timeout=10.0)
    try:
        w.dense_set_descriptors(B.bf16_bits(lv))
        for r, h in enumerate(w.handles):     # each rank holds exactly its 256-aligned column slice
            lo, hi = sharded.shard_bounds(n, world, r)
            if hi > lo:
                assert np.array_equal(h.pairdist_get_rows(0, n).astype(np.int64), d2[:, lo:hi])
        got, cost = w.dense_select(n_boxes, budget=150, first_index=5)
        assert got == ref
        ref2 = R.greedy_select(lambda i: d2[i].astype(np.float64), n, n_boxes, ref, -1, "300", forbid_sampled=True)
        got2, _ = w.dense_select(n_boxes, budget=300, sampled=ref, cost_start=R.cost_of_sampled(n_boxes, ref, 0.04, 0.12))
        assert got2 == ref2
        with pytest.raises(AssertionError):   # pool exhausted -> the reference's assertion, on every rank
            w.dense_select(n_boxes, budget=10 ** 9, first_index=5)
    finally:
        w.close()


def test_sharded_dense_path_with_empty_slices():
    """Fewer than 256 frames: every rank but the first owns an empty slice and builds no descriptor itself."""
    n = 40
    sizes = 3000 + (np.arange(n) * 131) % 301
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    single = native.Handle(0)
    single.bev_generate(off, n_sweeps=3, seed=3)
    single.bev_build()
    single.pairdist_build()
    n_boxes = (np.arange(n) % 7).astype(np.int32)
    want, _ = single.select_dense(n_boxes, budget=5, first_index=3)
    single.close()

    def load_batch(h, lo, hi):
        h.bev_generate(np.concatenate([[0], np.cumsum(sizes[lo:hi])]).astype(np.int64), n_sweeps=3, seed=3, first_frame=lo)

    w = sharded.LocalWorld(3, timeout=10.0)
    try:
        w.dense_build(n, 16, load_batch)
        got, _ = w.dense_select(n_boxes, budget=5, first_index=3)
        assert got == want
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_bev_pipeline_descriptors_gathered_by_the_histogram_kernel(world):
    """points -> histogram on the owner rank -> descriptors stored into every rank's arena -> column-sharded
    distances -> sharded greedy, against the oracle end to end."""
    n = 600
    sizes = 8_000 + (np.arange(n) * 131) % 3001
    single = native.Handle(0)
    off_all = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    single.bev_generate(off_all, n_sweeps=4, seed=3)
    pts = single.bev_get_points()
    single.close()
    ref_counts = B.bev_counts(pts, off_all)
    lv = B.density_levels(ref_counts)
    d2 = B.sq_distance_matrix(lv, lv)

    def load_batch(h, lo, hi):
        h.bev_set_points(pts[off_all[lo]:off_all[hi]], off_all[lo:hi + 1] - off_all[lo])

    n_boxes = (np.arange(n) * 7 % 40).astype(np.int32)
    ref = R.greedy_select(lambda i: d2[i].astype(np.float64), n, n_boxes, [], 11, "60", forbid_sampled=True)
    w = sharded.LocalWorld(world, timeout=10.0)
    try:
        w.dense_build(n, 100, load_batch)
        for r, h in enumerate(w.handles):
            _, levels, norm2 = h.bev_get_descriptors(counts=False)
            assert np.array_equal(levels, B.bf16_bits(lv))        # the full matrix, written by all ranks
            assert np.array_equal(norm2.astype(np.int64), (lv.astype(np.int64) ** 2).sum(1))
            lo, hi = sharded.shard_bounds(n, world, r)
            assert np.array_equal(h.pairdist_get_rows(0, n).astype(np.int64), d2[:, lo:hi])
        got, _ = w.dense_select(n_boxes, budget=60, first_index=11)
        assert got == ref
    finally:
        w.close()


def test_p281k_budget_4800_single_gpu_and_sharded(pools):
    """BASELINE.json config 5: the 281,300-frame pool.  The dense reference cannot run it (633 GB per map);
    the fixture comes from the streaming restatement (oracle/make_golden_p281k.py), itself pinned against the
    reference on the smaller pools."""
    golden = load_golden("P281K_st_restatement.json")
    from adsel_b200 import synthetic
    infos, _ = synthetic.generate_pool(synthetic.P281K)
    assert len(infos) == golden["n_frames"]
    loc, run_id, max_td, n_boxes = _arrays(infos)
    kw = dict(spatial="geodesic", temporal=True, normalize="exp", aggregate="sum", lambda_t=1.0, budget=4800,
              first_index=golden["first_index"])
    single = native.Handle(0)
    single.set_pool(loc, run_id, n_boxes)
    single.build_knn_graph(8)
    got, _ = single.select(**kw)
    single.close()
    assert got == golden["selected_index"]["4800"]
    w = sharded.LocalWorld(8, timeout=60.0)
    try:
        w.set_pool(loc, run_id, n_boxes)
        w.build_knn_graph(8)
        got8, _ = w.select(**kw)
        assert got8 == golden["selected_index"]["4800"]
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 3])
def test_replicated_schedule_sharded_build_whole_loop_on_every_rank(pools, world):
    """The default multi-GPU schedule: lists built by the rank owning the SOURCE frames and stored into every rank's
    arena, then every rank runs the whole (resident) greedy loop: same selection, cost and min-distance vector on every
    rank as on one GPU, also when a second round resumes from the first and when the arenas start too small."""
    infos = pools("P2K")["infos"]
    loc, run_id, max_td, n_boxes = _arrays(infos)
    single = native.Handle(0)
    single.set_pool(loc, run_id, n_boxes)
    single.build_knn_graph(8)
    kw = dict(spatial="geodesic", temporal=True, normalize="exp", aggregate="sum", lambda_t=1.0, budget=300, first_index=17)
    want, want_cost = single.select(**kw)
    want_fps = single.get_fps()
    want2, _ = single.select(**{**kw, "budget": 500, "sampled": want, "first_index": -1, "cost_start": want_cost})
    entries = single.counters()["list_entries"]
    single.close()
    for list_cap in (0, entries // 3):          # default size; too small -> collective retry with larger arenas
        w = sharded.LocalWorld(world, timeout=10.0, mode="replicate")
        w._list_cap = list_cap
        try:
            w.set_pool(loc, run_id, n_boxes)
            w.build_knn_graph(8)
            got, cost = w.select(**kw)
            assert got == want and cost == want_cost
            for h in w.handles:                  # every rank holds the whole vector and every list entry
                assert np.array_equal(h.get_fps().view(np.uint64), want_fps.view(np.uint64))
                assert h.counters()["list_entries"] == entries
            got2, _ = w.select(**{**kw, "budget": 500, "sampled": got, "first_index": -1, "cost_start": cost})
            assert got2 == want2
        finally:
            w.close()
