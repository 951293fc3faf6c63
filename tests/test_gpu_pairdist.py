"""GPU: tcgen05 pairwise squared distances (exact int32) and the dense greedy loop vs the oracle."""
import numpy as np
import pytest

import bev_oracle as B
import restatement as R
from adsel_b200 import native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = native.Handle(0)
    yield h
    h.close()


def _levels(rng, n, dim, sparsity=0.6):
    lv = rng.integers(0, 32, size=(n, dim))
    lv[rng.random((n, dim)) < sparsity] = 0
    return lv


@pytest.mark.parametrize("n,dim", [(1, 64), (7, 8), (300, 256), (129, 72), (257, 1000), (1000, 16384), (513, 4096)])
def test_distance_matrix_exact(handle, n, dim):
    rng = np.random.default_rng(n * 1000 + dim)
    lv = _levels(rng, n, dim)
    handle.pairdist_set_descriptors(B.bf16_bits(lv))
    handle.pairdist_build()
    got = handle.pairdist_get_rows(0, n)
    ref = B.sq_distance_matrix(lv, lv)
    assert np.array_equal(got.astype(np.int64), ref)
    assert np.all(np.diag(got) == 0)


def test_worst_case_magnitude_is_exact(handle):
    # all entries 31 over 16,384 cells: dot = 15,745,024 < 2^24, the largest value the accumulator sees
    n, dim = 130, 16384
    lv = np.full((n, dim), 31)
    lv[::2] = 0
    lv[1, :77] = 30
    handle.pairdist_set_descriptors(B.bf16_bits(lv))
    handle.pairdist_build()
    assert np.array_equal(handle.pairdist_get_rows(0, n).astype(np.int64), B.sq_distance_matrix(lv, lv))


def test_many_tiles_persistent_schedule(handle):
    # more tiles than SMs: exercises the smem ring / TMEM double buffering across tiles of one CTA
    n, dim = 5000, 512
    rng = np.random.default_rng(4)
    lv = _levels(rng, n, dim, 0.8)
    handle.pairdist_set_descriptors(B.bf16_bits(lv))
    handle.pairdist_build()
    for row0 in (0, 2431, 4872):
        got = handle.pairdist_get_rows(row0, 128)
        assert np.array_equal(got.astype(np.int64), B.sq_distance_matrix(lv[row0:row0 + 128], lv))


def test_dense_greedy_matches_oracle_loop(handle):
    n, dim = 700, 1024
    rng = np.random.default_rng(8)
    lv = _levels(rng, n, dim, 0.9)
    lv[100:140] = lv[99]                      # duplicated descriptors: exact ties, resolved by lowest index
    n_boxes = rng.integers(0, 60, size=n).astype(np.int32)
    handle.pairdist_set_descriptors(B.bf16_bits(lv))
    handle.pairdist_build()
    d2 = B.sq_distance_matrix(lv, lv).astype(np.float64)
    ref = R.greedy_select(lambda i: d2[i], n, n_boxes, [], 5, "150", forbid_sampled=True)
    got, cost = handle.select_dense(n_boxes, budget=150, first_index=5)
    assert got == ref
    # resume from a buffer (second active-learning round)
    ref2 = R.greedy_select(lambda i: d2[i], n, n_boxes, ref, -1, "300", forbid_sampled=True)
    got2, _ = handle.select_dense(n_boxes, budget=300, sampled=ref, cost_start=R.cost_of_sampled(n_boxes, ref, 0.04, 0.12))
    assert got2 == ref2
    with pytest.raises(AssertionError):       # pool exhausted -> the reference's assertion
        handle.select_dense(n_boxes, budget=10 ** 9, first_index=5)


def test_bev_to_selection_pipeline(handle):
    # points -> histogram -> levels -> tcgen05 distances -> greedy, against the oracle end to end
    sizes = 20_000 + (np.arange(200) * 131) % 3001
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    handle.bev_generate(off, n_sweeps=4, seed=3)
    handle.bev_build()
    counts, levels, norm2 = handle.bev_get_descriptors()
    pts = handle.bev_get_points()
    ref_counts = B.bev_counts(pts, off)
    assert np.array_equal(counts, ref_counts)
    lv = B.density_levels(ref_counts)
    handle.pairdist_build()
    d2 = B.sq_distance_matrix(lv, lv)
    assert np.array_equal(handle.pairdist_get_rows(0, len(sizes)).astype(np.int64), d2)
    n_boxes = (np.arange(len(sizes)) * 7 % 40).astype(np.int32)
    ref = R.greedy_select(lambda i: d2[i].astype(np.float64), len(sizes), n_boxes, [], 11, "60", forbid_sampled=True)
    got, _ = handle.select_dense(n_boxes, budget=60, first_index=11)
    assert got == ref


def test_bev_spatial_selector_drop_in(pools, tmp_path):
    """Registry-level use: files on disk -> multi-sweep load -> descriptors -> distances -> selection."""
    import json
    import pickle
    import random
    from adsel_b200 import synthetic
    from adsel_b200.selectors import build_selector
    from adsel_b200.selectors.bev_spatial_selector import load_multisweep_points
    pl = pools("P300")
    infos = [dict(i) for i in pl["infos"]]
    synthetic.write_lidar_files(infos, str(tmp_path), n_points=1200, nsweeps=3)
    infos_file = str(tmp_path / "infos.pkl")
    pickle.dump(infos, open(infos_file, "wb"), protocol=2)
    buf = str(tmp_path / "buffer.json")
    json.dump({"0": []}, open(buf, "w"))
    res = {}
    for budget in (40, 40):
        random.seed(3407)
        sel = build_selector(dict(type="BevSpatialSelector", budget=budget, buffer_file=buf, infos_origin=infos_file,
                                  data_root=str(tmp_path), nsweeps=3, batch_frames=128))
        sel.select_samples(local_rank=0)
        sel.dump_file()
        res.update(sel.get_selected_samples())
    # oracle pipeline on the same files
    frames = [load_multisweep_points(i, str(tmp_path), 3) for i in infos]
    off = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
    lv = B.density_levels(B.bev_counts(np.concatenate(frames), off))
    d2 = B.sq_distance_matrix(lv, lv).astype(np.float64)
    n_boxes = np.array([i["gt_names"].shape[0] for i in infos])
    random.seed(3407)
    first = random.choice(range(len(infos)))
    r1 = R.greedy_select(lambda i: d2[i], len(infos), n_boxes, [], first, "40", forbid_sampled=True)
    r2 = R.greedy_select(lambda i: d2[i], len(infos), n_boxes, r1, -1, "80", forbid_sampled=True)
    assert res == {"40": r1, "80": r1 + r2}
    assert json.load(open(buf)) == {"0": [], "40": r1, "80": r1 + r2}


def test_dense_greedy_generic_variant_beyond_32k_columns(handle):
    """More than 32,768 columns: the min-distance vector moves from registers to shared memory (<= 51,200 columns) or
    global memory; both must select like the oracle loop (short descriptors keep the matrix small: 40k x 40k int32)."""
    for n in (33_000, 52_000):
        dim = 8
        rng = np.random.default_rng(n)
        lv = rng.integers(0, 32, size=(n, dim))
        n_boxes = rng.integers(0, 60, size=n).astype(np.int32)
        handle.pairdist_set_descriptors(B.bf16_bits(lv))
        handle.pairdist_build()
        lv64 = lv.astype(np.int64)
        norms = (lv64 ** 2).sum(1)

        def row(i):
            return (norms + norms[i] - 2 * (lv64 @ lv64[i])).astype(np.float64)

        ref = R.greedy_select(row, n, n_boxes, [], 7, "40", forbid_sampled=True)
        got, _ = handle.select_dense(n_boxes, budget=40, first_index=7)
        assert got == ref
