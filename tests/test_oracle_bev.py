"""CPU: the BEV oracle's vectorised path equals a literal transcription of bev_ops.py:34-53."""
import numpy as np

import bev_oracle as B


def _points(rng, n):
    p = rng.uniform(-60, 60, size=(n, 5)).astype(np.float32)
    p[:, 2] = rng.uniform(-7, 5, size=n)
    return p


def test_vectorised_counts_equal_literal_loop():
    rng = np.random.default_rng(0)
    pts = _points(rng, 3000)
    # boundary and special values
    pts[:8, :3] = [[-51.2, -51.2, -5.0], [51.2, 0, 0], [51.19999, 51.19999, 2.99999], [0, 0, 3.0],
                   [np.nan, 0, 0], [0, np.inf, 0], [-0.0, -0.0, -0.0], [-51.200005, 0, 0]]
    got = B.bev_counts(pts, np.array([0, len(pts)]))[0]
    assert np.array_equal(got, B.bev_counts_literal(pts))
    assert got.sum() < len(pts)                      # some points fall outside


def test_grid_and_levels():
    assert B.grid_size(B.NUSC_RANGE, (0.8, 0.8, 8.0)).tolist() == [128, 128, 1]
    assert B.grid_size(B.NUSC_RANGE, (0.1, 0.1, 0.2)).tolist() == [1024, 1024, 40]
    c = np.array([0, 1, 2, 3, 4, 7, 8, 1023, 1024, 2 ** 31 - 1])
    assert B.density_levels(c).tolist() == [0, 1, 2, 2, 3, 3, 4, 10, 11, 31]
    lv = np.arange(32)
    assert np.array_equal((B.bf16_bits(lv).astype(np.uint32) << 16).view(np.float32), lv.astype(np.float32))


def test_sq_distance_matrix_is_exact_integer():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 32, size=(5, 64))
    d = B.sq_distance_matrix(a, a)
    assert np.array_equal(d, ((a[:, None, :] - a[None, :, :]) ** 2).sum(-1))
    assert np.all(np.diag(d) == 0)


def golden_frames():
    """(case id, voxel, frame number, points float32 [m,4], golden count vector) for every frame of
    tests/golden/bev_counts.json -- the reference's own ``points_to_bev`` count channel (oracle/make_golden_bev.py)."""
    from conftest import load_golden
    import make_golden_bev as G
    g = load_golden("bev_counts.json")
    assert tuple(g["range"]) == B.NUSC_RANGE
    for case, rec in g["cases"].items():
        for f, fr in enumerate(rec["frames"]):
            pts = G.case_points(case, f, rec["n_points"])
            assert len(pts) == fr["n_points"]
            want = np.zeros(fr["shape"][0] * fr["shape"][1], dtype=np.int32)
            want[np.asarray(fr["cells"], dtype=np.int64)] = fr["counts"]
            assert want.sum() == fr["inside"]
            yield case, tuple(rec["voxel"]), f, pts, want


def test_counts_equal_the_reference_bev_map():
    """Pins the binning (float32 subtract / divide / floor, half-open bounds, y-major cells, sum over z slices)
    to det3d/ops/point_cloud/bev_ops.py:27-53 as run by numba in this container."""
    seen = 0
    for case, voxel, f, pts, want in golden_frames():
        got = B.bev_counts(pts, np.array([0, len(pts)]), B.NUSC_RANGE, voxel)[0]
        assert np.array_equal(got, want), (case, f)
        seen += 1
    assert seen == 7
