"""GPU: BEV histogram / density descriptor, bit-exact (integer) against oracle/bev_oracle.py."""
import numpy as np
import pytest

import bev_oracle as B
from adsel_b200 import native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    h = native.Handle(0)
    yield h
    h.close()


RAGGED = [0, 1, 3, 4, 5, 127, 128, 129, 131, 640, 1000, 5003, 30001, 2, 0, 260]


@pytest.mark.parametrize("aggregate", [True, False])
def test_histogram_bit_exact_on_ragged_generated_frames(handle, aggregate):
    off = np.concatenate([[0], np.cumsum(RAGGED)]).astype(np.int64)      # unaligned frame starts, empty frames
    handle.bev_generate(off, n_sweeps=3, seed=5)
    pts = handle.bev_get_points()
    assert pts.shape == (off[-1], 5) and np.isfinite(pts).all()
    handle.bev_build(aggregate_atomics=aggregate)
    counts, levels, norm2 = handle.bev_get_descriptors()
    ref = B.bev_counts(pts, off)
    assert np.array_equal(counts, ref)
    lv = B.density_levels(ref)
    assert np.array_equal(levels, B.bf16_bits(lv))
    assert np.array_equal(norm2, (lv.astype(np.int64) ** 2).sum(1))
    assert 0 < counts.sum() < off[-1]                # some returns leave the range box


def test_histogram_special_values_and_boundaries(handle):
    rng = np.random.default_rng(2)
    pts = rng.uniform(-60, 60, size=(4099, 5)).astype(np.float32)
    pts[:, 2] = rng.uniform(-7, 5, size=len(pts))
    pts[:8, :3] = [[-51.2, -51.2, -5.0], [51.2, 0, 0], [51.19999, 51.19999, 2.99999], [0, 0, 3.0],
                   [np.nan, 0, 0], [0, np.inf, 0], [-0.0, -0.0, -0.0], [-51.200005, 0, 0]]
    pts[100:400, :3] = [10.3, -7.7, 0.1]             # 300 hits in one cell: heavy same-address contention
    off = np.array([0, 7, 2048, 4099], dtype=np.int64)
    handle.bev_set_points(pts, off)
    for agg in (True, False):
        handle.bev_build(aggregate_atomics=agg)
        counts, _, _ = handle.bev_get_descriptors()
        assert np.array_equal(counts, B.bev_counts(pts, off))
    assert np.array_equal(counts[0], B.bev_counts_literal(pts[:7]))


def test_other_grid(handle):
    off = np.array([0, 20000, 40001], dtype=np.int64)
    handle.bev_generate(off, n_sweeps=2, seed=9)
    pts = handle.bev_get_points()
    handle.bev_build(point_range=(-40.0, -40.0, -3.0, 40.0, 40.0, 1.0), voxel=(1.0, 0.5, 2.0))
    counts, _, _ = handle.bev_get_descriptors()
    assert counts.shape[1] == 80 * 160
    assert np.array_equal(counts, B.bev_counts(pts, off, (-40.0, -40.0, -3.0, 40.0, 40.0, 1.0), (1.0, 0.5, 2.0)))
    with pytest.raises(native.NativeError):
        handle.bev_build(voxel=(0.1, 0.1, 8.0))      # 1024x1024 cells do not fit the shared histogram


def test_pool_streamed_in_batches(handle):
    # descriptors of successive point batches land at their frame offset of one resident pool
    sizes = 5000 + (np.arange(10) * 37) % 211
    ref, all_pts = [], []
    for b, (lo, hi) in enumerate([(0, 4), (4, 7), (7, 10)]):
        off = np.concatenate([[0], np.cumsum(sizes[lo:hi])]).astype(np.int64)
        handle.bev_generate(off, n_sweeps=2, seed=100 + b)
        ref.append(B.bev_counts(handle.bev_get_points(), off))
        handle.bev_build(frame_base=lo, total_frames=10)
    counts, levels, norm2 = handle.bev_get_descriptors()
    want = np.concatenate(ref)
    assert np.array_equal(counts, want)
    assert np.array_equal(norm2, (B.density_levels(want).astype(np.int64) ** 2).sum(1))
    with pytest.raises(native.NativeError):
        handle.bev_build(frame_base=9, total_frames=10)      # batch of 3 does not fit at frame 9


def test_full_size_frames_conservation(handle):
    # BASELINE-sized frames (10 sweeps x ~30k returns), compared against numpy on the downloaded points
    sizes = 300_000 + (np.arange(6) * 977) % 4001 - 2000
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    handle.bev_generate(off, n_sweeps=10, seed=1)
    handle.bev_build()
    counts, levels, norm2 = handle.bev_get_descriptors()
    pts = handle.bev_get_points()
    assert np.array_equal(counts, B.bev_counts(pts, off))
    assert counts.sum(1).min() > 150_000
