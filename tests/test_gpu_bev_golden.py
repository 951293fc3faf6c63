"""GPU: the CUDA BEV histogram against the count channel of the reference's own ``points_to_bev``
(det3d/ops/point_cloud/bev_ops.py:27-53; golden tests/golden/bev_counts.json made by oracle/make_golden_bev.py),
through the C ABI, bit-exact."""
import numpy as np
import pytest

from adsel_b200 import native
from test_oracle_bev import golden_frames

pytestmark = pytest.mark.gpu


def test_histogram_equals_the_reference_bev_map():
    by_case = {}
    for case, voxel, f, pts, want in golden_frames():
        by_case.setdefault((case, voxel), []).append((pts, want))
    assert len(by_case) == 3
    h = native.Handle(0)
    try:
        checked = 0
        for (case, voxel), frames in by_case.items():
            pts5 = [np.concatenate([p, np.zeros((len(p), 1), dtype=np.float32)], axis=1) for p, _ in frames]  # + time lag
            off = np.concatenate([[0], np.cumsum([len(p) for p in pts5])]).astype(np.int64)
            h.bev_set_points(np.concatenate(pts5), off)
            if native.Handle.bev_cells(voxel=voxel) > 17920:      # larger grids are refused, never approximated
                with pytest.raises(native.NativeError):
                    h.bev_build(voxel=voxel)
                continue
            for agg in (False, True):
                h.bev_build(voxel=voxel, aggregate_atomics=agg)
                counts, _, _ = h.bev_get_descriptors()
                assert np.array_equal(counts, np.stack([w for _, w in frames])), case
                checked += 1
        assert checked == 2
    finally:
        h.close()
