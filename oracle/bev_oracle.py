"""oracle/bev_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the BEV density descriptor (north_star kernel 1).  No selector of the reference computes a
LiDAR descriptor, so the DESCRIPTOR (density levels, distances, what gets selected from
them) is PARITY UNPINNED BY THE REFERENCE.  The BINNING is pinned: it follows the reference's numba BEV map,
det3d/ops/point_cloud/bev_ops.py:27-53, and ``bev_counts`` is checked cell for cell against the count channel
``bev_map[-1]`` of the unmodified ``points_to_bev`` run in this container (tests/golden/bev_counts.json, made by
oracle/make_golden_bev.py) and against a literal pure-Python transcription of that loop on small inputs
(tests/test_oracle_bev.py):

  grid = round((hi - lo) / voxel)                     (:27-30, float32)
  c_j  = floor((p_j - lo_j) / voxel_j), j = x,y,z     (:37, float32 subtract and TRUE division)
  reject the point if c_j < 0 or c_j >= grid_j        (:38-40)
  count[c_y, c_x] += 1                                (:53, the bev_map[-1] channel)

Deviations (order-dependent or undefined in the reference): the ``max_voxels`` early ``break``
(:48-49) is dropped; non-finite coordinates are rejected.
"""
import numpy as np

NUSC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)


def grid_size(point_range, voxel):
    r = np.asarray(point_range, dtype=np.float32)
    v = np.asarray(voxel, dtype=np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int32)        # nx, ny, nz


def bev_counts(points, frame_off, point_range=NUSC_RANGE, voxel=(0.8, 0.8, 8.0)):
    """int32 [n_frames, ny*nx] point counts per BEV cell."""
    r = np.asarray(point_range, dtype=np.float32)
    v = np.asarray(voxel, dtype=np.float32)
    nx, ny, nz = grid_size(point_range, voxel)
    pts = np.asarray(points, dtype=np.float32)
    out = np.zeros((len(frame_off) - 1, ny * nx), dtype=np.int32)
    for f in range(len(frame_off) - 1):
        p = pts[frame_off[f]:frame_off[f + 1]]
        with np.errstate(invalid="ignore"):
            c = np.floor((p[:, :3] - r[:3]) / v)                  # float32 arithmetic throughout
            ok = (c[:, 0] >= 0) & (c[:, 0] < nx) & (c[:, 1] >= 0) & (c[:, 1] < ny) & (c[:, 2] >= 0) & (c[:, 2] < nz)
        cell = c[ok, 1].astype(np.int64) * nx + c[ok, 0].astype(np.int64)
        out[f] = np.bincount(cell, minlength=ny * nx).astype(np.int32)
    return out


def bev_counts_literal(points, point_range=NUSC_RANGE, voxel=(0.8, 0.8, 8.0)):
    """Pure-python transcription of the loop at bev_ops.py:34-53 for ONE frame (small inputs only)."""
    r = np.asarray(point_range, dtype=np.float32)
    v = np.asarray(voxel, dtype=np.float32)
    grid = grid_size(point_range, voxel)
    count = np.zeros((grid[1], grid[0]), dtype=np.int32)
    for p in np.asarray(points, dtype=np.float32):
        coor, failed = [0, 0, 0], False
        for j in range(3):
            c = np.floor((p[j] - r[j]) / v[j])
            if not np.isfinite(c) or c < 0 or c >= grid[j]:
                failed = True
                break
            coor[j] = int(c)
        if not failed:
            count[coor[1], coor[0]] += 1
    return count.reshape(-1)


def density_levels(counts):
    """level = bit_length(count) clipped to 31 (0,1,2,2,3,3,3,3,4,...): the descriptor fed to the distance tiles."""
    c = np.asarray(counts, dtype=np.int64)
    lv = np.zeros_like(c)
    nz = c > 0
    lv[nz] = np.floor(np.log2(c[nz])).astype(np.int64) + 1
    return np.minimum(lv, 31).astype(np.int32)


def bf16_bits(x):
    """bf16 bit pattern of small non-negative integers (exact for 0..256)."""
    f = np.asarray(x, dtype=np.float32)
    return (f.view(np.uint32) >> 16).astype(np.uint16)


def sq_distance_matrix(levels_a, levels_b):
    """Exact integer squared Euclidean distances between descriptor rows (int64)."""
    a = np.asarray(levels_a, dtype=np.int64)
    b = np.asarray(levels_b, dtype=np.int64)
    na, nb = (a * a).sum(1), (b * b).sum(1)
    return na[:, None] + nb[None, :] - 2 * (a @ b.T)
