"""``torch.library`` registrations of the selection entry points (last row of the boundary table).

For callers that live in torch -- an active-learning loop that already holds the pool arrays as tensors -- the
C-ABI calls behind ``SpatialTemporalSelector.select_samples`` are exposed as custom ops in the ``adsel``
namespace:

  torch.ops.adsel.select_frames(loc_xy, group_id, n_boxes, sampled, first_index, budget, ...) -> int64 [m]
      adsel_set_pool + adsel_build_knn_graph + adsel_select (det3d/selectors/spatial_temporal_selector.py:78-193;
      with other ``spatial`` / ``temporal`` / ``normalize`` values the sibling selectors' loops)
  torch.ops.adsel.bev_descriptors(points, frame_off, point_range, voxel) -> (counts, levels, norm2)
      adsel_bev_set_points + adsel_bev_build (binning of det3d/ops/point_cloud/bev_ops.py:27-53)

CUDA tensors stay where they are: the ops hand their ``data_ptr()`` to the device-pointer entry points of the C ABI
(``adsel_set_pool_device``, ``adsel_select_device``, ``adsel_bev_set_points_device``; validation, bounding box and the
frames-by-group table are computed by kernels) and the results are copied device-to-device into the returned tensors;
only scalars (how many frames were selected, the 64-byte pool summary) cross to the host.  CPU tensors go through the
host-array ABI, as the reference's selectors would.  The ops are registered with fake (meta) kernels, so they trace
under ``torch.compile`` / ``FakeTensorMode`` as opaque calls with a data-dependent output length.  There is no CPU
kernel: without the CUDA library or a device the call raises ``native.NativeError``.

Importing this module registers the ops (idempotent); the package does not import it on its own, so the ctypes
path has no torch dependency.
"""
from typing import Sequence, Tuple

import numpy as np
import torch

from . import native, sharded

NAMESPACE = "adsel"


def _host(t: torch.Tensor, dtype) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=dtype)


@torch.library.custom_op(f"{NAMESPACE}::select_frames", mutates_args=())
def select_frames(loc_xy: torch.Tensor, group_id: torch.Tensor, n_boxes: torch.Tensor, sampled: torch.Tensor,
                  first_index: int, budget: float, cost_start: float = 0.0, cost_b: float = 0.04,
                  cost_f: float = 0.12, spatial: str = "geodesic", temporal: bool = True, normalize: str = "exp",
                  aggregate: str = "sum", lambda_t: float = 1.0, max_temporal: float = 0.0, k: int = 8,
                  forbid_sampled: bool = False, device: int = 0) -> torch.Tensor:
    """Greedy k-center selection under the annotation-cost budget; returns the newly selected frame indices in
    selection order (the caller prepends ``sampled``, as spatial_temporal_selector.py:193 does).

    loc_xy float64 [n,2] ego locations, group_id int32 [n] (consecutive-run id, logfile-name id or city id,
    depending on the selector), n_boxes int32 [n], sampled int [s] frames already labelled (may be empty),
    first_index: the seed frame when ``sampled`` is empty (``random.choice(range(n))`` in the reference), else -1;
    budget: the cumulative budget ``int(current_budget)``; cost_start: cost of ``sampled``;
    max_temporal: the reference's ``max_temporal_distance`` (only read by normalize="linear")."""
    if spatial not in native.SPATIAL or normalize not in native.NORMALIZE or aggregate not in native.AGGREGATE:
        raise ValueError(f"adsel::select_frames: unknown spatial/normalize/aggregate {spatial!r}/{normalize!r}/{aggregate!r}")
    if loc_xy.ndim != 2 or loc_xy.shape[1] != 2:
        raise ValueError("adsel::select_frames: loc_xy must be [n, 2]")
    on_device = loc_xy.is_cuda
    if on_device:
        device = loc_xy.device.index
        torch.cuda.current_stream(loc_xy.device).synchronize()      # the library reads the tensors on its own stream
        loc_d = loc_xy.detach().to(torch.float64).contiguous()
        grp_d = group_id.detach().to(device=loc_xy.device, dtype=torch.int32).contiguous()
        box_d = n_boxes.detach().to(device=loc_xy.device, dtype=torch.int32).contiguous()
        smp_d = sampled.detach().to(device=loc_xy.device, dtype=torch.int32).contiguous()
        torch.cuda.current_stream(loc_xy.device).synchronize()
    engine = sharded.make_engine(device)
    try:
        if on_device and hasattr(engine, "set_pool_device"):
            engine.set_pool_device(loc_d.shape[0], loc_d.data_ptr(), grp_d.data_ptr(), box_d.data_ptr())
        else:
            on_device = False
            engine.set_pool(_host(loc_xy, np.float64), _host(group_id, np.int32), _host(n_boxes, np.int32))
        max_spatial = 0.0
        if spatial == "geodesic":
            engine.build_knn_graph(k)
            if normalize == "linear":
                max_spatial = engine.geodesic_max_finite()          # spatial_temporal_selector.py:139
        kw = dict(spatial=spatial, temporal=temporal, normalize=normalize, aggregate=aggregate, lambda_t=lambda_t,
                  max_spatial=max_spatial, max_temporal=max_temporal, cost_b=cost_b, cost_f=cost_f, cost_start=cost_start,
                  budget=budget, forbid_sampled=forbid_sampled, first_index=first_index)
        if on_device:
            m, _ = engine.select_device(smp_d.data_ptr(), smp_d.numel(), **kw)
            out32 = torch.empty(m, dtype=torch.int32, device=loc_xy.device)
            engine.copy_selected_device(out32.data_ptr(), m)
            return out32.to(torch.int64)
        selected, _ = engine.select(sampled=_host(sampled, np.int32).tolist(), **kw)
    finally:
        engine.close()
    return torch.tensor(selected, dtype=torch.int64, device=loc_xy.device)


@select_frames.register_fake
def _select_frames_fake(loc_xy, group_id, n_boxes, sampled, first_index, budget, cost_start=0.0, cost_b=0.04,
                        cost_f=0.12, spatial="geodesic", temporal=True, normalize="exp", aggregate="sum",
                        lambda_t=1.0, max_temporal=0.0, k=8, forbid_sampled=False, device=0):
    m = torch.library.get_ctx().new_dynamic_size()                  # how many frames fit the budget is data-dependent
    return loc_xy.new_empty((m,), dtype=torch.int64)


@torch.library.custom_op(f"{NAMESPACE}::bev_descriptors", mutates_args=())
def bev_descriptors(points: torch.Tensor, frame_off: torch.Tensor, point_range: Sequence[float],
                    voxel: Sequence[float], device: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-frame BEV density descriptors of float32 [P,5] points (x,y,z,intensity,dt); frame f owns rows
    ``frame_off[f]:frame_off[f+1]``.  Returns (counts int32 [F,ny*nx], levels int32 [F,ny*nx] = bit_length(count)
    clipped to 31, norm2 int32 [F] = sum(level^2))."""
    if len(point_range) != 6 or len(voxel) != 3:
        raise ValueError("adsel::bev_descriptors: point_range needs 6 values, voxel 3")
    off = _host(frame_off, np.int64)                                # per-frame offsets: host metadata (a few KB)
    dev = points.device
    if points.is_cuda:
        device = dev.index
    engine = native.Handle(device)                                  # descriptor build is per process, never sharded
    try:
        if points.is_cuda:
            pts_d = points.detach().to(torch.float32).contiguous()
            torch.cuda.current_stream(dev).synchronize()
            engine.bev_set_points_device(pts_d.data_ptr(), off)
        else:
            engine.bev_set_points(_host(points, np.float32), off)
        engine.bev_build(tuple(float(v) for v in point_range), tuple(float(v) for v in voxel))
        if points.is_cuda:                                           # results copied device-to-device out of the handle
            c_ptr, l_ptr, q_ptr, frames, cells = engine.bev_descriptors_device()
            counts = _from_device_ptr(c_ptr, (frames, cells), torch.int32, dev).clone()
            levels_bits = _from_device_ptr(l_ptr, (frames, cells), torch.int16, dev)
            norm2 = _from_device_ptr(q_ptr, (frames,), torch.int32, dev).clone()
            levels = (levels_bits.to(torch.int32) << 16).view(torch.float32).to(torch.int32)     # bf16 bit patterns of 0..31
            torch.cuda.current_stream(dev).synchronize()
            return counts, levels, norm2
        counts, levels_bf16, norm2 = engine.bev_get_descriptors()
    finally:
        engine.close()
    levels = (levels_bf16.astype(np.uint32) << 16).view(np.float32).astype(np.int32)     # bf16 bit patterns of 0..31
    return (torch.from_numpy(counts).to(dev), torch.from_numpy(levels).to(dev), torch.from_numpy(norm2).to(dev))


def _from_device_ptr(ptr: int, shape, dtype, device) -> torch.Tensor:
    """A tensor VIEW of device memory owned by the library (``__cuda_array_interface__``); clone before the handle closes."""
    typestr = {torch.int32: "<i4", torch.int16: "<i2", torch.float32: "<f4"}[dtype]

    class _Mem:
        __cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False),
                                    "version": 3, "strides": None}
    return torch.as_tensor(_Mem(), device=device)


@bev_descriptors.register_fake
def _bev_descriptors_fake(points, frame_off, point_range, voxel, device=0):
    frames = frame_off.shape[0] - 1
    cells = native.Handle.bev_cells(tuple(point_range), tuple(voxel))
    return (points.new_empty((frames, cells), dtype=torch.int32), points.new_empty((frames, cells), dtype=torch.int32),
            points.new_empty((frames,), dtype=torch.int32))
