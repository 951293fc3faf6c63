"""CPU: torch.library registrations -- schema, fake (meta) kernels, argument plumbing into the engine, and the
loud failure without a CUDA device (no CPU kernel exists)."""
import numpy as np
import pytest
import torch

from adsel_b200 import native, sharded, torch_ops  # noqa: F401  (importing registers torch.ops.adsel.*)


def test_schemas():
    s = str(torch.ops.adsel.select_frames.default._schema)
    assert s.startswith("adsel::select_frames(Tensor loc_xy, Tensor group_id, Tensor n_boxes, Tensor sampled, SymInt first_index, float budget")
    assert 'str spatial="geodesic"' in s and 'str normalize="exp"' in s and "SymInt k=8" in s and s.endswith("-> Tensor")
    b = str(torch.ops.adsel.bev_descriptors.default._schema)
    assert b == ("adsel::bev_descriptors(Tensor points, Tensor frame_off, float[] point_range, float[] voxel, "
                 "SymInt device=0) -> (Tensor, Tensor, Tensor)")


def test_fake_kernels_give_shapes_without_running():
    from torch._subclasses.fake_tensor import FakeTensorMode
    from torch.fx.experimental.symbolic_shapes import ShapeEnv
    with FakeTensorMode(shape_env=ShapeEnv()):
        loc, grp = torch.empty(100, 2, dtype=torch.float64), torch.empty(100, dtype=torch.int32)
        out = torch.ops.adsel.select_frames(loc, grp, grp, torch.empty(0, dtype=torch.int64), 5, 600.0)
        assert out.dtype == torch.int64 and out.dim() == 1 and not isinstance(out.shape[0], int)   # data-dependent length
        c, lv, n2 = torch.ops.adsel.bev_descriptors(torch.empty(1000, 5), torch.empty(4, dtype=torch.int64),
                                                    [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], [0.8, 0.8, 8.0])
        assert c.shape == (3, 16384) and lv.shape == (3, 16384) and n2.shape == (3,) and c.dtype == torch.int32


class _Recorder:
    def __init__(self):
        self.calls = []

    def set_pool(self, loc, group, boxes):
        self.calls.append(("set_pool", loc.dtype, loc.shape, group.dtype, boxes.dtype))

    def build_knn_graph(self, k):
        self.calls.append(("knn", k))

    def geodesic_max_finite(self):
        self.calls.append("max_finite")
        return 123.5

    def select(self, **kw):
        self.calls.append(("select", kw))
        return [7, 3, 5], 9.0

    def close(self):
        self.calls.append("close")


def test_arguments_reach_the_engine(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(sharded, "make_engine", lambda device: rec)
    loc = torch.arange(20, dtype=torch.float32).reshape(10, 2)        # converted to float64 on the way
    out = torch.ops.adsel.select_frames(loc, torch.zeros(10, dtype=torch.int64), torch.ones(10, dtype=torch.int64),
                                        torch.tensor([4, 2]), -1, 600.0, cost_start=1.5, normalize="linear",
                                        max_temporal=41.0, k=5)
    assert out.tolist() == [7, 3, 5] and out.dtype == torch.int64
    assert rec.calls[0] == ("set_pool", np.dtype("float64"), (10, 2), np.dtype("int32"), np.dtype("int32"))
    assert rec.calls[1:3] == [("knn", 5), "max_finite"] and rec.calls[-1] == "close"
    kw = rec.calls[3][1]
    assert kw["sampled"] == [4, 2] and kw["first_index"] == -1 and kw["budget"] == 600.0 and kw["cost_start"] == 1.5
    assert kw["max_spatial"] == 123.5 and kw["max_temporal"] == 41.0 and kw["normalize"] == "linear" and kw["temporal"] is True
    rec.calls.clear()
    torch.ops.adsel.select_frames(loc, torch.zeros(10, dtype=torch.int32), torch.ones(10, dtype=torch.int32),
                                  torch.empty(0, dtype=torch.int64), 3, 10.0, spatial="none", normalize="raw")
    assert [c if isinstance(c, str) else c[0] for c in rec.calls] == ["set_pool", "select", "close"]    # no graph for "none"
    with pytest.raises(ValueError):
        torch.ops.adsel.select_frames(loc, torch.zeros(10, dtype=torch.int32), torch.ones(10, dtype=torch.int32),
                                      torch.empty(0, dtype=torch.int64), 3, 10.0, aggregate="mean")
    with pytest.raises(ValueError):
        torch.ops.adsel.select_frames(torch.zeros(10, 3), torch.zeros(10, dtype=torch.int32), torch.ones(10, dtype=torch.int32),
                                      torch.empty(0, dtype=torch.int64), 3, 10.0)
    assert rec.calls[-1] == "close"


def test_no_cpu_kernel():
    try:
        if native.device_count() > 0:
            pytest.skip("a CUDA device is present")
    except native.NativeError:
        pass                                           # no driver / no device: exactly the case under test
    with pytest.raises(native.NativeError):
        torch.ops.adsel.select_frames(torch.zeros(10, 2, dtype=torch.float64), torch.zeros(10, dtype=torch.int32),
                                      torch.zeros(10, dtype=torch.int32), torch.empty(0, dtype=torch.int64), 0, 5.0)
    with pytest.raises(native.NativeError):
        torch.ops.adsel.bev_descriptors(torch.zeros(8, 5), torch.tensor([0, 8]), [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0],
                                        [0.8, 0.8, 8.0])
