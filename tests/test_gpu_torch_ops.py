"""GPU: the ``torch.library`` registrations (adsel_b200/torch_ops.py) return what the reference selects
(goldens) and what the reference's BEV map counts, from tensors."""
import numpy as np
import pytest

import restatement as R
from adsel_b200 import pool
from conftest import first_index_for, load_golden
from test_oracle_bev import golden_frames

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_ops():
    import torch
    from adsel_b200 import torch_ops as ops          # registers torch.ops.adsel.*
    assert ops.NAMESPACE == "adsel"
    return torch


def test_select_frames_equals_the_reference_goldens(torch_ops, pools):
    torch = torch_ops
    g = load_golden("P2K_all.json")
    infos = pools("P2K")["infos"]
    names = pool.logfile_names(infos)
    loc = torch.from_numpy(pool.ego_locations(infos))
    run_id, max_td = pool.consecutive_runs(names)
    boxes = torch.from_numpy(pool.box_counts(infos))
    none = torch.empty(0, dtype=torch.int64)
    first = first_index_for(len(infos))
    B = g["base_budget"]
    cases = {"st_exp_sum": dict(), "st_linear_max": dict(normalize="linear", aggregate="max", max_temporal=float(max_td)),
             "st_exp_sum_l05": dict(lambda_t=0.5),
             "spatial": dict(temporal=False, normalize="raw", forbid_sampled=True)}
    for cid, kw in cases.items():
        want = g["cases"][cid]["selected_index"][str(B)]
        got = torch.ops.adsel.select_frames(loc, torch.from_numpy(run_id), boxes, none, first, float(B), **kw)
        assert got.dtype == torch.int64 and got.tolist() == want, cid
    # temporal-only: groups by logfile NAME, tensors on the GPU are accepted too
    want = g["cases"]["temporal"]["selected_index"][str(B)]
    grp = torch.from_numpy(pool.name_groups(names)).cuda()
    got = torch.ops.adsel.select_frames(torch.zeros(len(infos), 2, dtype=torch.float64, device="cuda"), grp, boxes.cuda(),
                                        none, first, float(B), spatial="none", normalize="raw", forbid_sampled=True)
    assert got.is_cuda and got.tolist() == want
    # resume: the second round of the chained golden starts from the first round's buffer
    rounds = g["cases"]["st_exp_sum_rounds"]
    keys = list(rounds["selected_index"])
    prev, full = rounds["selected_index"][keys[0]], rounds["selected_index"][keys[1]]
    cost0 = R.cost_of_sampled(boxes.numpy(), prev, 0.04, 0.12)                # base_selector.py:78-86 order
    got = torch.ops.adsel.select_frames(loc, torch.from_numpy(run_id), boxes, torch.tensor(prev), -1, float(keys[1]),
                                        cost_start=float(cost0))
    assert prev + got.tolist() == full
    with pytest.raises(ValueError):
        torch.ops.adsel.select_frames(loc, torch.from_numpy(run_id), boxes, none, first, 10.0, normalize="log")


def test_bev_descriptors_equal_the_reference_bev_map(torch_ops):
    torch = torch_ops
    frames = [(pts, want) for case, voxel, f, pts, want in golden_frames() if case == "grid128"]
    pts5 = [np.concatenate([p, np.zeros((len(p), 1), dtype=np.float32)], axis=1) for p, _ in frames]
    off = np.concatenate([[0], np.cumsum([len(p) for p in pts5])]).astype(np.int64)
    counts, levels, norm2 = torch.ops.adsel.bev_descriptors(torch.from_numpy(np.concatenate(pts5)).cuda(), torch.from_numpy(off),
                                                            [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], [0.8, 0.8, 8.0])
    want = np.stack([w for _, w in frames])
    assert counts.is_cuda and np.array_equal(counts.cpu().numpy(), want)
    lv = np.where(want > 0, np.floor(np.log2(np.maximum(want, 1))).astype(np.int64) + 1, 0)
    assert np.array_equal(levels.cpu().numpy(), np.minimum(lv, 31))
    assert np.array_equal(norm2.cpu().numpy(), (np.minimum(lv, 31) ** 2).sum(1))


def test_cuda_tensor_inputs_never_visit_the_host(torch_ops, pools):
    """Inputs created on the device go through the device-pointer ABI (adsel_set_pool_device / adsel_select_device /
    adsel_bev_set_points_device): native.HOST_INPUT_BYTES -- bytes of caller arrays handed over from host memory -- stays
    put, and the results equal the reference goldens / the reference's BEV map."""
    torch = torch_ops
    from adsel_b200 import native
    g = load_golden("P2K_all.json")
    infos = pools("P2K")["infos"]
    names = pool.logfile_names(infos)
    run_id, max_td = pool.consecutive_runs(names)
    loc = torch.from_numpy(pool.ego_locations(infos)).cuda()
    grp = torch.from_numpy(run_id).cuda()
    boxes = torch.from_numpy(pool.box_counts(infos)).cuda()
    none = torch.empty(0, dtype=torch.int64, device="cuda")
    first = first_index_for(len(infos))
    B = g["base_budget"]
    before = native.HOST_INPUT_BYTES
    got = torch.ops.adsel.select_frames(loc, grp, boxes, none, first, float(B))
    assert got.is_cuda and got.tolist() == g["cases"]["st_exp_sum"]["selected_index"][str(B)]
    rounds = g["cases"]["st_exp_sum_rounds"]
    keys = list(rounds["selected_index"])
    prev, full = rounds["selected_index"][keys[0]], rounds["selected_index"][keys[1]]
    cost0 = R.cost_of_sampled(boxes.cpu().numpy(), prev, 0.04, 0.12)
    got = torch.ops.adsel.select_frames(loc, grp, boxes, torch.tensor(prev, device="cuda"), -1, float(keys[1]), cost_start=float(cost0))
    assert prev + got.tolist() == full
    # name groups (not contiguous in general): the device-side stable sort of (group, frame) pairs
    name_grp = torch.from_numpy(pool.name_groups(names)).cuda()
    got = torch.ops.adsel.select_frames(torch.zeros(len(infos), 2, dtype=torch.float64, device="cuda"), name_grp, boxes, none, first,
                                        float(B), spatial="none", normalize="raw", forbid_sampled=True)
    assert got.tolist() == g["cases"]["temporal"]["selected_index"][str(B)]
    assert native.HOST_INPUT_BYTES == before                              # three selections, not one input byte from the host
    shuffled = torch.from_numpy(np.random.default_rng(0).permutation(7)[pool.name_groups(names) % 7].astype(np.int32)).cuda()
    a = torch.ops.adsel.select_frames(loc, shuffled, boxes, none, first, 120.0, spatial="euclid", temporal=False, normalize="raw",
                                      forbid_sampled=True)
    b = torch.ops.adsel.select_frames(loc.cpu(), shuffled.cpu(), boxes.cpu(), none.cpu(), first, 120.0, spatial="euclid", temporal=False,
                                      normalize="raw", forbid_sampled=True)
    after_host_call = native.HOST_INPUT_BYTES
    assert a.tolist() == b.tolist() and after_host_call > before          # (the CPU-tensor call is the one that counts)
    frames = [(pts, want) for case, voxel, f, pts, want in golden_frames() if case == "grid128"]
    pts5 = [np.concatenate([p, np.zeros((len(p), 1), dtype=np.float32)], axis=1) for p, _ in frames]
    off = np.concatenate([[0], np.cumsum([len(p) for p in pts5])]).astype(np.int64)
    mid = native.HOST_INPUT_BYTES
    counts, levels, norm2 = torch.ops.adsel.bev_descriptors(torch.from_numpy(np.concatenate(pts5)).cuda(), torch.from_numpy(off),
                                                            [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], [0.8, 0.8, 8.0])
    assert native.HOST_INPUT_BYTES == mid and np.array_equal(counts.cpu().numpy(), np.stack([w for _, w in frames]))
    with pytest.raises(native.NativeError):
        torch.ops.adsel.select_frames(loc, grp, boxes, torch.tensor([len(infos) + 5], device="cuda"), -1, 50.0)
