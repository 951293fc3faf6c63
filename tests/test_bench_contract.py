"""CPU: the reference arm of bench.py (``--impl reference``) runs without a GPU and prints ONE JSON line with the
contract's keys (metric, value, unit, config.workload, cpu_baseline, e2e with zero copy bytes, impl)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, TQDM_DISABLE="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--port", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "pool_frames_scored_and_selected_per_sec"
    assert line["unit"] == "frames/s" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["config"]["workload"].startswith("P28K/SpatialTemporalSelector") and line["config"]["pool_frames"] == 28130
    # each step is a bounded sample of the workload: the first `sample_frames` frames of the pool through the same selector
    m = line["cpu_baseline"]["sample_frames"]
    assert 1000 <= m <= 28130 and line["cpu_baseline"]["sample_budget"] == max(60, round(4800 * m / 28130))
    assert line["value"] > 0 and abs(line["value"] - m / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    assert line["steps"] == 1 and line["warmup"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_profiled_traffic_table_names_kernels_of_this_library():
    """profiles/traffic.json (ncu dram bytes per launch) feeds roofline.traffic: every entry must name a kernel that
    exists in csrc/ and carry a byte count and the capture it came from."""
    import glob
    sys.path.insert(0, ROOT)
    import bench
    table = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    src = "".join(open(p).read() for p in glob.glob(os.path.join(ROOT, "exploring*", "csrc", "*.cu")))
    lookup = bench.profiled_traffic()
    for kernel, entry in table.items():
        if kernel.startswith("_"):
            continue
        assert kernel in src or kernel.replace("_kernel", "_tma_kernel") in src, kernel
        got = lookup(kernel)
        assert got["traffic"] == entry["bytes"] > 0 and got["traffic_launch"].split(":")[0].endswith(".csv")
        assert os.path.exists(os.path.join(ROOT, "profiles", got["traffic_launch"].split(":")[0]))
    assert lookup("no_such_kernel") == {"traffic": None}


def test_bev_extras_bookkeeping_with_a_stub_engine():
    """bench.bev_extras on a stub of native.Handle (no GPU): the roofline objects carry bound / achieved / peak / frac /
    traffic, and the flop / byte accounting follows the formulas stated in DESIGN.md."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench

    class Handle:
        def __init__(self, device):
            self.calls = []

        def bev_generate(self, off, **kw):
            self.calls.append(("gen", len(off) - 1, kw["first_frame"]))

        def bev_build(self, **kw):
            self.calls.append(("build", kw["frame_base"], kw["total_frames"]))

        def stage_ms(self):
            return {"bev_histogram": 1.0, "pairdist": 2.0, "dense_greedy": 3.0}

        def pairdist_build(self):
            pass

        def select_dense(self, n_boxes, **kw):
            assert kw["budget"] == 1200 and kw["first_index"] == 3
            return list(range(10)), 5.0

        def close(self):
            pass

    class Native:
        pass
    Native.Handle = Handle
    n = 5000
    out = bench.bev_extras(Native, 0, n, np.ones(n, np.int32), 3, {"hbm": 6441.0, "bf16": 1675.9, "bf16_source": "x"},
                           reps=1, batch=2048)
    hist, tiles = out["bev_descriptor_build"], out["pairdist_tiles"]
    sizes = 300_000 + (np.arange(n, dtype=np.int64) * 977) % 4001 - 2000
    assert hist["points"] == int(sizes.sum()) and hist["ms"] == 3.0            # three batches of <= 2,048 frames, 1 ms each
    want_gbs = (20.0 * sizes.sum() + n * (16384 * 6 + 4)) / 3e-3 / 1e9
    r = hist["roofline"]
    assert r["bound"] == "hbm" and abs(r["achieved"] - want_gbs) < 1e-6 * want_gbs and r["frac"] == r["achieved"] / r["peak"]
    assert r["traffic"] > 0 and "768 frames" in r["traffic_launch"]
    bm, bn = -(-n // 128), -(-n // 256)
    done = sum(max(0, bn - m // 2) for m in range(bm))
    assert tiles["tiles_computed"] == done and tiles["tiles_full_matrix"] == bm * bn
    assert abs(tiles["roofline"]["achieved"] - 2.0 * done * 128 * 256 * 16384 / 2e-3 / 1e12) < 1e-9
    assert tiles["roofline"]["bound"] == "tensor" and tiles["roofline"]["traffic"] > 0
    sel = out["bev_selection_budget1200"]
    assert sel["selected_frames"] == 10 and sel["ms_total"] == 3.0 + 2.0 + 3.0


def test_greedy_latency_floor_model():
    sys.path.insert(0, ROOT)
    import bench
    m = bench.greedy_latency_floor(9.02, 4047, {"sm_mhz": 1965.0}, "greedy_select_kernel")
    assert m["cycles_per_step"] == 4 * 295 + 3 * 71
    assert abs(m["floor_us_per_step"] - 1393 / 1965.0) < 1e-12 and abs(m["achieved_us_per_step"] - 9020.0 / 4047) < 1e-9
    assert abs(m["frac"] - m["floor_us_per_step"] / m["achieved_us_per_step"]) < 1e-12 and 0.25 < m["frac"] < 0.4
    r = bench.greedy_latency_floor(5.4, 4047, {"sm_mhz": 1965.0})                      # resident kernels: one L2 round trip
    assert r["cycles_per_step"] == 295 + 3 * 71 + 2 * (31 + 3 * 26) + 31 and r["cycles_per_step"] < m["cycles_per_step"]
    assert bench.greedy_latency_floor(9.02, 4047, None)["floor_us_per_step"] == r["floor_us_per_step"]   # clocks unavailable
    w = bench.greedy_latency_floor(4.1, 4047, {"sm_mhz": 1965.0}, "greedy_warp_kernel")                  # one warp: no barrier
    assert w["cycles_per_step"] == 295 + 2 * 31 + 4 * 20 + 3 * 26 and 0.2 < w["frac"] < 0.3
    assert bench.greedy_latency_floor(0.0, 0, {"sm_mhz": None})["frac"] is None
