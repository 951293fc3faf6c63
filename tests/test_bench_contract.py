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
    assert line["value"] > 0 and abs(line["value"] - 28130 / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] == 1
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
