"""GPU, several devices: one PROCESS per GPU (the deployment shape, ``torchrun``-like), arenas exchanged as
CUDA IPC handles over torch.distributed, kNN rows and the per-step exchange crossing NVLink.
Skipped on a single-GPU box (tests/test_gpu_sharded.py covers the same kernels with in-process ranks)."""
import json
import os
import random
import sys
import traceback

import pytest
import torch.multiprocessing as mp

from adsel_b200 import native
from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, init_file, pool_dir, out_dir):
    try:
        for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist
        import adsel_b200  # noqa: F401
        from adsel_b200 import sharded
        from adsel_b200.selectors import build_selector
        dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
        infos_path, logs_path = os.path.join(pool_dir, "infos_train_10sweeps_withvelo.pkl"), os.path.join(pool_dir, "log.json")
        buffer_file = os.path.join(out_dir, "buffer.json")
        if rank == 0:
            json.dump({"0": []}, open(buffer_file, "w"))
        dist.barrier()
        result = {}
        for budget in (300, 300, 600):                             # three rounds: each resumes from the buffer
            random.seed(3407)
            sel = build_selector(dict(type="SpatialTemporalSelector", budget=budget, buffer_file=buffer_file,
                                      infos_origin=infos_path, logs_file=logs_path,
                                      distance_store_file=os.path.join(out_dir, "unused.npy")))
            sel.select_samples(local_rank=rank)
            sel.dump_file()
            dist.barrier()
            result.update(sel.get_selected_samples())
        result["engine"] = type(sharded.make_engine(rank)._engine).__name__
        result["stats"] = sel.last_stats["stage_ms"]
        json.dump(result, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
        dist.barrier()
        sharded.release_engines()
        dist.destroy_process_group()
    except Exception:
        open(os.path.join(out_dir, f"rank{rank}.err"), "w").write(traceback.format_exc())
        raise


def test_one_process_per_gpu_selects_like_the_reference(pools, tmp_path):
    world = min(native.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    pl = pools("P8K")
    golden = load_golden("P8K_all.json")["cases"]["st_exp_sum_rounds"]
    assert golden["rounds"] == [300, 300, 600]
    out_dir = str(tmp_path)
    mp.spawn(_worker, args=(world, os.path.join(out_dir, "rendezvous"), pl["dir"], out_dir), nprocs=world, join=True)
    errs = [f for f in os.listdir(out_dir) if f.endswith(".err")]
    assert not errs, open(os.path.join(out_dir, errs[0])).read()
    for r in range(world):
        res = json.load(open(os.path.join(out_dir, f"rank{r}.json")))
        assert res["engine"] == "DistWorld"
        for key, want in golden["selected_index"].items():
            assert res[key] == want, f"rank {r} budget {key}"
    assert json.load(open(os.path.join(out_dir, "buffer.json"))) == {"0": [], **golden["selected_index"]}
