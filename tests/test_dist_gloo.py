"""CPU, world_size 2 over gloo: the host side of the pool-sharded path (sharded.DistWorld, the selectors'
engine switch, rank-0-only dump).  The CUDA library is replaced by tests/fake_native.py (oracle rows +
all_gather exchange); the device-side exchange itself is covered by tests/test_gpu_sharded.py."""
import json
import os
import random
import sys
import traceback

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from adsel_b200 import sharded
from conftest import ROOT, load_golden


def test_shard_bounds_cover_the_pool_in_aligned_contiguous_slices():
    for n in (1, 255, 256, 257, 2099, 28130, 281300):
        for world in (1, 2, 3, 4, 8):
            bounds = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(bounds, bounds[1:]))
            assert all(lo % 256 == 0 or lo == n for lo, _ in bounds)
            assert all(hi >= lo for lo, hi in bounds)


def test_pick_global_best_is_first_maximum_across_ranks():
    assert sharded.pick_global_best([(1.0, 7), (2.0, 300), (2.0, 12), (0.5, 1)]) == (2.0, 12)
    assert sharded.pick_global_best([(2.0, 300), (2.0, 301)]) == (2.0, 300)
    assert sharded.pick_global_best([(0.0, 2 ** 31 - 1), (0.0, 5)]) == (0.0, 5)      # empty slice loses to a real entry


def _worker(rank, world, init_file, pool_dir, out_dir):
    try:
        for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import adsel_b200  # noqa: F401
        from adsel_b200 import native, pool
        from adsel_b200.selectors import build_selector
        import fake_native
        dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
        native.Handle = fake_native.FakeShardHandle                     # no GPU in this test
        native.check_ego_locations = fake_native.check_ego_locations
        infos_path, logs_path = os.path.join(pool_dir, "infos_train_10sweeps_withvelo.pkl"), os.path.join(pool_dir, "log.json")
        result = {}

        # 1. the drop-in selector API under torch.distributed: ranks cooperate, rank 0 writes
        buffer_file = os.path.join(out_dir, "buffer.json")
        if rank == 0:
            json.dump({"0": []}, open(buffer_file, "w"))
        dist.barrier()
        for budget in (50, 50):
            random.seed(3407)
            sel = build_selector(dict(type="SpatialTemporalSelector", budget=budget, buffer_file=buffer_file,
                                      infos_origin=infos_path, logs_file=logs_path,
                                      distance_store_file=os.path.join(out_dir, "unused.npy")))
            sel.select_samples(local_rank=rank)
            sel.dump_file()
            dist.barrier()
            result.update(sel.get_selected_samples())
        engine = sharded.make_engine(rank)
        result["engine"] = type(engine._engine).__name__
        result["ego_checks"] = list(fake_native.EGO_CHECKS)
        result["handle_calls"] = [c if isinstance(c, str) else c[0] for c in fake_native.FakeShardHandle.instances[0].calls]

        # 2. DistWorld directly, one-term sibling (temporal groups by logfile name), tie-heavy integer rows
        from adsel_b200 import fileio
        infos = fileio.load(infos_path)
        w = sharded.DistWorld(rank)
        w.set_pool(np.zeros((len(infos), 2)), pool.name_groups(pool.logfile_names(infos)), pool.box_counts(infos))
        random.seed(3407)
        first = random.choice(range(len(infos)))
        result["temporal"], _ = w.select(spatial="none", temporal=True, normalize="raw", forbid_sampled=True, cost_b=0.04,
                                         cost_f=0.12, cost_start=0, budget=100, first_index=first)
        result["range"] = list(w.shard_range())

        # 3. a neighbour-list overflow is retried collectively: arenas re-created for the largest need, graph rebuilt
        fake_native.FakeShardHandle.overflow_once = True
        w2 = sharded.DistWorld(rank)
        loc = pool.ego_locations(infos)
        run_id, _ = pool.consecutive_runs(pool.logfile_names(infos))
        w2.set_pool(loc, run_id, pool.box_counts(infos))
        w2.build_knn_graph(8)
        result["retry"], _ = w2.select(spatial="geodesic", temporal=True, normalize="exp", aggregate="sum", lambda_t=1.0,
                                       cost_b=0.04, cost_f=0.12, cost_start=0, budget=50, first_index=first)
        result["retry_calls"] = [c if isinstance(c, str) else c[0] for c in w2.handle.calls]
        result["retry_list_cap"] = w2._list_cap
        fake_native.FakeShardHandle.overflow_once = False

        # 4. round-2 schedule and guards: sharded build + replicated loop by default, comm_check after every selection, the
        #    first frame of an empty buffer is rank 0's draw on every rank, a buffer that differs between ranks is refused
        result["mode"] = w2.mode
        result["replicate_flag"] = bool(getattr(w2.handle, "replicate", False))
        result["comm_checks"] = sum(1 for c in w2.handle.calls if (c if isinstance(c, str) else c[0]) == "comm_check")
        from adsel_b200 import dist as adist
        result["agreed"] = list(adist.agree_from_rank0((1000 + rank, 7)))
        result["max"] = adist.all_ranks_max(rank * 5)
        try:
            adist.check_same_on_all_ranks((rank,), "a value that differs")
            result["mismatch"] = "not detected"
        except RuntimeError as e:
            result["mismatch"] = str(e)[:30]
        random.seed(3407 + 17 * rank)                                    # unsynchronised per-rank RNG (a library user without the CLI)
        buffer2 = os.path.join(out_dir, "buffer2.json")
        if rank == 0:
            json.dump({"0": []}, open(buffer2, "w"))
        dist.barrier()
        sel = build_selector(dict(type="SpatialTemporalSelector", budget=50, buffer_file=buffer2, infos_origin=infos_path,
                                  logs_file=logs_path, distance_store_file=os.path.join(out_dir, "unused.npy")))
        sel.select_samples(local_rank=rank)
        result["unsynced_rng_selection"] = sel.get_selected_samples()["50"]
        os.environ["ADSEL_SHARD"] = "columns"
        result["mode_forced"] = sharded.shard_mode(300)
        os.environ.pop("ADSEL_SHARD")
        result["mode_big"] = sharded.shard_mode(500_000)
        result["lazy"] = [sharded.lazy_rows(28130, dict(spatial="geodesic", normalize="linear")),
                          sharded.lazy_rows(28130, dict(spatial="geodesic", normalize="exp")),
                          sharded.lazy_rows(2000, dict(spatial="geodesic", normalize="raw"))]
        json.dump(result, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:                                                  # surface the traceback in the parent
        open(os.path.join(out_dir, f"rank{rank}.err"), "w").write(traceback.format_exc())
        raise


def test_two_ranks_select_like_the_reference_and_only_rank0_writes(pools, tmp_path):
    pl = pools("P300")
    golden = load_golden("P300_all.json")["cases"]
    out_dir = str(tmp_path)
    init_file = os.path.join(out_dir, "rendezvous")
    mp.spawn(_worker, args=(2, init_file, pl["dir"], out_dir), nprocs=2, join=True)
    errs = [f for f in os.listdir(out_dir) if f.endswith(".err")]
    assert not errs, open(os.path.join(out_dir, errs[0])).read()
    r0, r1 = (json.load(open(os.path.join(out_dir, f"rank{r}.json"))) for r in (0, 1))
    want = golden["st_exp_sum_rounds"]["selected_index"]
    for r in (r0, r1):
        assert r["50"] == want["50"] and r["100"] == want["100"]            # identical on every rank
        assert r["temporal"] == golden["temporal"]["selected_index"]["100"]
        assert r["engine"] == "DistWorld"
        assert r["handle_calls"].count("comm_init") == 1                      # engine (arena, peer mappings) reused across rounds
    assert r0["range"] == [0, 256] and r1["range"] == [256, 300]
    # every rank had ITS half of the pool's ego locations redone by the (stand-in) device check, on its own device, per round
    assert [c[:2] for c in r0["ego_checks"]] == [[0, 150]] * 2 and [c[:2] for c in r1["ego_checks"]] == [[1, 150]] * 2
    for r in (r0, r1):
        assert r["retry"] == want["50"]                                      # same selection after the collective retry
        assert r["retry_calls"].count("comm_init") == 2                       # arenas re-created once ...
        assert r["retry_list_cap"] == int(1010 * 1.1) + 1024                  # ... sized for the largest need over the ranks
    assert json.load(open(os.path.join(out_dir, "buffer.json"))) == {"0": [], "50": want["50"], "100": want["100"]}
    assert os.path.exists(pl["paths"]["infos"].replace(".pkl", "_100.pkl"))
    for r in (r0, r1):
        assert r["mode"] == "replicate" and r["replicate_flag"] and r["comm_checks"] >= 1
        assert r["agreed"] == [1000, 7] and r["max"] == 5 and r["mismatch"].startswith("ranks disagree on")
        assert r["mode_forced"] == "columns" and r["mode_big"] == "columns" and r["lazy"] == [True, False, False]
    # both ranks started from rank 0's random.choice although their RNGs were seeded differently
    assert r0["unsynced_rng_selection"] == r1["unsynced_rng_selection"] and len(r0["unsynced_rng_selection"]) > 0


def _extract_worker(rank, world, init_file, pool_dir, out_dir):
    try:
        for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import adsel_b200  # noqa: F401
        from adsel_b200 import fileio, pool
        dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
        infos = fileio.load(os.path.join(pool_dir, "infos_train_10sweeps_withvelo.pkl"))
        known = {l["logfile"] for l in fileio.load(os.path.join(pool_dir, "log.json"))}
        out = {}
        # 1. whole pool, slices cut wherever n*r//world falls (inside runs and, after the edit below, exactly on one)
        loc, run_id, max_td, boxes = pool.extract_spatial_temporal(infos, known)
        out["a"] = [loc.view(np.uint64).tolist(), run_id.tolist(), int(max_td), boxes.tolist()]
        # 2. a prefix whose length puts slice boundaries exactly on run boundaries: runs of P2K start every ~40 frames
        names = pool.logfile_names(infos)
        starts = [i for i in range(1, len(names)) if names[i] != names[i - 1]]
        cut = starts[2] * world                         # n*r//world == r*starts[2]: rank 1's slice begins a new run
        sub = infos[:cut]
        loc, run_id, max_td, boxes = pool.extract_spatial_temporal(sub, known)
        out["b"] = [loc.view(np.uint64).tolist(), run_id.tolist(), int(max_td), boxes.tolist(), cut]
        # 3. a logfile unknown to logs_file, seen by the LAST rank's slice only: KeyError on every rank (:80)
        bad = [dict(i) for i in infos]
        bad[-3]["cam_front_path"] = bad[-3]["cam_front_path"].replace(names[-3], "n999-unknown-log")
        try:
            pool.extract_spatial_temporal(bad, known)
            out["c"] = "no error"
        except KeyError as e:
            out["c"] = e.args[0]
        # 4. fewer than 4 frames per rank: every rank parses everything itself (no collective)
        loc, run_id, max_td, boxes = pool.extract_spatial_temporal(infos[: 4 * world - 1], known)
        out["d"] = [len(loc), run_id.tolist()]
        json.dump(out, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        open(os.path.join(out_dir, f"rank{rank}.err"), "w").write(traceback.format_exc())
        raise


def test_three_ranks_assemble_the_same_pool_arrays_as_one_process(pools, tmp_path):
    """pool.extract_spatial_temporal under a world of 3 (uneven slices): every rank ends with exactly the arrays a
    single process extracts -- bit-equal ego locations, run ids merged across slice boundaries (inside a run and exactly
    on a run boundary), the reference's max_temporal_distance, box counts -- and a logfile missing from logs_file
    raises KeyError on every rank, whichever rank's slice holds the frame."""
    from adsel_b200 import pool
    pl = pools("P2K")
    out_dir = str(tmp_path)
    mp.spawn(_extract_worker, args=(3, os.path.join(out_dir, "rendezvous"), pl["dir"], out_dir), nprocs=3, join=True)
    errs = [f for f in os.listdir(out_dir) if f.endswith(".err")]
    assert not errs, open(os.path.join(out_dir, errs[0])).read()
    infos = pl["infos"]
    known = {l["logfile"] for l in pl["logs"]}
    res = [json.load(open(os.path.join(out_dir, f"rank{r}.json"))) for r in range(3)]
    loc, run_id, max_td, boxes = pool.extract_spatial_temporal(infos, known)          # this process: world 1
    want_a = [loc.view(np.uint64).tolist(), run_id.tolist(), int(max_td), boxes.tolist()]
    cut = res[0]["b"][4]
    loc, run_id, max_td, boxes = pool.extract_spatial_temporal(infos[:cut], known)
    want_b = [loc.view(np.uint64).tolist(), run_id.tolist(), int(max_td), boxes.tolist(), cut]
    names = pool.logfile_names(infos)
    assert names[cut // 3] != names[cut // 3 - 1]                                      # the boundary case is really hit
    for r in res:
        assert r["a"] == want_a and r["b"] == want_b
        assert r["c"] == "n999-unknown-log"
        assert r["d"] == [11, pool.consecutive_runs(names[:11])[0].tolist()]
