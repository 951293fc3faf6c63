"""Host-side part of the sharded end-to-end call, without a GPU: N gloo ranks run
``pool.extract_spatial_temporal`` on the weak-scaled bench pool and print where the time goes
(parse own slice / all-gather / assembly, max and min over ranks).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/host_extract_perf.py --dir /tmp/pool8
"""
import argparse
import gc
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

import adsel_b200  # noqa: E402,F401
from adsel_b200 import fileio, pool, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", required=True)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--no-settle", action="store_true", help="skip the allocator settling the selector's constructor performs")
    ap.add_argument("--no-gc", action="store_true", help="pause the cyclic GC around the timed call")
    args = ap.parse_args()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    base = synthetic.P28K
    spec = synthetic.PoolSpec(f"P28Kx{world}", base.n_scenes * world, base.n_long * world, base.long_len,
                              base.short_len, base.scenes_per_log, base.seed)
    os.makedirs(args.dir, exist_ok=True)
    paths = {"infos": os.path.join(args.dir, "infos_train_10sweeps_withvelo.pkl"), "logs": os.path.join(args.dir, "log.json")}
    if rank == 0 and not os.path.exists(paths["infos"]):
        synthetic.write_pool(spec, args.dir, box_m=1600.0 * world ** 0.5)
    dist.barrier()
    rows = []
    for it in range(args.iters):
        infos = fileio.load(paths["infos"])            # a fresh list per call, like a fresh selector
        if not args.no_settle:
            from adsel_b200.selectors import budget_buffer
            budget_buffer.settle_allocator(len(infos))   # what BaseSelector.__init__ does after the load
        dist.barrier()
        if args.no_gc:
            gc.disable()
        t0 = time.perf_counter()
        known = {l["logfile"] for l in fileio.load(paths["logs"])}
        t1 = time.perf_counter()
        pool.extract_spatial_temporal(infos, known)
        t2 = time.perf_counter()
        gc.enable()
        rows.append({"logs_ms": (t1 - t0) * 1e3, "extract_ms": (t2 - t1) * 1e3, **pool.LAST_TIMING})
        del infos
    out = [None] * world
    dist.all_gather_object(out, rows[1:])
    if rank == 0:
        keys = list(out[0][0].keys())
        for k in keys:
            vals = [[r[k] for r in per_rank] for per_rank in out]
            mean = [sum(v) / len(v) for v in vals]
            print(f"{k:22s} max over ranks {max(mean):8.2f} ms   min {min(mean):8.2f} ms")
        print("gc enabled:", gc.isenabled(), "frames:", spec.n_frames)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
