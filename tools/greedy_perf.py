#!/usr/bin/env python
"""Greedy-loop timing on one B200 for the resident (shared-memory) and the global-memory kernel (development aid).

usage: python tools/greedy_perf.py [POOL] [BUDGET]     env: ADSEL_GREEDY_CLUSTER=0/1/2/4/8/16, ADSEL_GREEDY_PROFILE=1
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adsel_b200  # noqa: E402,F401
from adsel_b200 import native, pool, synthetic  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "P28K"
    budget = int(sys.argv[2]) if len(sys.argv) > 2 else 4800
    if "x" in name:          # P28Kx2 ...: bench.py's weak-scaled pools (same frame density)
        base, world = getattr(synthetic, name.split("x")[0]), int(name.split("x")[1])
        spec = synthetic.PoolSpec(name, base.n_scenes * world, base.n_long * world, base.long_len, base.short_len,
                                  base.scenes_per_log, base.seed)
        infos, _ = synthetic.generate_pool(spec, box_m=1600.0 * world ** 0.5)
    else:
        infos, _ = synthetic.generate_pool(getattr(synthetic, name))
    loc = pool.ego_locations(infos)
    run_id, max_td = pool.consecutive_runs(pool.logfile_names(infos))
    nb = pool.box_counts(infos)
    random.seed(3407)
    first = random.choice(range(len(infos)))
    h = native.Handle(0)
    h.set_pool(loc, run_id, nb)
    h.build_knn_graph(8)
    golden = None
    gp = os.path.join(ROOT, "tests", "golden", "P28K_st.json")
    if name == "P28K" and os.path.exists(gp):
        cases = json.load(open(gp))["cases"]
        for c in cases.values():
            if c["rounds"] == [budget]:
                golden = c["selected_index"][str(budget)]
    variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "", "generic", "2"]
    for var in variants:
        os.environ.pop("ADSEL_GREEDY_CLUSTER", None)
        os.environ.pop("ADSEL_GREEDY_FAST", None)
        if var == "generic":          # one resident CTA, generic step
            os.environ["ADSEL_GREEDY_FAST"] = "0"
            os.environ["ADSEL_GREEDY_CLUSTER"] = "1"
        elif var != "":
            os.environ["ADSEL_GREEDY_CLUSTER"] = var
        best = None
        for _ in range(5):
            try:
                sel, cost = h.select(spatial="geodesic", temporal=True, normalize="exp", budget=budget, first_index=first)
            except Exception as e:      # noqa: BLE001
                print("variant", repr(var), "failed:", e)
                sel = None
                break
            st = h.stage_ms()
            g = st["greedy"] if isinstance(st, dict) and "greedy" in st else list(st.values())[3]
            best = g if best is None else min(best, g)
        if sel is None:
            continue
        steps = h.counters()["steps"] if "steps" in h.counters() else len(sel)
        print(f"variant {var or 'auto':>4}: greedy {best:.3f} ms, {len(sel)} frames, {steps} steps, "
              f"{best * 1e3 / max(steps, 1):.3f} us/step, stages {h.stage_ms()}, "
              f"golden {'n/a' if golden is None else sel == golden}", flush=True)
    h.close()


if __name__ == "__main__":
    main()
