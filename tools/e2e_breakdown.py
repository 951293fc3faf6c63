#!/usr/bin/env python
"""Host-side breakdown of one end-to-end ``select_samples()`` call (development aid, GPU box only)."""
import os
import random
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adsel_b200  # noqa: E402,F401
from adsel_b200 import fileio, native, pool, synthetic  # noqa: E402


def main():
    with tempfile.TemporaryDirectory() as tmp:
        paths = synthetic.write_pool(synthetic.P28K, tmp)
        infos = fileio.load(paths["infos"])
        logs = fileio.load(paths["logs"])
        for it in range(4):
            t = [time.perf_counter()]
            names = pool.logfile_names(infos); t.append(time.perf_counter())
            known = {l["logfile"] for l in logs}
            assert all(nm in known for nm in names); t.append(time.perf_counter())
            loc = pool.ego_locations(infos); t.append(time.perf_counter())
            run_id, max_td = pool.consecutive_runs(names); t.append(time.perf_counter())
            nb = pool.box_counts(infos); t.append(time.perf_counter())
            h = native.Handle(0); t.append(time.perf_counter())
            h.set_pool(loc, run_id, nb); t.append(time.perf_counter())
            h.build_knn_graph(8); t.append(time.perf_counter())
            random.seed(3407)
            sel, cost = h.select(spatial="geodesic", temporal=True, normalize="exp", budget=4800,
                                 first_index=random.choice(range(len(infos)))); t.append(time.perf_counter())
            st = h.stage_ms()
            h.close(); t.append(time.perf_counter())
            lab = ["names", "known", "ego_loc", "runs", "boxes", "Handle()", "set_pool", "knn_graph", "select", "close"]
            print(it, {l: round((b - a) * 1e3, 2) for l, a, b in zip(lab, t, t[1:])}, "total", round((t[-1] - t[0]) * 1e3, 1),
                  {k: round(v, 2) for k, v in st.items() if v})


if __name__ == "__main__":
    main()
