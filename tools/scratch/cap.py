import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import adsel_b200
from adsel_b200 import native, pool, synthetic
infos, _ = synthetic.generate_pool(synthetic.P300)
loc = pool.ego_locations(infos)
run_id, max_td = pool.consecutive_runs(pool.logfile_names(infos))
h = native.Handle(0)
h.set_pool(loc, run_id, pool.box_counts(infos))
h.build_knn_graph(8)
for env in ("0", None):
    if env is None:
        os.environ.pop("ADSEL_GREEDY_CLUSTER", None)
    else:
        os.environ["ADSEL_GREEDY_CLUSTER"] = env
    for cap in (300, 10):
        try:
            sel, cost = h.select(spatial="geodesic", temporal=True, normalize="exp", budget=1e6 if cap == 10 else 30, first_index=7, out_cap=cap)
            print(env, cap, "ok", len(sel), cost, h.greedy_info())
        except BaseException as e:
            print(env, cap, type(e).__name__, e, h.greedy_info(), h.counters())
