"""Full-size identity check of the passive-memory cluster kernel against the global-memory kernel (development aid)."""
import os, sys, random
sys.path.insert(0, "/root/repo")
import numpy as np
import adsel_b200
from adsel_b200 import native, pool, synthetic
for world in (2, 4, 8):
    base = synthetic.P28K
    spec = synthetic.PoolSpec(f"P28Kx{world}", base.n_scenes * world, base.n_long * world, base.long_len, base.short_len,
                              base.scenes_per_log, base.seed)
    infos, _ = synthetic.generate_pool(spec, box_m=1600.0 * world ** 0.5)
    loc = pool.ego_locations(infos)
    run_id, max_td = pool.consecutive_runs(pool.logfile_names(infos))
    h = native.Handle(0)
    h.set_pool(loc, run_id, pool.box_counts(infos))
    h.build_knn_graph(8)
    kw = dict(spatial="geodesic", temporal=True, normalize="exp", budget=4800, first_index=len(infos) // 3)
    os.environ["ADSEL_GREEDY_CLUSTER"] = "0"
    want, wc = h.select(**kw)
    wf = h.get_fps()
    k0 = h.greedy_info()
    os.environ.pop("ADSEL_GREEDY_CLUSTER")
    got, gc = h.select(**kw)
    gf = h.get_fps()
    k1 = h.greedy_info()
    got2, _ = h.select(**{**kw, "budget": 5400, "sampled": got, "first_index": -1, "cost_start": gc})
    os.environ["ADSEL_GREEDY_CLUSTER"] = "0"
    want2, _ = h.select(**{**kw, "budget": 5400, "sampled": want, "first_index": -1, "cost_start": wc})
    os.environ.pop("ADSEL_GREEDY_CLUSTER")
    print(len(infos), k0["kernel"], "vs", k1, "same selection", got == want, "same cost", gc == wc, "same fps",
          bool(np.array_equal(wf.view(np.uint64), gf.view(np.uint64))), "resumed round", got2 == want2, len(got), len(got2), flush=True)
    h.close()
