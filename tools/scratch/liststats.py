import os, sys, random, json
sys.path.insert(0, "/root/repo")
import numpy as np
import adsel_b200
from adsel_b200 import native, pool, synthetic
for name in ("P28K", "P8K"):
    infos, _ = synthetic.generate_pool(getattr(synthetic, name))
    loc = pool.ego_locations(infos)
    run_id, max_td = pool.consecutive_runs(pool.logfile_names(infos))
    h = native.Handle(0)
    h.set_pool(loc, run_id, pool.box_counts(infos))
    h.build_knn_graph(8)
    rng = np.random.default_rng(0)
    src = np.sort(rng.choice(len(infos), 2000, replace=False)).astype(np.int32)
    rows = h.geodesic_rows(src, radius=38.0)
    cnt = np.isfinite(rows).sum(axis=1)
    print(name, "list length over 2000 random frames: mean %.1f median %d p90 %d p99 %d max %d; frac>64 %.3f >96 %.3f >128 %.3f" % (
        cnt.mean(), np.median(cnt), np.percentile(cnt, 90), np.percentile(cnt, 99), cnt.max(), (cnt > 64).mean(), (cnt > 96).mean(), (cnt > 128).mean()))
    if name == "P28K":
        g = json.load(open("/root/repo/tests/golden/P28K_st.json"))["cases"]["st_exp_sum_4800"]["selected_index"]["4800"]
        rows = h.geodesic_rows(np.asarray(g[:2000], np.int32), radius=38.0)
        cnt = np.isfinite(rows).sum(axis=1)
        print(name, "selected frames: mean %.1f median %d p90 %d p99 %d max %d; frac>64 %.3f >96 %.3f >128 %.3f" % (
            cnt.mean(), np.median(cnt), np.percentile(cnt, 90), np.percentile(cnt, 99), cnt.max(), (cnt > 64).mean(), (cnt > 96).mean(), (cnt > 128).mean()))
    grp = np.bincount(run_id)
    print(name, "group sizes: mean %.1f max %d" % (grp.mean(), grp.max()))
    h.close()
