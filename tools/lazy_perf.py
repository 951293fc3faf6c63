#!/usr/bin/env python
"""Lazy-row loop timing on one B200: SpatialSelector (raw) and normalize="linear" at full size (development aid)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adsel_b200  # noqa: E402,F401
from adsel_b200 import native, pool, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "P28K"
infos, _ = synthetic.generate_pool(getattr(synthetic, name))
loc = pool.ego_locations(infos)
run_id, max_td = pool.consecutive_runs(pool.logfile_names(infos))
nb = pool.box_counts(infos)
random.seed(3407)
first = random.choice(range(len(infos)))
h = native.Handle(0)
h.set_pool(loc, run_id, nb)
h.build_knn_graph(8)
sel, _ = h.select(spatial="geodesic", temporal=False, normalize="raw", forbid_sampled=True, budget=1200, first_index=first)
st = h.stage_ms()
print(f"{name} SpatialSelector budget 1200: {len(sel)} frames, lazy loop {st['greedy']:.1f} ms = {st['greedy'] * 1e3 / len(sel):.1f} us/step ({h.greedy_info()})", flush=True)
msd = h.geodesic_max_finite()
print(f"geodesic_max_finite {h.stage_ms()['geodesic_max']:.1f} ms", flush=True)
for agg, budget in (("sum", 1200), ("min", 600)):
    sel, _ = h.select(spatial="geodesic", temporal=True, normalize="linear", aggregate=agg, max_spatial=msd, max_temporal=max_td,
                      budget=budget, first_index=first)
    st = h.stage_ms()
    print(f"{name} linear/{agg} budget {budget}: {len(sel)} frames, lazy loop {st['greedy']:.1f} ms = {st['greedy'] * 1e3 / len(sel):.1f} us/step", flush=True)
h.close()
