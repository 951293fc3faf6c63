#!/usr/bin/env python
"""Small end-to-end pass over every kernel, meant to be run under compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

(racecheck on the shared-memory histogram / block maxima, memcheck on the TMA tiles and the
peer-memory stores).  Sizes are tiny because the sanitizer slows kernels down by 10-100x.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adsel_b200  # noqa: E402,F401
from adsel_b200 import native, pool, sharded, synthetic  # noqa: E402


def main():
    infos, logs = synthetic.generate_pool(synthetic.P300)
    names = pool.logfile_names(infos)
    loc = pool.ego_locations(infos)
    run_id, max_td = pool.consecutive_runs(names)
    nb = pool.box_counts(infos)
    random.seed(3407)
    first = random.choice(range(len(infos)))
    h = native.Handle(0)
    h.set_pool(loc, run_id, nb)
    h.build_knn_graph(8)
    ref, _ = h.select(spatial="geodesic", temporal=True, normalize="exp", budget=60, first_index=first)
    msd = h.geodesic_max_finite()
    h.select(spatial="geodesic", temporal=True, normalize="linear", aggregate="min", max_spatial=msd, max_temporal=max_td,
             budget=40, first_index=first)
    h.select(spatial="geodesic", temporal=False, normalize="raw", forbid_sampled=True, budget=40, first_index=first)
    os.environ["ADSEL_KNN"] = "grid"
    h.build_knn_graph(8)
    again, _ = h.select(spatial="geodesic", temporal=True, normalize="exp", budget=60, first_index=first)
    assert again == ref
    os.environ.pop("ADSEL_KNN")
    os.environ["ADSEL_GREEDY_CLUSTER"] = "2"
    again, _ = h.select(spatial="geodesic", temporal=True, normalize="exp", budget=60, first_index=first)
    assert again == ref
    os.environ.pop("ADSEL_GREEDY_CLUSTER")
    h.set_pool(loc, pool.name_groups(names), nb)
    h.select(spatial="none", temporal=True, normalize="raw", forbid_sampled=True, budget=40, first_index=first)
    h.set_pool(loc, pool.city_groups(names, logs), nb)
    h.select(spatial="euclid", temporal=False, normalize="raw", forbid_sampled=True, budget=40, first_index=first)
    h.geodesic_rows([0, 5, 299])
    # BEV descriptors, tcgen05 tiles, dense greedy
    sizes = 3000 + (np.arange(40) * 131) % 301
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    h.bev_generate(off, n_sweeps=3, seed=3)
    h.bev_build()
    h.bev_build(aggregate_atomics=True)
    h.pairdist_build()
    dsel, _ = h.select_dense(np.arange(40, dtype=np.int32) % 7, budget=5, first_index=3)
    h.close()
    # pool sharded over two in-process ranks: peer stores, mailbox exchange, device barriers
    w = sharded.LocalWorld(2, timeout=120.0)
    w.set_pool(loc, run_id, nb)
    w.build_knn_graph(8)
    got, _ = w.select(spatial="geodesic", temporal=True, normalize="exp", budget=60, first_index=first)
    assert got == ref, "sharded selection differs"

    def load_batch(hh, lo, hi):
        o = np.concatenate([[0], np.cumsum(sizes[lo:hi])]).astype(np.int64)
        hh.bev_generate(o, n_sweeps=3, seed=3, first_frame=lo)

    w.dense_build(40, 16, load_batch)
    dgot, _ = w.dense_select(np.arange(40, dtype=np.int32) % 7, budget=5, first_index=3)
    assert dgot == dsel, "sharded dense selection differs"
    w.close()
    print("sanitize_smoke ok:", len(ref), "frames,", len(dsel), "dense frames")


if __name__ == "__main__":
    main()
