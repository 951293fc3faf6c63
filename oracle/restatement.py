"""oracle/restatement.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Streaming CPU restatement of the reference selectors' arithmetic.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.

The reference materialises four dense N x N float64 maps
(det3d/selectors/spatial_temporal_selector.py:95-155); this restatement produces the same ROWS
lazily -- one ``scipy.sparse.csgraph.dijkstra`` call per selected frame -- so it also runs at
N = 281,300 where the reference cannot allocate.  Third-party arithmetic it stands
on (the reference pins no versions; these are what defined the goldens): scipy 1.18.1
``cKDTree.query`` / ``csgraph.dijkstra``, numpy 2.3.5 ``exp`` (restated in oracle/npexp.c so that
the oracle does not depend on the host's SIMD dispatch).

PINNED: tests/test_oracle.py checks every function here against tests/golden/*.json, which were
produced by running the unmodified reference (oracle/make_goldens.py) in the build container.
"""
import ctypes
import os
import subprocess

import numpy as np
from scipy import sparse, spatial
from scipy.sparse import csgraph

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MARGIN = 1e6  # spatial_temporal_selector.py:112, temporal_selector.py:60, euclidean_spatial_selector.py:94


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/npexp.c -> oracle/_build/liboracle.so (gcc; contraction off, see npexp.c)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle.so")
    src = os.path.join(_HERE, "npexp.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        for name in ("oracle_npexp_array", "oracle_exp_normalize_array", "oracle_npexpf_array",
                     "oracle_exp_normalize_f32_array"):
            getattr(_LIB, name).argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
            getattr(_LIB, name).restype = None
    return _LIB


def npexp(x: np.ndarray) -> np.ndarray:
    """numpy's AVX512_SKX float64 exp, bit for bit (oracle/npexp.c)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _lib().oracle_npexp_array(x.ctypes.data, out.ctypes.data, x.size)
    return out


def npexp_f32(x: np.ndarray) -> np.ndarray:
    """numpy's float32 exp (its own AVX512F / AVX2 kernel), bit for bit (oracle/npexp.c)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    _lib().oracle_npexpf_array(x.ctypes.data, out.ctypes.data, x.size)
    return out


def exp_normalize_f32(d: np.ndarray) -> np.ndarray:
    """``1 - np.exp(-d)`` on a float32 array: float32 exp, float32 result (spatial_feature_selector.py:190)."""
    d = np.ascontiguousarray(d, dtype=np.float32)
    out = np.empty_like(d)
    _lib().oracle_exp_normalize_f32_array(d.ctypes.data, out.ctypes.data, d.size)
    return out


def exp_normalize(d: np.ndarray) -> np.ndarray:
    """``1 - np.exp(-d)`` (spatial_temporal_selector.py:141-143)."""
    d = np.ascontiguousarray(d, dtype=np.float64)
    out = np.empty_like(d)
    _lib().oracle_exp_normalize_array(d.ctypes.data, out.ctypes.data, d.size)
    return out


# --------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------
def logfile_of(info) -> str:
    """spatial_temporal_selector.py:79 -- basename of cam_front_path up to the first "__"."""
    return info["cam_front_path"].split("/")[-1].split("__")[0]


def ego_locations(infos) -> np.ndarray:
    """spatial_temporal_selector.py:83-91 -- ``-(cal[:3,3].T @ cal[:3,:3])[:2]`` per frame, float64 [N,2]."""
    out = []
    for info in infos:
        cal = info["car_from_global"]
        out.append((-(cal[:3, 3].T @ cal[:3, :3]))[:2])
    return np.stack(out)


def consecutive_runs(infos):
    """spatial_temporal_selector.py:114-129.

    Returns ``(run_id int32[N], max_temporal_distance)``; a run is a maximal range of consecutive
    indices with one logfile string; ``max_temporal_distance`` is the largest run LENGTH over all
    runs except the last one (the reference only updates it when a run ends, :127-129).
    """
    run_id = np.zeros(len(infos), dtype=np.int32)
    flag, count, max_td = 0, 0, 0
    logfile_flag = logfile_of(infos[0])
    for i, info in enumerate(infos):
        lf = logfile_of(info)
        if lf == logfile_flag:
            count += 1
        else:
            logfile_flag = lf
            flag += 1
            if count > max_td:
                max_td = count
            count = 1
        run_id[i] = flag
    return run_id, max_td


def logfile_groups(infos) -> np.ndarray:
    """temporal_selector.py:50-55 -- group id per frame by logfile NAME (not by consecutive run)."""
    ids, out = {}, np.zeros(len(infos), dtype=np.int32)
    for i, info in enumerate(infos):
        out[i] = ids.setdefault(logfile_of(info), len(ids))
    return out


def city_ids(infos, logs) -> np.ndarray:
    """euclidean_spatial_selector.py:57-60,75-76 -- ``location.split("-")[-1]`` of the frame's log."""
    log_to_loc = {l["logfile"]: l["location"].split("-")[-1] for l in logs}
    ids, out = {}, np.zeros(len(infos), dtype=np.int32)
    for i, info in enumerate(infos):
        out[i] = ids.setdefault(log_to_loc[logfile_of(info)], len(ids))
    return out


# --------------------------------------------------------------------------------------------
# spatial term
# --------------------------------------------------------------------------------------------
def knn_query(locations: np.ndarray, k: int):
    """spatial_temporal_selector.py:97-98 -- ``cKDTree(locations).query(locations, k+1)``."""
    tree = spatial.cKDTree(locations)
    return tree.query(locations, k + 1)


def knn_graph_csr(locations: np.ndarray, k: int = 8) -> sparse.csr_matrix:
    """spatial_temporal_selector.py:95-104 up to the CSR scipy's Dijkstra actually walks.

    The reference scatters the kNN distances symmetrically into a dense N x N matrix (:99-101) and
    ``shortest_path`` converts it with ``csgraph_from_dense(null_value=0)``, which masks entries
    with ``|w| <= 1e-8`` (np.ma.masked_values atol) as "no edge".  Equivalent sparse construction:
    union of (i,j) and (j,i), duplicates collapse (same value both ways), then drop w <= 1e-8.
    """
    n = len(locations)
    dist, ids = knn_query(locations, k)
    rows = np.repeat(np.arange(n), k + 1)
    cols = ids.reshape(-1)
    w = dist.reshape(-1)
    r = np.concatenate([rows, cols])
    c = np.concatenate([cols, rows])
    v = np.concatenate([w, w])
    key = r.astype(np.int64) * n + c
    _, first = np.unique(key, return_index=True)
    r, c, v = r[first], c[first], v[first]
    keep = ~np.isclose(v, 0.0, rtol=1e-5, atol=1e-8)
    return sparse.csr_matrix((v[keep], (r[keep], c[keep])), shape=(n, n))


def geodesic_rows(csr: sparse.csr_matrix, sources) -> np.ndarray:
    """Rows of ``shortest_path(dense, directed=False, method="D")`` (:104); inf when unreachable."""
    return csgraph.dijkstra(csr, directed=False, indices=np.asarray(sources, dtype=np.int64))


def geodesic_max_finite(csr: sparse.csr_matrix, chunk: int = 512) -> float:
    """``spatial_distance_map[spatial_distance_map != np.inf].max()`` (:139) without the dense map."""
    n, best = csr.shape[0], 0.0
    for lo in range(0, n, chunk):
        rows = geodesic_rows(csr, np.arange(lo, min(n, lo + chunk)))
        fin = rows[np.isfinite(rows)]
        if fin.size:
            best = max(best, float(fin.max()))
    return best


def euclid_row(locations: np.ndarray, city: np.ndarray, i: int) -> np.ndarray:
    """euclidean_spatial_selector.py:95-101."""
    d = np.sqrt(((locations - locations[i]) ** 2).sum(1))
    d[city != city[i]] = MARGIN
    return d


# --------------------------------------------------------------------------------------------
# temporal term
# --------------------------------------------------------------------------------------------
def temporal_row(group: np.ndarray, i: int) -> np.ndarray:
    """Row i of the temporal map (:112,131-133 / temporal_selector.py:60-63): |i-j| inside the group, else 1e6."""
    n = len(group)
    row = np.full(n, MARGIN)
    same = group == group[i]
    row[same] = np.abs(np.arange(n)[same] - i)
    return row


# --------------------------------------------------------------------------------------------
# rows of the aggregated distance map, per selector
# --------------------------------------------------------------------------------------------
class RowProvider:
    """``row(i)`` == ``distance_map[i]`` of the corresponding reference selector."""

    def __init__(self, kind, infos, logs=None, k=8, normalize="exp", lambda_t=1, aggregate="sum"):
        assert kind in ("spatial_temporal", "spatial", "temporal", "euclid")
        self.kind, self.normalize, self.lambda_t, self.aggregate = kind, normalize, lambda_t, aggregate
        self.n = len(infos)
        if kind in ("spatial_temporal", "spatial", "euclid"):
            self.locations = ego_locations(infos)
        if kind in ("spatial_temporal", "spatial"):
            self.csr = knn_graph_csr(self.locations, k)
        if kind == "spatial_temporal":
            assert normalize in ("linear", "exp") and aggregate in ("sum", "min", "max")
            self.group, self.max_td = consecutive_runs(infos)
            if normalize == "linear":
                self.max_sd = geodesic_max_finite(self.csr)
        if kind == "temporal":
            self.group = logfile_groups(infos)
        if kind == "euclid":
            self.city = city_ids(infos, logs)

    def row(self, i: int) -> np.ndarray:
        if self.kind == "spatial":
            return geodesic_rows(self.csr, [i])[0]
        if self.kind == "temporal":
            return temporal_row(self.group, i)
        if self.kind == "euclid":
            return euclid_row(self.locations, self.city, i)
        s = geodesic_rows(self.csr, [i])[0]
        t = temporal_row(self.group, i)
        if self.normalize == "linear":                    # :138-141
            s = s / self.max_sd
            t = t / self.max_td
        else:                                             # :142-144
            s = exp_normalize(s)
            t = exp_normalize(t)
        if self.aggregate == "sum":                       # :148-151
            return s + self.lambda_t * t
        if self.aggregate == "min":                       # :152-153
            return np.stack([s, t]).min(0)
        return np.stack([s, t]).max(0)                    # :154-155


# --------------------------------------------------------------------------------------------
# greedy loop
# --------------------------------------------------------------------------------------------
def cost_of_sampled(n_boxes, sampled, cost_b, cost_f) -> float:
    """base_selector.py:78-86 (python-float accumulation in that order)."""
    cost = 0
    cost += cost_f * len(sampled)
    for i in sampled:
        cost += int(n_boxes[i]) * cost_b
    return cost


def greedy_select(row_fn, n, n_boxes, sampled, first_index, current_budget, cost_b=0.04, cost_f=0.12,
                  forbid_sampled=False):
    """spatial_temporal_selector.py:157-193 (identical in the three sibling selectors).

    ``sampled``: indices already in the buffer; ``first_index``: the ``random.choice`` result used
    when ``sampled`` is empty.  Returns the new ``selected_index_list``.  Raises AssertionError like
    the reference when argmax lands on an already selected frame (:182).
    """
    if len(sampled) > 0:
        fps = np.stack([row_fn(i) for i in sampled]).min(0)
        selected = [int(np.argmax(fps))]
    else:
        selected = [int(first_index)]
        fps = row_fn(selected[-1])
    chosen = set(selected)
    if forbid_sampled:   # the three sibling selectors also assert against the buffer (e.g. spatial_selector.py:124)
        chosen |= set(sampled)
    cost = cost_of_sampled(n_boxes, sampled, cost_b, cost_f)
    cost += cost_f
    cost += int(n_boxes[selected[-1]]) * cost_b
    while True:
        fps = np.stack([fps, row_fn(selected[-1])]).min(0)
        nxt = int(np.argmax(fps))
        assert nxt not in chosen, f"id: {nxt} has been selected"
        cost += cost_f
        cost += int(n_boxes[nxt]) * cost_b
        if cost > int(current_budget):
            break
        selected.append(nxt)
        chosen.add(nxt)
    return selected


def select(kind, infos, logs, buffer, budget, first_index, cost_b=0.04, cost_f=0.12, **kw):
    """One ``select_samples()`` round; returns ``(current_budget_key, sampled + selected)``."""
    max_key = str(max(int(k) for k in buffer))                    # base_selector.py:47-49
    current_budget = str(budget + int(max_key))                   # base_selector.py:40
    sampled = list(buffer[max_key])
    prov = RowProvider(kind, infos, logs, **kw)
    n_boxes = np.array([info["gt_names"].shape[0] for info in infos], dtype=np.int64)
    selected = greedy_select(prov.row, prov.n, n_boxes, sampled, first_index, current_budget, cost_b, cost_f,
                             forbid_sampled=(kind != "spatial_temporal"))
    return current_budget, sampled + selected


# --------------------------------------------------------------------------------------------
# feature-distance selector (det3d/selectors/feature_selector.py:87-172), restated with the reference's own
# torch expressions (float32 on the CPU), rows computed lazily
# --------------------------------------------------------------------------------------------
def torch_sum_f32(x: np.ndarray) -> np.float32:
    """``torch.from_numpy(x).sum()`` of a contiguous float32 vector on the CPU, restated: the vectorised cascade sum of
    aten/src/ATen/native/cpu/SumKernel.cpp (``vectorized_inner_sum`` -> ``row_sum`` -> ``multi_row_sum``) with 8 float
    lanes, 4 interleaved accumulators and four cascade levels of 16 vectors; the scalar tail is added before the 8 lanes.
    Pinned bit for bit against torch 2.11 for lengths 16 .. 8192 (tests/test_oracle_feature.py); csrc/featdist.cu follows
    the same order on the device."""
    import math
    x = np.ascontiguousarray(x, dtype=np.float32)
    W, ILP, LEVELS = 8, 4, 4
    n = len(x)
    vec_size = n // W
    vecs = x[: vec_size * W].reshape(vec_size, W)
    size_ilp = vec_size // ILP
    ceil_log2 = 0 if size_ilp <= 1 else math.ceil(math.log2(size_ilp))
    level_power = max(4, ceil_log2 // LEVELS)
    level_step, level_mask = 1 << level_power, (1 << level_power) - 1
    acc = np.zeros((LEVELS, ILP, W), dtype=np.float32)
    i = 0
    while i + level_step <= size_ilp:
        for _ in range(level_step):
            for k in range(ILP):
                acc[0, k] = acc[0, k] + vecs[i * ILP + k]
            i += 1
        for j in range(1, LEVELS):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = 0
            if (i & (level_mask << (j * level_power))) != 0:
                break
    while i < size_ilp:
        for k in range(ILP):
            acc[0, k] = acc[0, k] + vecs[i * ILP + k]
        i += 1
    for j in range(1, LEVELS):
        acc[0] = acc[0] + acc[j]
    part = acc[0].copy()
    for v in range(size_ilp * ILP, vec_size):
        part[0] = part[0] + vecs[v]
    for k in range(1, ILP):
        part[0] = part[0] + part[k]
    s = np.float32(0)
    for e in x[vec_size * W:]:
        s = np.float32(s + e)
    for k in range(W):
        s = np.float32(s + part[0][k])
    return s


def feature_row(feats, i: int, p: int = 2) -> np.ndarray:
    """``distance_map[i]`` of feature_selector.py:96-101 as float32 numpy (``feats``: float32 [N, C] array)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(feats, dtype=np.float32))
    if p == 1:
        return torch.abs(t - t[i]).sum(1).numpy()
    return torch.sqrt((t - t[i]) ** 2).sum(1).numpy()


def feature_select(feats, n_boxes, buffer, budget, p=2, cost_b=0.04, cost_f=0.12):
    """One ``FeatureSelector.select_samples()`` round: ``(current_budget_key, new + sampled)`` (:143-172)."""
    max_key = str(max(int(k) for k in buffer))
    current_budget = str(budget + int(max_key))
    sampled = list(buffer[max_key])
    rows = {}

    def row(i):
        if i not in rows:
            rows[i] = feature_row(feats, i, p)
        return rows[i]

    fps = np.stack([row(i) for i in sampled]).min(0)
    selected = [int(np.argmax(fps))]
    cost = cost_of_sampled(n_boxes, sampled, cost_b, cost_f)
    cost += cost_f
    cost += int(n_boxes[selected[-1]]) * cost_b
    while True:
        fps = np.minimum(fps, row(selected[-1]))
        nxt = int(np.argmax(fps))
        assert nxt not in selected, f"id: {nxt} has been selected"
        cost += cost_f
        cost += int(n_boxes[nxt]) * cost_b
        if cost > int(current_budget):
            break
        selected.append(nxt)
    return current_budget, selected + sampled


# --------------------------------------------------------------------------------------------
# feature term fused with the spatial (/temporal) terms: det3d/selectors/spatial_feature_selector.py:113-236 and
# spatial_temporal_feature_selector.py:113-261, rows computed lazily
# --------------------------------------------------------------------------------------------
class FusedFeatureRows:
    """``distance_map[i]`` (and, for the spatial+feature selector, ``spatial_distance_map[i]`` after normalisation)
    of the two fused selectors.

    Types follow the reference: the geodesic and temporal maps are float64 and go through ``1 - np.exp(-x)`` in
    float64; the feature map comes out of torch as FLOAT32 (``get_feature_distance_map(...).cpu().numpy()``), so
    ``1 - np.exp(-feature)`` is numpy's float32 ``exp`` and a float32 result, promoted to float64 only when it is
    added to / stacked with the other terms.
    """

    def __init__(self, kind, infos, feats, p=2, k=8, aggregate="sum", lambda_t=1, lambda_f=1):
        assert kind in ("spatial_feature", "spatial_temporal_feature") and p in (1, 2)
        self.kind, self.p, self.aggregate, self.lambda_t, self.lambda_f = kind, p, aggregate, lambda_t, lambda_f
        self.n = len(infos)
        self.feats = np.ascontiguousarray(feats, dtype=np.float32)
        self.csr = knn_graph_csr(ego_locations(infos), k)
        if kind == "spatial_temporal_feature":
            self.group, _ = consecutive_runs(infos)                  # :183-207 (same run bookkeeping as :110-134 of ST)
        self._cache = {}

    def spatial(self, i: int) -> np.ndarray:
        return exp_normalize(geodesic_rows(self.csr, [i])[0])        # spatial_feature_selector.py:189

    def feature(self, i: int) -> np.ndarray:
        return exp_normalize_f32(feature_row(self.feats, i, self.p))  # :190, float32 in, float32 out

    def row(self, i: int) -> np.ndarray:
        if i not in self._cache:
            s, f = self.spatial(i), self.feature(i)
            assert s.dtype == np.float64 and f.dtype == np.float32
            if self.kind == "spatial_feature":
                if self.aggregate == "sum":                          # :192-193
                    r = s + f
                elif self.aggregate == "min":                        # :194-195
                    r = np.stack([s, f]).min(0)
                else:                                                # :196-197
                    r = np.stack([s, f]).max(0)
            else:                                                    # spatial_temporal_feature_selector.py:212-219
                t = exp_normalize(temporal_row(self.group, i))
                r = s + self.lambda_t * t + self.lambda_f * f
            self._cache[i] = r
        return self._cache[i]


def fused_feature_select(kind, infos, feats, buffer, budget, first_index, cost_b=0.04, cost_f=0.12, **kw):
    """One ``select_samples()`` round of SpatialFeatureSelector / SpatialTemporalFeatureSelector:
    ``(current_budget_key, sampled + selected)``.

    Quirk kept from the reference: the spatial+feature selector SEEDS the min-distance vector from the normalised
    SPATIAL rows only (spatial_feature_selector.py:201-209) and uses the aggregated map from the first update on
    (:217); it also asserts against the buffer (:222).  The spatial+temporal+feature selector seeds from the
    aggregated map (spatial_temporal_feature_selector.py:225-234) and only asserts against the new picks (:245).
    """
    max_key = str(max(int(k) for k in buffer))
    current_budget = str(budget + int(max_key))
    sampled = list(buffer[max_key])
    rows = FusedFeatureRows(kind, infos, feats, **kw)
    seed_row = rows.spatial if kind == "spatial_feature" else rows.row
    n_boxes = np.array([info["gt_names"].shape[0] for info in infos], dtype=np.int64)
    if len(sampled) > 0:
        fps = np.stack([seed_row(i) for i in sampled]).min(0)
        selected = [int(np.argmax(fps))]
    else:
        selected = [int(first_index)]
        fps = seed_row(selected[-1])
    cost = cost_of_sampled(n_boxes, sampled, cost_b, cost_f)
    cost += cost_f
    cost += int(n_boxes[selected[-1]]) * cost_b
    while True:
        fps = np.stack([fps, rows.row(selected[-1])]).min(0)
        nxt = int(np.argmax(fps))
        if kind == "spatial_feature":
            assert nxt not in sampled, f"id: {nxt} not in the left index list"
        assert nxt not in selected, f"id: {nxt} has been selected"
        cost += cost_f
        cost += int(n_boxes[nxt]) * cost_b
        if cost > int(current_budget):
            break
        selected.append(nxt)
    return current_budget, sampled + selected
