// geodesic.cu -- single-source shortest paths on the kNN graph, one CTA per source.
//
// Replaces scipy.sparse.csgraph.shortest_path(dense, directed=False, method="D")
// (det3d/selectors/spatial_temporal_selector.py:103-104).  scipy's Dijkstra row for source s is the
// least fixpoint of  d[v] = min_u fl(d[u] + w(u,v)),  d[s] = 0  in IEEE float64 (non-negative
// weights, monotone rounding), and a label-correcting relaxation reaches the same fixpoint
// bit-for-bit whatever the relaxation order -- which is what makes a parallel version exact.
//
// Truncation: with a finite radius only vertices with d < radius are expanded.  Every vertex whose
// true distance is < radius still receives its exact distance (all vertices on its shortest path
// are closer than itself); vertices at >= radius may be missing or over-estimated and are
// reported as "far".  For the "exp" normalisation any d >= 38 gives 1 - exp(-d) == 1.0 exactly
// (exp(-38) < 2^-54), so radius = 38 loses nothing.
#include <stdlib.h>

#include "common.cuh"
#include "sssp.cuh"

namespace adsel {

__global__ void fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// Dense rows for an explicit list of sources (tests; also the APSP maximum when out == nullptr).
__global__ void __launch_bounds__(SSSP_THREADS)
geodesic_rows_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const double *__restrict__ w,
                     int n, const int32_t *__restrict__ sources, int m, double radius,
                     unsigned long long *sc_dist, int32_t *sc_stamp, int32_t *sc_q0, int32_t *sc_q1, int32_t *sc_touched,
                     double *__restrict__ out, unsigned long long *__restrict__ max_finite_bits) {
    __shared__ SsspShared sh;
    __shared__ unsigned long long blk_max;
    if (threadIdx.x == 0) {
        sh.tag = 0;
        blk_max = 0ull;
    }
    const SsspScratch sc = scratch_of(sc_dist, sc_stamp, sc_q0, sc_q1, sc_touched, (size_t)n);
    // tags only grow within a launch; stale stamps from earlier launches are cleared by the host
    __syncthreads();
    for (int si = blockIdx.x; si < m; si += gridDim.x) {
        const int src = sources ? sources[si] : si;
        sssp_cta(row_ptr, col, w, src, radius, sc, sh);
        const int nt = sh.n_touched;
        unsigned long long local_max = 0ull;
        for (int t = threadIdx.x; t < nt; t += SSSP_THREADS) {
            const int v = sc.touched[t];
            const unsigned long long db = __ldcg(&sc.dist[v]);
            const double d = __longlong_as_double((long long)db);
            if (d < radius) {
                if (out) out[(size_t)si * n + v] = d;
                local_max = max(local_max, db);
            }
            sc.dist[v] = INF_BITS;
        }
        if (max_finite_bits) atomicMax(&blk_max, local_max);
        __syncthreads();
    }
    if (max_finite_bits && threadIdx.x == 0) atomicMax(max_finite_bits, blk_max);
}

// Neighbour lists: for every source of this rank's slice [src_lo, src_hi) the vertices with d < radius, each
// with the FINAL row value distance_map[src][v]: the normalised spatial term
// (spatial_temporal_selector.py:138-144) already aggregated with the temporal term of the pair
// (:131-133,148-155), so that a greedy step needs no further lookups.  An entry belongs to the rank that owns
// COLUMN v (the greedy loop is sharded by column): the searching CTA reserves room in that rank's pool with
// one system-scope atomicAdd on its cursor and stores the entries and the list's offset / length there --
// directly into peer memory when the owner is another GPU (the row -> column transposition of the sharded
// build is fused into the search kernel).  Entries beyond an owner's capacity are counted but not written
// (its greedy kernel sees cursor > capacity and reports it; the pool is grown and the build repeated).
struct ListArgs {
    int n, src_lo, src_hi;
    double radius;
    int norm_mode, temporal, aggregate;
    double divisor, lambda_t, t_far;
    const int32_t *group_id;
    const double *tlut;
    int world, per;                 // owner of column v = min(v / per, world - 1)
    int replicate;                  // 1: every rank receives every entry (sharded build, greedy loop replicated on each rank)
    ListStore dest[8];
};

__global__ void __launch_bounds__(SSSP_THREADS)
geodesic_lists_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const double *__restrict__ w,
                      const ListArgs la,
                      unsigned long long *sc_dist, int32_t *sc_stamp, int32_t *sc_q0, int32_t *sc_q1, int32_t *sc_touched) {
    __shared__ SsspShared sh;
    __shared__ int n_keep[8], n_written[8], fits[8];
    __shared__ unsigned long long base_off[8];
    if (threadIdx.x == 0) sh.tag = 0;
    const int n = la.n;
    const double radius = la.radius;
    const SsspScratch sc = scratch_of(sc_dist, sc_stamp, sc_q0, sc_q1, sc_touched, (size_t)n);
    __syncthreads();
    for (int src = la.src_lo + blockIdx.x; src < la.src_hi; src += gridDim.x) {
        sssp_cta(row_ptr, col, w, src, radius, sc, sh);
        const int nt = sh.n_touched;
        if (threadIdx.x < 8) {
            n_keep[threadIdx.x] = 0;
            n_written[threadIdx.x] = 0;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nt; t += SSSP_THREADS) {
            const int v = sc.touched[t];
            const double d = __longlong_as_double((long long)__ldcg(&sc.dist[v]));
            if (d < radius) atomicAdd(&n_keep[la.replicate ? 0 : min(v / la.per, la.world - 1)], 1);
        }
        __syncthreads();
        if (threadIdx.x < la.world) {
            const int o = threadIdx.x;
            const int cnt = n_keep[la.replicate ? 0 : o];
            const unsigned long long b = cnt ? atomicAdd_system(la.dest[o].cursor, (unsigned long long)cnt) : 0ull;
            base_off[o] = b;
            fits[o] = (long long)(b + cnt) <= la.dest[o].capacity;
        }
        __syncthreads();
        const int gsrc = la.temporal ? la.group_id[src] : 0;
        for (int t = threadIdx.x; t < nt; t += SSSP_THREADS) {
            const int v = sc.touched[t];
            const double d = __longlong_as_double((long long)__ldcg(&sc.dist[v]));
            if (d < radius) {
                const int o = la.replicate ? 0 : min(v / la.per, la.world - 1);
                if (la.replicate || fits[o]) {
                    const int pos = atomicAdd(&n_written[o], 1);
                    double val = normalize_value(d, la.norm_mode, la.divisor);
                    if (la.temporal) {
                        const double tv = la.group_id[v] == gsrc ? la.tlut[abs(v - src)] : la.t_far;
                        val = aggregate_value(val, tv, la.aggregate, la.lambda_t);
                    }
                    if (la.replicate) {   // the same position in every rank's pool that has room (peer stores over NVLink)
                        for (int r = 0; r < la.world; ++r)
                            if (fits[r]) {
                                la.dest[r].idx[base_off[r] + pos] = v;
                                la.dest[r].val[base_off[r] + pos] = val;
                            }
                    } else {
                        la.dest[o].idx[base_off[o] + pos] = v;
                        la.dest[o].val[base_off[o] + pos] = val;
                    }
                }
            }
            sc.dist[v] = INF_BITS;
        }
        if (threadIdx.x < la.world) {
            const int o = threadIdx.x;
            la.dest[o].off[src] = (int64_t)base_off[o];
            la.dest[o].len[src] = fits[o] ? n_keep[la.replicate ? 0 : o] : 0;
        }
        __syncthreads();
    }
}

int ensure_scratch(adsel_handle *h) {
    // Searches in flight per SM.  A search is a chain of dependent L2 round trips, so the kernel is latency-bound and
    // scales with concurrency (measured at P28K: 4 / 8 / 12 / 16 CTAs per SM -> 0.91 / 0.53 / 0.40 / 0.37 ms); every
    // resident CTA needs 24 N bytes of scratch, so the default of 16 is lowered for very large pools (<= 4 GB of scratch per handle).
    const char *e = getenv("ADSEL_SSSP_CTAS_PER_SM");
    int per_sm = 16;
    const long long fit = 4000000000ll / ((long long)h->num_sms * 24ll * (h->n > 0 ? h->n : 1));
    if (fit < per_sm) per_sm = fit < 4 ? 4 : (int)fit;
    if (e && atoi(e) > 0) per_sm = atoi(e);
    const int ctas = h->num_sms * per_sm;
    const size_t need = (size_t)ctas * h->n;
    if (h->sssp_ctas == ctas && h->sc_n == (size_t)h->n && h->sc_dist.p) return ADSEL_OK;
    ADSEL_TRY(h->sc_dist.reserve(need));
    ADSEL_TRY(h->sc_stamp.reserve(need));
    ADSEL_TRY(h->sc_q0.reserve(need));
    ADSEL_TRY(h->sc_q1.reserve(need));
    ADSEL_TRY(h->sc_touched.reserve(need));
    h->sssp_ctas = ctas;
    h->sc_n = h->n;
    fill_u64_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(h->sc_dist.p, need, INF_BITS);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    return ADSEL_OK;
}

// Stamps are only meaningful inside one launch (tags restart at 1), so they are cleared before each.
int clear_stamps(adsel_handle *h) {
    ADSEL_CUDA(cudaMemsetAsync(h->sc_stamp.p, 0, sizeof(int32_t) * (size_t)h->sssp_ctas * h->n, h->stream));
    return ADSEL_OK;
}

int geodesic_rows(adsel_handle *h, const int32_t *sources, int32_t m, double radius, double *out_host) {
    ADSEL_REQUIRE(h->graph_ready, "adsel_geodesic_rows: call adsel_build_knn_graph first");
    ADSEL_REQUIRE(m > 0 && sources && out_host, "adsel_geodesic_rows: empty request");
    for (int i = 0; i < m; ++i) ADSEL_REQUIRE(sources[i] >= 0 && sources[i] < h->n, "adsel_geodesic_rows: source out of range");
    ADSEL_TRY(ensure_scratch(h));
    ADSEL_TRY(clear_stamps(h));
    int32_t *d_src = nullptr;
    double *d_out = nullptr;
    const size_t cells = (size_t)m * h->n;
    ADSEL_CUDA(cudaMalloc(&d_src, sizeof(int32_t) * m));
    ADSEL_CUDA(cudaMalloc(&d_out, sizeof(double) * cells));
    ADSEL_CUDA(cudaMemcpyAsync(d_src, sources, sizeof(int32_t) * m, cudaMemcpyHostToDevice, h->stream));
    fill_u64_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>((unsigned long long *)d_out, cells, INF_BITS);
    geodesic_rows_kernel<<<min(h->sssp_ctas, m), SSSP_THREADS, 0, h->stream>>>(
        h->row_ptr.p, h->col.p, h->w.p, h->n, d_src, m, radius, h->sc_dist.p, h->sc_stamp.p, h->sc_q0.p, h->sc_q1.p,
        h->sc_touched.p, d_out, nullptr);
    h->launches += 2;
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_CUDA(cudaMemcpyAsync(out_host, d_out, sizeof(double) * cells, cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_src);
    cudaFree(d_out);
    return ADSEL_OK;
}

int geodesic_max_finite(adsel_handle *h, double *out) {
    ADSEL_REQUIRE(h->graph_ready, "adsel_geodesic_max_finite: call adsel_build_knn_graph first");
    ADSEL_TRY(ensure_scratch(h));
    ADSEL_TRY(clear_stamps(h));
    ADSEL_TRY(h->cursor.reserve(2));
    ADSEL_TRY(begin_stage(h));
    ADSEL_CUDA(cudaMemsetAsync(h->cursor.p, 0, 2 * sizeof(unsigned long long), h->stream));
    geodesic_rows_kernel<<<min(h->sssp_ctas, h->n), SSSP_THREADS, 0, h->stream>>>(
        h->row_ptr.p, h->col.p, h->w.p, h->n, nullptr, h->n, host_inf(), h->sc_dist.p, h->sc_stamp.p,
        h->sc_q0.p, h->sc_q1.p, h->sc_touched.p, nullptr, h->cursor.p + 1);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_TRY(end_stage(h, 4));
    unsigned long long bits = 0;
    ADSEL_CUDA(cudaMemcpy(&bits, h->cursor.p + 1, sizeof(bits), cudaMemcpyDeviceToHost));
    memcpy(out, &bits, sizeof(double));
    return ADSEL_OK;
}

ListStore arena_lists(adsel_handle *h, int w);
int comm_barrier(adsel_handle *h);

// Allocations of the list build (scratch, offsets, entry pool); never shrinks.  Selects the store in use.
int lists_prepare(adsel_handle *h, bool sharded) {
    const int n = h->n;
    ADSEL_TRY(ensure_scratch(h));
    if (sharded) {
        h->lists = arena_lists(h, h->rank);
        return ADSEL_OK;
    }
    ADSEL_TRY(h->cursor.reserve(2));
    ADSEL_TRY(h->list_off.reserve(n));
    ADSEL_TRY(h->list_len.reserve(n));
    if (h->list_capacity < (long long)n * 64) {
        const long long capacity = (long long)n * 64;
        ADSEL_TRY(h->list_idx.reserve(capacity));
        ADSEL_TRY(h->list_val.reserve(capacity));
        h->list_capacity = capacity;
    }
    h->lists = ListStore{h->list_off.p, h->list_len.p, h->list_idx.p, h->list_val.p, h->cursor.p, h->list_capacity};
    return ADSEL_OK;
}

int grow_lists(adsel_handle *h, long long needed) {
    ADSEL_REQUIRE(needed > h->list_capacity, "adsel_select: neighbour-list pool overflow after resize");
    const long long capacity = needed + needed / 16;
    ADSEL_TRY(h->list_idx.reserve(capacity));
    ADSEL_TRY(h->list_val.reserve(capacity));
    h->list_capacity = capacity;
    h->lists = ListStore{h->list_off.p, h->list_len.p, h->list_idx.p, h->list_val.p, h->cursor.p, h->list_capacity};
    h->lists_mode = -1;
    return ADSEL_OK;
}

// Enqueues the neighbour-list build for these selection parameters unless the resident lists were
// built for the same ones; radius is 38 for exp, inf otherwise.  h->tlut must already be enqueued.
// sharded: this rank searches from the sources of its slice and appends every entry to the pool of the rank
// that owns the entry's column -- or, with `replicate`, to EVERY rank's pool (each rank then runs the whole greedy
// loop on complete lists: no exchange per step); device barriers separate cursor reset / build / use across ranks.
int build_lists(adsel_handle *h, const adsel_select_params *p, double t_far, bool sharded, bool replicate) {
    ADSEL_REQUIRE(h->graph_ready, "adsel_select: call adsel_build_knn_graph first");
    const int n = h->n;
    const int lo = sharded ? h->shard_lo : 0, hi = sharded ? h->shard_hi : n;
    ListsKey key;
    key.norm_mode = p->normalize;
    key.temporal = p->temporal ? 1 : 0;
    key.aggregate = key.temporal ? p->aggregate : 0;
    key.lo = lo;
    key.hi = sharded ? -hi : hi;     // a sharded build never matches a single-GPU one
    key.pad = replicate ? 1 : 0;
    key.max_spatial = p->normalize == ADSEL_NORM_LINEAR ? p->max_spatial : 0.0;
    key.max_temporal = (key.temporal && p->normalize == ADSEL_NORM_LINEAR) ? p->max_temporal : 0.0;
    key.lambda_t = (key.temporal && p->aggregate == ADSEL_AGG_SUM) ? p->lambda_t : 0.0;
    if (h->lists_mode >= 0 && memcmp(&key, &h->lists_key, sizeof(key)) == 0) return ADSEL_OK;   // same decision on every rank
    ADSEL_TRY(lists_prepare(h, sharded));
    ListArgs la;
    memset(&la, 0, sizeof(la));
    la.n = n;
    la.src_lo = lo;
    la.src_hi = hi;
    la.radius = p->normalize == ADSEL_NORM_EXP ? 38.0 : host_inf();
    la.norm_mode = p->normalize;
    la.temporal = key.temporal;
    la.aggregate = p->aggregate;
    la.divisor = p->max_spatial;
    la.lambda_t = p->lambda_t;
    la.t_far = t_far;
    la.group_id = h->group_id.p;
    la.tlut = h->tlut.p;
    la.replicate = (sharded && replicate) ? 1 : 0;
    if (sharded) {
        la.world = h->world;
        long long per = ((long long)n + h->world - 1) / h->world;
        la.per = (int)((per + 255) / 256 * 256);
        for (int w = 0; w < h->world; ++w) la.dest[w] = arena_lists(h, w);
    } else {
        la.world = 1;
        la.per = n;
        la.dest[0] = h->lists;
    }
    ADSEL_TRY(clear_stamps(h));
    ADSEL_CUDA(cudaMemsetAsync(h->lists.cursor, 0, 2 * sizeof(unsigned long long), h->stream));
    if (sharded) ADSEL_TRY(comm_barrier(h));      // every rank's cursor is reset before anyone appends
    if (hi > lo) {
        geodesic_lists_kernel<<<min(h->sssp_ctas, hi - lo), SSSP_THREADS, 0, h->stream>>>(
            h->row_ptr.p, h->col.p, h->w.p, la, h->sc_dist.p, h->sc_stamp.p, h->sc_q0.p, h->sc_q1.p, h->sc_touched.p);
        h->launches += 1;
        ADSEL_CUDA(cudaGetLastError());
    }
    if (sharded) ADSEL_TRY(comm_barrier(h));      // every rank's entries have landed in this rank's pool
    h->lists_mode = p->normalize;
    h->lists_key = key;
    return ADSEL_OK;
}

int geodesic_preload() {
    cudaFuncAttributes attr;
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, geodesic_lists_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, geodesic_rows_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, fill_u64_kernel));
    return ADSEL_OK;
}

}  // namespace adsel
