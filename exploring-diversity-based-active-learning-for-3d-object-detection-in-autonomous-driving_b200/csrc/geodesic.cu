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
#include "common.cuh"

namespace adsel {

constexpr int SSSP_THREADS = 128;
constexpr int SSSP_LANES_PER_NODE = 8;   // lanes cooperating on one frontier vertex's edge list

struct SsspScratch {
    unsigned long long *dist;   // [n] ordered-u64 distance, INF_BITS when untouched
    int32_t *stamp;             // [n] round tag of the vertex's last frontier insertion
    int32_t *q0, *q1;           // [n] frontier queues
    int32_t *touched;           // [n] vertices whose dist was ever lowered for this source
};

struct SsspShared {
    int n_front, n_next, n_touched, tag;
};

// Runs the relaxation for `src`; on return sh.n_touched vertices are listed in sc.touched with
// their final distances in sc.dist.  The caller must reset sc.dist for the touched vertices.
__device__ __forceinline__ void sssp_cta(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                         const double *__restrict__ w, int src, double radius,
                                         const SsspScratch &sc, SsspShared &sh) {
    if (threadIdx.x == 0) {
        sc.dist[src] = 0ull;
        sc.q0[0] = src;
        sc.touched[0] = src;
        sh.n_front = 1;
        sh.n_touched = 1;
        sh.n_next = 0;
    }
    __syncthreads();
    int32_t *front = sc.q0, *next = sc.q1;
    const int sub = threadIdx.x % SSSP_LANES_PER_NODE;
    const int slot = threadIdx.x / SSSP_LANES_PER_NODE;
    constexpr int NODES_PER_PASS = SSSP_THREADS / SSSP_LANES_PER_NODE;
    while (true) {
        const int nf = sh.n_front;
        const int tag = sh.tag + 1;
        if (nf == 0) break;
        for (int base = 0; base < nf; base += NODES_PER_PASS) {
            const int qi = base + slot;
            if (qi < nf) {
                const int u = __ldcg(&front[qi]);
                const double du = __longlong_as_double((long long)__ldcg(&sc.dist[u]));
                if (du < radius) {
                    const int e1 = row_ptr[u + 1];
                    for (int e = row_ptr[u] + sub; e < e1; e += SSSP_LANES_PER_NODE) {
                        const int v = col[e];
                        const unsigned long long nd = (unsigned long long)__double_as_longlong(__dadd_rn(du, w[e]));
                        const unsigned long long old = atomicMin(&sc.dist[v], nd);
                        if (nd < old) {
                            if (old == INF_BITS) sc.touched[atomicAdd(&sh.n_touched, 1)] = v;
                            if (atomicExch(&sc.stamp[v], tag) != tag) next[atomicAdd(&sh.n_next, 1)] = v;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            sh.n_front = sh.n_next;
            sh.n_next = 0;
            sh.tag = tag;
        }
        int32_t *tmp = front;
        front = next;
        next = tmp;
        __syncthreads();
    }
}

__device__ __forceinline__ SsspScratch scratch_of(unsigned long long *dist, int32_t *stamp, int32_t *q0, int32_t *q1,
                                                  int32_t *touched, size_t n) {
    SsspScratch sc;
    const size_t o = (size_t)blockIdx.x * n;
    sc.dist = dist + o;
    sc.stamp = stamp + o;
    sc.q0 = q0 + o;
    sc.q1 = q1 + o;
    sc.touched = touched + o;
    return sc;
}

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

// Neighbour lists: for every source the vertices with d < radius and their NORMALISED spatial
// value (spatial_temporal_selector.py:138-144 applied to the row).  Lists are appended to one
// pool through a global cursor; `count_only` sizes the pool on a first pass.
__global__ void __launch_bounds__(SSSP_THREADS)
geodesic_lists_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const double *__restrict__ w,
                      int n, double radius, int norm_mode, double divisor,
                      unsigned long long *sc_dist, int32_t *sc_stamp, int32_t *sc_q0, int32_t *sc_q1, int32_t *sc_touched,
                      unsigned long long *__restrict__ cursor, long long capacity,
                      int64_t *__restrict__ list_off, int32_t *__restrict__ list_len,
                      int32_t *__restrict__ list_idx, double *__restrict__ list_val) {
    __shared__ SsspShared sh;
    __shared__ int n_keep, n_written;
    __shared__ unsigned long long base_off;
    if (threadIdx.x == 0) sh.tag = 0;
    const SsspScratch sc = scratch_of(sc_dist, sc_stamp, sc_q0, sc_q1, sc_touched, (size_t)n);
    __syncthreads();
    for (int src = blockIdx.x; src < n; src += gridDim.x) {
        sssp_cta(row_ptr, col, w, src, radius, sc, sh);
        const int nt = sh.n_touched;
        if (threadIdx.x == 0) {
            n_keep = 0;
            n_written = 0;
        }
        __syncthreads();
        int mine = 0;
        for (int t = threadIdx.x; t < nt; t += SSSP_THREADS) {
            const double d = __longlong_as_double((long long)__ldcg(&sc.dist[sc.touched[t]]));
            mine += d < radius ? 1 : 0;
        }
        if (mine) atomicAdd(&n_keep, mine);
        __syncthreads();
        if (threadIdx.x == 0) base_off = atomicAdd(cursor, (unsigned long long)n_keep);
        __syncthreads();
        const bool fits = (long long)(base_off + n_keep) <= capacity;
        for (int t = threadIdx.x; t < nt; t += SSSP_THREADS) {
            const int v = sc.touched[t];
            const double d = __longlong_as_double((long long)__ldcg(&sc.dist[v]));
            if (d < radius && fits) {
                const int pos = atomicAdd(&n_written, 1);
                list_idx[base_off + pos] = v;
                list_val[base_off + pos] = normalize_value(d, norm_mode, divisor);
            }
            sc.dist[v] = INF_BITS;
        }
        if (threadIdx.x == 0) {
            list_off[src] = (int64_t)base_off;
            list_len[src] = n_keep;
        }
        __syncthreads();
    }
}

static int ensure_scratch(adsel_handle *h) {
    const int ctas = h->num_sms * 4;
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
static int clear_stamps(adsel_handle *h) {
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

// Builds (or reuses) the neighbour lists for `norm_mode`; radius is 38 for exp, inf otherwise.
int build_lists(adsel_handle *h, int norm_mode, double max_spatial) {
    ADSEL_REQUIRE(h->graph_ready, "adsel_select: call adsel_build_knn_graph first");
    if (h->lists_mode == norm_mode && (norm_mode != ADSEL_NORM_LINEAR || h->lists_max_spatial == max_spatial)) return ADSEL_OK;
    const double inf = host_inf();
    const double radius = norm_mode == ADSEL_NORM_EXP ? 38.0 : inf;
    const int n = h->n;
    ADSEL_TRY(ensure_scratch(h));
    ADSEL_TRY(h->cursor.reserve(2));
    ADSEL_TRY(h->list_off.reserve(n));
    ADSEL_TRY(h->list_len.reserve(n));
    ADSEL_TRY(begin_stage(h));
    long long capacity = (long long)h->list_idx.cap;
    if (capacity < (long long)n * 64) {
        capacity = (long long)n * 64;
        ADSEL_TRY(h->list_idx.reserve(capacity));
        ADSEL_TRY(h->list_val.reserve(capacity));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        ADSEL_TRY(clear_stamps(h));
        ADSEL_CUDA(cudaMemsetAsync(h->cursor.p, 0, 2 * sizeof(unsigned long long), h->stream));
        geodesic_lists_kernel<<<min(h->sssp_ctas, n), SSSP_THREADS, 0, h->stream>>>(
            h->row_ptr.p, h->col.p, h->w.p, n, radius, norm_mode, max_spatial, h->sc_dist.p, h->sc_stamp.p, h->sc_q0.p,
            h->sc_q1.p, h->sc_touched.p, h->cursor.p, capacity, h->list_off.p, h->list_len.p, h->list_idx.p, h->list_val.p);
        h->launches += 1;
        ADSEL_CUDA(cudaGetLastError());
        unsigned long long total = 0;
        ADSEL_CUDA(cudaMemcpyAsync(&total, h->cursor.p, sizeof(total), cudaMemcpyDeviceToHost, h->stream));
        ADSEL_CUDA(cudaStreamSynchronize(h->stream));
        h->list_total = (int64_t)total;
        if ((long long)total <= capacity) break;
        ADSEL_REQUIRE(attempt == 0, "adsel_select: neighbour-list pool overflow after resize");
        capacity = (long long)total;
        ADSEL_TRY(h->list_idx.reserve(capacity));
        ADSEL_TRY(h->list_val.reserve(capacity));
    }
    ADSEL_TRY(end_stage(h, 2));
    h->lists_mode = norm_mode;
    h->lists_max_spatial = max_spatial;
    return ADSEL_OK;
}

}  // namespace adsel
