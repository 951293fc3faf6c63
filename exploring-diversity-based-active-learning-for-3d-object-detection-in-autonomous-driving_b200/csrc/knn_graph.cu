// knn_graph.cu -- exact float64 kNN over the ego locations and the symmetrised CSR graph that
// scipy's Dijkstra walks in the reference (det3d/selectors/spatial_temporal_selector.py:93-104).
//
//   knn_bruteforce_kernel   replaces cKDTree(locations).query(locations, k+1)          (:97-98)
//   count/fill edge kernels replace the symmetric dense scatter (:99-101) followed by scipy's
//                           dense->CSR conversion, which drops entries with |w| <= 1e-8.
//
// Arithmetic contract (bit-exact with scipy): d2 = dx*dx + dy*dy with separately
// rounded products (no FMA), neighbours ordered by (d2, index), reported distance sqrt(d2).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace adsel {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 1024;   // candidates staged per shared-memory tile (16 KB)

// One query per thread; candidates stream through shared memory as double2 (broadcast reads).
// Each thread keeps its K1 = k+1 best (d2, idx) sorted ascending in registers.
// SPLIT > 1: SPLIT consecutive threads share one query and scan interleaved candidate tiles; their
// partial lists are merged through shared memory (used when n is too small to fill the GPU).
// Destination tables of the kNN rows: one per rank.  In the sharded build every rank computes the rows
// of its own slice [q_lo, q_hi) and stores them straight into every rank's table (peer memory over
// NVLink): the all-gather of the kNN table is fused into the kernel's epilogue.
struct KnnOut {
    int world;
    int32_t *idx[8];
    double *dist[8];
};

template <int K1>
__global__ void __launch_bounds__(KNN_THREADS)
knn_bruteforce_kernel(const double2 *__restrict__ loc, int n, int q_lo, int q_hi, const KnnOut out) {
    __shared__ double2 tile[KNN_TILE];
    const int q = q_lo + blockIdx.x * KNN_THREADS + threadIdx.x;
    const double2 me = loc[q < q_hi ? q : q_hi - 1];
    double best_d[K1];
    int best_i[K1];
#pragma unroll
    for (int s = 0; s < K1; ++s) {
        best_d[s] = __longlong_as_double(INF_BITS);
        best_i[s] = -1;
    }
    for (int base = 0; base < n; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += KNN_THREADS) tile[t] = loc[base + t];
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            const double2 c = tile[t];
            const double dx = __dsub_rn(me.x, c.x), dy = __dsub_rn(me.y, c.y);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            // candidates arrive in ascending index: strict '<' keeps the lower index on a tie with the last entry;
            // the displaced entries carried down the list must be ordered by (d2, index) explicitly
            if (d2 < best_d[K1 - 1]) {
                double cd = d2;
                int ci = base + t;
#pragma unroll
                for (int s = 0; s < K1; ++s) {
                    if (cd < best_d[s] || (cd == best_d[s] && ci < best_i[s])) {
                        const double td = best_d[s];
                        const int ti = best_i[s];
                        best_d[s] = cd;
                        best_i[s] = ci;
                        cd = td;
                        ci = ti;
                    }
                }
            }
        }
    }
    if (q < q_hi) {
#pragma unroll
        for (int s = 0; s < K1; ++s) best_d[s] = __dsqrt_rn(best_d[s]);
        for (int wdst = 0; wdst < out.world; ++wdst) {
            int32_t *oi = out.idx[wdst] + (size_t)q * K1;
            double *od = out.dist[wdst] + (size_t)q * K1;
#pragma unroll
            for (int s = 0; s < K1; ++s) {
                oi[s] = best_i[s];
                od[s] = best_d[s];
            }
        }
    }
}

// Generic k (K1 up to 64) with the running list in local memory; same ordering contract.
__global__ void __launch_bounds__(KNN_THREADS)
knn_bruteforce_generic_kernel(const double2 *__restrict__ loc, int n, int q_lo, int q_hi, int k1, const KnnOut out) {
    __shared__ double2 tile[KNN_TILE];
    const int q = q_lo + blockIdx.x * KNN_THREADS + threadIdx.x;
    const double2 me = loc[q < q_hi ? q : q_hi - 1];
    double best_d[64];
    int best_i[64];
    for (int s = 0; s < k1; ++s) {
        best_d[s] = __longlong_as_double(INF_BITS);
        best_i[s] = -1;
    }
    for (int base = 0; base < n; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += KNN_THREADS) tile[t] = loc[base + t];
        __syncthreads();
        for (int t = 0; t < cnt; ++t) {
            const double2 c = tile[t];
            const double dx = __dsub_rn(me.x, c.x), dy = __dsub_rn(me.y, c.y);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            if (d2 < best_d[k1 - 1]) {
                int s = k1 - 1;
                while (s > 0 && d2 < best_d[s - 1]) {
                    best_d[s] = best_d[s - 1];
                    best_i[s] = best_i[s - 1];
                    --s;
                }
                best_d[s] = d2;
                best_i[s] = base + t;
            }
        }
    }
    if (q < q_hi) {
        for (int wdst = 0; wdst < out.world; ++wdst) {
            for (int s = 0; s < k1; ++s) {
                out.idx[wdst][(size_t)q * k1 + s] = best_i[s];
                out.dist[wdst][(size_t)q * k1 + s] = __dsqrt_rn(best_d[s]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Uniform-grid exact kNN (pools of more than a few thousand frames).  Same arithmetic and the same
// (d2, index) ordering as the brute-force kernel, but a query only visits the cells of growing square
// rings around its own cell until its current (k+1)-th distance is provably smaller than the distance to
// anything outside the visited block.  Ego poses lie on a few dozen drive tracks, so the k nearest are
// found within one or two rings: ~10^2 candidates per query instead of N.
// ---------------------------------------------------------------------------------------------
struct KnnGrid {
    double x0, y0, cell;
    int nx, ny;
};

__device__ __forceinline__ int grid_coord(double v, double v0, double cell, int nmax) {
    const int c = (int)floor((v - v0) / cell);
    return min(max(c, 0), nmax - 1);
}

__global__ void grid_count_kernel(const double2 *__restrict__ loc, int n, KnnGrid g, int32_t *__restrict__ cell_id,
                                  int32_t *__restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 p = loc[i];
    const int c = grid_coord(p.y, g.y0, g.cell, g.ny) * g.nx + grid_coord(p.x, g.x0, g.cell, g.nx);
    cell_id[i] = c;
    atomicAdd(&counts[c], 1);
}

// exclusive scan of counts[0..m) into start[0..m], single CTA (m <= a few million)
__global__ void __launch_bounds__(1024) scan_i32_kernel(const int32_t *__restrict__ counts, int m, int32_t *__restrict__ start) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < m; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < m ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const int before = carry + (wid > 0 ? warp_sum[wid - 1] : 0) + incl - v;
        if (i < m) start[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[m] = carry;
}

__global__ void grid_scatter_kernel(const double2 *__restrict__ loc, int n, const int32_t *__restrict__ cell_id,
                                    const int32_t *__restrict__ start, int32_t *__restrict__ fill,
                                    int32_t *__restrict__ sorted_idx, double2 *__restrict__ sorted_loc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_id[i];
    const int pos = start[c] + atomicAdd(&fill[c], 1);
    sorted_idx[pos] = i;
    sorted_loc[pos] = loc[i];
}

template <int K1>
__global__ void __launch_bounds__(KNN_THREADS)
knn_grid_kernel(const double2 *__restrict__ sorted_loc, const int32_t *__restrict__ sorted_idx,
                const int32_t *__restrict__ start, KnnGrid g, int n, int q_lo, int q_hi, const KnnOut out) {
    const int t = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (t >= n) return;
    const int q = sorted_idx[t];
    if (q < q_lo || q >= q_hi) return;          // another rank's row
    const double2 me = sorted_loc[t];
    const int cx = grid_coord(me.x, g.x0, g.cell, g.nx), cy = grid_coord(me.y, g.y0, g.cell, g.ny);
    double best_d[K1];
    int best_i[K1];
#pragma unroll
    for (int s = 0; s < K1; ++s) {
        best_d[s] = __longlong_as_double(INF_BITS);
        best_i[s] = 0x7fffffff;
    }
    // distance from the query to the border of its own cell (0 for points clamped from outside the box)
    const double fx = me.x - (g.x0 + cx * g.cell), fy = me.y - (g.y0 + cy * g.cell);
    const double inner = fmax(0.0, fmin(fmin(fx, g.cell - fx), fmin(fy, g.cell - fy)));
    const int r_max = max(max(cx, g.nx - 1 - cx), max(cy, g.ny - 1 - cy));
    for (int r = 0; r <= r_max; ++r) {
        for (int dy = -r; dy <= r; ++dy) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= g.ny) continue;
            const bool full_row = dy == -r || dy == r;
            // full row of the ring: one contiguous run of cells; otherwise only the two end cells
            for (int part = 0; part < (full_row || r == 0 ? 1 : 2); ++part) {
                int xa, xb;
                if (full_row) {
                    xa = max(cx - r, 0);
                    xb = min(cx + r, g.nx - 1);
                } else {
                    xa = xb = part == 0 ? cx - r : cx + r;
                    if (xa < 0 || xa >= g.nx) continue;
                }
                const int p0 = start[yy * g.nx + xa], p1 = start[yy * g.nx + xb + 1];
                for (int p = p0; p < p1; ++p) {
                    const double2 c = sorted_loc[p];
                    const double dx = __dsub_rn(me.x, c.x), dyv = __dsub_rn(me.y, c.y);
                    const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dyv, dyv));
                    if (d2 <= best_d[K1 - 1]) {
                        double cd = d2;
                        int ci = sorted_idx[p];
                        // candidates arrive in cell order: order by (d2, index) explicitly
#pragma unroll
                        for (int s = 0; s < K1; ++s) {
                            if (cd < best_d[s] || (cd == best_d[s] && ci < best_i[s])) {
                                const double td = best_d[s];
                                const int ti = best_i[s];
                                best_d[s] = cd;
                                best_i[s] = ci;
                                cd = td;
                                ci = ti;
                            }
                        }
                    }
                }
            }
        }
        // everything not yet visited is at least `bound` away (shrunk by 1e-9 against rounding of d2)
        const double bound = (r * g.cell + inner) * (1.0 - 1e-9);
        if (best_d[K1 - 1] < bound * bound) break;
    }
#pragma unroll
    for (int s = 0; s < K1; ++s) best_d[s] = __dsqrt_rn(best_d[s]);
    for (int wdst = 0; wdst < out.world; ++wdst) {
        int32_t *oi = out.idx[wdst] + (size_t)q * K1;
        double *od = out.dist[wdst] + (size_t)q * K1;
#pragma unroll
        for (int s = 0; s < K1; ++s) {
            oi[s] = best_i[s];
            od[s] = best_d[s];
        }
    }
}

// An entry of the kNN table is an edge of the reference's graph unless scipy's dense->CSR
// conversion masks it: np.isclose(w, 0, rtol=1e-5, atol=1e-8)  <=>  |w| <= 1e-8.
__device__ __forceinline__ bool is_edge(int nbr, double d) { return nbr >= 0 && !(fabs(d) <= NULL_EDGE_ATOL); }

// deg_own[i] = edges i keeps from its own kNN list; deg_rev[i] = edges (j -> i) whose reverse
// (i -> j) is not already in i's own list (the dense scatter writes both [i,j] and [j,i]).
__global__ void count_edges_kernel(const int32_t *__restrict__ knn_idx, const double *__restrict__ knn_dist,
                                   int n, int k1, int32_t *__restrict__ deg_own, int32_t *__restrict__ deg_rev) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k1) return;
    const int j = t / k1;
    const int i = knn_idx[t];
    const double d = knn_dist[t];
    if (!is_edge(i, d)) return;
    atomicAdd(&deg_own[j], 1);
    bool mutual = false;
    for (int s = 0; s < k1; ++s) mutual |= (knn_idx[(size_t)i * k1 + s] == j && is_edge(j, knn_dist[(size_t)i * k1 + s]));
    if (!mutual) atomicAdd(&deg_rev[i], 1);
}

// Single-CTA exclusive scan of deg_own + deg_rev -> row_ptr[0..n]; n <= a few 100k so one CTA
// walking the array in 1024-wide strips is a few microseconds and keeps the build graph-free.
__global__ void __launch_bounds__(1024)
scan_degrees_kernel(const int32_t *__restrict__ deg_own, const int32_t *__restrict__ deg_rev, int n,
                    int32_t *__restrict__ row_ptr) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? deg_own[i] + deg_rev[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const int before = carry + (wid > 0 ? warp_sum[wid - 1] : 0) + incl - v;
        if (i < n) row_ptr[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_ptr[n] = carry;
}

// Own edges land at row_ptr[j] + (rank among j's valid slots); reverse-only edges after them in
// arrival order (order inside a row never matters: relaxations take a min).
__global__ void fill_edges_kernel(const int32_t *__restrict__ knn_idx, const double *__restrict__ knn_dist,
                                  int n, int k1, const int32_t *__restrict__ row_ptr,
                                  const int32_t *__restrict__ deg_own, int32_t *__restrict__ rev_fill,
                                  int32_t *__restrict__ col, double *__restrict__ w) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k1) return;
    const int j = t / k1, slot = t - j * k1;
    const int i = knn_idx[t];
    const double d = knn_dist[t];
    if (!is_edge(i, d)) return;
    int rank = 0;
    for (int s = 0; s < slot; ++s) rank += is_edge(knn_idx[(size_t)j * k1 + s], knn_dist[(size_t)j * k1 + s]) ? 1 : 0;
    col[row_ptr[j] + rank] = i;
    w[row_ptr[j] + rank] = d;
    bool mutual = false;
    for (int s = 0; s < k1; ++s) mutual |= (knn_idx[(size_t)i * k1 + s] == j && is_edge(j, knn_dist[(size_t)i * k1 + s]));
    if (!mutual) {
        const int pos = row_ptr[i] + deg_own[i] + atomicAdd(&rev_fill[i], 1);
        col[pos] = j;
        w[pos] = d;
    }
}

template <int K1>
static void launch_knn(adsel_handle *h, int q_lo, int q_hi, const KnnOut &out) {
    const int grid = (q_hi - q_lo + KNN_THREADS - 1) / KNN_THREADS;
    knn_bruteforce_kernel<K1><<<grid, KNN_THREADS, 0, h->stream>>>(h->loc.p, h->n, q_lo, q_hi, out);
}

// Grid build + query; returns false when the brute-force kernel should be used instead.
template <int K1>
static void launch_knn_grid(adsel_handle *h, const KnnGrid &g, int q_lo, int q_hi, const KnnOut &out) {
    knn_grid_kernel<K1><<<(h->n + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, h->stream>>>(
        h->grid_loc.p, h->grid_idx.p, h->grid_start.p, g, h->n, q_lo, q_hi, out);
}

static bool use_grid(const adsel_handle *h, int k1) {
    const char *e = getenv("ADSEL_KNN");
    if (e && !strcmp(e, "brute")) return false;
    if (k1 != 9 && k1 != 5 && k1 != 17) return false;
    if (e && !strcmp(e, "grid")) return true;
    return h->n >= 1024;
}

static KnnGrid make_grid(const adsel_handle *h, int k1) {
    KnnGrid g;
    const double ex = h->bbox[2] - h->bbox[0], ey = h->bbox[3] - h->bbox[1];
    // about two frames per cell on average over the bounding box, at most 2048 x 2048 cells
    double cell = sqrt(fmax(ex, 1e-9) * fmax(ey, 1e-9) * 2.0 / (double)h->n);
    cell = fmax(cell, fmax(ex, ey) / 2048.0);
    if (!(cell > 0.0) || !isfinite(cell)) cell = 1.0;
    g.x0 = h->bbox[0];
    g.y0 = h->bbox[1];
    g.cell = cell;
    g.nx = (int)fmin(2048.0, floor(ex / cell) + 1.0);
    g.ny = (int)fmin(2048.0, floor(ey / cell) + 1.0);
    (void)k1;
    return g;
}

int knn_grid_prepare(adsel_handle *h, int k1) {
    if (!use_grid(h, k1)) return ADSEL_OK;
    const KnnGrid g = make_grid(h, k1);
    const size_t cells = (size_t)g.nx * g.ny;
    ADSEL_TRY(h->grid_cell.reserve(h->n));
    ADSEL_TRY(h->grid_idx.reserve(h->n));
    ADSEL_TRY(h->grid_loc.reserve(h->n));
    ADSEL_TRY(h->grid_start.reserve(2 * (cells + 1)));   // [0, cells] starts, then cells counters (count / fill)
    return ADSEL_OK;
}

static void enqueue_knn(adsel_handle *h, int k1, int q_lo, int q_hi, const KnnOut &out) {
    if (q_hi <= q_lo) return;
    if (use_grid(h, k1)) {
        const KnnGrid g = make_grid(h, k1);
        const int cells = g.nx * g.ny, n = h->n;
        int32_t *start = h->grid_start.p, *counter = h->grid_start.p + cells + 1;
        cudaMemsetAsync(counter, 0, sizeof(int32_t) * cells, h->stream);
        grid_count_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->loc.p, n, g, h->grid_cell.p, counter);
        scan_i32_kernel<<<1, 1024, 0, h->stream>>>(counter, cells, start);
        cudaMemsetAsync(counter, 0, sizeof(int32_t) * cells, h->stream);
        grid_scatter_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->loc.p, n, h->grid_cell.p, start, counter, h->grid_idx.p,
                                                                   h->grid_loc.p);
        switch (k1) {
            case 9: launch_knn_grid<9>(h, g, q_lo, q_hi, out); break;
            case 5: launch_knn_grid<5>(h, g, q_lo, q_hi, out); break;
            default: launch_knn_grid<17>(h, g, q_lo, q_hi, out); break;
        }
        h->launches += 4;
        return;
    }
    switch (k1) {
        case 9: launch_knn<9>(h, q_lo, q_hi, out); break;
        case 5: launch_knn<5>(h, q_lo, q_hi, out); break;
        case 17: launch_knn<17>(h, q_lo, q_hi, out); break;
        default: {
            const int grid = (q_hi - q_lo + KNN_THREADS - 1) / KNN_THREADS;
            knn_bruteforce_generic_kernel<<<grid, KNN_THREADS, 0, h->stream>>>(h->loc.p, h->n, q_lo, q_hi, k1, out);
        }
    }
    h->launches += 1;
}

// count -> scan -> fill on the kNN table in use.  col/w are sized for the worst case (every table entry an
// edge in both directions) so that no host read-back sits between the kernels.
static int enqueue_csr(adsel_handle *h) {
    const int n = h->n, k1 = h->k + 1;
    ADSEL_CUDA(cudaMemsetAsync(h->deg_own.p, 0, sizeof(int32_t) * n, h->stream));
    ADSEL_CUDA(cudaMemsetAsync(h->deg_rev.p, 0, sizeof(int32_t) * 2 * n, h->stream));
    const int threads = 256, blocks = (int)(((size_t)n * k1 + threads - 1) / threads);
    count_edges_kernel<<<blocks, threads, 0, h->stream>>>(h->knn_idx_p, h->knn_dist_p, n, k1, h->deg_own.p, h->deg_rev.p);
    scan_degrees_kernel<<<1, 1024, 0, h->stream>>>(h->deg_own.p, h->deg_rev.p, n, h->row_ptr.p);
    fill_edges_kernel<<<blocks, threads, 0, h->stream>>>(h->knn_idx_p, h->knn_dist_p, n, k1, h->row_ptr.p, h->deg_own.p,
                                                         h->deg_rev.p + n, h->col.p, h->w.p);
    h->launches += 3;
    ADSEL_CUDA(cudaGetLastError());
    h->nnz = -1;   // read back lazily (adsel_get_csr / adsel_get_counters)
    return ADSEL_OK;
}

int knn_prepare(adsel_handle *h, int32_t k, bool own_tables) {
    ADSEL_REQUIRE(h->n > 0, "adsel_build_knn_graph: call adsel_set_pool first");
    ADSEL_REQUIRE(k >= 1 && k + 1 <= 64 && k + 1 <= h->n, "adsel_build_knn_graph: need 1 <= k < min(n, 64)");
    const int n = h->n, k1 = k + 1;
    if (own_tables) {
        ADSEL_TRY(h->knn_idx.reserve((size_t)n * k1));
        ADSEL_TRY(h->knn_dist.reserve((size_t)n * k1));
    }
    ADSEL_TRY(h->row_ptr.reserve(n + 1));
    ADSEL_TRY(h->deg_own.reserve(n));
    ADSEL_TRY(h->deg_rev.reserve(2 * (size_t)n));   // [0,n) reverse counts, [n,2n) fill cursors
    ADSEL_TRY(h->col.reserve(2 * (size_t)n * k1));
    ADSEL_TRY(h->w.reserve(2 * (size_t)n * k1));
    ADSEL_TRY(knn_grid_prepare(h, k1));
    return ADSEL_OK;
}

int build_knn_graph(adsel_handle *h, int32_t k) {
    ADSEL_TRY(knn_prepare(h, k, true));
    const int n = h->n, k1 = k + 1;
    h->k = k;
    h->graph_ready = false;
    h->lists_mode = -1;
    h->knn_idx_p = h->knn_idx.p;
    h->knn_dist_p = h->knn_dist.p;
    KnnOut out;
    memset(&out, 0, sizeof(out));
    out.world = 1;
    out.idx[0] = h->knn_idx_p;
    out.dist[0] = h->knn_dist_p;
    ADSEL_TRY(begin_stage(h));
    enqueue_knn(h, k1, 0, n, out);
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_TRY(end_stage(h, 0));
    ADSEL_TRY(begin_stage(h));
    ADSEL_TRY(enqueue_csr(h));
    ADSEL_TRY(end_stage(h, 1));
    h->graph_ready = true;
    return ADSEL_OK;
}

int32_t *arena_knn_idx(adsel_handle *h, int w);
double *arena_knn_dist(adsel_handle *h, int w);
int comm_barrier(adsel_handle *h);

// Sharded build (enqueue only): this rank's kNN rows -> every rank's table, device barrier, CSR.
int build_knn_graph_sharded(adsel_handle *h, int32_t k) {
    ADSEL_REQUIRE(h->comm_ready, "adsel_build_knn_graph_sharded: call adsel_comm_init / adsel_comm_attach first");
    ADSEL_REQUIRE(h->arena_n == h->n && h->arena_k1 == k + 1, "adsel_build_knn_graph_sharded: arena was sized for another n / k");
    ADSEL_TRY(knn_prepare(h, k, false));
    const int k1 = k + 1;
    h->k = k;
    h->graph_ready = false;
    h->lists_mode = -1;
    h->knn_idx_p = arena_knn_idx(h, h->rank);
    h->knn_dist_p = arena_knn_dist(h, h->rank);
    KnnOut out;
    memset(&out, 0, sizeof(out));
    out.world = h->world;
    for (int w = 0; w < h->world; ++w) {
        out.idx[w] = arena_knn_idx(h, w);
        out.dist[w] = arena_knn_dist(h, w);
    }
    ADSEL_CUDA(cudaEventRecord(h->knn_ev[0], h->stream));
    enqueue_knn(h, k1, h->shard_lo, h->shard_hi, out);
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_TRY(comm_barrier(h));      // every rank's rows have landed in this rank's table
    ADSEL_TRY(enqueue_csr(h));
    ADSEL_TRY(comm_barrier(h));      // nobody overwrites a table (next build) while a peer still reads its copy
    ADSEL_CUDA(cudaEventRecord(h->knn_ev[1], h->stream));
    h->knn_pending_timing = true;
    h->graph_ready = true;
    return ADSEL_OK;
}

int knn_preload() {
    cudaFuncAttributes attr;
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, knn_bruteforce_kernel<9>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, knn_bruteforce_kernel<5>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, knn_bruteforce_kernel<17>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, knn_bruteforce_generic_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, knn_grid_kernel<9>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, knn_grid_kernel<5>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, knn_grid_kernel<17>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, grid_count_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, scan_i32_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, grid_scatter_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, count_edges_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, scan_degrees_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, fill_edges_kernel));
    return ADSEL_OK;
}

}  // namespace adsel
