// lazy.cu -- the greedy loop for rows WITHOUT a saturation radius: one geodesic row per SELECTED frame.
//
// With normalize="linear" (det3d/selectors/spatial_temporal_selector.py:139-141) or the raw distances of
// SpatialSelector (det3d/selectors/spatial_selector.py:95-138) the far value of the geodesic term is inf, so a frame's
// row differs from `far` on its whole connected component: precomputed neighbour lists (geodesic.cu) degenerate to the
// all-pairs map, sum of |component|^2 entries.  The reference computes that map (N x N float64, 82 % of its run time);
// only the rows of the frames it then SELECTS are ever read (:175-181), ~1 k-4 k of 28,130.  This kernel therefore runs
// one truncation-free single-source search per selected frame inside the greedy loop:
//
//   step:  d = SSSP(cur)  (sssp.cuh: the label-correcting search whose fixpoint is scipy's Dijkstra row, bit for bit)
//          one dense pass over the pool: row[j] = aggregate(normalize(d[j] or inf), temporal(cur, j));
//                                        fps[j] = min(fps[j], row[j]);  arg-max with np.argmax's first-maximum rule;
//                                        d[j] reset for the next search
//          cost bookkeeping and the :182 assertion exactly as in greedy.cu.
//
// One CTA (the searches are chains of dependent frontier rounds, like the list build), no N x N storage: memory is
// O(N), work O(steps x component).  Exactness follows from the same arguments as geodesic.cu / greedy.cu: identical
// IEEE operation sequence per entry, ordered-u64 keys, lowest index among equal maxima.
#include <stdlib.h>

#include "common.cuh"
#include "sssp.cuh"

namespace adsel {

constexpr int LAZY_NT = 512;

struct LazyArgs {
    int n, temporal, aggregate, norm_mode;
    double divisor, lambda_t, t_far, far;
    double cost_b, cost_f, cost_start, budget;
    int forbid_sampled, n_sampled, first_index, out_cap;
    const int32_t *row_ptr, *col;
    const double *w;
    const int32_t *group_id, *n_boxes;
    const double *tlut;
    unsigned long long *fps;
    unsigned char *is_selected;
    const int32_t *sampled;
    int32_t *out;
    double *result;   // [0] cost, [1] status, [2] m, [3] steps, [4] offending frame, [5] 0
    SsspScratch sc;
};

struct LazyShared {
    unsigned hi[LAZY_NT / 32], lo[LAZY_NT / 32];
    int idx[LAZY_NT / 32];
    int winner;
};

// fps = min(fps, distance_map[src]) for every frame, given src's search result in sc.dist; returns (to every
// thread) the arg-max of the updated vector.  sc.dist is reset to "untouched" on the way.
__device__ __forceinline__ int lazy_update(const LazyArgs &a, LazyShared &ls, int src, bool want_best) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gsrc = a.temporal ? a.group_id[src] : 0;
    unsigned bhi = 0u, blo = 0u;
    int bidx = 0x7fffffff;
    for (int j = tid; j < a.n; j += LAZY_NT) {
        const unsigned long long db = __ldcg(&a.sc.dist[j]);   // written by the search with L2 atomics
        a.sc.dist[j] = INF_BITS;
        // untouched entries hold +inf: normalize(inf) = inf for "raw" and "linear" (IEEE inf / divisor)
        const double s = normalize_value(__longlong_as_double((long long)db), a.norm_mode, a.divisor);
        double v = s;
        if (a.temporal) {
            const double tv = a.group_id[j] == gsrc ? a.tlut[abs(j - src)] : a.t_far;
            v = aggregate_value(s, tv, a.aggregate, a.lambda_t);
        }
        const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
        const unsigned long long old = a.fps[j];
        const unsigned long long nv = vb < old ? vb : old;
        if (vb < old) a.fps[j] = vb;
        const unsigned h = (unsigned)(nv >> 32), l = (unsigned)nv;
        if (h > bhi || (h == bhi && l > blo) || bidx == 0x7fffffff) {   // ascending j per thread: strict '>' keeps the first
            bhi = h;
            blo = l;
            bidx = j;
        }
    }
    if (!want_best) {
        __syncthreads();
        return -1;
    }
    {
        unsigned hi[1] = {bhi}, lo[1] = {blo};
        int idx[1] = {bidx};
        const Win x = warp_argmax<1>(hi, lo, idx);
        if (lane == 0) {
            ls.hi[wid] = x.hi;
            ls.lo[wid] = x.lo;
            ls.idx[wid] = x.idx;
        }
    }
    __syncthreads();
    if (wid == 0) {
        unsigned hi[1] = {lane < LAZY_NT / 32 ? ls.hi[lane] : 0u}, lo[1] = {lane < LAZY_NT / 32 ? ls.lo[lane] : 0u};
        int idx[1] = {lane < LAZY_NT / 32 ? ls.idx[lane] : 0x7fffffff};
        const Win x = warp_argmax<1>(hi, lo, idx);
        if (lane == 0) ls.winner = x.idx;
    }
    __syncthreads();
    return ls.winner;
}

__global__ void __launch_bounds__(LAZY_NT, 1) greedy_lazy_kernel(const LazyArgs a) {
    __shared__ SsspShared sh;
    __shared__ LazyShared ls;
    const int tid = threadIdx.x;
    const unsigned long long far_bits = (unsigned long long)__double_as_longlong(a.far);
    for (int j = tid; j < a.n; j += LAZY_NT) {
        a.fps[j] = far_bits;
        a.is_selected[j] = 0;
    }
    if (tid == 0) sh.tag = 0;
    __syncthreads();
    if (a.forbid_sampled)
        for (int t = tid; t < a.n_sampled; t += LAZY_NT) a.is_selected[a.sampled[t]] = 1;
    const double radius = __longlong_as_double((long long)INF_BITS);
    double cost = a.cost_start;   // every thread keeps the bookkeeping (identical values)
    int m = 0, cur = -1;
    long long steps = 0;
    int status = 0, offending = -1;
    bool seeding = a.n_sampled > 0;   // first pick = argmax without a preceding update (:165)
    if (a.n_sampled == 0) {           // empty buffer: the caller's random.choice is taken and costed up front (:168,172-174)
        cur = a.first_index;
        if (tid == 0) {
            a.out[0] = cur;
            a.is_selected[cur] = 1;
        }
        m = 1;
        cost = __dadd_rn(cost, a.cost_f);
        cost = __dadd_rn(cost, __dmul_rn((double)a.n_boxes[cur], a.cost_b));
    } else {
        for (int t = 0; t < a.n_sampled; ++t) {   // fps = min over the rows of the frames already in the buffer (:160-164)
            const int src = a.sampled[t];
            sssp_cta<LAZY_NT>(a.row_ptr, a.col, a.w, src, radius, a.sc, sh);
            lazy_update(a, ls, src, false);
        }
    }
    while (true) {
        int nxt;
        if (seeding) {
            // arg-max of the current vector without an update: a pass with an all-inf "row" leaves fps unchanged
            // (sc.dist is all untouched) -- but the temporal term would still lower entries, so scan fps directly
            unsigned bhi = 0u, blo = 0u;
            int bidx = 0x7fffffff;
            for (int j = tid; j < a.n; j += LAZY_NT) {
                const unsigned long long v = a.fps[j];
                const unsigned h = (unsigned)(v >> 32), l = (unsigned)v;
                if (h > bhi || (h == bhi && l > blo) || bidx == 0x7fffffff) {
                    bhi = h;
                    blo = l;
                    bidx = j;
                }
            }
            unsigned hi[1] = {bhi}, lo[1] = {blo};
            int idx[1] = {bidx};
            const Win x = warp_argmax<1>(hi, lo, idx);
            if ((tid & 31) == 0) {
                ls.hi[tid >> 5] = x.hi;
                ls.lo[tid >> 5] = x.lo;
                ls.idx[tid >> 5] = x.idx;
            }
            __syncthreads();
            if (tid < 32) {
                unsigned hi2[1] = {tid < LAZY_NT / 32 ? ls.hi[tid] : 0u}, lo2[1] = {tid < LAZY_NT / 32 ? ls.lo[tid] : 0u};
                int idx2[1] = {tid < LAZY_NT / 32 ? ls.idx[tid] : 0x7fffffff};
                const Win y = warp_argmax<1>(hi2, lo2, idx2);
                if (tid == 0) ls.winner = y.idx;
            }
            __syncthreads();
            nxt = ls.winner;
        } else {
            sssp_cta<LAZY_NT>(a.row_ptr, a.col, a.w, cur, radius, a.sc, sh);
            nxt = lazy_update(a, ls, cur, true);
        }
        ++steps;
        if (nxt < 0 || nxt >= a.n) {
            status = ADSEL_ERR_INVALID;
            break;
        }
        const double box_cost = __dmul_rn((double)a.n_boxes[nxt], a.cost_b);   // one rounded product (base_selector.py:85)
        if (seeding) {   // :165,172-174 -- always taken, no budget check
            seeding = false;
            if (tid == 0) {
                a.out[0] = nxt;
                a.is_selected[nxt] = 1;
            }
            m = 1;
            cost = __dadd_rn(cost, a.cost_f);
            cost = __dadd_rn(cost, box_cost);
            cur = nxt;
            __syncthreads();
            continue;
        }
        if (__ldcg(&a.is_selected[nxt])) {   // :182 assert (flags are written by thread 0 before a barrier of the next step)
            status = ADSEL_ERR_RESELECTED;
            offending = nxt;
            break;
        }
        cost = __dadd_rn(cost, a.cost_f);       // :184
        cost = __dadd_rn(cost, box_cost);       // :185
        if (cost > a.budget) break;             // :186
        if (m >= a.out_cap) {
            status = ADSEL_ERR_CAPACITY;
            offending = -2;
            break;
        }
        __syncthreads();   // every thread has read the flag before it is set
        if (tid == 0) {
            a.out[m] = nxt;
            __stcg(&a.is_selected[nxt], (unsigned char)1);
        }
        m += 1;
        cur = nxt;
    }
    if (tid == 0) {
        a.result[0] = cost;
        a.result[1] = (double)status;
        a.result[2] = (double)m;
        a.result[3] = (double)steps;
        a.result[4] = (double)offending;
        a.result[5] = 0.0;
    }
}

__global__ void temporal_lut_kernel(double *tlut, int n, int norm_mode, double divisor);
int ensure_scratch(adsel_handle *h);
int clear_stamps(adsel_handle *h);
int validate_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                    int32_t first_index, int32_t out_cap);
int finish_select(adsel_handle *h, const double *result, int32_t *out_idx, int32_t *out_m, double *out_cost, const char *who);
double far_value(const adsel_select_params *p, double *s_far_out, double *t_far_out);

// Rows without a saturation radius: normalize "raw" / "linear" with a geodesic term.  ADSEL_LAZY=1 forces the lazy
// loop, 0 forces precomputed lists; default: lazy from 16,384 frames (below, the all-pairs lists are small enough and
// make the loop itself faster).
bool lazy_wanted(const adsel_handle *h, const adsel_select_params *p) {
    if (p->spatial != ADSEL_SPATIAL_GEODESIC || p->normalize == ADSEL_NORM_EXP) return false;
    const char *e = getenv("ADSEL_LAZY");
    if (e && *e) return atoi(e) != 0;
    return h->n >= 16384;
}

int lazy_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_TRY(validate_select(h, p, sampled, n_sampled, first_index, out_cap));
    ADSEL_REQUIRE(out_idx && out_m, "adsel_select: null argument");
    ADSEL_REQUIRE(h->graph_ready, "adsel_select: call adsel_build_knn_graph first");
    const int n = h->n;
    ADSEL_TRY(ensure_scratch(h));
    ADSEL_TRY(h->fps.reserve(n));
    ADSEL_TRY(h->is_selected.reserve((size_t)n * 16));
    ADSEL_TRY(h->tlut.reserve(n));
    ADSEL_TRY(h->d_sampled.reserve(n_sampled > 0 ? n_sampled : 1));
    ADSEL_TRY(h->d_out.reserve(out_cap));
    ADSEL_TRY(h->d_result.reserve(8));
    ADSEL_TRY(begin_stage(h));
    ADSEL_TRY(clear_stamps(h));
    LazyArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    a.temporal = p->temporal ? 1 : 0;
    a.aggregate = p->aggregate;
    a.norm_mode = p->normalize;
    a.divisor = p->max_spatial;
    a.lambda_t = p->lambda_t;
    double s_far = 0.0, t_far = 0.0;
    a.far = far_value(p, &s_far, &t_far);
    a.t_far = t_far;
    a.cost_b = p->cost_b;
    a.cost_f = p->cost_f;
    a.cost_start = p->cost_start;
    a.budget = p->budget;
    a.forbid_sampled = p->forbid_sampled;
    a.n_sampled = n_sampled;
    a.first_index = first_index;
    a.out_cap = out_cap;
    a.row_ptr = h->row_ptr.p;
    a.col = h->col.p;
    a.w = h->w.p;
    a.group_id = h->group_id.p;
    a.n_boxes = h->n_boxes.p;
    a.tlut = h->tlut.p;
    a.fps = h->fps.p;
    a.is_selected = h->is_selected.p;
    a.sampled = h->d_sampled.p;
    a.out = h->d_out.p;
    a.result = h->d_result.p;
    a.sc.dist = h->sc_dist.p;       // the first resident CTA's scratch of the list build
    a.sc.stamp = h->sc_stamp.p;
    a.sc.q0 = h->sc_q0.p;
    a.sc.q1 = h->sc_q1.p;
    a.sc.touched = h->sc_touched.p;
    if (n_sampled > 0)
        ADSEL_CUDA(cudaMemcpyAsync(h->d_sampled.p, sampled, sizeof(int32_t) * n_sampled, cudaMemcpyDefault, h->stream));   // host or device list
    ADSEL_CUDA(cudaMemsetAsync(h->d_result.p, 0, sizeof(double) * 8, h->stream));
    if (p->temporal) {
        temporal_lut_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->tlut.p, n, p->normalize, p->max_temporal);
        h->launches += 1;
        ADSEL_CUDA(cudaGetLastError());
    }
    ADSEL_TRY(end_stage(h, 2));
    ADSEL_TRY(begin_stage(h));
    greedy_lazy_kernel<<<1, LAZY_NT, 0, h->stream>>>(a);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    double result[8];
    ADSEL_CUDA(cudaMemcpyAsync(result, h->d_result.p, sizeof(result), cudaMemcpyDeviceToHost, h->stream));
    ADSEL_TRY(end_stage(h, 3));
    h->greedy_cluster = 1;
    h->greedy_variant = 3;   // lazy rows
    return finish_select(h, result, out_idx, out_m, out_cost, "adsel_select");
}

int lazy_preload() {
    cudaFuncAttributes attr;
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_lazy_kernel));
    return ADSEL_OK;
}

}  // namespace adsel
