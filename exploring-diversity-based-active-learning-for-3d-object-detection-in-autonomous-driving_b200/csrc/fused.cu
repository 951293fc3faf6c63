// fused.cu -- the greedy loop of the selectors that add a detector-feature term to the geodesic (and temporal) terms:
// det3d/selectors/spatial_feature_selector.py:186-237 and spatial_temporal_feature_selector.py:208-258.
//
// Row of frame i (what the reference materialises as N x N float64):
//   SpatialFeatureSelector          agg( 1 - exp(-geo[i,j]) ,  float64( 1 - expf(-feat[i,j]) ) )       agg = sum | min | max
//   SpatialTemporalFeatureSelector  (1 - exp(-geo[i,j])) + lambda_t (1 - exp(-T[i,j])) + float64( float32(lambda_f) * (1 - expf(-feat[i,j])) )
// The feature map leaves torch as FLOAT32, so its normalisation is numpy's float32 exp (np_exp_f32, common.cuh) and,
// by numpy's promotion rules, a python scalar times that float32 array is a float32 product; both are reproduced
// operation by operation.  The geodesic / temporal part is the sparse row of greedy.cu: the neighbour list of i
// (precombined with the temporal term) where the geodesic distance is below the saturation radius, else the far value
// combined with the temporal term.  The feature term is dense, so a step is one pass over the pool, like lazy.cu:
//   scatter i's list into a scratch row, then for every j: base (list value or far/temporal) (+) feature -> fps min,
//   arg-max with np.argmax's first-maximum rule.
// Quirk kept: the spatial+feature selector SEEDS the min-distance vector from the normalised SPATIAL rows only
// (spatial_feature_selector.py:201-209) and uses the aggregated map from the first update on (:217).
#include "common.cuh"

namespace adsel {

constexpr int FUSED_NT = 1024;
constexpr unsigned long long FUSED_EMPTY = 0xffffffffffffffffull;   // scratch entry not covered by the row's list

struct FusedArgs {
    int n, temporal, aggregate, seed_spatial_only, ld;
    double lambda_t, s_far, t_far;
    float lambda_f;
    double far_seed, far_row;   // initial fps: the largest value a seeding row / an aggregated row can take
    double cost_b, cost_f, cost_start, budget;
    int forbid_sampled, n_sampled, first_index, out_cap;
    long long list_capacity;
    const unsigned long long *list_cursor;
    const int64_t *list_off;
    const int32_t *list_len, *list_idx;
    const double *list_val;
    const int32_t *group_id, *n_boxes;
    const double *tlut;
    const float *fmap;              // [n, ld] feature distances, float32
    unsigned long long *fps, *scratch;
    unsigned char *is_selected;
    const int32_t *sampled;
    int32_t *out;
    double *result;
};

struct FusedShared {
    unsigned hi[FUSED_NT / 32], lo[FUSED_NT / 32];
    int idx[FUSED_NT / 32];
    int winner;
};

// fps = min(fps, row(src)); with_feature = false: the spatial(+temporal) part only (seeding quirk).  Returns the arg-max.
__device__ __forceinline__ int fused_update(const FusedArgs &a, FusedShared &fs, int src, bool with_feature, bool want_best) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long off = a.list_off[src];
    const int len = a.list_len[src];
    for (int t = tid; t < len; t += FUSED_NT)
        __stcg(&a.scratch[a.list_idx[off + t]], (unsigned long long)__double_as_longlong(a.list_val[off + t]));
    __syncthreads();
    const int gsrc = a.temporal ? a.group_id[src] : 0;
    const float *frow = a.fmap + (size_t)src * a.ld;
    unsigned bhi = 0u, blo = 0u;
    int bidx = 0x7fffffff;
    for (int j = tid; j < a.n; j += FUSED_NT) {
        const unsigned long long sb = __ldcg(&a.scratch[j]);   // written by other threads of the CTA: read at L2
        __stcg(&a.scratch[j], FUSED_EMPTY);
        double base;
        if (sb != FUSED_EMPTY) {
            base = __longlong_as_double((long long)sb);
        } else if (a.temporal) {
            const double tv = a.group_id[j] == gsrc ? a.tlut[abs(j - src)] : a.t_far;
            base = __dadd_rn(a.s_far, __dmul_rn(a.lambda_t, tv));
        } else {
            base = a.s_far;
        }
        double v = base;
        if (with_feature) {
            const float nf = __fsub_rn(1.0f, np_exp_f32(-frow[j]));    // 1 - np.exp(-feature), float32
            if (a.temporal) v = __dadd_rn(base, (double)__fmul_rn(a.lambda_f, nf));     // + lambda_f * feature (float32 product)
            else v = a.aggregate == ADSEL_AGG_SUM ? __dadd_rn(base, (double)nf) : a.aggregate == ADSEL_AGG_MIN ? fmin(base, (double)nf)
                                                                                                               : fmax(base, (double)nf);
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
            fs.hi[wid] = x.hi;
            fs.lo[wid] = x.lo;
            fs.idx[wid] = x.idx;
        }
    }
    __syncthreads();
    if (wid == 0) {
        unsigned hi[1] = {fs.hi[lane]}, lo[1] = {fs.lo[lane]};
        int idx[1] = {fs.idx[lane]};
        const Win x = warp_argmax<1>(hi, lo, idx);
        if (lane == 0) fs.winner = x.idx;
    }
    __syncthreads();
    return fs.winner;
}

__device__ __forceinline__ int fused_argmax_only(const FusedArgs &a, FusedShared &fs) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned bhi = 0u, blo = 0u;
    int bidx = 0x7fffffff;
    for (int j = tid; j < a.n; j += FUSED_NT) {
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
    if (lane == 0) {
        fs.hi[wid] = x.hi;
        fs.lo[wid] = x.lo;
        fs.idx[wid] = x.idx;
    }
    __syncthreads();
    if (wid == 0) {
        unsigned hi2[1] = {fs.hi[lane]}, lo2[1] = {fs.lo[lane]};
        int idx2[1] = {fs.idx[lane]};
        const Win y = warp_argmax<1>(hi2, lo2, idx2);
        if (lane == 0) fs.winner = y.idx;
    }
    __syncthreads();
    return fs.winner;
}

__global__ void __launch_bounds__(FUSED_NT, 1) fused_feature_kernel(const FusedArgs a) {
    __shared__ FusedShared fs;
    const int tid = threadIdx.x;
    int status = 0, offending = -1;
    const unsigned long long needed = a.list_cursor ? *a.list_cursor : 0ull;
    if ((long long)needed > a.list_capacity) {   // neighbour-list pool too small: the host grows it and reruns
        if (tid == 0) {
            a.result[1] = (double)ADSEL_ERR_CAPACITY;
            a.result[4] = -1.0;
            a.result[5] = (double)needed;
        }
        return;
    }
    // the first rows applied are seeding rows when a buffer is resumed or (spatial+feature) always
    const bool seed_spatial = a.seed_spatial_only != 0;
    const unsigned long long far_bits = (unsigned long long)__double_as_longlong(seed_spatial ? a.far_seed : a.far_row);
    for (int j = tid; j < a.n; j += FUSED_NT) {
        a.fps[j] = far_bits;
        a.scratch[j] = FUSED_EMPTY;
        a.is_selected[j] = 0;
    }
    __syncthreads();
    if (a.forbid_sampled)
        for (int t = tid; t < a.n_sampled; t += FUSED_NT) a.is_selected[a.sampled[t]] = 1;
    double cost = a.cost_start;
    int m = 0, cur = -1;
    long long steps = 0;
    if (a.n_sampled == 0) {   // empty buffer: the caller's random.choice, seeded with its (spatial) row (:207-209)
        cur = a.first_index;
        fused_update(a, fs, cur, !seed_spatial, false);
    } else {                  // fps = min over the (spatial) rows of the frames in the buffer, first pick = its arg-max
        for (int t = 0; t < a.n_sampled; ++t) fused_update(a, fs, a.sampled[t], !seed_spatial, false);
        cur = fused_argmax_only(a, fs);
    }
    if (tid == 0) {
        a.out[0] = cur;
        __stcg(&a.is_selected[cur], (unsigned char)1);
    }
    m = 1;
    cost = __dadd_rn(cost, a.cost_f);
    cost = __dadd_rn(cost, __dmul_rn((double)a.n_boxes[cur], a.cost_b));
    __syncthreads();
    while (true) {
        const int nxt = fused_update(a, fs, cur, true, true);
        ++steps;
        if (nxt < 0 || nxt >= a.n) {
            status = ADSEL_ERR_INVALID;
            break;
        }
        if (__ldcg(&a.is_selected[nxt])) {   // :222-223 / :245 assert
            status = ADSEL_ERR_RESELECTED;
            offending = nxt;
            break;
        }
        cost = __dadd_rn(cost, a.cost_f);
        cost = __dadd_rn(cost, __dmul_rn((double)a.n_boxes[nxt], a.cost_b));
        if (cost > a.budget) break;
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
        a.result[5] = (double)needed;
    }
}

__global__ void temporal_lut_kernel(double *tlut, int n, int norm_mode, double divisor);
int build_lists(adsel_handle *h, const adsel_select_params *p, double t_far, bool sharded, bool replicate);
int lists_prepare(adsel_handle *h, bool sharded);
int grow_lists(adsel_handle *h, long long needed);
int validate_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                    int32_t first_index, int32_t out_cap);
int finish_select(adsel_handle *h, const double *result, int32_t *out_idx, int32_t *out_m, double *out_cost, const char *who);

// p: spatial = geodesic, normalize = exp; temporal / lambda_t select the spatial+temporal+feature selector (sum), else
// p->aggregate combines the spatial and the feature term.  The feature map must have been built (adsel_feature_distance_build).
int fused_feature_select(adsel_handle *h, const adsel_select_params *p, double lambda_f, int32_t seed_spatial_only,
                         const int32_t *sampled, int32_t n_sampled, int32_t first_index, int32_t *out_idx, int32_t out_cap,
                         int32_t *out_m, double *out_cost) {
    ADSEL_TRY(validate_select(h, p, sampled, n_sampled, first_index, out_cap));
    ADSEL_REQUIRE(out_idx && out_m, "adsel_select_fused_feature: null argument");
    ADSEL_REQUIRE(p->spatial == ADSEL_SPATIAL_GEODESIC && p->normalize == ADSEL_NORM_EXP,
                  "adsel_select_fused_feature: the fused selectors use the exp-normalised geodesic term");
    ADSEL_REQUIRE(!p->temporal || p->aggregate == ADSEL_AGG_SUM, "adsel_select_fused_feature: spatial+temporal+feature is a sum");
    ADSEL_REQUIRE(h->graph_ready, "adsel_select_fused_feature: call adsel_build_knn_graph first");
    ADSEL_REQUIRE(h->pd_n == h->n && h->d2_p && h->feat.p, "adsel_select_fused_feature: call adsel_feature_distance_build first");
    const int n = h->n;
    ADSEL_TRY(lists_prepare(h, false));
    ADSEL_TRY(h->fps.reserve(n));
    ADSEL_TRY(h->pd_fps.reserve(2 * (size_t)n));   // scratch row (8 bytes per frame)
    ADSEL_TRY(h->is_selected.reserve((size_t)n * 16));
    ADSEL_TRY(h->tlut.reserve(n));
    ADSEL_TRY(h->d_sampled.reserve(n_sampled > 0 ? n_sampled : 1));
    ADSEL_TRY(h->d_out.reserve(out_cap));
    ADSEL_TRY(h->d_result.reserve(8));
    for (int attempt = 0; attempt < 2; ++attempt) {
        ADSEL_TRY(begin_stage(h));
        if (p->temporal) {
            temporal_lut_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->tlut.p, n, p->normalize, p->max_temporal);
            h->launches += 1;
        }
        // list values: 1 - exp(-geo), already summed with lambda_t * temporal for the three-term selector
        ADSEL_TRY(build_lists(h, p, 1.0, false, false));
        ADSEL_TRY(end_stage(h, 2));
        ADSEL_TRY(begin_stage(h));
        FusedArgs a;
        memset(&a, 0, sizeof(a));
        a.n = n;
        a.temporal = p->temporal ? 1 : 0;
        a.aggregate = p->aggregate;
        a.seed_spatial_only = seed_spatial_only;
        a.ld = h->pd_ld;
        a.lambda_t = p->lambda_t;
        a.lambda_f = (float)lambda_f;
        a.s_far = 1.0;     // 1 - exp(-d) == 1.0 for d >= 38
        a.t_far = 1.0;
        const double base_far = p->temporal ? 1.0 + p->lambda_t * 1.0 : 1.0;
        a.far_seed = base_far;
        // an aggregated row entry never exceeds base_far (+) the largest feature term, 1.0 (float32 1 - exp(-x) <= 1)
        a.far_row = p->temporal ? base_far + (double)(a.lambda_f * 1.0f)
                                : (p->aggregate == ADSEL_AGG_SUM ? 2.0 : 1.0);
        a.cost_b = p->cost_b;
        a.cost_f = p->cost_f;
        a.cost_start = p->cost_start;
        a.budget = p->budget;
        a.forbid_sampled = p->forbid_sampled;
        a.n_sampled = n_sampled;
        a.first_index = first_index;
        a.out_cap = out_cap;
        a.list_capacity = h->lists.capacity;
        a.list_cursor = h->lists.cursor;
        a.list_off = h->lists.off;
        a.list_len = h->lists.len;
        a.list_idx = h->lists.idx;
        a.list_val = h->lists.val;
        a.group_id = h->group_id.p;
        a.n_boxes = h->n_boxes.p;
        a.tlut = h->tlut.p;
        a.fmap = reinterpret_cast<const float *>(h->d2_p);
        a.fps = h->fps.p;
        a.scratch = reinterpret_cast<unsigned long long *>(h->pd_fps.p);
        a.is_selected = h->is_selected.p;
        a.sampled = h->d_sampled.p;
        a.out = h->d_out.p;
        a.result = h->d_result.p;
        if (n_sampled > 0)
            ADSEL_CUDA(cudaMemcpyAsync(h->d_sampled.p, sampled, sizeof(int32_t) * n_sampled, cudaMemcpyDefault, h->stream));   // host or device list
        ADSEL_CUDA(cudaMemsetAsync(h->d_result.p, 0, sizeof(double) * 8, h->stream));
        fused_feature_kernel<<<1, FUSED_NT, 0, h->stream>>>(a);
        h->launches += 1;
        ADSEL_CUDA(cudaGetLastError());
        double result[8];
        ADSEL_CUDA(cudaMemcpyAsync(result, h->d_result.p, sizeof(result), cudaMemcpyDeviceToHost, h->stream));
        ADSEL_TRY(end_stage(h, 7));
        if ((int)result[1] == ADSEL_ERR_CAPACITY && result[4] == -1.0 && attempt == 0) {
            ADSEL_TRY(grow_lists(h, (long long)result[5]));
            continue;
        }
        h->greedy_variant = 4;
        h->greedy_cluster = 1;
        return finish_select(h, result, out_idx, out_m, out_cost, "adsel_select_fused_feature");
    }
    return ADSEL_ERR_CAPACITY;
}

}  // namespace adsel
