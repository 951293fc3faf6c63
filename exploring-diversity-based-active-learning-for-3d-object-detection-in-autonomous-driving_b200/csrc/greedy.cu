// greedy.cu -- the cost-budgeted greedy farthest-point (k-center) loop,
// det3d/selectors/spatial_temporal_selector.py:157-193 (identical in spatial_selector.py:101-138,
// temporal_selector.py:68-104, euclidean_spatial_selector.py:106-143).
//
// Reference step:  fps = min(fps, distance_map[idx]);  selected = int(np.argmax(fps))  (first index
// among equal maxima);  cost += cost_f;  cost += n_boxes*cost_b;  stop when cost > int(budget).
//
// distance_map[idx] differs from one constant `far` (the largest value a row can take) only on the
// "touched" set: idx's geodesic neighbour list and idx's temporal group.  Since every fps entry is
// already <= far, min(fps, row) only changes touched entries, so a step is a sparse update plus an
// arg-reduce over per-256-frame block maxima (recomputed only for dirtied blocks).  One CTA runs
// the whole loop: steps are strictly sequential and a step is a few microseconds, so a grid-wide
// barrier per step would cost more than the work.
//
// All values are non-negative doubles, handled as ordered u64 keys so that atomicMin / max are
// exact and "first index among bit-equal maxima" is well defined.
#include "common.cuh"

namespace adsel {

constexpr int GREEDY_THREADS = 1024;
constexpr int BLK = 256;            // frames per block-max entry
constexpr int MAX_BLKS = 16384;     // n <= 4,194,304

struct GreedyArgs {
    int n, nblk;
    int spatial, temporal, aggregate;
    double lambda_t, s_far, t_far;
    double cost_b, cost_f, cost_start, budget;
    int forbid_sampled, n_sampled, first_index, out_cap;
    const double2 *loc;
    const int32_t *group_id, *group_ptr, *group_members, *n_boxes;
    const int64_t *list_off;
    const int32_t *list_len, *list_idx;
    const double *list_val, *tlut;
    unsigned long long *fps, *blk_val;
    int32_t *blk_idx;
    unsigned char *is_selected;
    const int32_t *sampled;
    int32_t *out;
    double *result;   // [0] cost, [1] status, [2] m, [3] steps
};

struct Best {
    unsigned long long v;
    int i;
};
__device__ __forceinline__ Best better(Best a, Best b) {
    // larger value wins; bit-equal values resolve to the lower index (np.argmax returns the first maximum)
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = better(x, y);
    }
    return x;
}

__device__ __forceinline__ double combine(const GreedyArgs &a, double s, double t) {
    if (a.spatial != ADSEL_SPATIAL_NONE && a.temporal) return aggregate_value(s, t, a.aggregate, a.lambda_t);
    return a.spatial != ADSEL_SPATIAL_NONE ? s : t;
}

// fps[j] = min(fps[j], distance_map[src][j]) for every touched j.
__device__ __forceinline__ void apply_row(const GreedyArgs &a, int src, unsigned char *dirty) {
    const int g = a.group_id[src];
    if (a.spatial == ADSEL_SPATIAL_GEODESIC) {
        const int64_t off = a.list_off[src];
        const int len = a.list_len[src];
        for (int t = threadIdx.x; t < len; t += GREEDY_THREADS) {
            const int j = a.list_idx[off + t];
            const double s = a.list_val[off + t];
            double tv = 0.0;
            if (a.temporal) tv = a.group_id[j] == g ? a.tlut[abs(j - src)] : a.t_far;
            const double val = combine(a, s, tv);
            atomicMin(&a.fps[j], (unsigned long long)__double_as_longlong(val));
            dirty[j / BLK] = 1;
        }
    }
    if (a.temporal || a.spatial == ADSEL_SPATIAL_EUCLID) {
        // group members beyond the geodesic radius keep the far spatial value; members inside it were
        // already lowered above by a smaller-or-equal value (aggregate is monotone in s), so the
        // second, larger candidate is harmless.
        const int e0 = a.group_ptr[g], e1 = a.group_ptr[g + 1];
        const double2 me = a.loc[src];
        for (int t = e0 + threadIdx.x; t < e1; t += GREEDY_THREADS) {
            const int j = a.group_members[t];
            double s = a.s_far;
            if (a.spatial == ADSEL_SPATIAL_EUCLID) {   // euclidean_spatial_selector.py:96
                const double2 c = a.loc[j];
                const double dx = __dsub_rn(c.x, me.x), dy = __dsub_rn(c.y, me.y);
                s = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            }
            const double tv = a.temporal ? a.tlut[abs(j - src)] : 0.0;
            const double val = combine(a, s, tv);
            atomicMin(&a.fps[j], (unsigned long long)__double_as_longlong(val));
            dirty[j / BLK] = 1;
        }
    }
}

__global__ void __launch_bounds__(GREEDY_THREADS, 1) greedy_select_kernel(const GreedyArgs a) {
    __shared__ unsigned char dirty[MAX_BLKS];
    __shared__ Best warp_res[GREEDY_THREADS / 32];
    __shared__ int s_cur, s_stop, s_m;
    __shared__ double s_cost;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NWARPS = GREEDY_THREADS / 32;

    for (int b = threadIdx.x; b < a.nblk; b += GREEDY_THREADS) dirty[b] = 1;
    if (threadIdx.x == 0) {
        s_stop = 0;
        s_m = 0;
        s_cur = -1;
        s_cost = a.cost_start;
        if (a.n_sampled == 0) {   // empty buffer: the caller's random.choice is taken and costed up front (:168,172-174)
            s_cur = a.first_index;
            a.out[0] = s_cur;
            s_m = 1;
            a.is_selected[s_cur] = 1;
            s_cost = __dadd_rn(s_cost, a.cost_f);
            s_cost = __dadd_rn(s_cost, __dmul_rn((double)a.n_boxes[s_cur], a.cost_b));
        }
    }
    __syncthreads();
    if (a.forbid_sampled)
        for (int t = threadIdx.x; t < a.n_sampled; t += GREEDY_THREADS) a.is_selected[a.sampled[t]] = 1;
    for (int t = 0; t < a.n_sampled; ++t) {   // fps = min over the rows of the frames already in the buffer (:160-164)
        apply_row(a, a.sampled[t], dirty);
    }
    __syncthreads();
    long long steps = 0;
    bool seeding = a.n_sampled > 0;   // first pick = argmax without a preceding update (:165)
    while (true) {
        if (!seeding) {
            apply_row(a, s_cur, dirty);
            __syncthreads();
        }
        // refresh the maxima of dirtied blocks
        for (int b = wid; b < a.nblk; b += NWARPS) {
            if (!dirty[b]) continue;
            Best x{0ull, 0x7fffffff};
            const int j0 = b * BLK;
#pragma unroll
            for (int i = 0; i < BLK / 32; ++i) {
                const int j = j0 + lane + 32 * i;
                if (j < a.n) {
                    const unsigned long long v = __ldcg(&a.fps[j]);
                    if (v > x.v || x.i == 0x7fffffff) {
                        x.v = v;
                        x.i = j;
                    }
                }
            }
            x = warp_best(x);
            if (lane == 0) {
                a.blk_val[b] = x.v;
                a.blk_idx[b] = x.i;
                dirty[b] = 0;
            }
        }
        __syncthreads();
        // arg-reduce over block maxima
        Best x{0ull, 0x7fffffff};
        for (int b = threadIdx.x; b < a.nblk; b += GREEDY_THREADS) x = better(x, Best{a.blk_val[b], a.blk_idx[b]});
        x = warp_best(x);
        if (lane == 0) warp_res[wid] = x;
        __syncthreads();
        if (wid == 0) {
            Best y = lane < NWARPS ? warp_res[lane] : Best{0ull, 0x7fffffff};
            y = warp_best(y);
            if (lane == 0) {
                const int nxt = y.i;
                if (seeding) {   // :165,172-174 -- always taken, no budget check
                    a.out[0] = nxt;
                    s_m = 1;
                    a.is_selected[nxt] = 1;
                    s_cost = __dadd_rn(s_cost, a.cost_f);
                    s_cost = __dadd_rn(s_cost, __dmul_rn((double)a.n_boxes[nxt], a.cost_b));
                    s_cur = nxt;
                } else if (a.is_selected[nxt]) {   // :182 assert
                    a.result[1] = (double)ADSEL_ERR_RESELECTED;
                    a.result[4] = (double)nxt;
                    s_stop = 1;
                } else {
                    s_cost = __dadd_rn(s_cost, a.cost_f);                                                   // :184
                    s_cost = __dadd_rn(s_cost, __dmul_rn((double)a.n_boxes[nxt], a.cost_b));                // :185
                    if (s_cost > a.budget) {                                                                // :186
                        s_stop = 1;
                    } else if (s_m >= a.out_cap) {
                        a.result[1] = (double)ADSEL_ERR_CAPACITY;
                        s_stop = 1;
                    } else {
                        a.out[s_m] = nxt;
                        s_m += 1;
                        a.is_selected[nxt] = 1;
                        s_cur = nxt;
                    }
                }
            }
        }
        __syncthreads();
        ++steps;
        if (seeding) {
            seeding = false;
            continue;
        }
        if (s_stop) break;
    }
    if (threadIdx.x == 0) {
        a.result[0] = s_cost;
        a.result[2] = (double)s_m;
        a.result[3] = (double)steps;
    }
}

__global__ void temporal_lut_kernel(double *tlut, int n, int norm_mode, double divisor) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) tlut[d] = normalize_value((double)d, norm_mode, divisor);
}

__global__ void fill_u64_kernel2(unsigned long long *p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

static double host_normalize_far(int norm_mode, double d, double divisor) {
    // value of the normalisation at a "far" distance: exp -> exactly 1.0 (exp(-d) < 2^-54 for d >= 38),
    // linear -> d / divisor (IEEE division, identical on host and device), raw -> d.
    if (norm_mode == ADSEL_NORM_EXP) return 1.0;
    if (norm_mode == ADSEL_NORM_LINEAR) return d / divisor;
    return d;
}

int build_lists(adsel_handle *h, int norm_mode, double max_spatial);

int greedy_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                  int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_REQUIRE(h->n > 0, "adsel_select: call adsel_set_pool first");
    ADSEL_REQUIRE(p && out_idx && out_m && out_cap > 0, "adsel_select: null argument");
    ADSEL_REQUIRE(p->spatial >= 0 && p->spatial <= 2 && p->normalize >= 0 && p->normalize <= 2 && p->aggregate >= 0 &&
                      p->aggregate <= 2,
                  "adsel_select: bad mode");
    ADSEL_REQUIRE(p->spatial != ADSEL_SPATIAL_NONE || p->temporal, "adsel_select: no distance term");
    ADSEL_REQUIRE(!(p->spatial != ADSEL_SPATIAL_NONE && p->temporal && p->aggregate == ADSEL_AGG_SUM && !(p->lambda_t >= 0.0)),
                  "adsel_select: lambda_t must be >= 0 (row maximum argument, DESIGN.md §5)");
    ADSEL_REQUIRE(n_sampled >= 0 && (n_sampled == 0 || sampled), "adsel_select: bad sampled list");
    ADSEL_REQUIRE(n_sampled > 0 || (first_index >= 0 && first_index < h->n), "adsel_select: first_index out of range");
    for (int i = 0; i < n_sampled; ++i) ADSEL_REQUIRE(sampled[i] >= 0 && sampled[i] < h->n, "adsel_select: sampled index out of range");
    const int n = h->n;
    const int nblk = (n + BLK - 1) / BLK;
    ADSEL_REQUIRE(nblk <= MAX_BLKS, "adsel_select: pool larger than 4,194,304 frames");
    if (p->normalize == ADSEL_NORM_LINEAR) {
        ADSEL_REQUIRE(p->spatial != ADSEL_SPATIAL_GEODESIC || p->max_spatial > 0.0, "adsel_select: linear needs max_spatial");
        ADSEL_REQUIRE(!p->temporal || p->max_temporal > 0.0, "adsel_select: linear needs max_temporal");
    }
    if (p->spatial == ADSEL_SPATIAL_GEODESIC) ADSEL_TRY(build_lists(h, p->normalize, p->max_spatial));

    ADSEL_TRY(h->fps.reserve(n));
    ADSEL_TRY(h->blk_val.reserve(nblk));
    ADSEL_TRY(h->blk_idx.reserve(nblk));
    ADSEL_TRY(h->is_selected.reserve(n));
    ADSEL_TRY(h->tlut.reserve(n));
    ADSEL_TRY(h->d_sampled.reserve(n_sampled > 0 ? n_sampled : 1));
    ADSEL_TRY(h->d_out.reserve(out_cap));
    ADSEL_TRY(h->d_result.reserve(8));

    GreedyArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    a.nblk = nblk;
    a.spatial = p->spatial;
    a.temporal = p->temporal ? 1 : 0;
    a.aggregate = p->aggregate;
    a.lambda_t = p->lambda_t;
    const double inf = host_inf();
    // the value every untouched entry of a row takes
    a.s_far = p->spatial == ADSEL_SPATIAL_GEODESIC ? host_normalize_far(p->normalize, inf, p->max_spatial)
              : p->spatial == ADSEL_SPATIAL_EUCLID ? TEMPORAL_MARGIN : 0.0;
    a.t_far = a.temporal ? host_normalize_far(p->normalize, TEMPORAL_MARGIN, p->max_temporal) : 0.0;
    double far;
    if (a.spatial != ADSEL_SPATIAL_NONE && a.temporal) {
        if (a.aggregate == ADSEL_AGG_SUM) far = a.s_far + a.lambda_t * a.t_far;   // unfused on the host (no -ffast-math)
        else if (a.aggregate == ADSEL_AGG_MIN) far = a.s_far < a.t_far ? a.s_far : a.t_far;
        else far = a.s_far > a.t_far ? a.s_far : a.t_far;
    } else {
        far = a.spatial != ADSEL_SPATIAL_NONE ? a.s_far : a.t_far;
    }
    unsigned long long far_bits;
    memcpy(&far_bits, &far, 8);
    a.cost_b = p->cost_b;
    a.cost_f = p->cost_f;
    a.cost_start = p->cost_start;
    a.budget = p->budget;
    a.forbid_sampled = p->forbid_sampled;
    a.n_sampled = n_sampled;
    a.first_index = first_index;
    a.out_cap = out_cap;
    a.loc = h->loc.p;
    a.group_id = h->group_id.p;
    a.group_ptr = h->group_ptr.p;
    a.group_members = h->group_members.p;
    a.n_boxes = h->n_boxes.p;
    a.list_off = h->list_off.p;
    a.list_len = h->list_len.p;
    a.list_idx = h->list_idx.p;
    a.list_val = h->list_val.p;
    a.tlut = h->tlut.p;
    a.fps = h->fps.p;
    a.blk_val = h->blk_val.p;
    a.blk_idx = h->blk_idx.p;
    a.is_selected = h->is_selected.p;
    a.sampled = h->d_sampled.p;
    a.out = h->d_out.p;
    a.result = h->d_result.p;

    ADSEL_TRY(begin_stage(h));
    if (n_sampled > 0)
        ADSEL_CUDA(cudaMemcpyAsync(h->d_sampled.p, sampled, sizeof(int32_t) * n_sampled, cudaMemcpyHostToDevice, h->stream));
    ADSEL_CUDA(cudaMemsetAsync(h->is_selected.p, 0, n, h->stream));
    ADSEL_CUDA(cudaMemsetAsync(h->d_result.p, 0, sizeof(double) * 8, h->stream));
    fill_u64_kernel2<<<h->num_sms * 2, 256, 0, h->stream>>>(h->fps.p, (size_t)n, far_bits);
    h->launches += 1;
    if (a.temporal) {
        temporal_lut_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->tlut.p, n, p->normalize, p->max_temporal);
        h->launches += 1;
    }
    ADSEL_CUDA(cudaGetLastError());
    greedy_select_kernel<<<1, GREEDY_THREADS, 0, h->stream>>>(a);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    double result[8];
    ADSEL_CUDA(cudaMemcpyAsync(result, h->d_result.p, sizeof(result), cudaMemcpyDeviceToHost, h->stream));
    ADSEL_TRY(end_stage(h, 3));
    int m = (int)result[2];
    const int status = (int)result[1];
    h->last_steps = (int64_t)result[3];
    if (m > 0) ADSEL_CUDA(cudaMemcpy(out_idx, h->d_out.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost));
    *out_m = m;
    if (out_cost) *out_cost = result[0];
    if (status == ADSEL_ERR_RESELECTED) {
        set_error("id: " + std::to_string((int)result[4]) + " has been selected");
        return ADSEL_ERR_RESELECTED;
    }
    if (status == ADSEL_ERR_CAPACITY) {
        set_error("adsel_select: out_idx capacity exhausted");
        return ADSEL_ERR_CAPACITY;
    }
    return ADSEL_OK;
}

}  // namespace adsel
