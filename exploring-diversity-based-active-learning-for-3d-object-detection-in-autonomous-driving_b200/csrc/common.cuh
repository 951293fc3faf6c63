// common.cuh -- shared declarations of the sm_100a selection library (see include/adsel_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <limits>
#include <string>

#include "../../include/adsel_b200.h"

namespace adsel {

void set_error(const std::string &msg);

#define ADSEL_CUDA(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            adsel::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +        \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
            return ADSEL_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

#define ADSEL_REQUIRE(cond, msg)                                                                \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            adsel::set_error(std::string(msg));                                                 \
            return ADSEL_ERR_INVALID;                                                           \
        }                                                                                       \
    } while (0)

#define ADSEL_TRY(expr)                                                                         \
    do {                                                                                        \
        int _s = (expr);                                                                        \
        if (_s != ADSEL_OK) return _s;                                                          \
    } while (0)

constexpr unsigned long long INF_BITS = 0x7ff0000000000000ull;   // +inf as an ordered u64 key
constexpr double TEMPORAL_MARGIN = 1e6;                            // spatial_temporal_selector.py:112
constexpr double NULL_EDGE_ATOL = 1e-8;                            // scipy csgraph_from_dense / masked_values

// Device buffer with explicit capacity (grown geometrically; never shrinks).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return ADSEL_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        ADSEL_CUDA(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
        return ADSEL_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace adsel

namespace adsel {
// Everything the greedy kernel needs about the frame it just selected, fetched in one round trip.
// Two self-tagged 16-byte halves: a copy that is being overwritten asynchronously (cp.async into the greedy
// kernel's shared-memory cache) is valid exactly when both tags name the frame that is looked up.
struct __align__(16) FrameRec {
    long long list_off;   // neighbour list of this frame (entries inside this rank's column slice)
    int list_len;
    unsigned self;        // the frame this record describes
    int e0, e1;           // its temporal group as a range of group_members
    int n_boxes;          // box cost = n_boxes * cost_b, rounded once (the reference's per-frame product)
    unsigned self2;
};

// Where one rank's neighbour lists live: its own buffers, or the lists region of its arena (sharded, so that
// the ranks that run the searches can append to it over NVLink).
struct ListStore {
    int64_t *off;                  // [n] first entry of every source's list
    int32_t *len;                  // [n]
    int32_t *idx;                  // [capacity] column (frame) of the entry
    double *val;                   // [capacity] final row value distance_map[src][idx]
    unsigned long long *cursor;    // [0] entries requested so far (may exceed capacity: overflow)
    long long capacity;
};

// Parameters the precombined neighbour-list values depend on (lists are rebuilt when any changes).
struct ListsKey {
    int norm_mode = -1, temporal = 0, aggregate = 0, lo = 0, hi = 0, pad = 0;
    double max_spatial = 0.0, max_temporal = 0.0, lambda_t = 0.0;
};
}  // namespace adsel

struct adsel_handle {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t timer_ev[2] = {nullptr, nullptr};
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t launches = 0, last_steps = 0;

    // pool
    int32_t n = 0;
    adsel::DevBuf<double2> loc;
    adsel::DevBuf<int32_t> group_id, n_boxes;
    adsel::DevBuf<int32_t> group_ptr, group_members;   // CSR of frames by group (ascending index)
    int32_t n_groups = 0, max_group = 0;                  // number of groups, size of the largest

    // kNN + CSR graph
    int32_t k = 0;
    adsel::DevBuf<int32_t> knn_idx;
    adsel::DevBuf<double> knn_dist;
    adsel::DevBuf<int32_t> row_ptr, col, deg_own, deg_rev;
    adsel::DevBuf<double> w;
    int64_t nnz = 0;
    bool graph_ready = false;
    double bbox[4] = {0, 0, 0, 0};               // xmin, ymin, xmax, ymax of the ego locations
    adsel::DevBuf<int32_t> grid_cell, grid_idx, grid_start;   // uniform grid of the kNN search
    adsel::DevBuf<double2> grid_loc;

    // SSSP scratch (per resident CTA)
    int32_t sssp_ctas = 0;
    adsel::DevBuf<unsigned long long> sc_dist;
    adsel::DevBuf<int32_t> sc_stamp, sc_q0, sc_q1, sc_touched;
    size_t sc_n = 0;

    // neighbour lists (truncated geodesic neighbourhood of every frame, normalised)
    adsel::DevBuf<int64_t> list_off;
    adsel::DevBuf<int32_t> list_len, list_idx;
    adsel::DevBuf<double> list_val;
    adsel::DevBuf<unsigned long long> cursor;   // [0] entries reserved, [1] max finite bits
    int64_t list_total = 0;
    long long list_capacity = 0;                 // capacity of the own buffers
    adsel::ListStore lists = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};   // store in use (own buffers or arena)
    long long arena_list_cap = 0;                // entries of the arena's list region
    int lists_mode = -1;                         // normalize mode the lists were built for (-1: none)
    adsel::ListsKey lists_key;

    // BEV descriptors
    int32_t bev_frames = 0, bev_cells = 0, bev_batch = 0;   // pool frames, cells per frame, frames of the resident point batch
    int64_t bev_points = 0;
    adsel::DevBuf<float> bev_pts;
    adsel::DevBuf<int64_t> bev_off;
    adsel::DevBuf<int32_t> bev_counts, bev_norm2;
    adsel::DevBuf<uint16_t> bev_levels;          // bf16 bits
    uint16_t *desc_levels = nullptr;             // descriptors in use: own buffers, or the arena's copy (sharded)
    int32_t *desc_norm2 = nullptr;

    // pairwise squared distances of the descriptors (dense path)
    int32_t pd_n = 0, pd_ld = 0, pd_col0 = 0, pd_cols = 0;   // pool frames, row pitch, first / number of local columns
    adsel::DevBuf<int32_t> pd_d2, pd_norm, pd_fps, pd_boxes;
    int32_t *d2_p = nullptr;                     // distance matrix in use: pd_d2, or the arena's slice (sharded: peers store mirrored tiles into it)
    adsel::DevBuf<float> feat;                   // detector features [n, dim] (feature-distance map)
    adsel::DevBuf<int2> pd_tiles;                // (row block, column block) of every distance tile, in visiting order

    // greedy state
    adsel::DevBuf<unsigned long long> fps;       // ordered-u64 view of non-negative doubles
    adsel::DevBuf<adsel::FrameRec> frame_rec;    // per-frame record read once per greedy step
    adsel::DevBuf<unsigned char> slots;          // per-frame record + first entries of its neighbour list in one fixed-size slot (greedy_warp.cu)
    adsel::DevBuf<unsigned char> is_selected;
    adsel::DevBuf<double> tlut;
    adsel::DevBuf<int32_t> d_sampled, d_out;
    adsel::DevBuf<double> d_result;              // [0] cost, [1] status, [2] m, [3] steps, [4] offending frame
    int greedy_cluster = 0;                      // cluster size of the last greedy launch (0: global-memory variant)
    int greedy_variant = 0;                      // kernel of the last selection: 0 global-memory, 1 fast resident CTA, 2 resident cluster, 3 lazy rows, 4 fused feature, 5 one-warp resident
    bool groups_contiguous = false;              // every group is one contiguous index range (consecutive runs)

    // pool sharding across ranks (comm.cu): this rank owns frames [shard_lo, shard_hi)
    int32_t rank = 0, world = 1;
    int32_t shard_lo = 0, shard_hi = 0;
    unsigned char *arena = nullptr;              // peer-visible allocation: mailbox | kNN tables | descriptors
    size_t arena_bytes = 0;
    int32_t arena_n = 0, arena_k1 = 0, arena_dim = 0;
    bool pd_pending_timing = false;              // knn_ev[0..1] bracket an enqueued sharded distance build
    bool knn_pending_timing = false;             // ev[0..1] bracket an enqueued sharded graph build
    unsigned char *peer_arena[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool peer_ipc[8] = {false, false, false, false, false, false, false, false};   // opened with cudaIpcOpenMemHandle
    bool comm_ready = false;
    unsigned comm_seq = 0;                       // tag of the last completed exchange
    unsigned long long comm_barrier_seq = 0;
    double comm_timeout_s = 20.0;
    int32_t *knn_idx_p = nullptr;                // kNN table in use: own buffers or the arena's (sharded build)
    double *knn_dist_p = nullptr;
    // pending sharded selection (begin/end split so several ranks can be driven from one host thread)
    bool sampled_is_device = false;              // adsel_select_device: the `sampled` argument is a device pointer
    bool pending = false;
    bool pending_replicated = false;             // the pending selection runs the whole greedy loop on this rank
    int32_t pending_cap = 0;
    cudaEvent_t pending_ev[2] = {nullptr, nullptr};
    cudaEvent_t knn_ev[2] = {nullptr, nullptr};
};

namespace adsel {

static inline double host_inf() { return std::numeric_limits<double>::infinity(); }
int begin_stage(adsel_handle *h);
int end_stage(adsel_handle *h, int stage);

// ---------------------------------------------------------------------------------------------
// numpy float64 exp on AVX512_SKX hosts == Intel SVML __svml_exp8_ha main path (oracle/npexp.c
// documents how the sequence and the constants were obtained).  Bit-exact by construction: the
// same IEEE operations in the same order (one round-toward-zero FMA, then round-to-nearest FMAs).
// ---------------------------------------------------------------------------------------------
static __device__ __constant__ unsigned long long EXP_TH[16] = {
    0x3ff0000000000000ull, 0x3ff0b5586cf9890full, 0x3ff172b83c7d517bull, 0x3ff2387a6e756238ull,
    0x3ff306fe0a31b715ull, 0x3ff3dea64c123422ull, 0x3ff4bfdad5362a27ull, 0x3ff5ab07dd485429ull,
    0x3ff6a09e667f3bcdull, 0x3ff7a11473eb0187ull, 0x3ff8ace5422aa0dbull, 0x3ff9c49182a3f090ull,
    0x3ffae89f995ad3adull, 0x3ffc199bdd85529cull, 0x3ffd5818dcfba487ull, 0x3ffea4afa2a490daull};
static __device__ __constant__ unsigned long long EXP_TL[16] = {
    0x0000000000000000ull, 0x3c979aa65d837b6dull, 0xbc801b15eaa59348ull, 0x3c968efde3a8a894ull,
    0x3c834d754db0abb6ull, 0x3c859f48a72a4c6dull, 0x3c7690cebb7aafb0ull, 0x3c9063e1e21c5409ull,
    0xbc93b3efbf5e2228ull, 0xbc7b32dcb94da51dull, 0x3c8db72fc1f0eab4ull, 0x3c71affc2b91ce27ull,
    0x3c8c1a7792cb3387ull, 0x3c736eae30af0cb3ull, 0x3c74a385a63d07a7ull, 0xbc8ff7128fd391f0ull};

__device__ __forceinline__ double np_exp_f64(double x) {
    const double L2E = 0x1.71547652b82fep+0, SHIFTER = 0x1.8000000003ff0p+48;
    const double L2H = 0x1.62e42fefa39efp-1, L2L = 0x1.abc9e3b39803fp-56;
    const double C6 = 0x1.7411836940c04p-10, C5 = 0x1.1101cbbc265c0p-7, C4 = 0x1.55557242d68fep-5,
                 C3 = 0x1.5555553939732p-3, C2 = 0x1.000000000d008p-1, C1 = 0x1.fffffffffff70p-1;
    if (!(fabs(x) < 0x1.61da04cbafe44p+9)) return exp(x);   // SVML's scalar rare path: not bit-pinned
    const double z = __fma_rz(x, L2E, SHIFTER);
    const double N = __dsub_rn(z, SHIFTER);
    const unsigned j = (unsigned)(__double_as_longlong(z) & 0xf);
    double r = __fma_rn(-N, L2H, x);
    r = __fma_rn(-L2L, N, r);
    const double r2 = __dmul_rn(r, r);
    const double p65 = __fma_rn(C6, r, C5);
    const double p43 = __fma_rn(C4, r, C3);
    const double p21 = __fma_rn(C2, r, C1);
    double P = __fma_rn(r2, p65, p43);
    P = __fma_rn(r2, P, p21);
    const double th = __longlong_as_double(EXP_TH[j]), tl = __longlong_as_double(EXP_TL[j]);
    double y = __fma_rn(P, r, tl);
    y = __fma_rn(th, y, th);
    return ldexp(y, (int)floor(N));
}

// numpy's float32 exp (its own SIMD kernel, simd_exp_FLOAT: float32 does not go through SVML), bit for bit on every
// float32 input: the normalisation of the float32 feature-distance map in the fused feature selectors
// (det3d/selectors/spatial_feature_selector.py:190).  Restated and pinned in oracle/npexp.c::oracle_npexpf.
__device__ __forceinline__ float np_exp_f32(float x) {
    const float xmax = 88.72283935546875f, xmin = -103.97208404541015625f;
    const float log2e = 1.44269504088896341f, magic = 0x1.8p+23f;
    const float c1 = -6.93145752e-1f, c2 = -1.42860677e-6f;
    const float p0 = 9.999999999980870924916e-01f, p1 = 7.257664613233124478488e-01f, p2 = 2.473615434895520810817e-01f,
                p3 = 5.114512081637298353406e-02f, p4 = 6.757896990527504603057e-03f, p5 = 5.082762527590693718096e-04f;
    const float q0 = 1.0f, q1 = -2.742335390411667452936e-01f, q2 = 2.159509375685829852307e-02f;
    if (x != x) return x;
    if (x > xmax) return __int_as_float(0x7f800000);
    if (x < xmin) return 0.0f;
    float q = __fmul_rn(x, log2e);
    q = __fsub_rn(__fadd_rn(q, magic), magic);          // rint, ties to even
    float r = __fmaf_rn(q, c1, x);
    r = __fmaf_rn(q, c2, r);
    float num = __fmaf_rn(p5, r, p4);
    num = __fmaf_rn(num, r, p3);
    num = __fmaf_rn(num, r, p2);
    num = __fmaf_rn(num, r, p1);
    num = __fmaf_rn(num, r, p0);
    float den = __fmaf_rn(q2, r, q1);
    den = __fmaf_rn(den, r, q0);
    const float y = __fdiv_rn(num, den);                 // in (0.5, 2)
    // scalef(y, q), rounded once: powers of two are built from their exponent bits; a result below 2^-126 is
    // reached through an exact intermediate so that only the final multiplication rounds
    const int e = (int)q;                                // -150 .. 128
    if (e >= -126) {
        float v = __fmul_rn(y, __int_as_float((min(e, 127) + 127) << 23));
        if (e > 127) v = __fmul_rn(v, 2.0f);
        return v;
    }
    return __fmul_rn(__fmul_rn(y, __int_as_float((e + 64 + 127) << 23)), 0x1p-64f);
}

// Normalisation of one raw distance d >= 0 (spatial_temporal_selector.py:138-146).
__device__ __forceinline__ double normalize_value(double d, int mode, double divisor) {
    if (mode == ADSEL_NORM_EXP) return __dsub_rn(1.0, np_exp_f64(-d));
    if (mode == ADSEL_NORM_LINEAR) return __ddiv_rn(d, divisor);
    return d;
}

// Aggregate spatial s and temporal t (spatial_temporal_selector.py:148-155): sum is
// s + lambda_t*t as a rounded multiply followed by a rounded add (numpy evaluates it unfused).
__device__ __forceinline__ double aggregate_value(double s, double t, int agg, double lambda_t) {
    if (agg == ADSEL_AGG_SUM) return __dadd_rn(s, __dmul_rn(lambda_t, t));
    if (agg == ADSEL_AGG_MIN) return fmin(s, t);
    return fmax(s, t);
}

// Row value from its spatial and temporal terms (one-term selectors pass the other term through).
__device__ __forceinline__ double combine(int spatial, int temporal, int aggregate, double lambda_t, double s, double t) {
    if (spatial != ADSEL_SPATIAL_NONE && temporal) return aggregate_value(s, t, aggregate, lambda_t);
    return spatial != ADSEL_SPATIAL_NONE ? s : t;
}


// Arguments of the greedy kernels (greedy.cu, greedy_warp.cu).
struct GreedyArgs {
    int n, nblk, lo, hi, blk_lo, blk_hi, tl_n, sel_words;
    int csize, cta_cols, r_nblk;   // resident variant: CTAs per cluster, columns and blocks per CTA
    int spatial, temporal, aggregate, groups_contiguous, pull;
    double lambda_t, s_far, t_far;
    double cost_b, cost_f, cost_start, budget;
    int forbid_sampled, n_sampled, first_index, out_cap;
    long long list_capacity;
    unsigned long long far_bits;             // ordered key of the value every untouched row entry takes
    const unsigned long long *list_cursor;   // [0] entries the list build needed (overflow check)
    const FrameRec *rec;
    const double2 *loc;
    const int32_t *group_members;
    const int32_t *list_idx;
    const double *list_val, *tlut;
    unsigned long long *fps;
    unsigned char *is_selected;
    const int32_t *sampled;
    int32_t *out;
    double *result;   // [0] cost, [1] status, [2] m, [3] steps, [4] offending frame, [5] list entries needed
    unsigned long long *prof;   // PROFILE builds: clock64 cycles of thread 0 per phase
};

struct Win {
    unsigned hi, lo;
    int idx;
};

// Warp-wide arg-max over CNT (hi, lo, idx) candidates per lane: largest 64-bit key, lowest idx among equal keys.
template <int CNT>
__device__ __forceinline__ Win warp_argmax(const unsigned (&hi)[CNT], const unsigned (&lo)[CNT], const int (&idx)[CNT]) {
    const unsigned full = 0xffffffffu;
    unsigned h = hi[0];
#pragma unroll
    for (int i = 1; i < CNT; ++i) h = max(h, hi[i]);
    const unsigned mhi = __reduce_max_sync(full, h);
    unsigned l = 0u;
#pragma unroll
    for (int i = 0; i < CNT; ++i) l = max(l, hi[i] == mhi ? lo[i] : 0u);
    const unsigned mlo = __reduce_max_sync(full, l);
    int c = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < CNT; ++i) c = min(c, (hi[i] == mhi && lo[i] == mlo) ? idx[i] : 0x7fffffff);
    Win w;
    w.hi = mhi;
    w.lo = mlo;
    w.idx = __reduce_min_sync(full, c);
    return w;
}


}  // namespace adsel
