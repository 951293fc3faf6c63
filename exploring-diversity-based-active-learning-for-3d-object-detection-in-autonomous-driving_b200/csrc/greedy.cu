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
// arg-reduce over per-128-frame block maxima (recomputed only for dirtied blocks).  One CTA runs
// the whole loop: steps are strictly sequential and a step is about a microsecond, so a grid-wide
// barrier per step would cost more than the work.
//
// A step is a chain of dependent L2 round trips, so the kernel is organised to need as few as
// possible: (1) lane 0 fetches the selected frame's 32-byte FrameRec (list range, group range, box
// cost) -- an L1 hit thanks to the prefetch issued when its block was last refreshed; the selected-frame
// flags are a bitmap in shared memory; (2) the first two warps load the precombined (index, value)
// list entries while the others compute the group entries from a shared-memory LUT; (3) atomicMin
// into fps; (4) the dirtied 128-frame blocks are re-read, one warp per block, tree + REDUX arg-max.
// Block maxima, dirty flags and the dirty list live in shared memory; warp 0 alone does the final
// arg-reduce and the cost bookkeeping, so a step has three CTA barriers.
//
// SHARDED = true: the rank owns columns [lo, hi) only (its lists hold only those columns), and the
// step's local maximum is exchanged with the other ranks through peer-memory mailboxes (comm.cuh)
// before the bookkeeping, which every rank then performs identically.
//
// All values are non-negative doubles, handled as ordered u64 keys so that atomicMin / max are
// exact and "first index among bit-equal maxima" is well defined.
#include <stdlib.h>

#include <cooperative_groups.h>

#include "comm.cuh"
#include "common.cuh"

namespace adsel {

#ifndef ADSEL_BLK
#define ADSEL_BLK 128
#endif
constexpr int BLK = ADSEL_BLK;        // frames per block-max entry
constexpr int MAX_BLKS = 9800;        // 20 B of shared memory per block + 32 KB LUT: n <= 1,254,400 frames

struct GreedyArgs {
    int n, nblk, lo, hi, blk_lo, blk_hi, tl_n, sel_words;
    int csize, cta_cols;   // cluster variant: CTAs per cluster, columns per CTA (multiple of 256)
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

struct Best {
    unsigned long long v;
    int i;
};
__device__ __forceinline__ Best better(Best a, Best b) {
    // larger value wins; bit-equal values resolve to the lower index (np.argmax returns the first maximum)
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
// Warp-wide arg-max with np.argmax's tie rule in three REDUX operations (max of the high word, max of
// the low word among the lanes holding that high word, min index among the lanes holding both) instead
// of five shuffle rounds of 64-bit compare/select chains: on the single CTA that runs the greedy loop
// every dependent instruction of this reduction is exposed latency.
__device__ __forceinline__ Best warp_best(Best x) {
    const unsigned full = 0xffffffffu;
    const unsigned hi = (unsigned)(x.v >> 32), lo = (unsigned)x.v;
    const unsigned mhi = __reduce_max_sync(full, hi);
    const unsigned cand = __ballot_sync(full, hi == mhi);
    Best r;
    if ((cand & (cand - 1)) == 0) {   // a single lane holds the largest high word: fetch its record with two shuffles
        const int srcl = __ffs(cand) - 1;
        r.v = ((unsigned long long)mhi << 32) | __shfl_sync(full, lo, srcl);
        r.i = __shfl_sync(full, x.i, srcl);
        return r;
    }
    const unsigned mlo = __reduce_max_sync(full, hi == mhi ? lo : 0u);
    const int mi = __reduce_min_sync(full, (hi == mhi && lo == mlo) ? x.i : 0x7fffffff);
    r.v = ((unsigned long long)mhi << 32) | mlo;
    r.i = mi;
    return r;
}

__device__ __forceinline__ double combine(int spatial, int temporal, int aggregate, double lambda_t, double s, double t) {
    if (spatial != ADSEL_SPATIAL_NONE && temporal) return aggregate_value(s, t, aggregate, lambda_t);
    return spatial != ADSEL_SPATIAL_NONE ? s : t;
}

struct GreedyShared {
    FrameRec rec;
    double2 me;
    double cost;
    int cur, stop, m, ndirty;
};

__device__ __forceinline__ void mark_dirty(int *dirty, int *dirty_list, int *ndirty, int b) {
    if (((volatile int *)dirty)[b] == 0 && atomicExch(&dirty[b], 1) == 0) dirty_list[atomicAdd(ndirty, 1)] = b;
}

// Clipped temporal-group range of the row being applied, when groups are contiguous index ranges.
__device__ __forceinline__ void group_range(const GreedyArgs &a, const GreedyShared &sh, int &g0, int &g1) {
    g0 = g1 = 0;
    if (a.temporal || a.spatial == ADSEL_SPATIAL_EUCLID) {
        g0 = sh.rec.e0;
        g1 = sh.rec.e1;
        if (a.groups_contiguous) {   // group_members is the identity: the group is frames [e0, e1); clip to the slice
            g0 = max(g0, a.lo);
            g1 = max(min(g1, a.hi), g0);
        }
    }
}

// Value of distance_map[src][j] for a member j of src's group that lies beyond the geodesic radius.
__device__ __forceinline__ double group_value(const GreedyArgs &a, const GreedyShared &sh, const double *tl, int tl_n,
                                              int src, int j) {
    double s = a.s_far;
    if (a.spatial == ADSEL_SPATIAL_EUCLID) {   // euclidean_spatial_selector.py:96
        const double2 c = a.loc[j];
        const double dx = __dsub_rn(c.x, sh.me.x), dy = __dsub_rn(c.y, sh.me.y);
        s = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
    double tv = 0.0;
    if (a.temporal) {
        const int d = abs(j - src);
        tv = d < tl_n ? tl[d] : a.tlut[d];
    }
    return combine(a.spatial, a.temporal, a.aggregate, a.lambda_t, s, tv);
}

// First half of fps = min(fps, distance_map[src]) on this rank's slice (src's record is in sh):
//   * neighbour-list entries (sparse, precombined values): atomicMin, their blocks marked dirty;
//   * the temporal group: members beyond the geodesic radius keep the far spatial value; members inside
//     it are lowered from the list by a smaller-or-equal value (aggregate is monotone in s), so applying
//     both candidates is harmless.  When groups are contiguous index ranges only the group's BLOCKS are
//     marked here -- the values are folded into the block refresh that re-reads those blocks anyway
//     ("pull": no atomics, no second round trip); otherwise members are pushed with atomicMin.
template <int NT>
__device__ __forceinline__ void apply_row(const GreedyArgs &a, GreedyShared &sh, const double *tl, int tl_n, int *dirty,
                                          int *dirty_list) {
    const int src = sh.cur;
    const long long off = sh.rec.list_off;
    const int len = a.spatial == ADSEL_SPATIAL_GEODESIC ? sh.rec.list_len : 0;
    int g0, g1;
    group_range(a, sh, g0, g1);
    // Warp specialisation: the first LIST_WARPS warps walk the neighbour list (one dependent L2 round trip),
    // the others the temporal group (shared-memory LUT only), so the two phases overlap instead of chaining.
    constexpr int LIST_THREADS = NT >= 512 ? 64 : 32;
    const bool has_group = g1 > g0;
    if (threadIdx.x < LIST_THREADS || !has_group) {
        const int stride = has_group ? LIST_THREADS : NT;
        for (int t = threadIdx.x; t < len; t += stride) {
            const int j = a.list_idx[off + t];
            const double v = a.list_val[off + t];
            atomicMin(&a.fps[j], (unsigned long long)__double_as_longlong(v));
            mark_dirty(dirty, dirty_list, &sh.ndirty, j / BLK);
        }
        if (!has_group) return;
    }
    if (threadIdx.x >= LIST_THREADS) {
        const int gt = threadIdx.x - LIST_THREADS;
        constexpr int GROUP_THREADS = NT - LIST_THREADS;
        if (a.groups_contiguous) {
            if (a.pull) {
                const int b0 = g0 / BLK, b1 = (g1 - 1) / BLK;
                for (int b = b0 + gt; b <= b1; b += GROUP_THREADS) mark_dirty(dirty, dirty_list, &sh.ndirty, b);
                return;
            }
            // the group is the frame range [g0, g1): its blocks are known without looking at the entries
            const int b0 = g0 / BLK, b1 = (g1 - 1) / BLK;
            if (gt <= b1 - b0) mark_dirty(dirty, dirty_list, &sh.ndirty, b0 + gt);
            for (int j = g0 + gt; j < g1; j += GROUP_THREADS) {
                const double v = group_value(a, sh, tl, tl_n, src, j);
                atomicMin(&a.fps[j], (unsigned long long)__double_as_longlong(v));
            }
        } else {
            for (int t = g0 + gt; t < g1; t += GROUP_THREADS) {
                const int j = a.group_members[t];
                if (j >= a.lo && j < a.hi) {
                    const double v = group_value(a, sh, tl, tl_n, src, j);
                    atomicMin(&a.fps[j], (unsigned long long)__double_as_longlong(v));
                    mark_dirty(dirty, dirty_list, &sh.ndirty, j / BLK);
                }
            }
        }
    }
}

// Second half: re-read the dirtied BLK-frame blocks, fold in the pulled group values of the row being
// applied (`pull`), and recompute the block maxima.  One warp per block.
template <int NT>
__device__ __forceinline__ void refresh_blocks(const GreedyArgs &a, GreedyShared &sh, const double *tl, int tl_n,
                                               unsigned long long *blk_val, int *blk_idx, int *dirty, int *dirty_list,
                                               bool pull) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    const int nd = sh.ndirty;
    int g0 = 0, g1 = 0;
    if (pull && a.groups_contiguous && a.pull) group_range(a, sh, g0, g1);
    const int src = sh.cur;
    for (int q = wid; q < nd; q += NWARPS) {
        const int b = dirty_list[q];
        const int j0 = b * BLK;
        unsigned long long v[BLK / 32];
#pragma unroll
        for (int i = 0; i < BLK / 32; ++i) {
            const int j = j0 + lane + 32 * i;
            v[i] = (j >= a.lo && j < a.hi) ? __ldcg(&a.fps[j]) : 0ull;
        }
        Best e[BLK / 32];
#pragma unroll
        for (int i = 0; i < BLK / 32; ++i) {
            const int j = j0 + lane + 32 * i;
            e[i] = Best{0ull, 0x7fffffff};
            if (j >= a.lo && j < a.hi) {
                if (j >= g0 && j < g1) {
                    const unsigned long long c = (unsigned long long)__double_as_longlong(group_value(a, sh, tl, tl_n, src, j));
                    if (c < v[i]) {
                        v[i] = c;
                        __stcg(&a.fps[j], c);
                    }
                }
                e[i] = Best{v[i], j};
            }
        }
        // pairwise tree (depth 3) rather than a chain of 8 dependent compare/selects
#pragma unroll
        for (int w = 1; w < BLK / 32; w <<= 1)
#pragma unroll
            for (int i = 0; i + w < BLK / 32; i += 2 * w) e[i] = better(e[i], e[i + w]);
        Best x = e[0];
        x = warp_best(x);
        if (lane == 0) {
            blk_val[b] = x.v;
            blk_idx[b] = x.i;
            dirty[b] = 0;
            // if this block's best frame wins a later step, its record is then an L1 hit for lane 0
            if (x.i != 0x7fffffff) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.rec + x.i));
        }
    }
}

// Selected-frame flags: a bitmap in shared memory when it fits (no global round trip in the bookkeeping),
// else bytes in global memory.
__device__ __forceinline__ void set_selected(const GreedyArgs &a, unsigned *selbits, int idx) {
    if (a.sel_words) atomicOr(&selbits[idx >> 5], 1u << (idx & 31));
    else a.is_selected[idx] = 1;
}
__device__ __forceinline__ bool get_selected(const GreedyArgs &a, const unsigned *selbits, int idx) {
    if (a.sel_words) return (selbits[idx >> 5] >> (idx & 31)) & 1u;
    return a.is_selected[idx] != 0;
}

// lane 0 of warp 0: everything the next step needs about frame `idx`, in one round trip
__device__ __forceinline__ void fetch_frame(const GreedyArgs &a, GreedyShared &sh, int idx) {
    sh.rec = a.rec[idx];
    if (a.spatial == ADSEL_SPATIAL_EUCLID) sh.me = a.loc[idx];
    sh.cur = idx;
}

#define PROF_MARK(slot)                                   \
    if (PROFILE && threadIdx.x == 0) {                     \
        const int dep = *(volatile int *)&sh.stop;   /* consume shared memory: a deferred barrier blocks here */ \
        long long now;                                    \
        asm volatile("{\n.reg .b32 t;\nmov.b32 t, %1;\nmov.u64 %0, %%clock64;\n}" : "=l"(now) : "r"(dep) : "memory"); \
        prof_acc[slot] += (unsigned long long)(now - prof_t); \
        prof_t = now;                                     \
    }

template <int NT, bool SHARDED, bool PROFILE = false>
__global__ void __launch_bounds__(NT, 1) greedy_select_kernel(const GreedyArgs a, const CommView comm) {
    unsigned long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = PROFILE ? clock64() : 0;
    extern __shared__ unsigned long long greedy_smem[];
    unsigned long long *blk_val = greedy_smem;                      // [nblk]
    double *tl = reinterpret_cast<double *>(blk_val + a.nblk);      // [tl_n] head of the temporal LUT
    int *blk_idx = reinterpret_cast<int *>(tl + a.tl_n);            // [nblk]
    int *dirty = blk_idx + a.nblk;                                  // [nblk]
    int *dirty_list = dirty + a.nblk;                               // [nblk]
    unsigned *selbits = reinterpret_cast<unsigned *>(dirty_list + a.nblk);   // [sel_words] selected-frame bitmap (0 words: global bytes)
    for (int wd = threadIdx.x; wd < a.sel_words; wd += NT) selbits[wd] = 0u;
    const int tl_n = a.tl_n;
    for (int d = threadIdx.x; d < tl_n; d += NT) tl[d] = a.tlut[d];
    __shared__ GreedyShared sh;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ExchangeLane xlane;
    if (SHARDED) xlane = make_exchange_lane(comm);

    for (int b = threadIdx.x; b < a.nblk; b += NT) {
        const bool mine = b >= a.blk_lo && b < a.blk_hi;
        dirty[b] = mine ? 1 : 0;
        if (mine) dirty_list[b - a.blk_lo] = b;
        blk_val[b] = 0ull;
        blk_idx[b] = 0x7fffffff;
    }
    __syncthreads();   // bitmap zeroed before the first frame is flagged
    bool local_abort = false;
    if (threadIdx.x == 0) {
        sh.stop = 0;
        sh.m = 0;
        sh.cur = -1;
        sh.cost = a.cost_start;
        sh.ndirty = a.blk_hi - a.blk_lo;
        const unsigned long long needed = a.list_cursor ? *a.list_cursor : 0ull;
        a.result[5] = (double)needed;
        if ((long long)needed > a.list_capacity) {   // neighbour-list pool too small: the host grows it and reruns
            local_abort = true;
            a.result[1] = (double)ADSEL_ERR_CAPACITY;
            a.result[4] = -1.0;
            sh.stop = SHARDED ? 3 : 1;   // 3: skip the updates, tell the peers in the first exchange
        } else if (a.n_sampled == 0) {   // empty buffer: the caller's random.choice is taken and costed up front (:168,172-174)
            const int first = a.first_index;
            a.out[0] = first;
            sh.m = 1;
            set_selected(a, selbits, first);
            sh.cost = __dadd_rn(sh.cost, a.cost_f);
            fetch_frame(a, sh, first);
            sh.cost = __dadd_rn(sh.cost, sh.rec.box_cost);
        }
    }
    __syncthreads();
    if (!SHARDED && sh.stop) return;
    if (a.forbid_sampled)
        for (int t = threadIdx.x; t < a.n_sampled; t += NT) set_selected(a, selbits, a.sampled[t]);
    if (sh.stop == 0) {
        for (int t = 0; t < a.n_sampled; ++t) {   // fps = min over the rows of the frames already in the buffer (:160-164)
            if (threadIdx.x == 0) fetch_frame(a, sh, a.sampled[t]);
            __syncthreads();
            apply_row<NT>(a, sh, tl, tl_n, dirty, dirty_list);
            __syncthreads();
            if (a.groups_contiguous && a.pull) {   // pulled group values are folded in by the refresh, row by row
                refresh_blocks<NT>(a, sh, tl, tl_n, blk_val, blk_idx, dirty, dirty_list, true);
                __syncthreads();
                if (threadIdx.x == 0) sh.ndirty = 0;
            }
        }
    }
    long long steps = 0;
    bool seeding = a.n_sampled > 0;   // first pick = argmax without a preceding update (:165)
    while (true) {
        PROF_MARK(7)
        const bool applied = !seeding && sh.stop == 0;
        if (applied) apply_row<NT>(a, sh, tl, tl_n, dirty, dirty_list);
        PROF_MARK(0)
        __syncthreads();
        PROF_MARK(1)
        refresh_blocks<NT>(a, sh, tl, tl_n, blk_val, blk_idx, dirty, dirty_list, applied);
        PROF_MARK(2)
        __syncthreads();
        PROF_MARK(3)
        if (wid == 0) {
            // arg-reduce over this rank's block maxima
            Best y{0ull, 0x7fffffff};
            PROF_MARK(8)   // calibration: cost of one mark
            // 8 block maxima per lane at a time: independent shared-memory loads, then a depth-3 tree
            // (a serial compare/select chain over them would be exposed latency on this lone warp)
            for (int b0 = a.blk_lo + lane; b0 < a.blk_hi; b0 += 8 * 32) {
                Best e[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int b = b0 + 32 * q;
                    e[q] = b < a.blk_hi ? Best{blk_val[b], blk_idx[b]} : Best{0ull, 0x7fffffff};
                }
#pragma unroll
                for (int w = 1; w < 8; w <<= 1)
#pragma unroll
                    for (int q = 0; q + w < 8; q += 2 * w) e[q] = better(e[q], e[q + w]);
                y = better(y, e[0]);
            }
            PROF_MARK(9)
            y = warp_best(y);
            PROF_MARK(10)
            int comm_status = 0;
            if (SHARDED) {
                const bool abort_now = __shfl_sync(0xffffffffu, local_abort ? 1 : 0, 0) != 0;
                const ExchangeResult r = exchange_best(comm, xlane, comm.seq_base + 1u + (unsigned)steps, y.v,
                                                       abort_now ? COMM_ABORT_INDEX : (unsigned)y.i);
                y.v = r.key;
                y.i = r.index;
                comm_status = r.status;
            }
            PROF_MARK(4)
            if (lane == 0) {
                sh.ndirty = 0;
                const int nxt = y.i;
                if (comm_status != 0) {
                    if (a.result[1] == 0.0) a.result[1] = (double)(comm_status == 2 ? ADSEL_ERR_COMM : ADSEL_ERR_PEER_ABORT);
                    sh.stop = 2;
                } else if (nxt < 0 || nxt >= a.n) {   // empty pool slice everywhere: cannot happen with n > 0
                    a.result[1] = (double)ADSEL_ERR_INVALID;
                    sh.stop = 2;
                } else {
                    const FrameRec rec = a.rec[nxt];
                    const bool was_selected = get_selected(a, selbits, nxt);
                    const double2 me = a.spatial == ADSEL_SPATIAL_EUCLID ? a.loc[nxt] : make_double2(0.0, 0.0);
                    if (seeding) {   // :165,172-174 -- always taken, no budget check
                        a.out[0] = nxt;
                        sh.m = 1;
                        set_selected(a, selbits, nxt);
                        sh.cost = __dadd_rn(sh.cost, a.cost_f);
                        sh.cost = __dadd_rn(sh.cost, rec.box_cost);
                        sh.rec = rec;
                        sh.me = me;
                        sh.cur = nxt;
                    } else if (was_selected) {   // :182 assert
                        a.result[1] = (double)ADSEL_ERR_RESELECTED;
                        a.result[4] = (double)nxt;
                        sh.stop = 1;
                    } else {
                        sh.cost = __dadd_rn(sh.cost, a.cost_f);          // :184
                        sh.cost = __dadd_rn(sh.cost, rec.box_cost);      // :185 (n_boxes*cost_b, one rounded product)
                        if (sh.cost > a.budget) {                                                    // :186
                            sh.stop = 1;
                        } else if (sh.m >= a.out_cap) {
                            a.result[1] = (double)ADSEL_ERR_CAPACITY;
                            a.result[4] = -2.0;
                            sh.stop = 1;
                        } else {
                            a.out[sh.m] = nxt;
                            sh.m += 1;
                            set_selected(a, selbits, nxt);
                            sh.rec = rec;
                            sh.me = me;
                            sh.cur = nxt;
                        }
                    }
                }
            }
        }
        PROF_MARK(5)
        __syncthreads();
        PROF_MARK(6)
        ++steps;
        if (sh.stop == 2) break;
        if (seeding) {
            seeding = false;
            continue;
        }
        if (sh.stop) break;
    }
    if (threadIdx.x == 0) {
        a.result[0] = sh.cost;
        a.result[2] = (double)sh.m;
        a.result[3] = (double)steps;
        if (PROFILE)
            for (int i = 0; i < 12; ++i) a.prof[i] = prof_acc[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Cluster variant: fps lives in SHARED memory, split over the CTAs of one thread-block cluster.
//
// In the single-CTA kernel above the min-distance vector lives in L2 and every step pays an
// atomicMin (RED) round trip followed by a re-read of the dirtied blocks -- the longest link of the
// dependency chain.  Here the rank's columns are cut into `csize` contiguous pieces, one per CTA of a
// cluster (csize = 2 for the 28,130-frame pool: 110 KB of shared memory each), so that
//   * a row update is a plain shared-memory min (every touched column is written by exactly one thread:
//     list entries are unique per list, the group is applied in a second phase),
//   * the arg-max is a scan of the CTA's own shared-memory slice (no dirty tracking, no block maxima),
//   * the CTAs exchange their local maxima through distributed shared memory with the mailbox protocol of
//     comm.cuh (tagged self-validating 8-byte words, st.shared::cluster + poll; no cluster barrier in the
//     loop, so L1 -- which holds the prefetched frame records and the lists -- is never flushed),
//   * every CTA does the cost bookkeeping redundantly; cluster rank 0 writes the results and, when the
//     pool is sharded over GPUs, posts the rank's record to the other ranks while every CTA polls the
//     rank's mailbox.
// ---------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

constexpr int CLUSTER_MAX = 8;                  // portable cluster size
constexpr int XSLOT_WORDS = 4;

__device__ __forceinline__ unsigned long long ld_volatile_shared(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Warp-level exchange of the CTAs' local maxima through distributed shared memory.  xslot: this CTA's
// [2][CLUSTER_MAX][XSLOT_WORDS] slots; remote[r]: the same array in CTA r's shared memory.
__device__ __forceinline__ Best cluster_exchange(unsigned long long *xslot, unsigned long long *const *remote, int crank,
                                                 int csize, unsigned tag, Best local) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int peer = lane / 3, word = lane - 3 * peer;
    const size_t parity_off = (size_t)(tag & 1u) * CLUSTER_MAX * XSLOT_WORDS;
    const bool active = lane < 3 * csize;
    if (active) {
        const unsigned long long payload = word == 0 ? (local.v >> 32) : word == 1 ? (local.v & 0xffffffffull)
                                                                                  : (unsigned long long)(unsigned)local.i;
        remote[peer][parity_off + (size_t)crank * XSLOT_WORDS + word] = (payload << 32) | tag;   // st.shared::cluster
    }
    unsigned long long got = 0;
    if (active) {
        const unsigned long long *p = xslot + parity_off + (size_t)peer * XSLOT_WORDS + word;
        while (true) {
            got = ld_volatile_shared(p);
            if ((unsigned)got == tag) break;
        }
    }
    const int src = lane < csize ? 3 * lane : 0;
    const unsigned long long hi = __shfl_sync(full, got, src) >> 32;
    const unsigned long long lo = __shfl_sync(full, got, src + 1) >> 32;
    const unsigned idx = (unsigned)(__shfl_sync(full, got, src + 2) >> 32);
    Best x{0ull, 0x7fffffff};
    if (lane < csize) {
        x.v = (hi << 32) | lo;
        x.i = (int)idx;
    }
    return warp_best(x);
}

template <int NT, bool SHARDED>
__global__ void __launch_bounds__(NT, 1) greedy_cluster_kernel(const GreedyArgs a, const CommView comm) {
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank(), csize = a.csize;
    const int c_lo = min(a.lo + crank * a.cta_cols, a.hi), c_hi = min(c_lo + a.cta_cols, a.hi);
    const int c_len = c_hi - c_lo;
    extern __shared__ unsigned long long greedy_smem[];
    unsigned long long *fps_s = greedy_smem;                                       // [cta_cols]
    double *tl = reinterpret_cast<double *>(fps_s + a.cta_cols);                   // [tl_n]
    unsigned long long *xslot = reinterpret_cast<unsigned long long *>(tl + a.tl_n);   // [2][CLUSTER_MAX][XSLOT_WORDS]
    unsigned *selbits = reinterpret_cast<unsigned *>(xslot + 2 * CLUSTER_MAX * XSLOT_WORDS);   // [sel_words]
    __shared__ GreedyShared sh;
    __shared__ Best wres[NT / 32];
    __shared__ unsigned long long *remote[CLUSTER_MAX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NWARPS = NT / 32;
    const int tl_n = a.tl_n;
    ExchangeLane xlane;
    if (SHARDED) xlane = make_exchange_lane(comm);

    const unsigned long long far_bits = a.far_bits;
    for (int i = threadIdx.x; i < a.cta_cols; i += NT) fps_s[i] = far_bits;
    for (int d = threadIdx.x; d < tl_n; d += NT) tl[d] = a.tlut[d];
    for (int wd = threadIdx.x; wd < a.sel_words; wd += NT) selbits[wd] = 0u;
    for (int i = threadIdx.x; i < 2 * CLUSTER_MAX * XSLOT_WORDS; i += NT) xslot[i] = 0ull;
    if (threadIdx.x < CLUSTER_MAX)
        remote[threadIdx.x] = threadIdx.x < csize ? cluster.map_shared_rank(xslot, threadIdx.x) : nullptr;
    __syncthreads();
    bool local_abort = false;
    if (threadIdx.x == 0) {
        sh.stop = 0;
        sh.m = 0;
        sh.cur = -1;
        sh.cost = a.cost_start;
        const unsigned long long needed = a.list_cursor ? *a.list_cursor : 0ull;
        if (crank == 0) a.result[5] = (double)needed;
        if ((long long)needed > a.list_capacity) {   // neighbour-list pool too small: the host grows it and reruns
            local_abort = true;
            if (crank == 0) {
                a.result[1] = (double)ADSEL_ERR_CAPACITY;
                a.result[4] = -1.0;
            }
            sh.stop = SHARDED ? 3 : 1;
        } else if (a.n_sampled == 0) {
            const int first = a.first_index;
            if (crank == 0) a.out[0] = first;
            sh.m = 1;
            set_selected(a, selbits, first);
            sh.cost = __dadd_rn(sh.cost, a.cost_f);
            fetch_frame(a, sh, first);
            sh.cost = __dadd_rn(sh.cost, sh.rec.box_cost);
        }
    }
    cluster.sync();   // every CTA's slots are zeroed and mapped before anyone posts
    if (!SHARDED && sh.stop) {
        cluster.sync();
        return;
    }
    if (a.forbid_sampled)
        for (int t = threadIdx.x; t < a.n_sampled; t += NT) set_selected(a, selbits, a.sampled[t]);

    // one row update on this CTA's slice: list entries, barrier, group members
    auto apply = [&]() {
        const int src = sh.cur;
        const long long off = sh.rec.list_off;
        const int len = a.spatial == ADSEL_SPATIAL_GEODESIC ? sh.rec.list_len : 0;
        for (int t = threadIdx.x; t < len; t += NT) {
            const int j = a.list_idx[off + t];
            const unsigned long long v = (unsigned long long)__double_as_longlong(a.list_val[off + t]);   // same round trip
            if (j >= c_lo && j < c_hi && v < fps_s[j - c_lo]) fps_s[j - c_lo] = v;
        }
        __syncthreads();
        int g0 = 0, g1 = 0;
        if (a.temporal || a.spatial == ADSEL_SPATIAL_EUCLID) {
            g0 = sh.rec.e0;
            g1 = sh.rec.e1;
        }
        if (a.groups_contiguous) {
            g0 = max(g0, c_lo);
            g1 = min(g1, c_hi);
            for (int j = g0 + threadIdx.x; j < g1; j += NT) {
                const unsigned long long v = (unsigned long long)__double_as_longlong(group_value(a, sh, tl, tl_n, src, j));
                if (v < fps_s[j - c_lo]) fps_s[j - c_lo] = v;
            }
        } else {
            for (int t = g0 + threadIdx.x; t < g1; t += NT) {
                const int j = a.group_members[t];
                if (j >= c_lo && j < c_hi) {
                    const unsigned long long v = (unsigned long long)__double_as_longlong(group_value(a, sh, tl, tl_n, src, j));
                    if (v < fps_s[j - c_lo]) fps_s[j - c_lo] = v;
                }
            }
        }
    };

    if (sh.stop == 0) {
        for (int t = 0; t < a.n_sampled; ++t) {   // fps = min over the rows of the frames already in the buffer (:160-164)
            __syncthreads();
            if (threadIdx.x == 0) fetch_frame(a, sh, a.sampled[t]);
            __syncthreads();
            apply();
        }
    }
    long long steps = 0;
    bool seeding = a.n_sampled > 0;
    while (true) {
        if (!seeding && sh.stop == 0) apply();
        __syncthreads();
        // arg-max of this CTA's slice, first index among equal maxima
        Best acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = Best{0ull, 0x7fffffff};
        for (int i0 = threadIdx.x; i0 < c_len; i0 += 4 * NT) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * NT;
                if (i < c_len) {
                    const unsigned long long v = fps_s[i];
                    if (v > acc[q].v || acc[q].i == 0x7fffffff) {   // ascending i per accumulator: strict '>' keeps the first
                        acc[q].v = v;
                        acc[q].i = c_lo + i;
                    }
                }
            }
        }
        Best x = better(better(acc[0], acc[1]), better(acc[2], acc[3]));
        x = warp_best(x);
        if (lane == 0) wres[wid] = x;
        __syncthreads();
        if (wid == 0) {
            Best y = lane < NWARPS ? wres[lane] : Best{0ull, 0x7fffffff};
            y = warp_best(y);
            if (y.i != 0x7fffffff && lane == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.rec + y.i));
            const unsigned tag = 1u + (unsigned)steps;
            if (csize > 1) y = cluster_exchange(xslot, remote, crank, csize, tag, y);
            int comm_status = 0;
            if (SHARDED) {
                const bool abort_now = __shfl_sync(0xffffffffu, local_abort ? 1 : 0, 0) != 0;
                const ExchangeResult r = exchange_best(comm, xlane, comm.seq_base + tag, y.v,
                                                       abort_now ? COMM_ABORT_INDEX : (unsigned)y.i, crank == 0);
                y.v = r.key;
                y.i = r.index;
                comm_status = r.status;
            }
            if (lane == 0) {
                const int nxt = y.i;
                if (comm_status != 0) {
                    if (crank == 0 && a.result[1] == 0.0)
                        a.result[1] = (double)(comm_status == 2 ? ADSEL_ERR_COMM : ADSEL_ERR_PEER_ABORT);
                    sh.stop = 2;
                } else if (nxt < 0 || nxt >= a.n) {
                    if (crank == 0) a.result[1] = (double)ADSEL_ERR_INVALID;
                    sh.stop = 2;
                } else {
                    const FrameRec rec = a.rec[nxt];
                    const bool was_selected = get_selected(a, selbits, nxt);
                    const double2 me = a.spatial == ADSEL_SPATIAL_EUCLID ? a.loc[nxt] : make_double2(0.0, 0.0);
                    if (seeding) {
                        if (crank == 0) a.out[0] = nxt;
                        sh.m = 1;
                        set_selected(a, selbits, nxt);
                        sh.cost = __dadd_rn(sh.cost, a.cost_f);
                        sh.cost = __dadd_rn(sh.cost, rec.box_cost);
                        sh.rec = rec;
                        sh.me = me;
                        sh.cur = nxt;
                    } else if (was_selected) {   // :182 assert
                        if (crank == 0) {
                            a.result[1] = (double)ADSEL_ERR_RESELECTED;
                            a.result[4] = (double)nxt;
                        }
                        sh.stop = 1;
                    } else {
                        sh.cost = __dadd_rn(sh.cost, a.cost_f);          // :184
                        sh.cost = __dadd_rn(sh.cost, rec.box_cost);      // :185
                        if (sh.cost > a.budget) {                        // :186
                            sh.stop = 1;
                        } else if (sh.m >= a.out_cap) {
                            if (crank == 0) {
                                a.result[1] = (double)ADSEL_ERR_CAPACITY;
                                a.result[4] = -2.0;
                            }
                            sh.stop = 1;
                        } else {
                            if (crank == 0) a.out[sh.m] = nxt;
                            sh.m += 1;
                            set_selected(a, selbits, nxt);
                            sh.rec = rec;
                            sh.me = me;
                            sh.cur = nxt;
                        }
                    }
                }
            }
        }
        __syncthreads();
        ++steps;
        if (sh.stop == 2) break;
        if (seeding) {
            seeding = false;
            continue;
        }
        if (sh.stop) break;
    }
    for (int i = threadIdx.x; i < c_len; i += NT) a.fps[c_lo + i] = fps_s[i];   // adsel_get_fps
    if (threadIdx.x == 0 && crank == 0) {
        a.result[0] = sh.cost;
        a.result[2] = (double)sh.m;
        a.result[3] = (double)steps;
    }
    cluster.sync();   // no CTA leaves while a peer may still store into its slots
}

__global__ void temporal_lut_kernel(double *tlut, int n, int norm_mode, double divisor) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) tlut[d] = normalize_value((double)d, norm_mode, divisor);
}

__global__ void fill_u64_kernel2(unsigned long long *p, size_t n, unsigned long long v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void frame_rec_kernel(FrameRec *rec, int n, const int64_t *list_off, const int32_t *list_len,
                                 const int32_t *group_id, const int32_t *group_ptr, const int32_t *n_boxes, double cost_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FrameRec r;
    r.list_off = list_off ? list_off[i] : 0;
    r.list_len = list_len ? list_len[i] : 0;
    const int g = group_id[i];
    r.e0 = group_ptr[g];
    r.e1 = group_ptr[g + 1];
    r.n_boxes = n_boxes[i];
    r.box_cost = __dmul_rn((double)r.n_boxes, cost_b);   // anno["gt_names"].shape[0] * cost_b (base_selector.py:85)
    rec[i] = r;
}

static double host_normalize_far(int norm_mode, double d, double divisor) {
    // value of the normalisation at a "far" distance: exp -> exactly 1.0 (exp(-d) < 2^-54 for d >= 38),
    // linear -> d / divisor (IEEE division, identical on host and device), raw -> d.
    if (norm_mode == ADSEL_NORM_EXP) return 1.0;
    if (norm_mode == ADSEL_NORM_LINEAR) return d / divisor;
    return d;
}

int build_lists(adsel_handle *h, const adsel_select_params *p, double t_far, bool sharded);
int lists_prepare(adsel_handle *h, bool sharded);
int grow_lists(adsel_handle *h, long long needed);
int comm_view(adsel_handle *h, CommView *view);

static unsigned long long *greedy_prof_ptr() {
    static unsigned long long *buf = nullptr;
    if (!buf) cudaMalloc(&buf, 12 * sizeof(unsigned long long));
    return buf;
}

static int greedy_threads() {
    const char *e = getenv("ADSEL_GREEDY_THREADS");
    const int t = e ? atoi(e) : 0;
    return (t == 256 || t == 512 || t == 1024) ? t : 512;
}

// Cluster size for the shared-memory-resident variant, 0 when it is not used.
// MEASURED SLOWER than the single-CTA kernel on B200 (P28K, budget 4800: 16.2 / 15.3 / 13.6 ms with 2 / 4 / 8 CTAs
// against 10.4 ms): what the shared-memory fps saves (the RED -> ld.cg round trip) is spent again on the scan of the
// whole slice, the two-level arg-max, the DSMEM exchange (~215 cycles one way) and the frame-record fetch that now
// misses L1 in every CTA but the winner's.  It is therefore OFF unless ADSEL_GREEDY_CLUSTER=1/2/4/8 asks for it
// (negative or unset: off); it passes the same parity tests and is kept as the measured alternative of DESIGN.md §5.
static int pick_cluster(const GreedyArgs &a, int nt, size_t *smem_out, int *cta_cols_out, int *sel_words_out) {
    const char *e = getenv("ADSEL_GREEDY_CLUSTER");
    const int forced = e ? atoi(e) : 0;
    if (forced <= 0) return 0;
    const long long cols = a.hi - a.lo;
    const size_t budget = 220 * 1024;   // of the 227 KB opt-in maximum: room for the static shared variables
    for (int c = 1; c <= CLUSTER_MAX; c *= 2) {
        if (forced > 0 && c != forced) continue;
        long long per = (cols + c - 1) / c;
        per = (per + 255) / 256 * 256;
        if (per < 256) per = 256;
        const size_t base = (size_t)per * 8 + (size_t)a.tl_n * 8 + 2 * CLUSTER_MAX * XSLOT_WORDS * 8 + 64;
        if (base > budget) continue;
        // the selected-frame bitmap goes to shared memory when it still fits, else global bytes are used
        const size_t bitmap = (size_t)((a.n + 31) / 32) * 4;
        const int sel_words = (a.n <= 524288 && base + bitmap <= budget) ? (a.n + 31) / 32 : 0;
        // default: the smallest cluster that fits, but no more than 16 slice entries per thread
        if (forced <= 0 && c < CLUSTER_MAX && per > 16LL * nt && ((cols + 2 * c - 1) / (2 * c)) >= 256) continue;
        *smem_out = base + (size_t)sel_words * 4;
        *cta_cols_out = (int)per;
        *sel_words_out = sel_words;
        return c;
    }
    return 0;
}

template <int NT, bool SHARDED>
static cudaError_t launch_cluster_nt(adsel_handle *h, const GreedyArgs &a, const CommView &view, size_t smem) {
    cudaFuncSetAttribute(greedy_cluster_kernel<NT, SHARDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)a.csize, 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, greedy_cluster_kernel<NT, SHARDED>, a, view);
}

template <bool SHARDED>
static void launch_greedy(adsel_handle *h, const GreedyArgs &a, const CommView &view, size_t smem) {
    if (a.prof) {   // ADSEL_GREEDY_PROFILE=1: per-phase cycle counters (development aid)
        cudaFuncSetAttribute(greedy_select_kernel<512, SHARDED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        greedy_select_kernel<512, SHARDED, true><<<1, 512, smem, h->stream>>>(a, view);
        return;
    }
    switch (greedy_threads()) {
        case 256:
            cudaFuncSetAttribute(greedy_select_kernel<256, SHARDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            greedy_select_kernel<256, SHARDED><<<1, 256, smem, h->stream>>>(a, view);
            break;
        case 1024:
            cudaFuncSetAttribute(greedy_select_kernel<1024, SHARDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            greedy_select_kernel<1024, SHARDED><<<1, 1024, smem, h->stream>>>(a, view);
            break;
        default:
            cudaFuncSetAttribute(greedy_select_kernel<512, SHARDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            greedy_select_kernel<512, SHARDED><<<1, 512, smem, h->stream>>>(a, view);
    }
}

// Loads the kernels of the sharded path up front: with lazy module loading a first-use load can need
// the device idle, which a spinning persistent kernel waiting for a peer would never grant.
int greedy_preload() {
    cudaFuncAttributes attr;
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_select_kernel<256, true>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_select_kernel<512, true>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_select_kernel<1024, true>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_select_kernel<256, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_select_kernel<512, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_select_kernel<1024, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_cluster_kernel<256, true>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_cluster_kernel<512, true>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_cluster_kernel<1024, true>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_cluster_kernel<256, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_cluster_kernel<512, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_cluster_kernel<1024, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, temporal_lut_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, fill_u64_kernel2));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, frame_rec_kernel));
    return ADSEL_OK;
}

static int validate_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                           int32_t first_index, int32_t out_cap) {
    ADSEL_REQUIRE(h->n > 0, "adsel_select: call adsel_set_pool first");
    ADSEL_REQUIRE(p && out_cap > 0, "adsel_select: null argument");
    ADSEL_REQUIRE(p->spatial >= 0 && p->spatial <= 2 && p->normalize >= 0 && p->normalize <= 2 && p->aggregate >= 0 &&
                      p->aggregate <= 2,
                  "adsel_select: bad mode");
    ADSEL_REQUIRE(p->spatial != ADSEL_SPATIAL_NONE || p->temporal, "adsel_select: no distance term");
    ADSEL_REQUIRE(!(p->spatial != ADSEL_SPATIAL_NONE && p->temporal && p->aggregate == ADSEL_AGG_SUM && !(p->lambda_t >= 0.0)),
                  "adsel_select: lambda_t must be >= 0 (row maximum argument, DESIGN.md §5)");
    ADSEL_REQUIRE(n_sampled >= 0 && (n_sampled == 0 || sampled), "adsel_select: bad sampled list");
    ADSEL_REQUIRE(n_sampled > 0 || (first_index >= 0 && first_index < h->n), "adsel_select: first_index out of range");
    for (int i = 0; i < n_sampled; ++i) ADSEL_REQUIRE(sampled[i] >= 0 && sampled[i] < h->n, "adsel_select: sampled index out of range");
    ADSEL_REQUIRE((h->n + BLK - 1) / BLK <= MAX_BLKS, "adsel_select: pool larger than 1,254,400 frames");
    if (p->normalize == ADSEL_NORM_LINEAR) {
        ADSEL_REQUIRE(p->spatial != ADSEL_SPATIAL_GEODESIC || p->max_spatial > 0.0, "adsel_select: linear needs max_spatial");
        ADSEL_REQUIRE(!p->temporal || p->max_temporal > 0.0, "adsel_select: linear needs max_temporal");
    }
    return ADSEL_OK;
}

// Allocations of a selection (separate from the enqueue so that a host thread driving several ranks
// can finish every cudaMalloc/cudaFree -- which may synchronise the device -- before any rank's
// persistent kernel starts waiting for its peers).
int greedy_prepare(adsel_handle *h, const adsel_select_params *p, int32_t n_sampled, int32_t out_cap, bool sharded) {
    const int n = h->n;
    if (p->spatial == ADSEL_SPATIAL_GEODESIC) ADSEL_TRY(lists_prepare(h, sharded));
    ADSEL_TRY(h->fps.reserve(n));
    ADSEL_TRY(h->frame_rec.reserve(n));
    ADSEL_TRY(h->is_selected.reserve(n));
    ADSEL_TRY(h->tlut.reserve(n));
    ADSEL_TRY(h->d_sampled.reserve(n_sampled > 0 ? n_sampled : 1));
    ADSEL_TRY(h->d_out.reserve(out_cap));
    ADSEL_TRY(h->d_result.reserve(8));
    return ADSEL_OK;
}

static double far_temporal(const adsel_select_params *p) {
    return p->temporal ? host_normalize_far(p->normalize, TEMPORAL_MARGIN, p->max_temporal) : 0.0;
}

// Enqueues temporal LUT -> neighbour lists (skipped when the resident ones match) on the handle's stream.
static int enqueue_lists(adsel_handle *h, const adsel_select_params *p, bool sharded) {
    const int n = h->n;
    if (p->temporal) {
        temporal_lut_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->tlut.p, n, p->normalize, p->max_temporal);
        h->launches += 1;
        ADSEL_CUDA(cudaGetLastError());
    }
    if (p->spatial == ADSEL_SPATIAL_GEODESIC)
        ADSEL_TRY(build_lists(h, p, far_temporal(p), sharded));
    return ADSEL_OK;
}

// Enqueues frame records -> fps fill -> greedy kernel on the handle's stream.
static int enqueue_greedy(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                          int32_t first_index, int32_t out_cap, bool sharded) {
    const int n = h->n;
    const int nblk = (n + BLK - 1) / BLK;
    GreedyArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    a.nblk = nblk;
    a.lo = sharded ? h->shard_lo : 0;
    a.hi = sharded ? h->shard_hi : n;
    a.blk_lo = a.lo / BLK;
    a.blk_hi = a.hi > a.lo ? (a.hi + BLK - 1) / BLK : a.blk_lo;
    a.spatial = p->spatial;
    a.temporal = p->temporal ? 1 : 0;
    a.tl_n = a.temporal ? (h->max_group < 4096 ? h->max_group : 4096) : 0;   // |i-j| < group size inside a group
    a.aggregate = p->aggregate;
    a.groups_contiguous = h->groups_contiguous ? 1 : 0;
    a.pull = getenv("ADSEL_GREEDY_PULL") ? atoi(getenv("ADSEL_GREEDY_PULL")) : 0;
    a.lambda_t = p->lambda_t;
    const double inf = host_inf();
    // the value every untouched entry of a row takes
    a.s_far = p->spatial == ADSEL_SPATIAL_GEODESIC ? host_normalize_far(p->normalize, inf, p->max_spatial)
              : p->spatial == ADSEL_SPATIAL_EUCLID ? TEMPORAL_MARGIN : 0.0;
    a.t_far = far_temporal(p);
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

    const bool lists = p->spatial == ADSEL_SPATIAL_GEODESIC;
    frame_rec_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->frame_rec.p, n, lists ? h->lists.off : nullptr,
                                                             lists ? h->lists.len : nullptr, h->group_id.p, h->group_ptr.p,
                                                             h->n_boxes.p, p->cost_b);
    h->launches += 1;

    a.cost_b = p->cost_b;
    a.cost_f = p->cost_f;
    a.cost_start = p->cost_start;
    a.budget = p->budget;
    a.forbid_sampled = p->forbid_sampled;
    a.n_sampled = n_sampled;
    a.first_index = first_index;
    a.out_cap = out_cap;
    a.list_capacity = lists ? h->lists.capacity : 0;
    a.list_cursor = lists ? h->lists.cursor : nullptr;
    a.rec = h->frame_rec.p;
    a.loc = h->loc.p;
    a.group_members = h->group_members.p;
    a.list_idx = h->lists.idx;
    a.list_val = h->lists.val;
    a.tlut = h->tlut.p;
    a.fps = h->fps.p;
    a.is_selected = h->is_selected.p;
    a.sampled = h->d_sampled.p;
    a.out = h->d_out.p;
    a.result = h->d_result.p;
    a.prof = getenv("ADSEL_GREEDY_PROFILE") ? greedy_prof_ptr() : nullptr;

    if (n_sampled > 0)
        ADSEL_CUDA(cudaMemcpyAsync(h->d_sampled.p, sampled, sizeof(int32_t) * n_sampled, cudaMemcpyHostToDevice, h->stream));
    ADSEL_CUDA(cudaMemsetAsync(h->is_selected.p, 0, n, h->stream));
    ADSEL_CUDA(cudaMemsetAsync(h->d_result.p, 0, sizeof(double) * 8, h->stream));
    fill_u64_kernel2<<<h->num_sms * 2, 256, 0, h->stream>>>(h->fps.p, (size_t)n, far_bits);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    a.sel_words = n <= 524288 ? (n + 31) / 32 : 0;
    a.far_bits = far_bits;
    CommView view;
    if (sharded) ADSEL_TRY(comm_view(h, &view));
    size_t csmem = 0;
    int cta_cols = 0, csel = 0;
    const int nt = greedy_threads();
    const int csize = a.prof ? 0 : pick_cluster(a, nt, &csmem, &cta_cols, &csel);
    if (csize > 0) {   // shared-memory-resident fps, one CTA per slice piece, exchange over distributed shared memory
        a.csize = csize;
        a.cta_cols = cta_cols;
        a.sel_words = csel;
        cudaError_t e;
        if (sharded) e = nt == 256 ? launch_cluster_nt<256, true>(h, a, view, csmem) : nt == 1024 ? launch_cluster_nt<1024, true>(h, a, view, csmem)
                                                                                                  : launch_cluster_nt<512, true>(h, a, view, csmem);
        else e = nt == 256 ? launch_cluster_nt<256, false>(h, a, view, csmem) : nt == 1024 ? launch_cluster_nt<1024, false>(h, a, view, csmem)
                                                                                           : launch_cluster_nt<512, false>(h, a, view, csmem);
        ADSEL_CUDA(e);
        h->greedy_cluster = csize;
    } else {
        const size_t smem = (size_t)nblk * 20 + (size_t)a.tl_n * 8 + (size_t)a.sel_words * 4 + 16;
        if (sharded) launch_greedy<true>(h, a, view, smem);
        else launch_greedy<false>(h, a, view, smem);
        h->greedy_cluster = 0;
    }
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    return ADSEL_OK;
}

static int finish_select(adsel_handle *h, const double *result, int32_t *out_idx, int32_t *out_m, double *out_cost,
                         const char *who) {
    const int m = (int)result[2];
    const int status = (int)result[1];
    h->last_steps = (int64_t)result[3];
    h->list_total = (int64_t)result[5];
    if (m > 0) ADSEL_CUDA(cudaMemcpy(out_idx, h->d_out.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost));
    *out_m = m;
    if (out_cost) *out_cost = result[0];
    if (status == ADSEL_ERR_RESELECTED) {
        set_error("id: " + std::to_string((int)result[4]) + " has been selected");
        return ADSEL_ERR_RESELECTED;
    }
    if (status == ADSEL_ERR_CAPACITY) {
        set_error(std::string(who) + (result[4] == -1.0 ? ": neighbour-list pool too small" : ": out_idx capacity exhausted"));
        return ADSEL_ERR_CAPACITY;
    }
    if (status == ADSEL_ERR_COMM) {
        set_error(std::string(who) + ": timed out waiting for a peer rank's mailbox record");
        return ADSEL_ERR_COMM;
    }
    if (status == ADSEL_ERR_PEER_ABORT) {
        set_error(std::string(who) + ": a peer rank aborted the selection");
        return ADSEL_ERR_PEER_ABORT;
    }
    if (status != 0) {
        set_error(std::string(who) + ": device-side failure " + std::to_string(status));
        return status;
    }
    return ADSEL_OK;
}

int greedy_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                  int32_t first_index, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_TRY(validate_select(h, p, sampled, n_sampled, first_index, out_cap));
    ADSEL_REQUIRE(out_idx && out_m, "adsel_select: null argument");
    ADSEL_TRY(greedy_prepare(h, p, n_sampled, out_cap, false));
    for (int attempt = 0; attempt < 2; ++attempt) {
        ADSEL_TRY(begin_stage(h));
        ADSEL_TRY(enqueue_lists(h, p, false));
        ADSEL_TRY(end_stage(h, 2));
        ADSEL_TRY(begin_stage(h));
        ADSEL_TRY(enqueue_greedy(h, p, sampled, n_sampled, first_index, out_cap, false));
        double result[8];
        ADSEL_CUDA(cudaMemcpyAsync(result, h->d_result.p, sizeof(result), cudaMemcpyDeviceToHost, h->stream));
        ADSEL_TRY(end_stage(h, 3));
        if (getenv("ADSEL_GREEDY_PROFILE")) {
            unsigned long long prof[12];
            cudaMemcpy(prof, greedy_prof_ptr(), sizeof(prof), cudaMemcpyDeviceToHost);
            const double steps = result[3] > 0 ? result[3] : 1;
            fprintf(stderr, "[greedy profile] cycles/step: apply %.0f | wait B1 %.0f | refresh %.0f | wait B2 %.0f | reduce(+exchange) %.0f | "
                            "fetch+bookkeeping %.0f | wait B3 %.0f | loop %.0f\n",
                    prof[0] / steps, prof[1] / steps, prof[2] / steps, prof[3] / steps, prof[4] / steps, prof[5] / steps,
                    prof[6] / steps, prof[7] / steps);
            fprintf(stderr, "[greedy profile]   one mark %.0f | block loop %.0f | warp_best %.0f\n", prof[8] / steps, prof[9] / steps,
                    prof[10] / steps);
        }
        if ((int)result[1] == ADSEL_ERR_CAPACITY && result[4] == -1.0 && attempt == 0) {
            ADSEL_TRY(grow_lists(h, (long long)result[5]));
            continue;
        }
        return finish_select(h, result, out_idx, out_m, out_cost, "adsel_select");
    }
    return ADSEL_ERR_CAPACITY;
}

// ---- pool-sharded selection: enqueue on every rank, then collect ------------------------------------
int greedy_select_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                                int32_t first_index, int32_t out_cap) {
    ADSEL_TRY(validate_select(h, p, sampled, n_sampled, first_index, out_cap));
    ADSEL_REQUIRE(h->comm_ready, "adsel_select_sharded_begin: call adsel_comm_init / adsel_comm_attach first");
    ADSEL_REQUIRE(!h->pending, "adsel_select_sharded_begin: a sharded selection is already pending");
    ADSEL_TRY(greedy_prepare(h, p, n_sampled, out_cap, true));
    ADSEL_CUDA(cudaEventRecord(h->pending_ev[0], h->stream));
    ADSEL_TRY(enqueue_lists(h, p, true));
    ADSEL_TRY(enqueue_greedy(h, p, sampled, n_sampled, first_index, out_cap, true));
    ADSEL_CUDA(cudaEventRecord(h->pending_ev[1], h->stream));
    h->pending = true;
    h->pending_cap = out_cap;
    return ADSEL_OK;
}

int greedy_select_sharded_end(adsel_handle *h, int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost) {
    ADSEL_REQUIRE(h->pending, "adsel_select_sharded_end: no pending sharded selection");
    ADSEL_REQUIRE(out_idx && out_m && out_cap >= h->pending_cap, "adsel_select_sharded_end: bad output buffer");
    h->pending = false;
    double result[8];
    ADSEL_CUDA(cudaMemcpyAsync(result, h->d_result.p, sizeof(result), cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    ADSEL_CUDA(cudaEventElapsedTime(&ms, h->pending_ev[0], h->pending_ev[1]));
    h->stage_ms[2] = 0.f;
    h->stage_ms[3] = ms;   // lists + greedy + exchanges of this rank
    const int status = (int)result[1];
    // every rank ran the same number of exchanges unless the launch failed; after a failure the tags are
    // moved to a fresh epoch (all ranks observe the same failure through the exchanged abort record)
    if (status == 0 || status == ADSEL_ERR_RESELECTED) h->comm_seq += (unsigned)result[3];
    else h->comm_seq = ((h->comm_seq >> 24) + 1u) << 24;
    // an overflowing list region of the arena cannot grow in place: the caller re-creates the arenas with
    // adsel_comm_init(list_cap = the largest adsel_get_counters()[2] over the ranks) and rebuilds
    if (status != 0) h->lists_mode = -1;
    return finish_select(h, result, out_idx, out_m, out_cost, "adsel_select_sharded");
}

}  // namespace adsel
