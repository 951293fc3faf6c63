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
            sh.cost = __dadd_rn(sh.cost, __dmul_rn((double)sh.rec.n_boxes, a.cost_b));
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
                        sh.cost = __dadd_rn(sh.cost, __dmul_rn((double)rec.n_boxes, a.cost_b));
                        sh.rec = rec;
                        sh.me = me;
                        sh.cur = nxt;
                    } else if (was_selected) {   // :182 assert
                        a.result[1] = (double)ADSEL_ERR_RESELECTED;
                        a.result[4] = (double)nxt;
                        sh.stop = 1;
                    } else {
                        sh.cost = __dadd_rn(sh.cost, a.cost_f);          // :184
                        sh.cost = __dadd_rn(sh.cost, __dmul_rn((double)rec.n_boxes, a.cost_b));      // :185 (n_boxes*cost_b, one rounded product)
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
// Resident variant (the default whenever the columns fit): fps lives in SHARED memory.
//
// In the kernel above the min-distance vector lives in L2, so a step pays four dependent L2 round
// trips (frame record, list entries, RED, re-read of the dirtied blocks) and three CTA barriers behind
// which one warp does the serial work.  Here the columns are held in shared memory -- by one CTA when
// they fit (28,130 frames * 8 B = 225 KB of the 227 KB a CTA may use, 256-frame blocks), else split over
// the CTAs of a thread-block cluster (up to 16: 440 k frames) -- and a step is organised so that its
// dependent chain is ONE L2 round trip (the winner's list entries) plus shared-memory work:
//   * every thread keeps the bookkeeping state (cost, count, current frame) in registers and every warp
//     does the final arg-max over the block maxima redundantly, so the winner needs no broadcast and no
//     barrier of its own: two CTA barriers per step plus one that hides under the list load;
//   * a row update is a plain shared-memory min.  Group (temporal) entries are written while the list
//     loads are in flight, list entries after a barrier (list values of same-group frames are <= the
//     group values, so the order makes the plain writes exact); with the "exp" normalisation the group
//     phase only visits |j - src| < 38, the part of the temporal LUT that differs from the far value;
//   * a block's maximum can only change when the entry that IS its maximum is lowered (entries only
//     decrease), so only those blocks are re-reduced: 1-2 per step instead of every touched block;
//   * the record of every block's best frame (list range, group range, box count) is cached in shared
//     memory when the block is refreshed -- fetched asynchronously and stored one phase later, validated
//     by the frame index -- so the winner's record is a shared-memory read, not an L2 round trip;
//   * cluster: warp 0 of every CTA posts its local maximum to every CTA through distributed shared
//     memory (tagged self-validating words, comm.cuh's protocol) and every warp polls its own CTA's
//     slots; every CTA does the bookkeeping redundantly, CTA 0 writes the results.
// ---------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

constexpr int CLUSTER_MAX = 16;                 // non-portable cluster size (opt-in), 8 is the portable limit
constexpr int XSLOT_WORDS = 4;
constexpr int RES_NT = 256;                     // threads of a resident CTA

struct BlkRec {           // cached FrameRec of a block's best frame (shared memory, 28 B)
    unsigned idx;         // frame the record belongs to (0xffffffff: none)
    int list_len;
    unsigned off_lo, off_hi;
    int e0, e1, n_boxes;
};

__device__ __forceinline__ unsigned long long ld_volatile_shared(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Issues the two 16-byte loads of a frame record without consuming them (the caller stores them a phase later).
__device__ __forceinline__ void ldg_rec_async(const FrameRec *p, uint4 &lo, uint4 &hi) {
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(p));
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(reinterpret_cast<const char *>(p) + 16));
}

// Exchange of the CTAs' local maxima through distributed shared memory.  xslot: this CTA's
// [2][CLUSTER_MAX][XSLOT_WORDS] slots; remote[r]: the same array in CTA r's shared memory.  Warp 0 posts,
// EVERY warp polls the own slots and reduces, so all threads of all CTAs hold the winner without a barrier.
__device__ __forceinline__ Best cluster_exchange(unsigned long long *xslot, unsigned long long *const *remote, int crank,
                                                 int csize, unsigned tag, Best local, bool poster) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t parity_off = (size_t)(tag & 1u) * CLUSTER_MAX * XSLOT_WORDS;
    const int nwords = 3 * csize;   // <= 48
    if (poster) {
        for (int l = lane; l < nwords; l += 32) {
            const int peer = l / 3, word = l - 3 * peer;
            const unsigned long long payload = word == 0 ? (local.v >> 32) : word == 1 ? (local.v & 0xffffffffull)
                                                                                      : (unsigned long long)(unsigned)local.i;
            remote[peer][parity_off + (size_t)crank * XSLOT_WORDS + word] = (payload << 32) | tag;   // st.shared::cluster
        }
    }
    unsigned long long got0 = 0, got1 = 0;
    if (lane < nwords) {
        const unsigned long long *p = xslot + parity_off + (size_t)(lane / 3) * XSLOT_WORDS + (lane % 3);
        do got0 = ld_volatile_shared(p);
        while ((unsigned)got0 != tag);
    }
    if (lane + 32 < nwords) {
        const int l = lane + 32;
        const unsigned long long *p = xslot + parity_off + (size_t)(l / 3) * XSLOT_WORDS + (l % 3);
        do got1 = ld_volatile_shared(p);
        while ((unsigned)got1 != tag);
    }
    // lane r < csize assembles CTA r's record from words 3r, 3r+1, 3r+2 (word w sits in lane w & 31, register w >> 5)
    unsigned part[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int w = lane < csize ? 3 * lane + q : 0;
        const unsigned long long a0 = __shfl_sync(full, got0, w & 31), a1 = __shfl_sync(full, got1, w & 31);
        part[q] = (unsigned)((w >> 5 ? a1 : a0) >> 32);
    }
    Best x{0ull, 0x7fffffff};
    if (lane < csize) {
        x.v = ((unsigned long long)part[0] << 32) | part[1];
        x.i = (int)part[2];
    }
    return warp_best(x);
}

#define RPROF_MARK(slot)                                  \
    if (PROFILE && threadIdx.x == 0) {                     \
        const int dep = *(volatile int *)&sh.err;    /* consume shared memory: a deferred barrier blocks here */ \
        long long now;                                    \
        asm volatile("{\n.reg .b32 t;\nmov.b32 t, %1;\nmov.u64 %0, %%clock64;\n}" : "=l"(now) : "r"(dep) : "memory"); \
        prof_acc[slot] += (unsigned long long)(now - prof_t); \
        prof_t = now;                                     \
    }

#define FPROF(slot)                                       \
    if (PROFILE && threadIdx.x == 0 && blockIdx.x == 0) { \
        long long now;                                    \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(now) :: "memory"); \
        prof_acc[slot] += (unsigned long long)(now - prof_t); \
        prof_t = now;                                     \
    }

struct ResidentShared {
    int ndirty[2];
    int err, err_frame, sat;
};

template <int RB, bool CLUSTER, bool PROFILE>
__global__ void __launch_bounds__(RES_NT, 1) greedy_resident_kernel(const GreedyArgs a) {
    constexpr int NT = RES_NT, NW = NT / 32;
    unsigned long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = PROFILE ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int crank = 0;
    if (CLUSTER) crank = (int)cg::this_cluster().block_rank();
    const int csize = CLUSTER ? a.csize : 1;
    const int c_lo = min(a.lo + crank * a.cta_cols, a.hi), c_hi = min(c_lo + a.cta_cols, a.hi), c_len = c_hi - c_lo;
    const int nblk = a.r_nblk;   // blocks of a CTA's slice (cta_cols / RB); those at or beyond c_len are empty
    const bool writer = crank == 0 && tid == 0;

    extern __shared__ unsigned long long greedy_smem[];
    unsigned long long *fps_s = greedy_smem;                                         // [cta_cols]
    unsigned long long *blk_val = fps_s + a.cta_cols;                                // [nblk]
    unsigned long long *xslot = blk_val + nblk;                                      // CLUSTER: [2][CLUSTER_MAX][XSLOT_WORDS]
    unsigned long long **remote = reinterpret_cast<unsigned long long **>(xslot + (CLUSTER ? 2 * CLUSTER_MAX * XSLOT_WORDS : 0));
    double *tl = reinterpret_cast<double *>(remote + (CLUSTER ? CLUSTER_MAX : 0));   // [tl_n] head of the temporal LUT
    int *blk_idx = reinterpret_cast<int *>(tl + a.tl_n);                             // [nblk]
    int *dirty = blk_idx + nblk;                                                     // [nblk]
    int *dirty_list = dirty + nblk;                                                  // [nblk]
    BlkRec *blk_rec = reinterpret_cast<BlkRec *>(dirty_list + nblk);                 // [nblk]
    unsigned *selbits = reinterpret_cast<unsigned *>(blk_rec + nblk);                // [sel_words] (0 words: private global bytes)
    unsigned char *sel_bytes = a.is_selected + (size_t)crank * a.n;                  // this CTA's private copy of the flags
    __shared__ ResidentShared sh;

    const int tl_n = a.tl_n;
    const unsigned long long far_bits = a.far_bits;
    for (int i = tid; i < a.cta_cols; i += NT) fps_s[i] = far_bits;
    for (int d = tid; d < tl_n; d += NT) tl[d] = a.tlut[d];
    for (int wd = tid; wd < a.sel_words; wd += NT) selbits[wd] = 0u;
    if (CLUSTER) {
        for (int i = tid; i < 2 * CLUSTER_MAX * XSLOT_WORDS; i += NT) xslot[i] = 0ull;
        if (tid < CLUSTER_MAX) remote[tid] = tid < csize ? cg::this_cluster().map_shared_rank(xslot, tid) : nullptr;
    }
    // every fps entry starts at `far`, so a block's best is its first frame: no initial refresh
    for (int lb = tid; lb < nblk; lb += NT) {
        const bool nonempty = lb * RB < c_len;
        const int first = c_lo + lb * RB;
        blk_val[lb] = nonempty ? far_bits : 0ull;
        blk_idx[lb] = nonempty ? first : 0x7fffffff;
        dirty[lb] = 0;
        BlkRec br;
        br.idx = 0xffffffffu;
        br.list_len = br.e0 = br.e1 = br.n_boxes = 0;
        br.off_lo = br.off_hi = 0u;
        if (nonempty) {
            const FrameRec r = a.rec[first];
            br.idx = (unsigned)first;
            br.list_len = r.list_len;
            br.off_lo = (unsigned)r.list_off;
            br.off_hi = (unsigned)((unsigned long long)r.list_off >> 32);
            br.e0 = r.e0;
            br.e1 = r.e1;
            br.n_boxes = r.n_boxes;
        }
        blk_rec[lb] = br;
    }
    if (tid == 0) {
        sh.ndirty[0] = sh.ndirty[1] = 0;
        sh.err = 0;
        sh.err_frame = -1;
        sh.sat = 0;
    }
    __syncthreads();
    // Saturation of the temporal LUT: for d >= sat, tlut[d] equals the far value bit for bit (exp: 38), so group
    // members further away than that keep the row's far value and need not be visited.
    int tl_sat = 0x7fffffff;
    if (a.temporal && a.groups_contiguous && a.spatial != ADSEL_SPATIAL_EUCLID) {
        const int lim = min(a.n, 4096);
        const unsigned long long tfar = (unsigned long long)__double_as_longlong(a.t_far);
        int mx = 0;
        for (int d = tid; d < lim; d += NT)
            if ((unsigned long long)__double_as_longlong(a.tlut[d]) != tfar) mx = d + 1;
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0 && mx > 0) atomicMax(&sh.sat, mx);
        __syncthreads();
        tl_sat = sh.sat >= lim ? 0x7fffffff : sh.sat;
    }

    // ---- state replicated in the registers of every thread (of every CTA) ----
    double cost = a.cost_start;
    int m = 0, cur = -1, stop = 0;   // stop: 1 budget reached, 2 error
    long long l_off = 0;
    int l_len = 0, e0 = 0, e1 = 0;
    double2 me = make_double2(0.0, 0.0);
    int par = 0;                     // parity of the dirty counter in use (the other one is being reset)
    int to_flag = -1;                // thread 0: frame whose selected flag is checked / set in the next apply
    bool flag_check = false;         // thread 0: flag_seen holds the frame's flag as loaded when it was picked
    bool flag_seen = false;

    auto set_flag = [&](int idx) {   // thread 0 only
        if (a.sel_words) selbits[idx >> 5] |= 1u << (idx & 31);
        else sel_bytes[idx] = 1;
    };
    auto get_flag = [&](int idx) -> bool {   // thread 0 only
        if (a.sel_words) return (selbits[idx >> 5] >> (idx & 31)) & 1u;
        return __ldcg(sel_bytes + idx) != 0;
    };
    auto load_rec_global = [&](int idx) -> int {   // all threads, same address; returns n_boxes
        const FrameRec r = a.rec[idx];
        l_off = r.list_off;
        l_len = r.list_len;
        e0 = r.e0;
        e1 = r.e1;
        return r.n_boxes;
    };
    auto lower = [&](int j, unsigned long long v) {
        const int lj = j - c_lo;
        if (v < fps_s[lj]) {
            fps_s[lj] = v;
            const int lb = lj / RB;
            // only the entry that IS the block's maximum can change it by decreasing
            if (blk_idx[lb] == j && atomicExch(&dirty[lb], 1) == 0) dirty_list[atomicAdd(&sh.ndirty[par], 1)] = lb;
        }
    };
    // pending record of a refreshed block's best frame (lane 0 of every warp)
    int pend_lb = -1, pend_idx = 0;
    uint4 pend_lo = make_uint4(0, 0, 0, 0), pend_hi = make_uint4(0, 0, 0, 0);

    // fps = min(fps, distance_map[cur]) on this CTA's slice; ends with the list writes (caller barriers)
    auto apply = [&]() {
        const int src = cur;
        const int len = a.spatial == ADSEL_SPATIAL_GEODESIC ? l_len : 0;
        // 1. list loads in flight (one L2 round trip, the only one of a step)
        int j0 = -1, j1 = -1;
        unsigned long long v0 = 0, v1 = 0;
        if (tid < len) {
            j0 = a.list_idx[l_off + tid];
            v0 = (unsigned long long)__double_as_longlong(a.list_val[l_off + tid]);
        }
        if (tid + NT < len) {
            j1 = a.list_idx[l_off + tid + NT];
            v1 = (unsigned long long)__double_as_longlong(a.list_val[l_off + tid + NT]);
        }
        // 2. temporal group (or Euclid city group) while they fly
        GreedyShared gs;   // group_value() reads sh.me only
        gs.me = me;
        if (a.temporal || a.spatial == ADSEL_SPATIAL_EUCLID) {
            if (a.groups_contiguous) {
                int g0 = max(e0, c_lo), g1 = min(e1, c_hi);
                if (tl_sat != 0x7fffffff) {
                    g0 = max(g0, src - tl_sat + 1);
                    g1 = min(g1, src + tl_sat);
                }
                for (int j = g0 + tid; j < g1; j += NT)
                    lower(j, (unsigned long long)__double_as_longlong(group_value(a, gs, tl, tl_n, src, j)));
            } else {
                for (int t = e0 + tid; t < e1; t += NT) {
                    const int j = a.group_members[t];
                    if (j >= c_lo && j < c_hi)
                        lower(j, (unsigned long long)__double_as_longlong(group_value(a, gs, tl, tl_n, src, j)));
                }
            }
        }
        __syncthreads();   // group writes before list writes (hidden under the list loads)
        // 3. list entries: unique columns, plain writes
        if (j0 >= c_lo && j0 < c_hi) lower(j0, v0);
        if (j1 >= c_lo && j1 < c_hi) lower(j1, v1);
        for (int t = tid + 2 * NT; t < len; t += NT) {
            const int j = a.list_idx[l_off + t];
            if (j >= c_lo && j < c_hi) lower(j, (unsigned long long)__double_as_longlong(a.list_val[l_off + t]));
        }
        // 4. deferred work that must not race with the bookkeeping reads before the barrier above
        if (lane == 0 && pend_lb >= 0) {
            BlkRec br;
            br.idx = (unsigned)pend_idx;
            br.off_lo = pend_lo.x;      // FrameRec: list_off (8), list_len, self | e0, e1, n_boxes, self2
            br.off_hi = pend_lo.y;
            br.list_len = (int)pend_lo.z;
            br.e0 = (int)pend_hi.x;
            br.e1 = (int)pend_hi.y;
            br.n_boxes = (int)pend_hi.z;
            blk_rec[pend_lb] = br;
            pend_lb = -1;
        }
        if (tid == 0 && to_flag >= 0) {   // :182 assert, evaluated a phase late so that the flag load is off the chain
            if (flag_check && flag_seen) {
                sh.err = 1;
                sh.err_frame = to_flag;
            } else {
                set_flag(to_flag);
            }
            to_flag = -1;
        }
    };

    bool local_abort = false;
    {
        const unsigned long long needed = a.list_cursor ? *a.list_cursor : 0ull;
        if (writer) a.result[5] = (double)needed;
        if ((long long)needed > a.list_capacity) {   // neighbour-list pool too small: the host grows it and reruns
            local_abort = true;
            if (writer) {
                a.result[1] = (double)ADSEL_ERR_CAPACITY;
                a.result[4] = -1.0;
            }
        }
    }
    if (CLUSTER) cg::this_cluster().sync();   // every CTA's slots are zeroed and mapped before anyone posts
    long long steps = 0;
    bool seeding = a.n_sampled > 0;   // first pick = argmax without a preceding update (:165)
    if (!local_abort) {
        if (a.forbid_sampled && tid == 0)
            for (int t = 0; t < a.n_sampled; ++t) set_flag(a.sampled[t]);
        if (a.n_sampled == 0) {   // empty buffer: the caller's random.choice is taken and costed up front (:168,172-174)
            cur = a.first_index;
            const int nb = load_rec_global(cur);
            if (a.spatial == ADSEL_SPATIAL_EUCLID) me = a.loc[cur];
            if (writer) a.out[0] = cur;
            m = 1;
            to_flag = cur;
            cost = __dadd_rn(cost, a.cost_f);
            cost = __dadd_rn(cost, __dmul_rn((double)nb, a.cost_b));
        } else {
            for (int t = 0; t < a.n_sampled; ++t) {   // fps = min over the rows of the frames already in the buffer (:160-164)
                cur = a.sampled[t];
                load_rec_global(cur);
                if (a.spatial == ADSEL_SPATIAL_EUCLID) me = a.loc[cur];
                apply();
                __syncthreads();
            }
        }
        while (true) {
            RPROF_MARK(7)
            if (!seeding) apply();
            RPROF_MARK(0)
            __syncthreads();
            RPROF_MARK(1)
            {   // re-reduce the blocks whose maximum was lowered, one warp per block
                const int nd = sh.ndirty[par];
                for (int q = wid; q < nd; q += NW) {
                    const int lb = dirty_list[q];
                    Best e[RB / 32];
#pragma unroll
                    for (int i = 0; i < RB / 32; ++i) {
                        const int lj = lb * RB + lane + 32 * i;
                        e[i] = lj < c_len ? Best{fps_s[lj], c_lo + lj} : Best{0ull, 0x7fffffff};
                    }
#pragma unroll
                    for (int w = 1; w < RB / 32; w <<= 1)
#pragma unroll
                        for (int i = 0; i + w < RB / 32; i += 2 * w) e[i] = better(e[i], e[i + w]);
                    const Best x = warp_best(e[0]);
                    if (lane == 0) {
                        blk_val[lb] = x.v;
                        blk_idx[lb] = x.i;
                        dirty[lb] = 0;
                        if (x.i != 0x7fffffff) {   // its record: in flight now, stored into blk_rec during the next apply
                            pend_lb = lb;
                            pend_idx = x.i;
                            ldg_rec_async(a.rec + x.i, pend_lo, pend_hi);
                        }
                    }
                }
            }
            RPROF_MARK(2)
            __syncthreads();
            RPROF_MARK(3)
            if (tid == 0) sh.ndirty[par] = 0;   // the next apply counts in the other word (zeroed a step ago)
            par ^= 1;
            // arg-reduce over this CTA's block maxima, by every warp
            Best y{0ull, 0x7fffffff};
            for (int b0 = lane; b0 < nblk; b0 += 4 * 32) {
                Best e[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int b = b0 + 32 * q;
                    e[q] = b < nblk ? Best{blk_val[b], blk_idx[b]} : Best{0ull, 0x7fffffff};
                }
                y = better(y, better(better(e[0], e[1]), better(e[2], e[3])));
            }
            y = warp_best(y);
            if (CLUSTER) y = cluster_exchange(xslot, remote, crank, csize, 1u + (unsigned)steps, y, wid == 0);
            RPROF_MARK(4)
            ++steps;
            if (sh.err) {   // :182 assert -- the frame applied in this step had been selected before
                stop = 2;
                break;
            }
            const int nxt = y.i;
            if (nxt < 0 || nxt >= a.n) {   // empty pool: cannot happen with n > 0
                if (writer) a.result[1] = (double)ADSEL_ERR_INVALID;
                stop = 2;
                break;
            }
            // the winner's record: cached with its block's maximum, else one global round trip
            int nb = 0;
            bool have = false;
            if (nxt >= c_lo && nxt < c_hi) {
                const BlkRec br = blk_rec[(nxt - c_lo) / RB];
                if (br.idx == (unsigned)nxt) {
                    have = true;
                    l_off = (long long)(((unsigned long long)br.off_hi << 32) | br.off_lo);
                    l_len = br.list_len;
                    e0 = br.e0;
                    e1 = br.e1;
                    nb = br.n_boxes;
                }
            }
            if (!have) nb = load_rec_global(nxt);
            if (a.spatial == ADSEL_SPATIAL_EUCLID) me = a.loc[nxt];
            const double box_cost = __dmul_rn((double)nb, a.cost_b);   // n_boxes * cost_b, one rounded product (base_selector.py:85)
            if (seeding) {   // :165,172-174 -- always taken, no budget check
                seeding = false;
                if (writer) a.out[0] = nxt;
                m = 1;
                cost = __dadd_rn(cost, a.cost_f);
                cost = __dadd_rn(cost, box_cost);
                cur = nxt;
                to_flag = nxt;
                continue;
            }
            if (tid == 0) flag_seen = get_flag(nxt);   // consumed in the next apply, or below when the loop ends here
            cost = __dadd_rn(cost, a.cost_f);      // :184
            cost = __dadd_rn(cost, box_cost);      // :185
            const bool over = cost > a.budget;     // :186
            const bool full = !over && m >= a.out_cap;
            if (over || full) {
                if (writer) {
                    if (flag_seen) {               // :182 comes before the cost test
                        a.result[1] = (double)ADSEL_ERR_RESELECTED;
                        a.result[4] = (double)nxt;
                    } else if (full) {
                        a.result[1] = (double)ADSEL_ERR_CAPACITY;
                        a.result[4] = -2.0;
                    }
                }
                stop = over ? 1 : 2;
                break;
            }
            if (writer) a.out[m] = nxt;
            m += 1;
            cur = nxt;
            to_flag = nxt;
            flag_check = true;
            RPROF_MARK(5)
        }
    }
    __syncthreads();
    if (stop == 2 && sh.err && writer) {
        a.result[1] = (double)ADSEL_ERR_RESELECTED;
        a.result[4] = (double)sh.err_frame;
    }
    for (int i = tid; i < c_len; i += NT) a.fps[c_lo + i] = fps_s[i];   // adsel_get_fps
    if (writer) {
        a.result[0] = cost;
        a.result[2] = (double)m;
        a.result[3] = (double)steps;
        if (PROFILE)
            for (int i = 0; i < 12; ++i) a.prof[i] = prof_acc[i];
    }
    if (CLUSTER) cg::this_cluster().sync();   // no CTA leaves while a peer may still store into its slots
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
    r.n_boxes = n_boxes[i];   // anno["gt_names"].shape[0]; the kernels form n_boxes * cost_b as one rounded product (base_selector.py:85)
    r.self = r.self2 = (unsigned)i;
    rec[i] = r;
}

static double host_normalize_far(int norm_mode, double d, double divisor) {
    // value of the normalisation at a "far" distance: exp -> exactly 1.0 (exp(-d) < 2^-54 for d >= 38),
    // linear -> d / divisor (IEEE division, identical on host and device), raw -> d.
    if (norm_mode == ADSEL_NORM_EXP) return 1.0;
    if (norm_mode == ADSEL_NORM_LINEAR) return d / divisor;
    return d;
}

int build_lists(adsel_handle *h, const adsel_select_params *p, double t_far, bool sharded, bool replicate);
bool lazy_wanted(const adsel_handle *h, const adsel_select_params *p);
int lazy_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled, int32_t first_index,
                int32_t *out_idx, int32_t out_cap, int32_t *out_m, double *out_cost);
int lists_prepare(adsel_handle *h, bool sharded);
int grow_lists(adsel_handle *h, long long needed);
int comm_view(adsel_handle *h, CommView *view);
int warp_prepare(adsel_handle *h, const adsel_select_params *p, bool sharded);
int warp_preload();
int warp_enqueue(adsel_handle *h, const adsel_select_params *p, GreedyArgs &a, bool sharded, bool *launched);

static unsigned long long *greedy_prof_ptr() {
    static unsigned long long *buf = nullptr;
    if (!buf) cudaMalloc(&buf, 16 * sizeof(unsigned long long));
    return buf;
}

static int greedy_threads() {
    const char *e = getenv("ADSEL_GREEDY_THREADS");
    const int t = e ? atoi(e) : 0;
    return (t == 256 || t == 512 || t == 1024) ? t : 512;
}

// ---------------------------------------------------------------------------------------------
// Fast variant of the resident kernel: ONE CTA, the configuration of the headline workload family
// (no Euclid term, temporal groups that are contiguous index ranges), written for the LENGTH OF THE
// INSTRUCTION CHAIN of a step.  ncu on the generic resident kernel (profiles/r02_resident_v1_hotspots.txt):
// 2,464 warp-instructions per step, IPC per warp 0.08 -- the step is a serial chain of constant-bank reloads
// (LDC/LDCU 226 per step), mode tests and 64-bit compare/select trees, not of memory round trips.  Here
//   * everything a step needs from the kernel arguments is pinned in registers (no LDC in the loop) and
//     no mode is tested in the loop: the group value is a shared-memory table gval[|j - src|] that already
//     holds aggregate(s_far, normalize(|j - src|)) -- any normalisation and any aggregate cost one LDS;
//   * block refresh and final arg-max are three 32-bit REDUX operations each (max high word, max low word
//     among the lanes that hold it, min frame index among those that hold both) around VIMNMX trees: no
//     64-bit compare/select, no branches;
//   * the selected frame's record (list range, group range, box cost) comes from the shared-memory cache of
//     its block's best frame (two LDS.128), its list loads are issued before anything else, the cost
//     bookkeeping and the group phase run in their shadow, the re-selection flag is checked a phase late.
// Chain of a step: LDS record -> LDG list (the one L2 round trip) -> LDS/STS min -> BAR -> refresh of the
// 1-2 blocks whose best was lowered -> BAR -> arg-max over the block maxima by every warp.
// ---------------------------------------------------------------------------------------------
#define ADSEL_PIN32(x) asm volatile("" : "+r"(x))
#define ADSEL_PIN64(x) asm volatile("" : "+l"(x))
#define ADSEL_PINF64(x) asm volatile("" : "+d"(x))

struct __align__(16) FastShared {
    int4 nd[2];                  // per step parity: x = blocks listed for a re-reduce, y = the step's frame had been selected before,
                                 // z = an entry of the previous step's block was lowered (that block is re-reduced in full)
    uint4 sub_rec[8][2];         // record of every sub-block's best frame of the row's own block, copied in asynchronously
    unsigned long long sub_val[8];
    int sub_idx[8];
    int err, err_frame, sat, winner;
};

// Exchange state of the cluster variant only (the one-CTA variant has no shared memory to spare for it).
template <bool CLUSTER>
struct FastExchange {};
template <>
struct __align__(16) FastExchange<true> {
    uint4 xrecs[2][CLUSTER_MAX];   // every CTA's (value, frame) maximum of the step, stored here by its owner
    unsigned long long xbar[2];    // mbarriers (csize arrivals per phase), one per step parity
};

__device__ __forceinline__ int atom_inc_shared(int *p) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return old;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint4 lds128_volatile(const void *p) {   // one 16-byte access: a half is seen whole or not at all
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p))
                 : "memory");
    return v;
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int RB, int NT, bool CLUSTER, bool PROFILE>
__global__ void __launch_bounds__(NT, 1) greedy_fast_kernel(const GreedyArgs a) {
    // Blocks of RB frames; the row's own block is re-reduced as NSUB sub-blocks of SUBF frames dealt out to the NW warps.
    // RB = 512 halves the number of block maxima every warp scans at the end of a step (55 at 28,130 frames: two per lane).
    constexpr int NW = NT / 32, PER = RB / 32, SUBF = RB == 512 ? 64 : 32, NSUB = RB / SUBF, SPL = SUBF / 32;
    constexpr int SHIFT = RB == 512 ? 9 : RB == 256 ? 8 : 7;
    constexpr int QB = RB == 512 ? 2 : 4;   // block maxima per lane and pass of the final arg-max
    static_assert(RB == 128 || RB == 256 || RB == 512, "block size");
    static_assert(NSUB <= 8, "FastShared holds eight sub-block slots");
    unsigned long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = PROFILE ? clock64() : 0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // CLUSTER: the columns are split over the CTAs of a thread-block cluster; this CTA holds frames [c_lo, c_hi).  Every CTA
    // applies the part of a row that falls into its slice, reduces its own block maxima, and the CTAs' maxima are exchanged
    // through distributed shared memory: remote stores of the records followed by a release-arrive on every CTA's mbarrier
    // (a mailbox of plain tagged words polled by the CTAs, 414 cycles in isolation -- tools/microbench/dsmem.cu -- cost
    // 4,600 cycles per step inside this loop, the hardware cluster barrier ~1,500); the bookkeeping is replicated, CTA 0
    // writes the results.
    int crank = 0, csize = 1;
    if (CLUSTER) {
        crank = (int)cg::this_cluster().block_rank();
        csize = a.csize;
    }
    const bool writer = crank == 0 && tid == 0;

    // ---- hot arguments, pinned in registers ----
    int n = a.n, cols = a.cta_cols, nblk = a.r_nblk, G = a.tl_n, out_cap = a.out_cap;
    const int32_t *lidx = a.list_idx;
    const double *lval = a.list_val;
    const FrameRec *recs = a.rec;
    int32_t *out = a.out;
    unsigned char *sel_bytes = a.is_selected + (size_t)crank * a.n;   // this CTA's private copy of the flags
    int c_lo = min(crank * a.cta_cols, a.n), c_len = min(a.cta_cols, a.n - c_lo);   // this CTA's frames [c_lo, c_lo + c_len)
    double budget = a.budget, cost_f = a.cost_f, cost_b = a.cost_b;
    const bool has_list = a.spatial == ADSEL_SPATIAL_GEODESIC, has_group = a.temporal != 0;
    ADSEL_PIN32(n);
    ADSEL_PIN32(nblk);
    ADSEL_PIN32(G);
    ADSEL_PIN32(out_cap);
    ADSEL_PIN32(c_lo);
    ADSEL_PIN32(c_len);
    ADSEL_PIN64(lidx);
    ADSEL_PIN64(lval);
    ADSEL_PIN64(recs);
    ADSEL_PIN64(out);
    ADSEL_PIN64(sel_bytes);
    ADSEL_PINF64(budget);
    ADSEL_PINF64(cost_f);
    ADSEL_PINF64(cost_b);

    extern __shared__ unsigned long long greedy_smem[];
    unsigned long long *fps_s = greedy_smem;                              // [cols]
    unsigned long long *blk_val = fps_s + cols;                           // [nblk]
    unsigned long long *gval = blk_val + nblk;                            // [G] row value of a group member at index gap d
    uint4 *blk_rec = reinterpret_cast<uint4 *>(gval + G);                 // [nblk][2] FrameRec of the block's best frame (self-tagged halves)
    int *blk_idx = reinterpret_cast<int *>(blk_rec + 2 * (size_t)nblk);   // [nblk]
    int *dirty_list = blk_idx + nblk;                                     // [2 * nblk]
    int *blk_step = dirty_list + 2 * nblk;                                // [nblk] step that last re-reduced the block (its cached record may still be landing)
    __shared__ FastShared sh;
    __shared__ FastExchange<CLUSTER> xs;

    const unsigned long long far_bits = a.far_bits;
    for (int i = tid; i < cols; i += NT) fps_s[i] = i < c_len ? far_bits : 0ull;
    for (int d = tid; d < G; d += NT)
        gval[d] = (unsigned long long)__double_as_longlong(combine(a.spatial, a.temporal, a.aggregate, a.lambda_t, a.s_far, a.tlut[d]));
    // every fps entry starts at `far`, so a block's best is its first frame: no initial refresh
    for (int lb = tid; lb < nblk; lb += NT) {
        const int first = c_lo + (lb << SHIFT);
        const bool nonempty = (lb << SHIFT) < c_len;
        blk_val[lb] = nonempty ? far_bits : 0ull;
        blk_idx[lb] = nonempty ? first : 0x7fffffff;
        blk_step[lb] = -1;
        uint4 lo = make_uint4(0, 0, 0, 0xffffffffu), hi = make_uint4(0, 0, 0, 0xffffffffu);
        if (nonempty) ldg_rec_async(recs + first, lo, hi);
        blk_rec[2 * lb] = lo;
        blk_rec[2 * lb + 1] = hi;
    }
    if (tid == 0) {
        sh.nd[0] = sh.nd[1] = make_int4(0, 0, 0, 0);
        sh.err = 0;
        sh.err_frame = -1;
        sh.sat = 0;
    }
    __syncthreads();
    // Saturation of the temporal LUT: for d >= sat, tlut[d] equals the far value bit for bit (exp: 38), so group
    // members further away than that keep the row's far value and are not visited.
    int sat = 0x3fffffff;
    if (has_group) {
        const int lim = min(n, 4096);
        const unsigned long long tfar = (unsigned long long)__double_as_longlong(a.t_far);
        int mx = 0;
        for (int d = tid; d < lim; d += NT)
            if ((unsigned long long)__double_as_longlong(a.tlut[d]) != tfar) mx = d + 1;
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0 && mx > 0) atomicMax(&sh.sat, mx);
        __syncthreads();
        sat = sh.sat >= lim ? 0x3fffffff : sh.sat;
    }

    // ---- state replicated in the registers of every thread ----
    double cost = a.cost_start;
    int m = 0, cur = -1, cur_blk = -1, prev_blk = -1, stop = 0, par = 0;
    long long l_off = 0;
    int l_len = 0, e0 = 0, e1 = 0;
    int to_flag = -1;                // thread 0: frame whose selected flag is checked / set in the next apply
    bool flag_check = false;
    unsigned flag_seen = 0;
    bool marking = true;             // false while the rows of the buffer are applied (all blocks are re-reduced afterwards)

    auto lower = [&](int j, unsigned long long v) {   // j: frame of this CTA's slice
        const int lj = j - c_lo;
        if (v < fps_s[lj]) {
            fps_s[lj] = v;
            const int lb = lj >> SHIFT;
            // Only the entry that IS a block's maximum can change it by decreasing.  The block of the row's own
            // frame is always re-reduced (by all warps), so it is not listed; the block of the PREVIOUS row's frame gets
            // its maximum written while this row is applied, so it is re-reduced in full when anything in it is lowered;
            // another block can be listed twice per step (group phase, list phase): re-reduced twice, same result.
            if (lb != cur_blk && marking) {
                if (lb == prev_blk) sh.nd[par].z = 1;
                else if (blk_idx[lb] == j) dirty_list[atom_inc_shared(&sh.nd[par].x)] = lb;
            }
        }
    };
    auto take_rec = [&](const uint4 &lo, const uint4 &hi) -> double {   // FrameRec words -> row state; returns the box cost
        l_off = (long long)(((unsigned long long)lo.y << 32) | lo.x);
        l_len = has_list ? (int)lo.z : 0;
        e0 = (int)hi.x;
        e1 = (int)hi.y;
        return __dmul_rn((double)(int)hi.z, cost_b);   // n_boxes * cost_b, one rounded product (base_selector.py:85)
    };
    auto apply = [&]() {
        // 1. list loads in flight (the one L2 round trip of a step)
        const int32_t *pi = lidx + l_off + tid;
        const double *pv = lval + l_off + tid;
        int j0 = -1, j1 = -1;
        unsigned long long v0 = 0, v1 = 0;
        if (tid < l_len) {
            j0 = *pi;
            v0 = (unsigned long long)__double_as_longlong(*pv);
        }
        if (tid + NT < l_len) {
            j1 = pi[NT];
            v1 = (unsigned long long)__double_as_longlong(pv[NT]);
        }
        // 2. temporal group in their shadow: frames [e0, e1) within the LUT's unsaturated reach of cur
        if (has_group) {
            const int g0 = max(max(e0, cur - sat + 1), c_lo), g1 = min(min(e1, cur + sat), c_lo + c_len);
            for (int j = g0 + tid; j < g1; j += NT) {
                const int d = abs(j - cur);
                const unsigned long long v = d < G ? gval[d]
                    : (unsigned long long)__double_as_longlong(combine(a.spatial, a.temporal, a.aggregate, a.lambda_t, a.s_far, a.tlut[d]));
                lower(j, v);
            }
        }
        FPROF(0)
        __syncthreads();   // group writes before list writes (list values of same-group frames are <= the group values); hidden under the list loads
        RPROF_MARK(1)
        // 3. list entries: unique columns, plain writes
        if (CLUSTER) {   // entries of other CTAs' columns are theirs to apply
            if ((unsigned)(j0 - c_lo) < (unsigned)c_len) lower(j0, v0);
            if ((unsigned)(j1 - c_lo) < (unsigned)c_len) lower(j1, v1);
            for (int t = tid + 2 * NT; t < l_len; t += NT)
                if ((unsigned)(pi[t - tid] - c_lo) < (unsigned)c_len) lower(pi[t - tid], (unsigned long long)__double_as_longlong(pv[t - tid]));
        } else {
            if (j0 >= 0) lower(j0, v0);
            if (j1 >= 0) lower(j1, v1);
            for (int t = tid + 2 * NT; t < l_len; t += NT) lower(pi[t - tid], (unsigned long long)__double_as_longlong(pv[t - tid]));
        }
        // records copied into blk_rec by last step's refresh: landed long ago; waited for here, where none of this thread's
        // own loads is in flight (the wait is a scoreboard barrier that would also wait for those)
        if (lane == 0) cp_async_wait_all();
        FPROF(2)
        // 4. the selected flag of the row's frame: :182 assert, evaluated a phase late so that the flag load is off the chain
        if (tid == 0 && to_flag >= 0) {
            if (flag_check && flag_seen) {
                sh.err_frame = to_flag;
                sh.nd[par].y = 1;                       // read with the block count after the next barrier
            } else {
                sel_bytes[to_flag] = 1;
            }
            to_flag = -1;
        }
    };
    // Re-reduce one block: every lane PER entries, three REDUX; lane 0 stores the maximum and starts the asynchronous
    // copy of the best frame's record into the block's cache slot (valid once both halves carry the frame's tag).
    auto refresh_block = [&](int lb, int step) {
        const int base = (lb << SHIFT) + lane;
        unsigned hi[PER], lo[PER];
        int idx[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const unsigned long long v = fps_s[base + 32 * i];
            hi[i] = (unsigned)(v >> 32);
            lo[i] = (unsigned)v;
            idx[i] = base + 32 * i < c_len ? c_lo + base + 32 * i : 0x7fffffff;
        }
        if (PROFILE) asm volatile("" ::"r"(hi[0] + hi[PER - 1] + lo[0]));
        FPROF(5)
        const Win x = warp_argmax<PER>(hi, lo, idx);
        if (PROFILE) asm volatile("" ::"r"(x.idx));
        FPROF(6)
        if (lane == 0) {
            blk_val[lb] = ((unsigned long long)x.hi << 32) | x.lo;
            blk_idx[lb] = x.idx;
            blk_step[lb] = step;
            if (x.idx != 0x7fffffff) {
                cp_async16(blk_rec + 2 * lb, recs + x.idx);
                cp_async16(blk_rec + 2 * lb + 1, reinterpret_cast<const char *>(recs + x.idx) + 16);
            }
        }
    };

    bool local_abort = false;
    {
        const unsigned long long needed = a.list_cursor ? *a.list_cursor : 0ull;
        if (writer) a.result[5] = (double)needed;
        if ((long long)needed > a.list_capacity) {   // neighbour-list pool too small: the host grows it and reruns
            local_abort = true;
            if (writer) {
                a.result[1] = (double)ADSEL_ERR_CAPACITY;
                a.result[4] = -1.0;
            }
        }
    }
    if constexpr (CLUSTER) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&xs.xbar[0])), "r"(csize) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&xs.xbar[1])), "r"(csize) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cg::this_cluster().sync();   // every CTA is resident and its mbarriers are initialised before the first exchange
    }
    long long steps = 0;
    bool seeding = a.n_sampled > 0;   // first pick = argmax without a preceding update (:165)
    if (!local_abort) {
        if (a.forbid_sampled && tid == 0)
            for (int t = 0; t < a.n_sampled; ++t) sel_bytes[a.sampled[t]] = 1;
        if (a.n_sampled == 0) {   // empty buffer: the caller's random.choice is taken and costed up front (:168,172-174)
            cur = a.first_index;
            cur_blk = (unsigned)(cur - c_lo) < (unsigned)c_len ? (cur - c_lo) >> SHIFT : -1;
            uint4 lo, hi;
            ldg_rec_async(recs + cur, lo, hi);
            const double box_cost = take_rec(lo, hi);
            if (writer) out[0] = cur;
            m = 1;
            to_flag = cur;
            cost = __dadd_rn(cost, cost_f);
            cost = __dadd_rn(cost, box_cost);
        } else {
            marking = false;
            for (int t = 0; t < a.n_sampled; ++t) {   // fps = min over the rows of the frames already in the buffer (:160-164)
                cur = a.sampled[t];
                uint4 lo, hi;
                ldg_rec_async(recs + cur, lo, hi);
                take_rec(lo, hi);
                apply();
                __syncthreads();
            }
            marking = true;
            cur_blk = -1;
            // every block maximum and its cached record from scratch
            for (int lb = wid; lb < nblk; lb += NW) refresh_block(lb, -1);
            if (lane == 0) cp_async_wait_all();
            __syncthreads();
        }
        while (true) {
            if (!seeding) apply();
            FPROF(3)
            __syncthreads();
            RPROF_MARK(4)
            // Re-reduce (a) the row's own block, cooperatively: warp w takes its w-th 32 frames (one per lane, three REDUX) and
            // starts the asynchronous copy of that sub-block's best record -- half of all steps pick their next frame from
            // this very block; (b) rarely (4 % of the steps) other blocks whose maximum was lowered, one warp each, in full.
            const int4 nde = sh.nd[par];
            const int nd = nde.x;
            if (nde.y) {   // :182 assert -- the frame applied in this step had been selected before
                stop = 2;
                break;
            }
            if (PROFILE && tid == 0) {
                prof_acc[13] += nd > 0 ? 1 : 0;
                prof_acc[14] += nd;
            }
            const int ref_blk = cur_blk;   // -1: nothing was applied in this iteration (seeding)
            for (int sb = wid; sb < NSUB && ref_blk >= 0; sb += NW) {
                const int lj0 = (ref_blk << SHIFT) + SUBF * sb + lane;
                unsigned hi[SPL], lo[SPL];
                int idx[SPL];
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const int lj = lj0 + 32 * i;
                    const unsigned long long v = fps_s[lj];
                    hi[i] = (unsigned)(v >> 32);
                    lo[i] = (unsigned)v;
                    idx[i] = lj < c_len ? c_lo + lj : 0x7fffffff;
                }
                const Win x = warp_argmax<SPL>(hi, lo, idx);
                if (lane == 0) {
                    sh.sub_val[sb] = ((unsigned long long)x.hi << 32) | x.lo;
                    sh.sub_idx[sb] = x.idx;
                    if (x.idx != 0x7fffffff) {
                        cp_async_wait_all();   // the slot's previous copy (a step ago) has landed: no reordering of two copies
                        cp_async16(&sh.sub_rec[sb][0], recs + x.idx);
                        cp_async16(&sh.sub_rec[sb][1], reinterpret_cast<const char *>(recs + x.idx) + 16);
                    }
                }
            }
            if (wid > 0) {
                if (wid == 1 && nde.z && prev_blk >= 0 && prev_blk != ref_blk) refresh_block(prev_blk, (int)steps);
                for (int q = wid - 1; q < nd; q += NW - 1) refresh_block(dirty_list[q], (int)steps);
            }
            FPROF(7)
            __syncthreads();
            RPROF_MARK(8)
            if (tid == 0) sh.nd[par] = make_int4(0, 0, 0, 0);   // the next apply counts in the other word (zeroed a step ago)
            par ^= 1;
            // arg-max over the block maxima (the own block's entry replaced by its fresh sub-block maxima), by every warp:
            // no broadcast, no barrier
            unsigned shi = 0u, slo = 0u;
            int sidx = 0x7fffffff;
            if (ref_blk >= 0 && lane < NSUB) {
                const unsigned long long v = sh.sub_val[lane];
                shi = (unsigned)(v >> 32);
                slo = (unsigned)v;
                sidx = sh.sub_idx[lane];
            }
            Win y;
            y.hi = y.lo = 0u;
            y.idx = 0x7fffffff;
            for (int b0 = 0; b0 < nblk; b0 += QB * 32) {   // uniform trip count: the REDUX inside need every lane
                unsigned hi[QB + 1], lo[QB + 1];
                int idx[QB + 1];
#pragma unroll
                for (int q = 0; q < QB; ++q) {
                    const int b = b0 + lane + 32 * q;
                    const bool ok = b < nblk && b != ref_blk;
                    const unsigned long long v = ok ? blk_val[b] : 0ull;
                    hi[q] = (unsigned)(v >> 32);
                    lo[q] = (unsigned)v;
                    idx[q] = ok ? blk_idx[b] : 0x7fffffff;
                }
                hi[QB] = b0 == 0 ? shi : 0u;
                lo[QB] = b0 == 0 ? slo : 0u;
                idx[QB] = b0 == 0 ? sidx : 0x7fffffff;
                if (PROFILE) asm volatile("" ::"r"(hi[0] + hi[QB - 1] + idx[0] + idx[QB - 1]));
                FPROF(9)
                const Win w = warp_argmax<QB + 1>(hi, lo, idx);
                const bool take = w.hi > y.hi || (w.hi == y.hi && (w.lo > y.lo || (w.lo == y.lo && w.idx < y.idx)));
                if (take) y = w;
            }
            if constexpr (CLUSTER) {   // the CTAs' maxima -> the cluster's maximum, on every warp of every CTA
                // Warp 0 stores this CTA's (value, frame) record into EVERY CTA's shared memory and then arrives, with release
                // semantics at cluster scope, on that CTA's mbarrier of the step's parity; everybody waits (acquire) on the
                // own mbarrier, which completes when all csize records have landed, and reads them locally.  Records and
                // mbarriers are double-buffered by step parity: a step's arrivals can never be counted in another step's phase.
                const int xp = (int)(steps & 1);
                if (wid == 0 && lane < csize) {
                    unsigned dst_rec, dst_bar;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst_rec) : "r"((unsigned)__cvta_generic_to_shared(&xs.xrecs[xp][crank])), "r"(lane));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst_bar) : "r"((unsigned)__cvta_generic_to_shared(&xs.xbar[xp])), "r"(lane));
                    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst_rec), "r"(y.hi), "r"(y.lo), "r"((unsigned)y.idx), "r"(0u) : "memory");
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(dst_bar) : "memory");
                }
                {
                    const unsigned bar = (unsigned)__cvta_generic_to_shared(&xs.xbar[xp]);
                    const unsigned parity = (unsigned)((steps >> 1) & 1);
                    unsigned done = 0;
                    while (!done)
                        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                }
                unsigned hi1[1] = {0u}, lo1[1] = {0u};
                int idx1[1] = {0x7fffffff};
                if (lane < csize) {
                    const uint4 r = xs.xrecs[xp][lane];
                    hi1[0] = r.x;
                    lo1[0] = r.y;
                    idx1[0] = (int)r.z;
                }
                y = warp_argmax<1>(hi1, lo1, idx1);
            }
            FPROF(10)
            const int this_step = (int)steps;
            ++steps;
            const int nxt = y.idx;
            if (nxt < 0 || nxt >= n) {   // empty pool: cannot happen with n > 0
                if (writer) a.result[1] = (double)ADSEL_ERR_INVALID;
                stop = 2;
                break;
            }
            // The last warp folds the sub-block maxima into the own block's entry while the others go on (it is next read two
            // barriers from here; lower() does not consult it for this block during the next row, see prev_blk).
            if (wid == NW - 1 && ref_blk >= 0) {
                unsigned hi[1] = {shi}, lo[1] = {slo};
                int idx[1] = {sidx};
                const Win x = warp_argmax<1>(hi, lo, idx);
                if (lane == 0) {
                    blk_val[ref_blk] = ((unsigned long long)x.hi << 32) | x.lo;
                    blk_idx[ref_blk] = x.idx;
                    if (x.idx != 0x7fffffff) {
                        cp_async_wait_all();
                        cp_async16(blk_rec + 2 * ref_blk, recs + x.idx);
                        cp_async16(blk_rec + 2 * ref_blk + 1, reinterpret_cast<const char *>(recs + x.idx) + 16);
                    }
                }
            }
            // The winner's record.  From its own block's sub-block copies when the next frame comes from the block just
            // re-reduced: that copy was started before the barrier and is polled until both halves carry the frame's tag.
            // Else from the cache slot of its block (landed a step or more ago); a block re-reduced in full in THIS step
            // may still be receiving its record and is not trusted: one global round trip instead.
            prev_blk = ref_blk;
            cur = nxt;
            cur_blk = (unsigned)(nxt - c_lo) < (unsigned)c_len ? (nxt - c_lo) >> SHIFT : -1;   // -1: another CTA's frame
            uint4 rlo, rhi;
            if (cur_blk < 0) {   // CLUSTER only: one global round trip (the owner's cache is not visible from here)
                ldg_rec_async(recs + nxt, rlo, rhi);
            } else if (cur_blk == ref_blk) {
                const int w = ((nxt - c_lo) & (RB - 1)) / SUBF;
                while (true) {
                    rlo = lds128_volatile(&sh.sub_rec[w][0]);
                    rhi = lds128_volatile(&sh.sub_rec[w][1]);
                    if (rlo.w == (unsigned)nxt && rhi.w == (unsigned)nxt) break;
                    __nanosleep(20);   // 256 threads spinning would occupy the port the copy arrives through
                }
                if (PROFILE && tid == 0) prof_acc[15] += 1;
            } else {
                rlo = blk_rec[2 * cur_blk];
                rhi = blk_rec[2 * cur_blk + 1];
                if (blk_step[cur_blk] == this_step || rlo.w != (unsigned)nxt || rhi.w != (unsigned)nxt) {
                    ldg_rec_async(recs + nxt, rlo, rhi);
                    if (PROFILE && tid == 0) prof_acc[12] += 1;
                }
            }
            const double box_cost = take_rec(rlo, rhi);
            if (seeding) {   // :165,172-174 -- always taken, no budget check
                seeding = false;
                if (writer) out[0] = nxt;
                m = 1;
                cost = __dadd_rn(cost, cost_f);
                cost = __dadd_rn(cost, box_cost);
                to_flag = nxt;
                continue;
            }
            if (tid == 0) flag_seen = __ldcg(sel_bytes + nxt);   // consumed in the next apply, or below when the loop ends here
            cost = __dadd_rn(cost, cost_f);        // :184
            cost = __dadd_rn(cost, box_cost);      // :185
            const bool over = cost > budget;       // :186
            const bool full = !over && m >= out_cap;
            if (over || full) {
                if (writer) {
                    if (flag_seen) {               // :182 comes before the cost test
                        a.result[1] = (double)ADSEL_ERR_RESELECTED;
                        a.result[4] = (double)nxt;
                    } else if (full) {
                        a.result[1] = (double)ADSEL_ERR_CAPACITY;
                        a.result[4] = -2.0;
                    }
                }
                stop = over ? 1 : 2;
                break;
            }
            if (writer) out[m] = nxt;
            m += 1;
            to_flag = nxt;
            flag_check = true;
            FPROF(11)
        }
    }
    __syncthreads();
    if (lane == 0) cp_async_wait_all();
    if (stop == 2 && sh.err_frame >= 0 && writer) {
        a.result[1] = (double)ADSEL_ERR_RESELECTED;
        a.result[4] = (double)sh.err_frame;
    }
    for (int i = tid; i < c_len; i += NT) a.fps[c_lo + i] = fps_s[i];   // adsel_get_fps
    if (writer) {
        a.result[0] = cost;
        a.result[2] = (double)m;
        a.result[3] = (double)steps;
        if (PROFILE)
            for (int i = 0; i < 16; ++i) a.prof[i] = prof_acc[i];
    }
    if (CLUSTER) cg::this_cluster().sync();   // no CTA leaves while a peer may still read its record
}

// ---- host side of the resident variant -----------------------------------------------------------
struct ResidentPlan {
    int csize = 0, rb = 0, cta_cols = 0, nblk = 0, tl_n = 0, sel_words = 0;
    size_t smem = 0;
};

static size_t resident_fixed_bytes(long long cta_cols, long long nblk, bool cluster) {
    size_t b = (size_t)cta_cols * 8 + (size_t)nblk * 8;
    if (cluster) b += (size_t)(2 * CLUSTER_MAX * XSLOT_WORDS + CLUSTER_MAX) * 8;
    return b + (size_t)nblk * 12 + (size_t)nblk * sizeof(BlkRec);
}

template <int RB, bool CLUSTER, bool PROFILE>
static cudaError_t launch_resident(adsel_handle *h, const GreedyArgs &a, size_t smem, bool dry_run, int *max_clusters) {
    auto kernel = greedy_resident_kernel<RB, CLUSTER, PROFILE>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (CLUSTER && a.csize > 8) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)a.csize, 1, 1);
    cfg.blockDim = dim3(RES_NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CLUSTER ? 1 : 0;
    if (dry_run) {   // can a cluster of this size and shared-memory footprint be scheduled at all?
        *max_clusters = 1;
        return CLUSTER ? cudaOccupancyMaxActiveClusters(max_clusters, kernel, &cfg) : cudaSuccess;
    }
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

static cudaError_t dispatch_resident(adsel_handle *h, const GreedyArgs &a, const ResidentPlan &pl, bool dry_run, int *max_clusters) {
    if (pl.csize > 1) return launch_resident<128, true, false>(h, a, pl.smem, dry_run, max_clusters);
    if (a.prof) return pl.rb == 128 ? launch_resident<128, false, true>(h, a, pl.smem, dry_run, max_clusters)
                                    : launch_resident<256, false, true>(h, a, pl.smem, dry_run, max_clusters);
    return pl.rb == 128 ? launch_resident<128, false, false>(h, a, pl.smem, dry_run, max_clusters)
                        : launch_resident<256, false, false>(h, a, pl.smem, dry_run, max_clusters);
}

// Cluster size / block size / shared-memory layout of the resident variant; false when the columns do not fit
// (more than 16 CTAs' worth: the global-memory kernel above takes over) or ADSEL_GREEDY_CLUSTER=0 asks for that.
// ADSEL_GREEDY_CLUSTER=1/2/4/8/16 forces a cluster size (tests, measurements); unset: the smallest that fits.
static bool plan_resident(adsel_handle *h, GreedyArgs &a, int tl_want, ResidentPlan *out) {
    const char *e = getenv("ADSEL_GREEDY_CLUSTER");
    const int forced = (e && *e) ? atoi(e) : -1;
    if (forced == 0) return false;
    if (a.list_capacity >= (1ll << 40)) return false;
    const size_t budget = 232448 - 512;   // 227 KB opt-in maximum per CTA minus the static shared variables
    const long long cols = a.hi - a.lo;
    for (int c = 1; c <= CLUSTER_MAX; c *= 2) {
        if (forced > 0 && c != forced) continue;
        if (c > 1 && a.prof) continue;
        for (int rb = 128; rb <= 256; rb *= 2) {
            if (c > 1 && rb != 128) continue;
            long long per = (cols + c - 1) / c;
            per = (per + rb - 1) / rb * rb;
            if (per < rb) per = rb;
            const long long nblk = per / rb;
            const size_t fixed = resident_fixed_bytes(per, nblk, c > 1);
            if (fixed > budget) continue;
            size_t room = budget - fixed;
            ResidentPlan pl;
            pl.csize = c;
            pl.rb = rb;
            pl.cta_cols = (int)per;
            pl.nblk = (int)nblk;
            pl.tl_n = (int)((size_t)tl_want * 8 <= room ? tl_want : room / 8);
            room -= (size_t)pl.tl_n * 8;
            const size_t words = (size_t)(a.n + 31) / 32;
            pl.sel_words = words * 4 <= room ? (int)words : 0;
            pl.smem = fixed + (size_t)pl.tl_n * 8 + (size_t)pl.sel_words * 4;
            a.csize = pl.csize;
            a.cta_cols = pl.cta_cols;
            a.r_nblk = pl.nblk;
            a.tl_n = pl.tl_n;
            a.sel_words = pl.sel_words;
            int max_clusters = 0;
            if (dispatch_resident(h, a, pl, true, &max_clusters) != cudaSuccess || max_clusters < 1) {
                cudaGetLastError();
                continue;
            }
            *out = pl;
            return true;
        }
    }
    return false;
}


// ---- host side of the fast variant ---------------------------------------------------------------
struct FastPlan {
    int rb = 0, cols = 0, nblk = 0, g = 0, csize = 0;
    size_t smem = 0;
};

template <int RB, int NT, bool CLUSTER, bool PROFILE>
static cudaError_t launch_fast(adsel_handle *h, const GreedyArgs &a, size_t smem, bool dry_run, int *max_clusters) {
    auto kernel = greedy_fast_kernel<RB, NT, CLUSTER, PROFILE>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (CLUSTER && a.csize > 8) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(CLUSTER ? a.csize : 1), 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)(CLUSTER ? a.csize : 1);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CLUSTER ? 1 : 0;
    if (dry_run) {   // can a cluster of this size and shared-memory footprint be scheduled at all?
        *max_clusters = 1;
        return CLUSTER ? cudaOccupancyMaxActiveClusters(max_clusters, kernel, &cfg) : cudaSuccess;
    }
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

static int fast_threads() {   // ADSEL_FAST_THREADS=128/256: threads of a fast CTA (measurement aid; default 256)
    const char *e = getenv("ADSEL_FAST_THREADS");
    return (e && atoi(e) == 128) ? 128 : 256;
}

static cudaError_t dispatch_fast(adsel_handle *h, const GreedyArgs &a, const FastPlan &pl, bool dry_run, int *max_clusters) {
    const bool t128 = fast_threads() == 128;
    if (pl.csize > 1) {
        if (a.prof) return launch_fast<256, 256, true, true>(h, a, pl.smem, dry_run, max_clusters);
        return t128 ? launch_fast<256, 128, true, false>(h, a, pl.smem, dry_run, max_clusters)
                    : launch_fast<256, 256, true, false>(h, a, pl.smem, dry_run, max_clusters);
    }
    if (a.prof) return pl.rb == 128   ? launch_fast<128, 256, false, true>(h, a, pl.smem, dry_run, max_clusters)
                       : pl.rb == 256 ? launch_fast<256, 256, false, true>(h, a, pl.smem, dry_run, max_clusters)
                                      : launch_fast<512, 256, false, true>(h, a, pl.smem, dry_run, max_clusters);
    if (t128) return pl.rb == 128   ? launch_fast<128, 128, false, false>(h, a, pl.smem, dry_run, max_clusters)
                     : pl.rb == 256 ? launch_fast<256, 128, false, false>(h, a, pl.smem, dry_run, max_clusters)
                                    : launch_fast<512, 128, false, false>(h, a, pl.smem, dry_run, max_clusters);
    return pl.rb == 128   ? launch_fast<128, 256, false, false>(h, a, pl.smem, dry_run, max_clusters)
           : pl.rb == 256 ? launch_fast<256, 256, false, false>(h, a, pl.smem, dry_run, max_clusters)
                          : launch_fast<512, 256, false, false>(h, a, pl.smem, dry_run, max_clusters);
}

// The lean step (greedy_fast_kernel): no Euclid term, contiguous temporal groups, the whole pool on this device, its
// columns resident in the shared memory of ONE CTA when they fit (<= 28.4 k frames), else of the smallest thread-block
// cluster that holds them (2 / 4 / 8 / 16 CTAs: up to ~450 k frames).  ADSEL_GREEDY_FAST=0 disables it;
// ADSEL_GREEDY_CLUSTER=1/2/4/8/16 forces the cluster size.
static bool plan_fast(adsel_handle *h, GreedyArgs &a, bool sharded, int g_want, FastPlan *out) {
    const char *e = getenv("ADSEL_GREEDY_FAST");
    if (e && atoi(e) == 0) return false;
    const char *c = getenv("ADSEL_GREEDY_CLUSTER");
    const int forced = (c && *c) ? atoi(c) : -1;
    if (forced == 0) return false;               // the global-memory kernel was asked for
    if (sharded || a.spatial == ADSEL_SPATIAL_EUCLID || (a.temporal && !a.groups_contiguous)) return false;
    if (a.list_capacity >= (1ll << 40)) return false;
    const size_t budget = 232448 - 512;   // 227 KB opt-in maximum per CTA minus the static shared variables (FastShared: 400 B)
    for (int cs = 1; cs <= CLUSTER_MAX; cs *= 2) {
        if (forced > 0 && cs != forced) continue;
        const char *rbe = getenv("ADSEL_FAST_BLOCK");   // 128/256/512: frames per block (measurement aid)
        const int rb_forced = rbe ? atoi(rbe) : 0;
        for (int rb = 128; rb <= 512; rb *= 2) {
            if (cs > 1 && rb != 256) continue;
            long long cols = ((long long)a.n + cs - 1) / cs;
            cols = (cols + rb - 1) / rb * rb;
            const long long nblk = cols / rb;
            // one CTA: the smallest block size whose maxima are at most two per lane in the step's final arg-max (<= 64 blocks)
            if (cs == 1 && (rb_forced ? rb != rb_forced : (rb < 512 && nblk > 64))) continue;
            const size_t fixed = (size_t)cols * 8 + (size_t)nblk * (8 + 32 + 4 + 8 + 4) + 8;
            if (fixed > budget) continue;
            long long g = (long long)((budget - fixed) / 8);
            if (g > g_want) g = g_want;
            if ((nblk + g) & 1) g += ((size_t)(g + 1) * 8 <= budget - fixed) ? 1 : -1;   // cached records are 16-byte aligned
            if (g < 0) continue;
            FastPlan pl;
            pl.rb = rb;
            pl.cols = (int)cols;
            pl.nblk = (int)nblk;
            pl.g = (int)g;
            pl.csize = cs;
            pl.smem = fixed + (size_t)g * 8;
            a.csize = cs;
            a.cta_cols = pl.cols;
            a.r_nblk = pl.nblk;
            a.tl_n = pl.g;
            a.sel_words = 0;
            int max_clusters = 0;
            if (dispatch_fast(h, a, pl, true, &max_clusters) != cudaSuccess || max_clusters < 1) {
                cudaGetLastError();
                continue;
            }
            *out = pl;
            return true;
        }
    }
    return false;
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
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_fast_kernel<128, 256, false, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_fast_kernel<256, 256, false, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_fast_kernel<512, 256, false, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_fast_kernel<256, 256, true, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_resident_kernel<128, false, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_resident_kernel<256, false, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_resident_kernel<128, true, false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, temporal_lut_kernel));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, fill_u64_kernel2));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, frame_rec_kernel));
    ADSEL_TRY(warp_preload());
    return ADSEL_OK;
}

int validate_select(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
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
    if (!h->sampled_is_device)   // (a device-resident list is range-checked by a kernel in adsel_select_device)
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
    ADSEL_TRY(h->is_selected.reserve((size_t)n * CLUSTER_MAX));   // one private copy per CTA of a resident cluster
    ADSEL_TRY(h->tlut.reserve(n));
    ADSEL_TRY(h->d_sampled.reserve(n_sampled > 0 ? n_sampled : 1));
    ADSEL_TRY(h->d_out.reserve(out_cap));
    ADSEL_TRY(h->d_result.reserve(8));
    ADSEL_TRY(warp_prepare(h, p, false));   // (a replicated multi-rank selection runs the single-device kernels too)
    return ADSEL_OK;
}

static double far_temporal(const adsel_select_params *p) {
    return p->temporal ? host_normalize_far(p->normalize, TEMPORAL_MARGIN, p->max_temporal) : 0.0;
}

// The value every untouched entry of a row takes: aggregate(normalize_s(inf), normalize_t(1e6)) -- the largest value a
// row entry can have (normalize and aggregate are monotone, lambda_t >= 0).
double far_value(const adsel_select_params *p, double *s_far_out, double *t_far_out) {
    const double inf = host_inf();
    const double s_far = p->spatial == ADSEL_SPATIAL_GEODESIC ? host_normalize_far(p->normalize, inf, p->max_spatial)
                         : p->spatial == ADSEL_SPATIAL_EUCLID ? TEMPORAL_MARGIN : 0.0;
    const double t_far = far_temporal(p);
    double far;
    if (p->spatial != ADSEL_SPATIAL_NONE && p->temporal) {
        if (p->aggregate == ADSEL_AGG_SUM) far = s_far + p->lambda_t * t_far;   // unfused on the host (no -ffast-math)
        else if (p->aggregate == ADSEL_AGG_MIN) far = s_far < t_far ? s_far : t_far;
        else far = s_far > t_far ? s_far : t_far;
    } else {
        far = p->spatial != ADSEL_SPATIAL_NONE ? s_far : t_far;
    }
    if (s_far_out) *s_far_out = s_far;
    if (t_far_out) *t_far_out = t_far;
    return far;
}

// Enqueues temporal LUT -> neighbour lists (skipped when the resident ones match) on the handle's stream.
static int enqueue_lists(adsel_handle *h, const adsel_select_params *p, bool sharded, bool replicate = false) {
    const int n = h->n;
    if (p->temporal) {
        temporal_lut_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->tlut.p, n, p->normalize, p->max_temporal);
        h->launches += 1;
        ADSEL_CUDA(cudaGetLastError());
    }
    if (p->spatial == ADSEL_SPATIAL_GEODESIC)
        ADSEL_TRY(build_lists(h, p, far_temporal(p), sharded, replicate));
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
    const double far = far_value(p, &a.s_far, &a.t_far);   // the value every untouched entry of a row takes
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
        ADSEL_CUDA(cudaMemcpyAsync(h->d_sampled.p, sampled, sizeof(int32_t) * n_sampled, cudaMemcpyDefault, h->stream));   // host or device list
    ADSEL_CUDA(cudaMemsetAsync(h->d_result.p, 0, sizeof(double) * 8, h->stream));
    a.sel_words = n <= 524288 ? (n + 31) / 32 : 0;
    a.far_bits = far_bits;
    CommView view;
    if (sharded) ADSEL_TRY(comm_view(h, &view));
    ResidentPlan plan;
    FastPlan fast;
    const int tl_want = a.temporal ? (p->normalize == ADSEL_NORM_EXP && a.tl_n > 64 ? 64 : a.tl_n) : 0;
    const int tl_all = a.tl_n, sel_all = a.sel_words;
    bool warp_launched = false;
    ADSEL_TRY(warp_enqueue(h, p, a, sharded, &warp_launched));   // one warp on resident columns (greedy_warp.cu)
    if (warp_launched) {
    } else if (plan_fast(h, a, sharded, tl_want, &fast)) {   // resident in one CTA or a cluster, lean step
        ADSEL_CUDA(cudaMemsetAsync(h->is_selected.p, 0, (size_t)n * fast.csize, h->stream));
        int unused = 0;
        ADSEL_CUDA(dispatch_fast(h, a, fast, false, &unused));
        h->greedy_cluster = fast.csize;
        h->greedy_variant = 1;
    } else if ((a.tl_n = tl_all, a.sel_words = sel_all, !sharded) && plan_resident(h, a, tl_want, &plan)) {   // fps resident in shared memory (one CTA or a cluster)
        ADSEL_CUDA(cudaMemsetAsync(h->is_selected.p, 0, (size_t)n * plan.csize, h->stream));
        int unused = 0;
        ADSEL_CUDA(dispatch_resident(h, a, plan, false, &unused));
        h->greedy_cluster = plan.csize;
        h->greedy_variant = 2;
    } else {
        ADSEL_CUDA(cudaMemsetAsync(h->is_selected.p, 0, n, h->stream));
        fill_u64_kernel2<<<h->num_sms * 2, 256, 0, h->stream>>>(h->fps.p, (size_t)n, far_bits);
        h->launches += 1;
        ADSEL_CUDA(cudaGetLastError());
        a.csize = 0;
        a.tl_n = a.temporal ? (h->max_group < 4096 ? h->max_group : 4096) : 0;
        a.sel_words = n <= 524288 ? (n + 31) / 32 : 0;
        const size_t smem = (size_t)nblk * 20 + (size_t)a.tl_n * 8 + (size_t)a.sel_words * 4 + 16;
        if (sharded) launch_greedy<true>(h, a, view, smem);
        else launch_greedy<false>(h, a, view, smem);
        h->greedy_cluster = 0;
        h->greedy_variant = 0;
    }
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    return ADSEL_OK;
}

int finish_select(adsel_handle *h, const double *result, int32_t *out_idx, int32_t *out_m, double *out_cost,
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
    if (lazy_wanted(h, p))   // rows without a saturation radius on a large pool: one search per selected frame (lazy.cu)
        return lazy_select(h, p, sampled, n_sampled, first_index, out_idx, out_cap, out_m, out_cost);
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
            unsigned long long prof[16];
            cudaMemcpy(prof, greedy_prof_ptr(), sizeof(prof), cudaMemcpyDeviceToHost);
            const double steps = result[3] > 0 ? result[3] : 1;
            if (h->greedy_variant == 5 || h->greedy_variant == 6)
                fprintf(stderr, "[warp profile] cycles/step: issue %.0f | finalise %.0f | spec group %.0f | slot wait %.0f | accept+group %.0f | list %.0f | "
                                "dirty blocks %.0f | loads+tree %.0f | redux %.0f\n",
                        prof[6] / steps, prof[7] / steps, prof[0] / steps, prof[1] / steps, prof[8] / steps, prof[2] / steps, prof[3] / steps,
                        prof[4] / steps, prof[5] / steps);
            if (h->greedy_variant == 1)
                fprintf(stderr, "[fast profile] cycles/step: issue+group %.0f | bar0 %.0f | list lower %.0f | pend+flags %.0f | bar1 %.0f | "
                                "refresh loads %.0f | refresh argmax %.0f | refresh store %.0f | bar2 %.0f | reduce loads %.0f | reduce argmax %.0f | bookkeeping %.0f\n",
                        prof[0] / steps, prof[1] / steps, prof[2] / steps, prof[3] / steps, prof[4] / steps, prof[5] / steps, prof[6] / steps,
                        prof[7] / steps, prof[8] / steps, prof[9] / steps, prof[10] / steps, prof[11] / steps);
            if (h->greedy_variant == 1)
                fprintf(stderr, "[fast profile] record-cache misses %.3f per step, next frame from the own block %.3f; steps with other blocks to re-reduce %.3f, blocks per step %.3f\n",
                        prof[12] / steps, prof[15] / steps, prof[13] / steps, prof[14] / steps);
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
// replicate: the lists are built by the rank that owns the SOURCE frames and stored into every rank's arena; each rank
// then runs the whole greedy loop (resident in its shared memory) on complete lists -- no exchange per step, the same
// result on every rank (like the reference, where every rank computes the whole selection, base_selector.py:57).
int greedy_select_sharded_begin(adsel_handle *h, const adsel_select_params *p, const int32_t *sampled, int32_t n_sampled,
                                int32_t first_index, int32_t out_cap, bool replicate) {
    ADSEL_TRY(validate_select(h, p, sampled, n_sampled, first_index, out_cap));
    ADSEL_REQUIRE(h->comm_ready, "adsel_select_sharded_begin: call adsel_comm_init / adsel_comm_attach first");
    ADSEL_REQUIRE(!h->pending, "adsel_select_sharded_begin: a sharded selection is already pending");
    ADSEL_TRY(greedy_prepare(h, p, n_sampled, out_cap, true));
    ADSEL_CUDA(cudaEventRecord(h->pending_ev[0], h->stream));
    ADSEL_TRY(enqueue_lists(h, p, true, replicate));
    ADSEL_TRY(enqueue_greedy(h, p, sampled, n_sampled, first_index, out_cap, !replicate));
    h->pending_replicated = replicate;
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
    if (h->pending_replicated) {
        // no mailbox exchange ran: the tags stay where they are
    } else if (status == 0 || status == ADSEL_ERR_RESELECTED) h->comm_seq += (unsigned)result[3];
    else h->comm_seq = ((h->comm_seq >> 24) + 1u) << 24;
    // an overflowing list region of the arena cannot grow in place: the caller re-creates the arenas with
    // adsel_comm_init(list_cap = the largest adsel_get_counters()[2] over the ranks) and rebuilds
    if (status != 0) h->lists_mode = -1;
    return finish_select(h, result, out_idx, out_m, out_cost, "adsel_select_sharded");
}

}  // namespace adsel
