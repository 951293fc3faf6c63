// greedy_warp.cu -- the greedy k-center loop (spatial_temporal_selector.py:157-193) as ONE WARP on resident columns.
//
// What a step of the headline configuration touches is tiny: the selected frame's truncated geodesic neighbourhood
// (24 entries on average at 28,130 frames, 3.5 % of the rows longer than 64), at most 2 x 38 - 1 frames of its temporal
// group (the exp-normalised LUT saturates at 38 frames), one 512-frame block to re-reduce and 55 block maxima.  The
// multi-warp kernels (greedy.cu) spend their step on keeping 8 warps in step: three CTA barriers, a record cache so that
// the row's list address is known without a second L2 round trip, redundant reductions in every warp -- 2,670 cycles
// per step of which ~350 are the one L2 round trip.  Here
//   * one warp owns the loop: no CTA barrier, no broadcast, every piece of state is a register of all 32 lanes;
//   * the frame record and the first 64 entries of its neighbour list sit in ONE fixed-size slot per frame
//     (slot_build_kernel, 800 B): address = frame x 800, so the step's only L2 round trip starts the moment the
//     winner is known -- no record cache, no cp.async, no polling.  Longer lists continue in the list pool (a second,
//     dependent round trip, on 3.5 % of the steps);
//   * everything that does not need the slot runs in the shadow of its load: the candidates of the temporal group
//     (index window around the winner, values from a shared-memory LUT, the comparison with fps and the "is this
//     the block's maximum" test) are evaluated speculatively and only masked by the group bounds when the header
//     arrives; cost + cost_f is formed there too (the box cost n_boxes * cost_b comes rounded from the slot);
//   * the final arg-max is one pass: 16 entries of the row's own block + 2 block maxima per lane, a per-lane
//     compare/select tree, then TWO interleaved chains of three REDUX: the step's winner, and the own block's new
//     maximum (stored for the steps in which the block is no longer the row's own) -- the second chain hides in the
//     latency of the first.
// Bit-exact like the other kernels: the same minima, np.argmax's first-maximum rule, the same rounded cost sums.
#include "common.cuh"

namespace adsel {

constexpr int WARP_K = 64;                          // list entries held in a slot (two per lane)
constexpr int WARP_SLOT = 32 + WARP_K * 12;         // header {list_off, list_len, frame | e0, e1, box cost f64} | idx[K] int32 | val[K] float64
#ifndef ADSEL_WARP_RB
#define ADSEL_WARP_RB 512
#endif
constexpr int WARP_RB = ADSEL_WARP_RB, WARP_SHIFT = WARP_RB == 512 ? 9 : 8, WARP_PER = WARP_RB / 32;   // frames per block
constexpr int WARP_NB = WARP_RB == 512 ? 2 : 4;     // block maxima per lane in the final arg-max: up to 64 / 128 blocks
static_assert(WARP_RB == 512 || WARP_RB == 256, "block size");
constexpr int WARP_SPEC = 3;                        // speculative group candidates per lane: window of 96 >= 2 * sat - 1

__global__ void slot_build_kernel(unsigned char *__restrict__ slots, const FrameRec *__restrict__ rec,
                                  const int32_t *__restrict__ lidx, const double *__restrict__ lval, int n, long long capacity,
                                  const unsigned long long *__restrict__ cursor, double cost_b) {
    const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n) return;
    const FrameRec r = rec[w];
    unsigned char *s = slots + (size_t)w * WARP_SLOT;
    if (lane == 0) {
        *reinterpret_cast<uint4 *>(s) = make_uint4((unsigned)r.list_off, (unsigned)((unsigned long long)r.list_off >> 32),
                                                    (unsigned)r.list_len, (unsigned)w);
        // n_boxes * cost_b as ONE rounded product (base_selector.py:85), formed here so that the greedy step adds it as is
        const unsigned long long bc = (unsigned long long)__double_as_longlong(__dmul_rn((double)r.n_boxes, cost_b));
        *reinterpret_cast<uint4 *>(s + 16) = make_uint4((unsigned)r.e0, (unsigned)r.e1, (unsigned)bc, (unsigned)(bc >> 32));
    }
    if (lidx == nullptr || (cursor && (long long)*cursor > capacity)) return;   // overflowed pool: the greedy kernel reports it
    int32_t *si = reinterpret_cast<int32_t *>(s + 32);
    double *sv = reinterpret_cast<double *>(s + 32 + 4 * WARP_K);
    const int len = min(r.list_len, WARP_K);
    for (int t = lane; t < len; t += 32) {
        si[t] = lidx[r.list_off + t];
        sv[t] = lval[r.list_off + t];
    }
}

__device__ __forceinline__ uint4 ldg_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_s32(const void *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long ldg_u64(const void *p) {
    unsigned long long v;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// shared memory by 32-bit shared-window address; a conditional store is ONE predicated STS
__device__ __forceinline__ unsigned long long lds64(unsigned addr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64_if(unsigned addr, unsigned long long v, bool p) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.shared.u64 [%0], %1;\n}" ::"r"(addr), "l"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void sts32_if(unsigned addr, int v, bool p) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.shared.s32 [%0], %1;\n}" ::"r"(addr), "r"(v), "r"((unsigned)p) : "memory");
}

struct Cand {   // a 64-bit key with its frame
    unsigned long long k;
    int i;
};
// b comes after a in index order: it wins only with a strictly larger key (np.argmax keeps the first maximum)
__device__ __forceinline__ Cand later_wins_if_larger(const Cand a, const Cand b) { return b.k > a.k ? b : a; }
__device__ __forceinline__ Cand better_of(const Cand a, const Cand b) { return (b.k > a.k || (b.k == a.k && b.i < a.i)) ? b : a; }

#define WPROF(slot) WPROF_DEP(slot, 0)
// the clock is read after `dep` is available (a register the preceding phase produces)
#define WPROF_DEP(slot, dep)                                                                                         \
    if (PROFILE) {                                                                                                   \
        long long now;                                                                                               \
        asm volatile("{\n.reg .b32 t;\nmov.b32 t, %1;\nmov.u64 %0, %%clock64;\n}" : "=l"(now) : "r"((int)(dep)) : "memory"); \
        prof_acc[slot] += (unsigned long long)(now - prof_t);                                                        \
        prof_t = now;                                                                                                \
    }

template <bool PROFILE>
__global__ void __launch_bounds__(32, 1) greedy_warp_kernel(const GreedyArgs a, const unsigned char *__restrict__ slots) {
    constexpr int SHIFT = WARP_SHIFT, PER = WARP_PER, K = WARP_K;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    unsigned long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = PROFILE ? clock64() : 0;

    const int n = a.n, cols = a.cta_cols, nblk = a.r_nblk, G = a.tl_n, out_cap = a.out_cap;
    const int32_t *const lidx = a.list_idx;
    const double *const lval = a.list_val;
    int32_t *const out = a.out;
    unsigned char *const sel_bytes = a.is_selected;
    const double budget = a.budget, cost_f = a.cost_f;
    const bool has_group = a.temporal != 0;

    extern __shared__ unsigned long long warp_smem[];
    unsigned long long *fps_s = warp_smem;              // [cols]   min distance to the selected set, ordered-u64 keys
    unsigned long long *blk_val = fps_s + cols;         // [nblk]   maximum of every 512-frame block
    unsigned long long *gval = blk_val + nblk;          // [G]      row value of a group member at index gap d
    int *blk_idx = reinterpret_cast<int *>(gval + G);   // [nblk]   frame holding it (lowest index among equals)

    const unsigned long long far_bits = a.far_bits;
    for (int i = lane; i < cols; i += 32) fps_s[i] = i < n ? far_bits : 0ull;
    for (int d = lane; d < G; d += 32)
        gval[d] = (unsigned long long)__double_as_longlong(combine(a.spatial, a.temporal, a.aggregate, a.lambda_t, a.s_far, a.tlut[d]));
    for (int lb = lane; lb < nblk; lb += 32) {   // every entry starts at `far`: a block's best is its first frame
        const bool nonempty = (lb << SHIFT) < n;
        blk_val[lb] = nonempty ? far_bits : 0ull;
        blk_idx[lb] = nonempty ? (lb << SHIFT) : 0x7fffffff;
    }
    // Saturation of the temporal LUT: for d >= sat, tlut[d] equals the far value bit for bit (exp: 38).
    int sat = 0x3fffffff;
    if (has_group) {
        const int lim = min(n, 4096);
        const unsigned long long tfar = (unsigned long long)__double_as_longlong(a.t_far);
        int mx = 0;
        for (int d = lane; d < lim; d += 32)
            if ((unsigned long long)__double_as_longlong(a.tlut[d]) != tfar) mx = d + 1;
        mx = __reduce_max_sync(FULL, mx);
        sat = mx >= lim ? 0x3fffffff : mx;
    }
    const bool spec_group = has_group && 2 * sat - 1 <= 32 * WARP_SPEC && sat <= G;
    __syncwarp();

    double cost = a.cost_start;
    int m = 0, stop = 0, err_frame = -1;
    long long steps = 0;
    bool marking = true;

    // Re-reduce one block from shared memory (other blocks whose maximum was lowered: 2 % of the steps; seeding).
    auto refresh_block = [&](int lb) {
        const int base = (lb << SHIFT) + lane;
        unsigned hi[PER], lo[PER];
        int idx[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const unsigned long long v = fps_s[base + 32 * i];
            hi[i] = (unsigned)(v >> 32);
            lo[i] = (unsigned)v;
            idx[i] = base + 32 * i < n ? base + 32 * i : 0x7fffffff;
        }
        const Win x = warp_argmax<PER>(hi, lo, idx);
        if (lane == 0) {
            blk_val[lb] = ((unsigned long long)x.hi << 32) | x.lo;
            blk_idx[lb] = x.idx;
        }
    };

    bool local_abort = false;
    {
        const unsigned long long needed = a.list_cursor ? *a.list_cursor : 0ull;
        if (lane == 0) a.result[5] = (double)needed;
        if ((long long)needed > a.list_capacity) {   // neighbour-list pool too small: the host grows it and reruns
            local_abort = true;
            if (lane == 0) {
                a.result[1] = (double)ADSEL_ERR_CAPACITY;
                a.result[4] = -1.0;
            }
        }
    }

    // mode of the frame at the top of the loop: 0 = row of a frame already in the buffer (applied only), 1 = taken
    // unconditionally (the caller's first frame / the first pick after seeding, :165-174), 2 = candidate (:182-189)
    int mode = 1, cur = a.first_index, t_sampled = 0;
    if (!local_abort) {
        if (a.forbid_sampled && lane == 0)
            for (int t = 0; t < a.n_sampled; ++t) sel_bytes[a.sampled[t]] = 1;
        if (a.n_sampled > 0) {
            mode = 0;
            cur = a.sampled[0];
            marking = false;
        }
    }
    int cur_blk = cur >> SHIFT, prev_blk = -1;
    bool cur_key_zero = false;
    // Shared memory through explicit 32-bit shared-window addresses: generic pointers make the compiler rebuild the window
    // base (S2UR SR_CgaCtaId + ULEA) in front of every access, and `if (...) smem[j] = v` becomes a divergent branch;
    // here a conditional store is one predicated STS and the loop body is straight-line code.
    unsigned s_fps = (unsigned)__cvta_generic_to_shared(fps_s), s_bval = (unsigned)__cvta_generic_to_shared(blk_val),
             s_gval = (unsigned)__cvta_generic_to_shared(gval), s_bidx = (unsigned)__cvta_generic_to_shared(blk_idx);
    asm volatile("" : "+r"(s_fps), "+r"(s_bval), "+r"(s_gval), "+r"(s_bidx));   // opaque: kept in registers, not rebuilt per access
    const int sat1 = sat - 1;

    while (!local_abort) {
        unsigned long long dirty = 0ull, dirty1 = 0ull;   // blocks 0..63 / 64..127 whose maximum was lowered
        auto mark = [&](int lb) {
            const unsigned long long bit = 1ull << (lb & 63);
            if (WARP_NB == 2 || lb < 64) dirty |= bit;
            else dirty1 |= bit;
        };
        // ---- 1. the slot: header + the first 64 list entries, one L2 round trip ----
        const unsigned char *sp = slots + (size_t)cur * WARP_SLOT;
        const uint4 h0 = ldg_v4(sp), h1 = ldg_v4(sp + 16);
        const int j0 = ldg_s32(sp + 32 + 4 * lane), j1 = ldg_s32(sp + 32 + 4 * (lane + 32));
        const unsigned long long v0 = ldg_u64(sp + 32 + 4 * K + 8 * lane), v1 = ldg_u64(sp + 32 + 4 * K + 8 * (lane + 32));
        const double cost_pre = __dadd_rn(cost, cost_f);   // :184
        unsigned flag_seen = 0;
        // A frame that was selected before has fps == 0 (its own row holds 0 for itself), so only a winner whose key is 0 can be a
        // re-selection (:182): the flag is read on those steps only (pool exhausted, duplicates of a selected frame).
        if (mode == 2 && cur_key_zero) flag_seen = __ldcg(sel_bytes + cur);
        WPROF(6)
        // ---- 2. in its shadow ----
        WPROF(7)
        // (b) temporal group, speculatively: the frames within the LUT's reach of cur, whatever the group's bounds are
        int gj[WARP_SPEC];
        unsigned long long gv[WARP_SPEC];
        bool gneed[WARP_SPEC] = {false, false, false}, gmax[WARP_SPEC] = {false, false, false};
        if (spec_group) {
            unsigned long long old[WARP_SPEC];
            int bi[WARP_SPEC];
#pragma unroll
            for (int q = 0; q < WARP_SPEC; ++q) {
                const int j = cur - sat1 + lane + 32 * q;
                const int d = abs(j - cur);
                const bool valid = (unsigned)j < (unsigned)n && d < sat;
                gj[q] = valid ? j : cur;
                gv[q] = lds64(s_gval + 8u * (valid ? d : 0));
                old[q] = lds64(s_fps + 8u * gj[q]);
                bi[q] = lds32(s_bidx + 4u * (gj[q] >> SHIFT));
                gneed[q] = valid;
            }
#pragma unroll
            for (int q = 0; q < WARP_SPEC; ++q) {
                gneed[q] = gneed[q] && gv[q] < old[q];
                gmax[q] = marking && (gj[q] >> SHIFT) != cur_blk && bi[q] == gj[q];
            }
        }
        WPROF_DEP(0, (int)gneed[0] + (int)gmax[WARP_SPEC - 1])
        // ---- 3. the header is here ----
        // (ptxas places the first use of the header right after the loads' issue; forcing it behind the shadow work with a
        //  run-time-zero dependency was measured slower, 1.03 vs 0.96 us per step, and is not done)
        const long long l_off = (long long)(((unsigned long long)h0.y << 32) | h0.x);
        const int l_len = (int)h0.z, e0 = (int)h1.x, e1 = (int)h1.y;
        const double box_cost = __longlong_as_double((long long)(((unsigned long long)h1.w << 32) | h1.z));   // n_boxes * cost_b (slot_build_kernel)
        asm volatile("" ::"r"(h0.w));   // (keeps the unused word's register from being recycled while the load is in flight)
        WPROF_DEP(1, l_len + e0 + j0 + j1 + (int)v0 + (int)v1)
        if (mode == 2) {
            cost = __dadd_rn(cost_pre, box_cost);  // :184-185 (cost + cost_f was formed while the slot was in flight)
            const bool over = cost > budget;       // :186
            const bool full = !over && m >= out_cap;
            unsigned verdict = (flag_seen != 0 ? 1u : 0u) | (over ? 2u : 0u) | (full ? 4u : 0u);
            asm volatile("" : "+r"(verdict));      // opaque: cicc 12.9 was seen to drop the capacity term from a fused condition
            if (verdict != 0u) {
                if (verdict & 1u) {                // :182 comes before the cost test
                    err_frame = cur;
                    stop = 2;
                } else {
                    if ((verdict & 4u) && lane == 0) {
                        a.result[1] = (double)ADSEL_ERR_CAPACITY;
                        a.result[4] = -2.0;
                    }
                    stop = (verdict & 2u) ? 1 : 2;
                }
                break;
            }
        } else if (mode == 1) {                    // :168,172-174 -- taken and costed up front, no budget check
            cost = __dadd_rn(cost_pre, box_cost);
        }
        if (mode != 0) {
            if (lane == 0) {
                out[m] = cur;
                sel_bytes[cur] = 1;
            }
            m += 1;
        }
        // group entries first, list entries second: list values of same-group frames are <= the group values
        if (spec_group) {
#pragma unroll
            for (int q = 0; q < WARP_SPEC; ++q) {
                const bool w = gneed[q] && gj[q] >= e0 && gj[q] < e1;
                sts64_if(s_fps + 8u * gj[q], gv[q], w);
                if (w && gmax[q]) mark(gj[q] >> SHIFT);
            }
        } else if (has_group) {
            const int g0 = max(e0, cur - sat + 1), g1 = min(e1, cur + sat);
            for (int j = g0 + lane; j < g1; j += 32) {
                const int d = abs(j - cur);
                const unsigned long long v = d < G ? gval[d]
                    : (unsigned long long)__double_as_longlong(combine(a.spatial, a.temporal, a.aggregate, a.lambda_t, a.s_far, a.tlut[d]));
                if (v < fps_s[j]) {
                    fps_s[j] = v;
                    const int lb = j >> SHIFT;
                    if (lb != cur_blk && marking && blk_idx[lb] == j) mark(lb);
                }
            }
        }
        __syncwarp();
        WPROF_DEP(8, (int)dirty)
        {   // the slot's 64 entries: two per lane, straight-line
            const bool p0 = lane < l_len, p1 = lane + 32 < l_len;
            const int c0 = p0 ? j0 : cur, c1 = p1 ? j1 : cur;
            const unsigned long long o0 = lds64(s_fps + 8u * c0), o1 = lds64(s_fps + 8u * c1);
            const int b0 = lds32(s_bidx + 4u * (c0 >> SHIFT)), b1 = lds32(s_bidx + 4u * (c1 >> SHIFT));
            const bool w0 = p0 && v0 < o0, w1 = p1 && v1 < o1;
            sts64_if(s_fps + 8u * c0, v0, w0);
            sts64_if(s_fps + 8u * c1, v1, w1);
            if (w0 && marking && (c0 >> SHIFT) != cur_blk && b0 == c0) mark(c0 >> SHIFT);
            if (w1 && marking && (c1 >> SHIFT) != cur_blk && b1 == c1) mark(c1 >> SHIFT);
        }
        if (l_len > K) {   // rows longer than a slot continue in the list pool (3.5 % of the rows)
            for (int t = K + lane; t < l_len; t += 32) {
                const int j = lidx[l_off + t];
                const unsigned long long v = (unsigned long long)__double_as_longlong(lval[l_off + t]);
                if (v < fps_s[j]) {
                    fps_s[j] = v;
                    const int lb = j >> SHIFT;
                    if (lb != cur_blk && marking && blk_idx[lb] == j) mark(lb);
                }
            }
        }
        __syncwarp();
        WPROF_DEP(2, (int)dirty)
        if (mode == 0) {   // next row of the buffer; after the last one every block maximum from scratch, then the first pick
            t_sampled += 1;
            if (t_sampled < a.n_sampled) {
                cur = a.sampled[t_sampled];
                cur_blk = cur >> SHIFT;
                continue;
            }
            marking = true;
            for (int lb = 0; lb < nblk; ++lb) refresh_block(lb);
            __syncwarp();
            cur_blk = -1;
            prev_blk = -1;
        }
        // ---- 4. other blocks whose maximum was lowered (2 % of the steps) ----
        if (__any_sync(FULL, (dirty | dirty1) != 0ull)) {
            unsigned long long words[2] = {dirty, dirty1};
#pragma unroll
            for (int w = 0; w < (WARP_NB == 2 ? 1 : 2); ++w) {
                unsigned dlo = __reduce_or_sync(FULL, (unsigned)words[w]), dhi = __reduce_or_sync(FULL, (unsigned)(words[w] >> 32));
                while (dlo) {
                    refresh_block(64 * w + __ffs(dlo) - 1);
                    dlo &= dlo - 1;
                }
                while (dhi) {
                    refresh_block(64 * w + 32 + __ffs(dhi) - 1);
                    dhi &= dhi - 1;
                }
            }
            __syncwarp();
        }
        WPROF(3)
        // ---- 5. arg-max: the row's own block from its 512 entries, every other block from its maximum ----
        Cand bm[WARP_NB];
#pragma unroll
        for (int q = 0; q < WARP_NB; ++q) {
            const int b = lane + 32 * q;
            const bool ok = b < nblk && b != cur_blk;
            const unsigned long long k = lds64(s_bval + 8u * (ok ? b : 0));
            const int i = lds32(s_bidx + 4u * (ok ? b : 0));
            bm[q].k = ok ? k : 0ull;
            bm[q].i = ok ? i : 0x7fffffff;
        }
        Cand own{0ull, 0x7fffffff};
        if (cur_blk >= 0) {
            const int base = (cur_blk << SHIFT) + lane;
            Cand c[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                c[i].k = lds64(s_fps + 8u * (base + 32 * i));
                c[i].i = base + 32 * i;
            }
#pragma unroll
            for (int w = 1; w < PER; w *= 2)
#pragma unroll
                for (int i = 0; i + w < PER; i += 2 * w) c[i] = later_wins_if_larger(c[i], c[i + w]);
            own = c[0];
            if (own.i >= n) own.i = 0x7fffffff;   // padding of the last block (key 0): never before a real frame
        }
        Cand best = better_of(bm[0], bm[1]);
        if (WARP_NB == 4) best = better_of(best, better_of(bm[2], bm[3]));
        best = better_of(own, best);
        WPROF_DEP(4, (int)best.k + best.i)
        // Two arg-max chains interleaved, three REDUX each: (A) the step's winner over everything, (B) the own block's new
        // maximum over its 512 entries, stored for the steps in which the block is no longer the row's own.
        Win y;
        {
            const unsigned ah = (unsigned)(best.k >> 32), al = (unsigned)best.k, bh = (unsigned)(own.k >> 32), bl = (unsigned)own.k;
            const unsigned mah = __reduce_max_sync(FULL, ah);
            const unsigned mbh = __reduce_max_sync(FULL, bh);
            const unsigned mal = __reduce_max_sync(FULL, ah == mah ? al : 0u);
            const unsigned mbl = __reduce_max_sync(FULL, bh == mbh ? bl : 0u);
            y.idx = __reduce_min_sync(FULL, (ah == mah && al == mal) ? best.i : 0x7fffffff);
            const int bidx = __reduce_min_sync(FULL, (bh == mbh && bl == mbl) ? own.i : 0x7fffffff);
            y.hi = mah;
            y.lo = mal;
            sts64_if(s_bval + 8u * max(cur_blk, 0), ((unsigned long long)mbh << 32) | mbl, lane == 0 && cur_blk >= 0);
            sts32_if(s_bidx + 4u * max(cur_blk, 0), bidx, lane == 0 && cur_blk >= 0);
        }
        WPROF_DEP(5, y.idx)
        ++steps;
        const int nxt = y.idx;
        cur_key_zero = (y.hi | y.lo) == 0u;
        if (nxt < 0 || nxt >= n) {   // empty pool: cannot happen with n > 0
            if (lane == 0) a.result[1] = (double)ADSEL_ERR_INVALID;
            stop = 2;
            break;
        }
        prev_blk = cur_blk;
        mode = mode == 0 ? 1 : 2;
        cur = nxt;
        cur_blk = nxt >> SHIFT;
    }
    __syncwarp();
    // the last row's own block never got its maximum; adsel_get_fps reads the values only
    if (stop == 2 && err_frame >= 0 && lane == 0) {
        a.result[1] = (double)ADSEL_ERR_RESELECTED;
        a.result[4] = (double)err_frame;
    }
    for (int i = lane; i < n; i += 32) a.fps[i] = fps_s[i];
    if (lane == 0) {
        a.result[0] = cost;
        a.result[2] = (double)m;
        a.result[3] = (double)steps;
        if (PROFILE)
            for (int i = 0; i < 12; ++i) a.prof[i] = prof_acc[i];
    }
}

// ---------------------------------------------------------------------------------------------
// The same loop for pools beyond one CTA's shared memory: ONE warp of CTA 0 runs it, the other CTAs of a thread-block
// cluster (2 / 4 / 8 / 16) are PASSIVE MEMORY -- each holds a slice of the min-distance vector in its shared memory and
// waits at the cluster barrier; the warp reaches the slices through distributed shared memory (mapa + ld/st.shared::cluster).
// The multi-warp cluster kernel (greedy.cu) exchanges every CTA's maximum once per step (~1,100 cycles of a 2.2-2.5 us
// step); here a step pays two exposed remote round trips instead (the list entries' and the own block's loads; the group
// window's loads hide under the slot load) and no exchange.  Block maxima live in CTA 0; with up to 768 blocks they get
// a second level: super-blocks of 32 blocks (one per lane in the final arg-max, next to the own super-block's 31 sibling
// blocks, one per lane) whose maxima are kept by a third interleaved REDUX chain.
// ---------------------------------------------------------------------------------------------
constexpr int WC_DIRTY_CAP = 1024;   // blocks whose maximum was lowered in one step (listed; beyond: every block is re-reduced)

__device__ __forceinline__ unsigned long long ldc64(unsigned addr) {
    unsigned long long v;
    asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void stc64_if(unsigned addr, unsigned long long v, bool p) {
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %2, 0;\n@q st.shared::cluster.u64 [%0], %1;\n}" ::"r"(addr), "l"(v), "r"((unsigned)p) : "memory");
}

template <bool PROFILE>
__global__ void __launch_bounds__(32, 1) greedy_warp_cluster_kernel(const GreedyArgs a, const unsigned char *__restrict__ slots,
                                                                     int b0, int bpc, unsigned magic) {
    constexpr int SHIFT = 9, PER = 16, K = WARP_K, SUP = 5;   // 512-frame blocks, 32 blocks per super-block
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    unsigned crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    unsigned long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_t = PROFILE ? clock64() : 0;

    // CTA 0 holds blocks [0, b0) next to the block tables, CTA r >= 1 blocks [b0 + (r-1) bpc, b0 + r bpc): the slices of the
    // passive CTAs can use (almost) all of their shared memory
    const int n = a.n, nblk = a.r_nblk, nsup = (nblk + 31) >> SUP, G = a.tl_n, out_cap = a.out_cap;
    const int first_blk = crank == 0 ? 0 : b0 + ((int)crank - 1) * bpc, cols = (crank == 0 ? b0 : bpc) << SHIFT;
    const long long first = (long long)first_blk << SHIFT;
    const int32_t *const lidx = a.list_idx;
    const double *const lval = a.list_val;
    int32_t *const out = a.out;
    unsigned char *const sel_bytes = a.is_selected;
    const double budget = a.budget, cost_f = a.cost_f;
    const bool has_group = a.temporal != 0;

    extern __shared__ unsigned long long warp_smem[];
    unsigned long long *fps_s = warp_smem;                 // [cols]  this CTA's slice of the min-distance vector
    unsigned long long *blk_val = fps_s + cols;            // CTA 0: [nblk] block maxima
    unsigned long long *sup_val = blk_val + nblk;          //        [32]   super-block maxima
    unsigned long long *gval = sup_val + 32;               //        [G]    row value of a group member at index gap d
    int *blk_idx = reinterpret_cast<int *>(gval + G);      //        [nblk]
    int *sup_idx = blk_idx + nblk;                         //        [32]
    int *dirty_list = sup_idx + 32;                        //        [WC_DIRTY_CAP]
    int *dirty_cnt = dirty_list + WC_DIRTY_CAP;            //        [1]

    const unsigned long long far_bits = a.far_bits;
    for (int i = lane; i < cols; i += 32) fps_s[i] = first + i < n ? far_bits : 0ull;
    int sat = 0x3fffffff;
    if (crank == 0) {
        for (int d = lane; d < G; d += 32)
            gval[d] = (unsigned long long)__double_as_longlong(combine(a.spatial, a.temporal, a.aggregate, a.lambda_t, a.s_far, a.tlut[d]));
        for (int lb = lane; lb < nblk; lb += 32) {   // every entry starts at `far`: a block's best is its first frame
            blk_val[lb] = far_bits;
            blk_idx[lb] = lb << SHIFT;
        }
        if (lane < 32) {
            sup_val[lane] = lane < nsup ? far_bits : 0ull;
            sup_idx[lane] = lane < nsup ? lane << (SHIFT + SUP) : 0x7fffffff;
        }
        if (lane == 0) *dirty_cnt = 0;
        if (has_group) {   // saturation of the temporal LUT: for d >= sat, tlut[d] equals the far value bit for bit (exp: 38)
            const int lim = min(n, 4096);
            const unsigned long long tfar = (unsigned long long)__double_as_longlong(a.t_far);
            int mx = 0;
            for (int d = lane; d < lim; d += 32)
                if ((unsigned long long)__double_as_longlong(a.tlut[d]) != tfar) mx = d + 1;
            mx = __reduce_max_sync(FULL, mx);
            sat = mx >= lim ? 0x3fffffff : mx;
        }
    }
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // every slice is initialised
    if (crank != 0) {   // passive memory: wait for CTA 0 to finish, then hand the slice back (adsel_get_fps)
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        for (int i = lane; i < cols; i += 32)
            if (first + i < n) a.fps[first + i] = fps_s[i];
        return;
    }

    const bool spec_group = has_group && 2 * sat - 1 <= 32 * WARP_SPEC && sat <= G;
    double cost = a.cost_start;
    int m = 0, stop = 0, err_frame = -1;
    long long steps = 0;
    bool marking = true;

    unsigned s_fps = (unsigned)__cvta_generic_to_shared(fps_s), s_bval = (unsigned)__cvta_generic_to_shared(blk_val),
             s_gval = (unsigned)__cvta_generic_to_shared(gval), s_bidx = (unsigned)__cvta_generic_to_shared(blk_idx),
             s_sval = (unsigned)__cvta_generic_to_shared(sup_val), s_sidx = (unsigned)__cvta_generic_to_shared(sup_idx),
             s_dlist = (unsigned)__cvta_generic_to_shared(dirty_list), s_dcnt = (unsigned)__cvta_generic_to_shared(dirty_cnt);
    asm volatile("" : "+r"(s_fps), "+r"(s_bval), "+r"(s_gval), "+r"(s_bidx), "+r"(s_sval), "+r"(s_sidx), "+r"(s_dlist), "+r"(s_dcnt));
    // frame j -> address of its fps entry in the cluster's shared memory: block lb = j >> 9 lives in CTA 0 when lb < b0, else
    // in CTA 1 + (lb - b0) / bpc (magic division, verified for every block by the host); every slice starts at the same offset
    auto fps_at = [&](int j) -> unsigned {
        const int t = (j >> SHIFT) - b0;
        const unsigned q = t < 0 ? 0u : ((unsigned)t * magic) >> 16;
        const unsigned rank = t < 0 ? 0u : q + 1u;
        const unsigned local = (unsigned)j - (t < 0 ? 0u : ((unsigned)b0 + q * (unsigned)bpc) << SHIFT);
        unsigned ra;
        asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(s_fps + 8u * local), "r"(rank));
        return ra;
    };
    bool marked = false;
    auto mark = [&](int lb) {   // rare: the entry that IS block lb's maximum was lowered
        unsigned pos;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(s_dcnt) : "memory");
        if (pos < (unsigned)WC_DIRTY_CAP) dirty_list[pos] = lb;
        marked = true;
    };
    auto refresh_block = [&](int lb) {
        const unsigned base = fps_at(lb << SHIFT) + 8u * lane;
        unsigned hi[PER], lo[PER];
        int idx[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const unsigned long long v = ldc64(base + 256u * i);
            hi[i] = (unsigned)(v >> 32);
            lo[i] = (unsigned)v;
            const int j = (lb << SHIFT) + lane + 32 * i;
            idx[i] = j < n ? j : 0x7fffffff;
        }
        const Win x = warp_argmax<PER>(hi, lo, idx);
        if (lane == 0) {
            blk_val[lb] = ((unsigned long long)x.hi << 32) | x.lo;
            blk_idx[lb] = x.idx;
        }
    };
    auto refresh_super = [&](int s) {
        const int b = (s << SUP) + lane;
        const bool ok = b < nblk;
        unsigned hi1[1] = {0u}, lo1[1] = {0u};
        int idx1[1] = {0x7fffffff};
        if (ok) {
            const unsigned long long v = blk_val[b];
            hi1[0] = (unsigned)(v >> 32);
            lo1[0] = (unsigned)v;
            idx1[0] = blk_idx[b];
        }
        const Win x = warp_argmax<1>(hi1, lo1, idx1);
        if (lane == 0) {
            sup_val[s] = ((unsigned long long)x.hi << 32) | x.lo;
            sup_idx[s] = x.idx;
        }
    };

    bool local_abort = false;
    {
        const unsigned long long needed = a.list_cursor ? *a.list_cursor : 0ull;
        if (lane == 0) a.result[5] = (double)needed;
        if ((long long)needed > a.list_capacity) {   // neighbour-list pool too small: the host grows it and reruns
            local_abort = true;
            if (lane == 0) {
                a.result[1] = (double)ADSEL_ERR_CAPACITY;
                a.result[4] = -1.0;
            }
        }
    }
    int mode = 1, cur = a.first_index, t_sampled = 0;
    if (!local_abort) {
        if (a.forbid_sampled && lane == 0)
            for (int t = 0; t < a.n_sampled; ++t) sel_bytes[a.sampled[t]] = 1;
        if (a.n_sampled > 0) {
            mode = 0;
            cur = a.sampled[0];
            marking = false;
        }
    }
    int cur_blk = cur >> SHIFT;
    bool cur_key_zero = false;
    const int sat1 = sat - 1;

    while (!local_abort) {
        marked = false;
        // ---- 1. the slot ----
        const unsigned char *sp = slots + (size_t)cur * WARP_SLOT;
        const uint4 h0 = ldg_v4(sp), h1 = ldg_v4(sp + 16);
        const int j0 = ldg_s32(sp + 32 + 4 * lane), j1 = ldg_s32(sp + 32 + 4 * (lane + 32));
        const unsigned long long v0 = ldg_u64(sp + 32 + 4 * K + 8 * lane), v1 = ldg_u64(sp + 32 + 4 * K + 8 * (lane + 32));
        const double cost_pre = __dadd_rn(cost, cost_f);   // :184
        unsigned flag_seen = 0;
        if (mode == 2 && cur_key_zero) flag_seen = __ldcg(sel_bytes + cur);   // only a zero-key winner can be a re-selection (:182)
        WPROF(6)
        // ---- 2. in its shadow: the temporal group, speculatively (remote loads) ----
        int gj[WARP_SPEC];
        unsigned ga[WARP_SPEC];
        unsigned long long gv[WARP_SPEC];
        bool gneed[WARP_SPEC] = {false, false, false}, gmax[WARP_SPEC] = {false, false, false};
        if (spec_group) {
            unsigned long long old[WARP_SPEC];
            int bi[WARP_SPEC];
#pragma unroll
            for (int q = 0; q < WARP_SPEC; ++q) {
                const int j = cur - sat1 + lane + 32 * q;
                const int d = abs(j - cur);
                const bool valid = (unsigned)j < (unsigned)n && d < sat;
                gj[q] = valid ? j : cur;
                ga[q] = fps_at(gj[q]);
                gv[q] = lds64(s_gval + 8u * (valid ? d : 0));
                old[q] = ldc64(ga[q]);
                bi[q] = lds32(s_bidx + 4u * (gj[q] >> SHIFT));
                gneed[q] = valid;
            }
#pragma unroll
            for (int q = 0; q < WARP_SPEC; ++q) {
                gneed[q] = gneed[q] && gv[q] < old[q];
                gmax[q] = marking && (gj[q] >> SHIFT) != cur_blk && bi[q] == gj[q];
            }
        }
        WPROF_DEP(0, (int)gneed[0] + (int)gmax[WARP_SPEC - 1])
        // ---- 3. the header is here ----
        const long long l_off = (long long)(((unsigned long long)h0.y << 32) | h0.x);
        const int l_len = (int)h0.z, e0 = (int)h1.x, e1 = (int)h1.y;
        const double box_cost = __longlong_as_double((long long)(((unsigned long long)h1.w << 32) | h1.z));   // n_boxes * cost_b (slot_build_kernel)
        asm volatile("" ::"r"(h0.w));
        WPROF_DEP(1, l_len + e0 + j0 + j1 + (int)v0 + (int)v1)
        if (mode == 2) {
            cost = __dadd_rn(cost_pre, box_cost);  // :184-185
            const bool over = cost > budget;       // :186
            const bool full = !over && m >= out_cap;
            unsigned verdict = (flag_seen != 0 ? 1u : 0u) | (over ? 2u : 0u) | (full ? 4u : 0u);
            asm volatile("" : "+r"(verdict));
            if (verdict != 0u) {
                if (verdict & 1u) {                // :182 comes before the cost test
                    err_frame = cur;
                    stop = 2;
                } else {
                    if ((verdict & 4u) && lane == 0) {
                        a.result[1] = (double)ADSEL_ERR_CAPACITY;
                        a.result[4] = -2.0;
                    }
                    stop = (verdict & 2u) ? 1 : 2;
                }
                break;
            }
        } else if (mode == 1) {                    // :168,172-174 -- taken and costed up front, no budget check
            cost = __dadd_rn(cost_pre, box_cost);
        }
        if (mode != 0) {
            if (lane == 0) {
                out[m] = cur;
                sel_bytes[cur] = 1;
            }
            m += 1;
        }
        // group entries first, list entries second: list values of same-group frames are <= the group values
        if (spec_group) {
#pragma unroll
            for (int q = 0; q < WARP_SPEC; ++q) {
                const bool w = gneed[q] && gj[q] >= e0 && gj[q] < e1;
                stc64_if(ga[q], gv[q], w);
                if (w && gmax[q]) mark(gj[q] >> SHIFT);
            }
        } else if (has_group) {
            const int g0 = max(e0, cur - sat + 1), g1 = min(e1, cur + sat);
            for (int j = g0 + lane; j < g1; j += 32) {
                const int d = abs(j - cur);
                const unsigned long long v = d < G ? gval[d]
                    : (unsigned long long)__double_as_longlong(combine(a.spatial, a.temporal, a.aggregate, a.lambda_t, a.s_far, a.tlut[d]));
                const unsigned ad = fps_at(j);
                if (v < ldc64(ad)) {
                    stc64_if(ad, v, true);
                    const int lb = j >> SHIFT;
                    if (lb != cur_blk && marking && blk_idx[lb] == j) mark(lb);
                }
            }
        }
        __syncwarp();
        WPROF(8)
        {   // the slot's 64 entries: two per lane, straight-line
            const bool p0 = lane < l_len, p1 = lane + 32 < l_len;
            const int c0 = p0 ? j0 : cur, c1 = p1 ? j1 : cur;
            const unsigned a0 = fps_at(c0), a1 = fps_at(c1);
            const unsigned long long o0 = ldc64(a0), o1 = ldc64(a1);
            const int b0 = lds32(s_bidx + 4u * (c0 >> SHIFT)), b1 = lds32(s_bidx + 4u * (c1 >> SHIFT));
            const bool w0 = p0 && v0 < o0, w1 = p1 && v1 < o1;
            stc64_if(a0, v0, w0);
            stc64_if(a1, v1, w1);
            if (w0 && marking && (c0 >> SHIFT) != cur_blk && b0 == c0) mark(c0 >> SHIFT);
            if (w1 && marking && (c1 >> SHIFT) != cur_blk && b1 == c1) mark(c1 >> SHIFT);
        }
        if (l_len > K) {   // rows longer than a slot continue in the list pool
            for (int t = K + lane; t < l_len; t += 32) {
                const int j = lidx[l_off + t];
                const unsigned long long v = (unsigned long long)__double_as_longlong(lval[l_off + t]);
                const unsigned ad = fps_at(j);
                if (v < ldc64(ad)) {
                    stc64_if(ad, v, true);
                    const int lb = j >> SHIFT;
                    if (lb != cur_blk && marking && blk_idx[lb] == j) mark(lb);
                }
            }
        }
        __syncwarp();
        WPROF(2)
        if (mode == 0) {   // next row of the buffer; after the last one every maximum from scratch, then the first pick
            t_sampled += 1;
            if (t_sampled < a.n_sampled) {
                cur = a.sampled[t_sampled];
                cur_blk = cur >> SHIFT;
                continue;
            }
            marking = true;
            for (int lb = 0; lb < nblk; ++lb) refresh_block(lb);
            __syncwarp();
            for (int s = 0; s < nsup; ++s) refresh_super(s);
            __syncwarp();
            cur_blk = -1;
        }
        // ---- 4. other blocks whose maximum was lowered (rare) ----
        if (__any_sync(FULL, marked)) {
            __syncwarp();
            const int cnt = lds32(s_dcnt);
            if (cnt > WC_DIRTY_CAP) {
                for (int lb = 0; lb < nblk; ++lb) refresh_block(lb);
                __syncwarp();
                for (int s = 0; s < nsup; ++s) refresh_super(s);
            } else {
                for (int q = 0; q < cnt; ++q) refresh_block(lds32(s_dlist + 4u * q));
                __syncwarp();
                for (int q = 0; q < cnt; ++q) refresh_super(lds32(s_dlist + 4u * q) >> SUP);
            }
            __syncwarp();
            if (lane == 0) *dirty_cnt = 0;
            __syncwarp();
        }
        WPROF(3)
        // ---- 5. arg-max: own block from its 512 entries, its super-block's other blocks and every other super-block from their maxima ----
        const int cur_sup = cur_blk >= 0 ? cur_blk >> SUP : -1;
        Cand sib{0ull, 0x7fffffff}, sup{0ull, 0x7fffffff};
        {
            const int b = (max(cur_sup, 0) << SUP) + lane;
            const bool ok = cur_sup >= 0 && b < nblk && b != cur_blk;
            const unsigned long long k = lds64(s_bval + 8u * (ok ? b : 0));
            const int i = lds32(s_bidx + 4u * (ok ? b : 0));
            sib.k = ok ? k : 0ull;
            sib.i = ok ? i : 0x7fffffff;
            const bool oks = lane < nsup && lane != cur_sup;
            const unsigned long long ks = lds64(s_sval + 8u * lane);
            const int is = lds32(s_sidx + 4u * lane);
            sup.k = oks ? ks : 0ull;
            sup.i = oks ? is : 0x7fffffff;
        }
        Cand own{0ull, 0x7fffffff};
        if (cur_blk >= 0) {
            const unsigned base = fps_at(cur_blk << SHIFT) + 8u * lane;
            const int j0b = (cur_blk << SHIFT) + lane;
            Cand c[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                c[i].k = ldc64(base + 256u * i);
                c[i].i = j0b + 32 * i;
            }
#pragma unroll
            for (int w = 1; w < PER; w *= 2)
#pragma unroll
                for (int i = 0; i + w < PER; i += 2 * w) c[i] = later_wins_if_larger(c[i], c[i + w]);
            own = c[0];
            if (own.i >= n) own.i = 0x7fffffff;
        }
        const Cand in_sup = better_of(own, sib);      // this lane's best inside the own super-block
        const Cand best = better_of(in_sup, sup);
        WPROF_DEP(4, (int)best.k + best.i)
        // three interleaved chains of three REDUX: (A) the winner, (B) the own block's new maximum, (C) the own super-block's
        Win y;
        {
            const unsigned ah = (unsigned)(best.k >> 32), al = (unsigned)best.k, bh = (unsigned)(own.k >> 32), bl = (unsigned)own.k,
                           ch = (unsigned)(in_sup.k >> 32), cl = (unsigned)in_sup.k;
            const unsigned mah = __reduce_max_sync(FULL, ah);
            const unsigned mbh = __reduce_max_sync(FULL, bh);
            const unsigned mch = __reduce_max_sync(FULL, ch);
            const unsigned mal = __reduce_max_sync(FULL, ah == mah ? al : 0u);
            const unsigned mbl = __reduce_max_sync(FULL, bh == mbh ? bl : 0u);
            const unsigned mcl = __reduce_max_sync(FULL, ch == mch ? cl : 0u);
            y.idx = __reduce_min_sync(FULL, (ah == mah && al == mal) ? best.i : 0x7fffffff);
            const int bidx = __reduce_min_sync(FULL, (bh == mbh && bl == mbl) ? own.i : 0x7fffffff);
            const int cidx = __reduce_min_sync(FULL, (ch == mch && cl == mcl) ? in_sup.i : 0x7fffffff);
            y.hi = mah;
            y.lo = mal;
            const bool st = lane == 0 && cur_blk >= 0;
            sts64_if(s_bval + 8u * max(cur_blk, 0), ((unsigned long long)mbh << 32) | mbl, st);
            sts32_if(s_bidx + 4u * max(cur_blk, 0), bidx, st);
            sts64_if(s_sval + 8u * max(cur_sup, 0), ((unsigned long long)mch << 32) | mcl, st);
            sts32_if(s_sidx + 4u * max(cur_sup, 0), cidx, st);
        }
        WPROF_DEP(5, y.idx)
        ++steps;
        const int nxt = y.idx;
        cur_key_zero = (y.hi | y.lo) == 0u;
        if (nxt < 0 || nxt >= n) {   // empty pool: cannot happen with n > 0
            if (lane == 0) a.result[1] = (double)ADSEL_ERR_INVALID;
            stop = 2;
            break;
        }
        mode = mode == 0 ? 1 : 2;
        cur = nxt;
        cur_blk = nxt >> SHIFT;
    }
    __syncwarp();
    if (stop == 2 && err_frame >= 0 && lane == 0) {
        a.result[1] = (double)ADSEL_ERR_RESELECTED;
        a.result[4] = (double)err_frame;
    }
    if (lane == 0) {
        a.result[0] = cost;
        a.result[2] = (double)m;
        a.result[3] = (double)steps;
        if (PROFILE)
            for (int i = 0; i < 12; ++i) a.prof[i] = prof_acc[i];
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // the passive CTAs may leave
    for (int i = lane; i < cols; i += 32)
        if (i < n) a.fps[i] = fps_s[i];
}

// ---- host side -----------------------------------------------------------------------------------
static bool warp_enabled() {
    const char *e = getenv("ADSEL_GREEDY_WARP");
    if (e && atoi(e) == 0) return false;
    const char *f = getenv("ADSEL_GREEDY_FAST");       // the switches of the multi-warp kernels ask for those kernels
    const char *c = getenv("ADSEL_GREEDY_CLUSTER");
    const char *b = getenv("ADSEL_FAST_BLOCK");
    return !(f && *f) && !(c && *c) && !(b && *b);
}

// The one-warp kernel serves the configuration its slot layout is made for: truncated geodesic lists (exp
// normalisation: radius 38, a few dozen entries), temporal groups that are contiguous index ranges (or none), the whole
// pool on this device and resident in one CTA's shared memory.
bool warp_eligible(const adsel_handle *h, const adsel_select_params *p, bool sharded) {
    if (!warp_enabled() || sharded) return false;
    if (p->spatial != ADSEL_SPATIAL_GEODESIC || p->normalize != ADSEL_NORM_EXP) return false;
    if (p->temporal && !h->groups_contiguous) return false;
    const long long cols = ((long long)h->n + WARP_RB - 1) / WARP_RB * WARP_RB, nblk = cols / WARP_RB;
    return nblk <= 32 * WARP_NB && (size_t)cols * 8 + (size_t)nblk * 12 + 64 * 8 + 16 <= 232448 - 64;
}

// Geometry of the passive-memory cluster variant: the smallest cluster of 2 / 4 / 8 / 16 CTAs whose slices of at most 48
// blocks (24,576 frames, 192 KB) hold the pool -- up to 393 k frames.  ADSEL_WARP_CLUSTER=0 disables the variant,
// ADSEL_WARP_CLUSTER_BPC=b forces b blocks per CTA (tests: clusters on small pools).
struct WarpClusterPlan {
    int csize = 0, b0 = 0, bpc = 0, nblk = 0;
    unsigned magic = 0;
    size_t smem = 0;
};

static bool warp_cluster_mode_ok(const adsel_handle *h, const adsel_select_params *p, bool sharded) {
    const char *e = getenv("ADSEL_WARP_CLUSTER");
    if (e && atoi(e) == 0) return false;
    if (!warp_enabled() || sharded) return false;
    if (p->spatial != ADSEL_SPATIAL_GEODESIC || p->normalize != ADSEL_NORM_EXP) return false;
    return !(p->temporal && !h->groups_contiguous);
}

// Geometry of the passive-memory cluster variant: the smallest cluster of 2 / 4 / 8 CTAs that holds the pool, CTA 0 with a
// smaller slice next to the block tables, the passive CTAs with up to 56 blocks (28,672 frames, 224 KB) each: up to ~229 k
// frames.  (A 16-CTA cluster measured SLOWER than the multi-warp cluster kernel -- remote shared memory gets farther with
// the cluster's size: 2.37 vs 2.17 us per step at 281,300 frames -- and is left to that kernel.)
// ADSEL_WARP_CLUSTER=0 disables the variant, ADSEL_WARP_CLUSTER_BPC=b caps a slice at b blocks (tests: clusters, up to
// 16 CTAs, on small pools).
static bool warp_cluster_plan(const adsel_handle *h, const adsel_select_params *p, bool sharded, int g, WarpClusterPlan *out) {
    if (!warp_cluster_mode_ok(h, p, sharded)) return false;
    const char *fb = getenv("ADSEL_WARP_CLUSTER_BPC");
    const int forced = (fb && *fb) ? atoi(fb) : 0;
    if (forced <= 0 && warp_eligible(h, p, sharded)) return false;   // one CTA holds the pool
    const int nblk = (h->n + 511) / 512;
    const size_t budget = 232448 - 64;
    const size_t extras = (size_t)nblk * 12 + 32 * 12 + (size_t)g * 8 + (size_t)(WC_DIRTY_CAP + 4) * 4 + 16;
    if (extras + 4096 > budget) return false;
    const int cap_passive = forced > 0 ? forced : (int)(budget / 4096);                  // 56 blocks
    const int cap_leader = forced > 0 ? forced : (int)((budget - extras) / 4096);
    for (int c = 2; c <= (forced > 0 ? 16 : 8); c *= 2) {
        int b0 = (nblk + c - 1) / c;
        if (b0 > cap_leader) b0 = cap_leader;
        int bpc = (nblk - b0 + c - 2) / (c - 1);
        if (bpc < 1) bpc = 1;
        if (bpc > cap_passive) continue;
        unsigned magic = 65536u / (unsigned)bpc + 1u;
        bool exact = true;
        for (int t = 0; t < (c - 1) * bpc && exact; ++t) exact = (((unsigned)t * magic) >> 16) == (unsigned)(t / bpc);
        if (!exact) return false;
        out->csize = c;
        out->b0 = b0;
        out->bpc = bpc;
        out->nblk = nblk;
        out->magic = magic;
        const size_t leader = (size_t)b0 * 4096 + extras, passive = (size_t)bpc * 4096;
        out->smem = leader > passive ? leader : passive;
        return out->smem <= budget;
    }
    return false;
}

template <bool PROFILE>
static cudaError_t launch_warp_cluster(adsel_handle *h, const GreedyArgs &a, const WarpClusterPlan &pl, bool dry_run, int *max_clusters) {
    auto kernel = greedy_warp_cluster_kernel<PROFILE>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    if (pl.csize > 8) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)pl.csize, 1, 1);
    cfg.blockDim = dim3(32, 1, 1);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)pl.csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (dry_run) return cudaOccupancyMaxActiveClusters(max_clusters, kernel, &cfg);
    return cudaLaunchKernelEx(&cfg, kernel, a, (const unsigned char *)h->slots.p, pl.b0, pl.bpc, pl.magic);
}

int warp_prepare(adsel_handle *h, const adsel_select_params *p, bool sharded) {
    WarpClusterPlan pl;
    if (warp_eligible(h, p, sharded) || warp_cluster_plan(h, p, sharded, 64, &pl)) ADSEL_TRY(h->slots.reserve((size_t)h->n * WARP_SLOT));
    return ADSEL_OK;
}

int warp_preload() {
    cudaFuncAttributes attr;
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_warp_kernel<false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, greedy_warp_cluster_kernel<false>));
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, slot_build_kernel));
    return ADSEL_OK;
}

// Enqueues slot build + the one-warp loop (one CTA, or a cluster of passive-memory CTAs); *launched = false when this
// selection is not for it.
int warp_enqueue(adsel_handle *h, const adsel_select_params *p, GreedyArgs &a, bool sharded, bool *launched) {
    *launched = false;
    const int n = h->n;
    if (h->slots.cap < (size_t)n * WARP_SLOT || a.list_capacity >= (1ll << 40)) return ADSEL_OK;
    const int g = a.temporal ? (a.tl_n < 64 ? a.tl_n : 64) : 0;
    WarpClusterPlan pl;
    const bool single = warp_eligible(h, p, sharded) && !(getenv("ADSEL_WARP_CLUSTER_BPC") && *getenv("ADSEL_WARP_CLUSTER_BPC"));
    if (!single) {
        if (!warp_cluster_plan(h, p, sharded, g, &pl)) return ADSEL_OK;
        GreedyArgs probe = a;
        int max_clusters = 0;
        const cudaError_t e = a.prof ? launch_warp_cluster<true>(h, probe, pl, true, &max_clusters)
                                     : launch_warp_cluster<false>(h, probe, pl, true, &max_clusters);
        if (e != cudaSuccess || max_clusters < 1) {   // a cluster of this size and footprint cannot be scheduled: the multi-warp kernels take over
            cudaGetLastError();
            return ADSEL_OK;
        }
    }
    ADSEL_CUDA(cudaMemsetAsync(h->is_selected.p, 0, (size_t)n, h->stream));
    slot_build_kernel<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, h->stream>>>(h->slots.p, a.rec, a.list_idx, a.list_val, n,
                                                                                       a.list_capacity, a.list_cursor, a.cost_b);
    ADSEL_CUDA(cudaGetLastError());
    h->launches += 1;
    a.tl_n = g;
    a.sel_words = 0;
    if (!single) {
        a.csize = pl.csize;
        a.cta_cols = pl.b0 * 512;
        a.r_nblk = pl.nblk;
        int unused = 0;
        ADSEL_CUDA(a.prof ? launch_warp_cluster<true>(h, a, pl, false, &unused) : launch_warp_cluster<false>(h, a, pl, false, &unused));
        h->greedy_cluster = pl.csize;
        h->greedy_variant = 6;
        *launched = true;
        return ADSEL_OK;
    }
    const long long cols = ((long long)n + WARP_RB - 1) / WARP_RB * WARP_RB, nblk = cols / WARP_RB;
    a.csize = 1;
    a.cta_cols = (int)cols;
    a.r_nblk = (int)nblk;
    const size_t smem = (size_t)cols * 8 + (size_t)nblk * 8 + (size_t)g * 8 + (size_t)nblk * 4 + 16;
    if (a.prof) {
        ADSEL_CUDA(cudaFuncSetAttribute(greedy_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        greedy_warp_kernel<true><<<1, 32, smem, h->stream>>>(a, h->slots.p);
    } else {
        ADSEL_CUDA(cudaFuncSetAttribute(greedy_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        greedy_warp_kernel<false><<<1, 32, smem, h->stream>>>(a, h->slots.p);
    }
    ADSEL_CUDA(cudaGetLastError());
    h->greedy_cluster = 1;
    h->greedy_variant = 5;
    *launched = true;
    return ADSEL_OK;
}

}   // namespace adsel
