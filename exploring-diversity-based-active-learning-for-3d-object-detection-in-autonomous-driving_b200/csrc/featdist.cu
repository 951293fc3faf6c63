// featdist.cu -- the feature-distance map of the detector-feature selectors:
// det3d/selectors/feature_selector.py:87-109 (same loop in spatial_feature / spatial_temporal_feature /
// badge / uwe selectors).  The reference fills an N x N float32 map row by row on the CPU with
//     p == 1:  torch.abs(feats - feats[i]).sum(1)
//     p == 2:  torch.sqrt((feats - feats[i]) ** 2).sum(1)        -- an L1 distance as well ("p=2" bug kept)
// and then runs the same greedy loop as the other selectors on it.  Here the whole map is one tiled
// kernel; the rows feed the dense greedy kernel of pairdist.cu unchanged, because non-negative float32
// values order exactly like their int32 bit patterns.
//
// Arithmetic: per element d = fl(a - b) and the term |d| (p == 2: sqrt(fl(d*d)), which in binary
// floating point equals |d| whenever d*d neither underflows nor overflows; outside that range the
// square root is evaluated literally).
//
// Summation order.  torch's CPU `.sum(1)` of a contiguous float32 [N, C] tensor is a vectorised cascade sum
// (aten/src/ATen/native/cpu/SumKernel.cpp: 8 float lanes x 4 interleaved accumulators, blocks of 16 vectors
// folded through four levels, then the four accumulators, then the scalar tail, then the 8 lanes in order);
// oracle/restatement.py::torch_sum_f32 restates it and tests/test_oracle_feature.py pins it bit for bit against
// torch 2.11 for C = 16 .. 8192.  feature_l1_torch_kernel follows exactly that order whenever C is a multiple of
// 32 (the reference's C is 512): lane l of a warp IS accumulator (l / 8, l % 8), so the map equals the
// reference's BIT FOR BIT and every comparison of the greedy loop resolves the same way.  Other C fall back to
// feature_l1_kernel (index-order sum: rows agree to float32 summation accuracy, 1e-5 relative, not bitwise).
// The features themselves need the trained detector (out of scope), tests use synthetic ones.
#include <stdlib.h>

#include "common.cuh"

namespace adsel {

constexpr int FD_TILE = 128, FD_K = 16, FD_THREADS = 256;

// literal term of the reference expression (used when a staged tile holds values for which d*d may leave
// float32's normal range)
__device__ __forceinline__ float feat_term_literal(float a, float b, int p2) {
    const float d = __fsub_rn(a, b);
    return p2 ? __fsqrt_rn(__fmul_rn(d, d)) : fabsf(d);
}

// A staged value is "plain" when it is 0 or 2^-39 <= |v| <= 2^62: the difference d of two plain values is
// either 0 or 2^-63 <= |d| <= 2^63 (a non-zero difference is at least one ulp of the larger operand), so
// d*d stays normal and sqrt(fl(d*d)) == |d| exactly.
__device__ __forceinline__ bool plain_value(float v) {
    const float a = fabsf(v);
    return a == 0.0f || (a >= 1.8189894e-12f && a <= 4.6116860e18f);
}

// out[i][j] = sum_k term(x[j][k], x[i][k])   (row i of the reference: feats - feats[i]); 128 x 128 outputs per
// CTA, 8 x 8 per thread on a 16-strided layout (shared-memory reads are broadcasts / conflict-free), two FP32
// adds per term on the fast path.
__global__ void __launch_bounds__(FD_THREADS)
feature_l1_kernel(const float *__restrict__ x, int n, int dim, int p2, float *__restrict__ out, int ld) {
    __shared__ float sa[FD_K][FD_TILE], sb[FD_K][FD_TILE];
    const int i0 = blockIdx.y * FD_TILE, j0 = blockIdx.x * FD_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // outputs (ty + 16 r, tx + 16 c)
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
    for (int k0 = 0; k0 < dim; k0 += FD_K) {
        // stage 128 rows x 16 features of both operands, transposed to [k][row]
        int odd = 0;
        for (int e = threadIdx.x; e < FD_TILE * FD_K; e += FD_THREADS) {
            const int row = e / FD_K, k = e - row * FD_K;
            const int gi = i0 + row, gj = j0 + row, gk = k0 + k;
            const float va = (gi < n && gk < dim) ? x[(size_t)gi * dim + gk] : 0.0f;
            const float vb = (gj < n && gk < dim) ? x[(size_t)gj * dim + gk] : 0.0f;
            sa[k][row] = va;
            sb[k][row] = vb;
            odd |= (plain_value(va) && plain_value(vb)) ? 0 : 1;
        }
        const int literal = __syncthreads_or(p2 ? odd : 0);   // also the barrier that publishes the tile
        if (!literal) {
#pragma unroll
            for (int k = 0; k < FD_K; ++k) {
                float a[8], b[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) a[r] = sa[k][ty + 16 * r];
#pragma unroll
                for (int c = 0; c < 8; ++c) b[c] = sb[k][tx + 16 * c];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = __fadd_rn(acc[r][c], fabsf(__fsub_rn(b[c], a[r])));
            }
        } else {
            for (int k = 0; k < FD_K; ++k) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        acc[r][c] = __fadd_rn(acc[r][c], feat_term_literal(sb[k][tx + 16 * c], sa[k][ty + 16 * r], p2));
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int gi = i0 + ty + 16 * r;
        if (gi >= n) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int gj = j0 + tx + 16 * c;
            if (gj < ld) out[(size_t)gi * ld + gj] = gj < n ? acc[r][c] : 0.0f;
        }
    }
}


// ---- torch's summation order, bit for bit ---------------------------------------------------------------------
// A warp works on one "i" row against 32 "j" rows at a time.  Lane l accumulates, for each of the 32 outputs, the terms
// c = 32 pos + l (pos = 0 .. C/32 - 1): it IS torch's accumulator (k = l / 8, lane = l % 8) of the 4 x 8 interleaved vector
// accumulators, and every 16 terms the running sums are folded into the next level exactly as multi_row_sum does -- 32
// independent accumulation chains per lane.  The 32 x 32 partial sums then go through a padded shared-memory transpose so
// that lane j holds all 32 accumulators of output j and finishes it alone, in torch's order: ((a0 + a1) + a2) + a3 over the
// four accumulators of each of the 8 vector lanes, then the 8 lanes in order.  A CTA stages 64 "j" rows in shared memory
// (fewer when 64 do not fit) and reuses them for a chunk of up to 4,096 "i" rows: each of its 8 warps takes every 8th row
// of the chunk and keeps it in registers.
constexpr int FT_ROWS = 64, FT_THREADS = 256, FT_MAXREG = 32;   // C <= 32 * FT_MAXREG = 1024 keeps the i row in registers

// One (i row, 32 j rows) batch: returns nothing, leaves the 32 x 32 partial sums in `tile`.  LITERAL: evaluate the reference's
// sqrt(fl(d*d)) term as written (needed only when a staged value may make d*d leave float32's normal range -- see
// plain_value above; otherwise sqrt(fl(d*d)) == |d| exactly and the IEEE square root, ~20 instructions per term, is skipped).
template <bool LITERAL>
__device__ __forceinline__ void ft_batch(const float (&a)[FT_MAXREG], const float *b, int dim, int nvec, float *tile, int lane) {
    float acc0[32], acc1[32], acc2[32], acc3[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc0[o] = acc1[o] = acc2[o] = acc3[o] = 0.f;
    int pos = 0;
#pragma unroll 1
    for (; pos + 16 <= nvec;) {
#pragma unroll
        for (int q = 0; q < FT_MAXREG; ++q) {
            if (q >= pos && q < pos + 16) {
#pragma unroll
                for (int o = 0; o < 32; ++o) {
                    const float d = __fsub_rn(b[(size_t)o * dim + 32 * q], a[q]);
                    acc0[o] = __fadd_rn(acc0[o], LITERAL ? __fsqrt_rn(__fmul_rn(d, d)) : fabsf(d));
                }
            }
        }
        pos += 16;
#pragma unroll
        for (int o = 0; o < 32; ++o) {   // fold through the levels like multi_row_sum (level_power = 4)
            acc1[o] = __fadd_rn(acc1[o], acc0[o]);
            acc0[o] = 0.f;
            if ((pos & 0xf0) == 0) {
                acc2[o] = __fadd_rn(acc2[o], acc1[o]);
                acc1[o] = 0.f;
                if ((pos & 0xf00) == 0) {
                    acc3[o] = __fadd_rn(acc3[o], acc2[o]);
                    acc2[o] = 0.f;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < FT_MAXREG; ++q) {
        if (q >= pos && q < nvec) {
#pragma unroll
            for (int o = 0; o < 32; ++o) {
                const float d = __fsub_rn(b[(size_t)o * dim + 32 * q], a[q]);
                acc0[o] = __fadd_rn(acc0[o], LITERAL ? __fsqrt_rn(__fmul_rn(d, d)) : fabsf(d));
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float t = __fadd_rn(acc0[o], acc1[o]);
        t = __fadd_rn(t, acc2[o]);
        t = __fadd_rn(t, acc3[o]);
        tile[o * 33 + lane] = t;            // accumulator `lane` of output o
    }
    __syncwarp();
}

template <int P2>
__global__ void __launch_bounds__(FT_THREADS)
feature_l1_torch_kernel(const float *__restrict__ x, int n, int dim, float *__restrict__ out, int ld, int rows, int ichunk) {
    extern __shared__ float ft_smem[];   // [rows][dim] j rows | [8 warps][32][33] transpose tiles
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int j0 = blockIdx.x * rows, i0 = blockIdx.y * ichunk;   // the staged j tile is reused by `ichunk` i rows
    // The map is symmetric BIT FOR BIT (|a - b| == |b - a| term by term, same summation order), so only the rows i >= j0 of
    // a j tile are computed and every result below the diagonal tile is stored twice: out[i][j] and out[j][i].
    if (i0 + ichunk <= j0) return;
    const int nvec = dim / 32;            // torch's size_ilp: vectors per interleaved accumulator
    float *tile = ft_smem + (size_t)rows * dim + (size_t)wid * 32 * 33;
    int odd = 0;
    for (int e = threadIdx.x; e < rows * dim; e += FT_THREADS) {
        const int r = e / dim, c = e - r * dim;
        const float v = j0 + r < n ? x[(size_t)(j0 + r) * dim + c] : 0.0f;
        ft_smem[e] = v;
        odd |= plain_value(v) ? 0 : 1;
    }
    const bool plain_j = __syncthreads_or(odd) == 0;   // also the barrier that publishes the tile
    for (int ii = wid; ii < ichunk; ii += FT_THREADS / 32) {
        const int i = i0 + ii;
        if (i >= n) break;
        if (i < j0) continue;
        float a[FT_MAXREG];
#pragma unroll
        for (int q = 0; q < FT_MAXREG; ++q) a[q] = q < nvec ? x[(size_t)i * dim + 32 * q + lane] : 0.0f;
        int odd_i = 0;
#pragma unroll
        for (int q = 0; q < FT_MAXREG; ++q) odd_i |= plain_value(a[q]) ? 0 : 1;
        const bool plain_i = __any_sync(0xffffffffu, odd_i) == 0;
        for (int jj = 0; jj < rows; jj += 32) {
            const float *b = ft_smem + (size_t)jj * dim + lane;
            if (P2 && !(plain_j && plain_i)) ft_batch<true>(a, b, dim, nvec, tile, lane);
            else ft_batch<false>(a, b, dim, nvec, tile, lane);
            // lane o finishes output o: partial_sums[0] += partial_sums[k] per vector lane, then the 8 lanes in order
            float s = 0.0f;
#pragma unroll
            for (int l8 = 0; l8 < 8; ++l8) {
                float v = tile[lane * 33 + l8];
                v = __fadd_rn(v, tile[lane * 33 + l8 + 8]);
                v = __fadd_rn(v, tile[lane * 33 + l8 + 16]);
                v = __fadd_rn(v, tile[lane * 33 + l8 + 24]);
                s = __fadd_rn(s, v);
            }
            const int j = j0 + jj + lane;
            if (jj + lane < rows && j < ld) out[(size_t)i * ld + j] = j < n ? s : 0.0f;
            if (jj + lane < rows && j < n && i >= j0 + rows) out[(size_t)j * ld + i] = s;   // mirror image
        }
    }
}

__global__ void feature_finite_kernel(const float *__restrict__ x, size_t count, int *bad) {
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < count; e += (size_t)gridDim.x * blockDim.x)
        if (!isfinite(x[e])) atomicExch(bad, 1);
}

int feature_distance_build(adsel_handle *h, const float *feats_host, int32_t n, int32_t dim, int32_t p) {
    ADSEL_REQUIRE(feats_host && n > 0 && dim > 0, "adsel_feature_distance_build: null/empty argument");
    ADSEL_REQUIRE(p == 1 || p == 2, "adsel_feature_distance_build: p must be 1 or 2 (feature_selector.py:47)");
    const int ld = (n + 3) / 4 * 4;
    ADSEL_TRY(h->feat.reserve((size_t)n * dim));
    ADSEL_TRY(h->pd_d2.reserve((size_t)n * ld));
    ADSEL_CUDA(cudaMemcpyAsync(h->feat.p, feats_host, sizeof(float) * (size_t)n * dim, cudaMemcpyHostToDevice, h->stream));
    h->d2_p = h->pd_d2.p;
    h->pd_n = n;
    h->pd_ld = ld;
    h->pd_col0 = 0;
    h->pd_cols = n;
    // NaN / inf features would break the ordered-bit-pattern view the greedy loop takes of the map: refused
    ADSEL_TRY(h->d_result.reserve(8));
    ADSEL_CUDA(cudaMemsetAsync(h->d_result.p, 0, sizeof(int), h->stream));
    feature_finite_kernel<<<h->num_sms * 4, 256, 0, h->stream>>>(h->feat.p, (size_t)n * dim, reinterpret_cast<int *>(h->d_result.p));
    int bad = 0;
    ADSEL_CUDA(cudaMemcpyAsync(&bad, h->d_result.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    h->launches += 1;
    ADSEL_REQUIRE(bad == 0, "adsel_feature_distance_build: the features contain NaN or inf");
    ADSEL_TRY(begin_stage(h));
    const bool exact = dim % 32 == 0 && dim <= 32 * FT_MAXREG && !getenv("ADSEL_FEATDIST_INDEX_ORDER");
    if (exact) {   // torch's summation order: the map is the reference's, bit for bit
        int rows = FT_ROWS;     // a multiple of 32: a warp finishes 32 outputs at a time
        const size_t tiles = sizeof(float) * (FT_THREADS / 32) * 32 * 33;
        while (rows > 32 && (size_t)rows * dim * sizeof(float) + tiles > 220 * 1024) rows -= 32;
        // every CTA stages one tile of j rows and reuses it for `ichunk` i rows: enough CTAs to fill the GPU a few times over
        int ichunk = 4096;
        while (ichunk > 256 && (long long)((n + rows - 1) / rows) * ((n + ichunk - 1) / ichunk) < 8LL * h->num_sms) ichunk /= 2;
        dim3 grid((n + rows - 1) / rows, (n + ichunk - 1) / ichunk);
        const size_t smem = sizeof(float) * rows * dim + tiles;
        if (p == 2) {
            ADSEL_CUDA(cudaFuncSetAttribute(feature_l1_torch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            feature_l1_torch_kernel<1><<<grid, FT_THREADS, smem, h->stream>>>(h->feat.p, n, dim, reinterpret_cast<float *>(h->d2_p), ld, rows, ichunk);
        } else {
            ADSEL_CUDA(cudaFuncSetAttribute(feature_l1_torch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            feature_l1_torch_kernel<0><<<grid, FT_THREADS, smem, h->stream>>>(h->feat.p, n, dim, reinterpret_cast<float *>(h->d2_p), ld, rows, ichunk);
        }
    } else {
        dim3 grid((n + FD_TILE - 1) / FD_TILE, (n + FD_TILE - 1) / FD_TILE);
        feature_l1_kernel<<<grid, FD_THREADS, 0, h->stream>>>(h->feat.p, n, dim, p == 2 ? 1 : 0, reinterpret_cast<float *>(h->d2_p), ld);
    }
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_TRY(end_stage(h, 6));
    return ADSEL_OK;
}

}  // namespace adsel
