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
// square root is evaluated literally).  The float32 sum over the feature dimension runs in index order
// here, whereas torch uses a vectorised cascade sum whose order depends on the CPU's vector width, so
// the map agrees with the reference to float32 summation accuracy (tests: 1e-5 relative), not bitwise;
// the features themselves need the trained detector (out of scope), tests use synthetic ones.
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
    ADSEL_TRY(begin_stage(h));
    dim3 grid((n + FD_TILE - 1) / FD_TILE, (n + FD_TILE - 1) / FD_TILE);
    feature_l1_kernel<<<grid, FD_THREADS, 0, h->stream>>>(h->feat.p, n, dim, p == 2 ? 1 : 0, reinterpret_cast<float *>(h->d2_p), ld);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    ADSEL_TRY(end_stage(h, 6));
    return ADSEL_OK;
}

}  // namespace adsel
