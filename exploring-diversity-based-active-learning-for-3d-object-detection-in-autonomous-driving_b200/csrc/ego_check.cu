// ego_check.cu -- every frame's ego location recomputed on the device in the reference's literal order.
//
// The reference takes  -(cal[:3,3].T @ cal[:3,:3])[:2]  per frame (spatial_temporal_selector.py:83-91; also
// spatial_selector.py, euclidean_spatial_selector.py).  The host extraction (csrc/hostext.c, pool.py) evaluates that
// 3-term dot product in C for the whole pool and pins WHICH rounding numpy's vector-matrix product uses on this host
// (fused multiply-add chain or separately rounded products) on a few probe frames.  This kernel extends the pin to ALL
// frames: one thread per frame redoes both coordinates in the pinned order from the nine pose entries involved and
// compares bit patterns with what the host produced (a compiler that contracted the unfused expression, a pose read
// through the wrong stride, a special value handled differently would all show up here).  Two NaNs count as equal: their
// payload/sign is not defined by the reference either.
#include <mutex>

#include "common.cuh"

namespace adsel {

__global__ void ego_check_kernel(const double *__restrict__ pose9, const double *__restrict__ loc, long long m, int fused,
                                 unsigned long long *__restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m) return;
    // pose9[i] = {c00, c01, c03, c10, c11, c13, c20, c21, c23}: columns 0, 1 (rotation) and 3 (translation) of rows 0..2
    const double *c = pose9 + 9 * i;
    const double t0 = c[2], t1 = c[5], t2 = c[8];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const double r0 = c[j], r1 = c[3 + j], r2 = c[6 + j];
        double v;
        if (fused) v = __fma_rn(t2, r2, __fma_rn(t1, r1, __dmul_rn(t0, r0)));
        else v = __dadd_rn(__dadd_rn(__dmul_rn(t0, r0), __dmul_rn(t1, r1)), __dmul_rn(t2, r2));
        v = -v;
        const double got = loc[2 * i + j];
        if (__double_as_longlong(v) != __double_as_longlong(got) && !(v != v && got != got)) bad = true;
    }
    if (bad) {
        atomicAdd(&out[0], 1ull);
        atomicMin(&out[1], (unsigned long long)i);
    }
}

namespace {
struct EgoScratch {
    DevBuf<double> pose, loc;
    DevBuf<unsigned long long> out;
};
std::mutex g_ego_mutex;
EgoScratch g_ego[64];   // per device, grow-only (a few MB), kept for the life of the process
}   // namespace

}   // namespace adsel

using namespace adsel;

extern "C" int adsel_check_ego_locations(int device, const double *pose9, const double *loc_xy, int64_t m, int fused,
                                         int64_t *n_bad, int64_t *first_bad) {
    ADSEL_REQUIRE(n_bad && first_bad, "adsel_check_ego_locations: null result pointer");
    ADSEL_REQUIRE(m >= 0 && (m == 0 || (pose9 && loc_xy)), "adsel_check_ego_locations: null input");
    int count = 0;
    ADSEL_CUDA(cudaGetDeviceCount(&count));
    ADSEL_REQUIRE(device >= 0 && device < count && device < 64,
                  "adsel_check_ego_locations: no such CUDA device (this library has no CPU fallback)");
    *n_bad = 0;
    *first_bad = -1;
    if (m == 0) return ADSEL_OK;
    std::lock_guard<std::mutex> lock(g_ego_mutex);
    ADSEL_CUDA(cudaSetDevice(device));
    EgoScratch &s = g_ego[device];
    ADSEL_TRY(s.pose.reserve((size_t)m * 9));
    ADSEL_TRY(s.loc.reserve((size_t)m * 2));
    ADSEL_TRY(s.out.reserve(2));
    const unsigned long long init[2] = {0ull, ~0ull};
    ADSEL_CUDA(cudaMemcpy(s.pose.p, pose9, (size_t)m * 9 * sizeof(double), cudaMemcpyHostToDevice));
    ADSEL_CUDA(cudaMemcpy(s.loc.p, loc_xy, (size_t)m * 2 * sizeof(double), cudaMemcpyHostToDevice));
    ADSEL_CUDA(cudaMemcpy(s.out.p, init, sizeof(init), cudaMemcpyHostToDevice));
    const int threads = 256;
    ego_check_kernel<<<(unsigned)((m + threads - 1) / threads), threads>>>(s.pose.p, s.loc.p, m, fused != 0, s.out.p);
    ADSEL_CUDA(cudaGetLastError());
    unsigned long long res[2];
    ADSEL_CUDA(cudaMemcpy(res, s.out.p, sizeof(res), cudaMemcpyDeviceToHost));
    *n_bad = (int64_t)res[0];
    *first_bad = res[0] ? (int64_t)res[1] : -1;
    return ADSEL_OK;
}
