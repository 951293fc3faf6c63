// Latency micro-benchmarks on one SM of a B200 (development aid for the greedy loop):
// dependent global loads over working sets of different sizes, LDS, REDUX, shuffle, CTA barrier, smem atomics.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void chase(const unsigned *next, int iters, unsigned start, long long *out, unsigned *sink) {
    unsigned p = start;
    for (int i = 0; i < 64; ++i) p = next[p];          // warm the TLB a little
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) p = next[p];
    long long t1 = clock64();
    out[0] = (t1 - t0) / iters;
    *sink = p;
}

__global__ void chase_cg(const unsigned *next, int iters, unsigned start, long long *out, unsigned *sink) {
    unsigned p = start;
    for (int i = 0; i < 64; ++i) p = __ldcg(next + p);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) p = __ldcg(next + p);
    long long t1 = clock64();
    out[0] = (t1 - t0) / iters;
    *sink = p;
}

__global__ void misc(long long *out, unsigned *sink, int nwarps_active) {
    __shared__ unsigned long long sm[1024];
    __shared__ int cnt;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 1024; i += blockDim.x) sm[i] = (i * 37 + 11) & 1023;
    if (tid == 0) cnt = 0;
    __syncthreads();
    const int N = 2000;
    // dependent LDS.64
    unsigned long long p = lane;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) p = sm[p];
    long long t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0) / N;
    // dependent REDUX.MAX
    unsigned v = (unsigned)p + lane;
    t0 = clock64();
    for (int i = 0; i < N; ++i) v = __reduce_max_sync(0xffffffffu, v ^ lane) + 1;
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0) / N;
    // dependent shuffle
    t0 = clock64();
    for (int i = 0; i < N; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0) / N;
    // ballot + ffs
    t0 = clock64();
    for (int i = 0; i < N; ++i) v = __ffs(__ballot_sync(0xffffffffu, (v & 1) == 0)) + v;
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0) / N;
    // CTA barrier, all warps
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
        __syncthreads();
        v += sm[(v + i) & 1023] & 1;      // a shared-memory consumer after every barrier
    }
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0) / N;
    // smem atomic add with return, one lane
    t0 = clock64();
    if (lane == 0)
        for (int i = 0; i < N; ++i) v += atomicAdd(&cnt, 1) & 1;
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0) / N;
    // dependent DADD
    double d = (double)v;
    t0 = clock64();
    for (int i = 0; i < N; ++i) d = __dadd_rn(d, 1.25);
    t1 = clock64();
    if (tid == 0) out[7] = (t1 - t0) / N;
    // dependent IMAD
    t0 = clock64();
    for (int i = 0; i < N; ++i) v = v * 3 + 1;
    t1 = clock64();
    if (tid == 0) out[8] = (t1 - t0) / N;
    // 64-bit compare + select chain
    unsigned long long a = p, b = p * 3;
    t0 = clock64();
    for (int i = 0; i < N; ++i) {
        a = a > b ? a + 1 : b + 3;
        b ^= a;
    }
    t1 = clock64();
    if (tid == 0) out[9] = (t1 - t0) / N;
    if (tid == 0) *sink = v + (unsigned)d + (unsigned)a;
}

static void run_chase(size_t bytes, bool cg) {
    const size_t n = bytes / 4;
    const size_t stride = 32;                 // one sector-sized hop: 128 B lines
    unsigned *h = (unsigned *)malloc(n * 4);
    const size_t hops = n / stride;
    // random cyclic permutation over line starts
    size_t *perm = (size_t *)malloc(hops * sizeof(size_t));
    for (size_t i = 0; i < hops; ++i) perm[i] = i;
    srand(7);
    for (size_t i = hops - 1; i > 0; --i) {
        size_t j = ((size_t)rand() * 32768 + rand()) % (i + 1);
        size_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
    for (size_t i = 0; i < n; ++i) h[i] = 0;
    for (size_t i = 0; i < hops; ++i) h[perm[i] * stride] = (unsigned)(perm[(i + 1) % hops] * stride);
    unsigned *d, *sink;
    long long *out;
    cudaMalloc(&d, n * 4);
    cudaMalloc(&sink, 4);
    cudaMalloc(&out, 16 * 8);
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    const int iters = (int)(hops < 20000 ? hops : 20000);
    for (int rep = 0; rep < 2; ++rep) {
        if (cg) chase_cg<<<1, 1>>>(d, iters, (unsigned)(perm[0] * stride), out, sink);
        else chase<<<1, 1>>>(d, iters, (unsigned)(perm[0] * stride), out, sink);
        cudaDeviceSynchronize();
    }
    long long r;
    cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost);
    printf("dependent %s over %8.1f MB: %lld cycles/hop\n", cg ? "ld.cg" : "ld   ", bytes / 1e6, r);
    cudaFree(d); cudaFree(sink); cudaFree(out); free(h); free(perm);
}

int main() {
    for (size_t mb : {1, 8, 32, 64, 100, 200, 512}) run_chase(mb << 20 >> 0, false);
    run_chase(32 << 20, true);
    run_chase(16 << 10, false);
    long long *out, h[16];
    unsigned *sink;
    cudaMalloc(&out, 16 * 8);
    cudaMalloc(&sink, 4);
    for (int nt : {32, 256, 512}) {
        misc<<<1, nt>>>(out, sink, nt / 32);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %3d: LDS.64 %lld | REDUX %lld | SHFL %lld | ballot+ffs %lld | BAR+LDS %lld | ATOMS.ADD ret %lld | DADD %lld | IMAD %lld | cmp64+sel %lld\n",
               nt, h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
    }
    return 0;
}
