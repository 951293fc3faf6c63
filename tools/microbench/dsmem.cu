// Cost of one all-to-all exchange of an 8-byte tagged word between the CTAs of a thread-block cluster (development aid):
// (a) st.shared::cluster + poll on the own copy, (b) hardware cluster barrier + remote loads.
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void exch_poll(long long *out, int iters, int sleep_ns) {
    __shared__ unsigned long long slot[2][16];
    cg::cluster_group cl = cg::this_cluster();
    const int crank = cl.block_rank(), csize = cl.num_blocks();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) slot[threadIdx.x >> 4][threadIdx.x & 15] = 0;
    cl.sync();
    long long t0 = clock64();
    if (threadIdx.x < 32) {
        for (unsigned s = 1; s <= (unsigned)iters; ++s) {
            const unsigned base = (unsigned)__cvta_generic_to_shared(&slot[s & 1][0]);
            if (lane < csize) {
                unsigned dst;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(base + crank * 8), "r"(lane));
                asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(dst), "l"(((unsigned long long)crank << 32) | s) : "memory");
            }
            if (lane < csize) {
                unsigned long long got;
                while (true) {
                    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(got) : "r"(base + lane * 8) : "memory");
                    if ((unsigned)got == s) break;
                    if (sleep_ns) __nanosleep(sleep_ns);
                }
            }
            __syncwarp();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && crank == 0) out[0] = (t1 - t0) / iters;
    cl.sync();
}

__global__ void exch_barrier(long long *out, int iters) {
    __shared__ unsigned long long slot[2];
    cg::cluster_group cl = cg::this_cluster();
    const int crank = cl.block_rank(), csize = cl.num_blocks();
    cl.sync();
    long long t0 = clock64();
    unsigned long long acc = 0;
    for (unsigned s = 1; s <= (unsigned)iters; ++s) {
        if (threadIdx.x == 0) slot[s & 1] = ((unsigned long long)crank << 32) | s;
        asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (threadIdx.x < csize) {
            const unsigned long long *p = cl.map_shared_rank(&slot[s & 1], threadIdx.x);
            acc += *p;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && crank == 0) out[0] = (t1 - t0) / iters + (acc == 1);
    cl.sync();
}

template <typename K, typename... A>
static long long run(K kernel, int csize, int threads, A... args) {
    long long *out;
    cudaMalloc(&out, 8);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(threads);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, out, args...);
    cudaDeviceSynchronize();
    long long r = -1;
    if (e == cudaSuccess && cudaGetLastError() == cudaSuccess) cudaMemcpy(&r, out, 8, cudaMemcpyDeviceToHost);
    cudaFree(out);
    return r;
}

int main() {
    for (int c : {1, 2, 4, 8, 16}) {
        printf("cluster %2d: st.shared::cluster + poll %lld cycles/exchange (spin), %lld (nanosleep 20); cluster barrier + remote loads (256 threads) %lld\n",
               c, run(exch_poll, c, 256, 20000, 0), run(exch_poll, c, 256, 20000, 20), run(exch_barrier, c, 256, 20000));
    }
    return 0;
}
