// sssp.cuh -- frontier-based label-correcting single-source shortest paths by one CTA (shared by geodesic.cu and lazy.cu).
//
// scipy's Dijkstra row for source s is the least fixpoint of  d[v] = min_u fl(d[u] + w(u,v)),  d[s] = 0  in IEEE
// float64 (non-negative weights, monotone rounding); a label-correcting relaxation reaches the same fixpoint bit for
// bit whatever the relaxation order (det3d/selectors/spatial_temporal_selector.py:103-104).
#pragma once
#include "common.cuh"

namespace adsel {

constexpr int SSSP_THREADS = 128;
constexpr int SSSP_LANES_PER_NODE = 8;   // lanes cooperating on one frontier vertex's edge list

struct SsspScratch {
    unsigned long long *dist;   // [n] ordered-u64 distance, INF_BITS when untouched
    int32_t *stamp;             // [n] round tag of the vertex's last frontier insertion
    int32_t *q0, *q1;           // [n] frontier queues
    int32_t *touched;           // [n] vertices whose dist was ever lowered for this source
};

struct SsspShared {
    int n_front, n_next, n_touched, tag;
};

// Runs the relaxation for `src`; on return sh.n_touched vertices are listed in sc.touched with
// their final distances in sc.dist.  The caller must reset sc.dist for the touched vertices.
template <int THREADS = SSSP_THREADS>
__device__ __forceinline__ void sssp_cta(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                         const double *__restrict__ w, int src, double radius,
                                         const SsspScratch &sc, SsspShared &sh) {
    if (threadIdx.x == 0) {
        sc.dist[src] = 0ull;
        sc.q0[0] = src;
        sc.touched[0] = src;
        sh.n_front = 1;
        sh.n_touched = 1;
        sh.n_next = 0;
    }
    __syncthreads();
    int32_t *front = sc.q0, *next = sc.q1;
    const int sub = threadIdx.x % SSSP_LANES_PER_NODE;
    const int slot = threadIdx.x / SSSP_LANES_PER_NODE;
    constexpr int NODES_PER_PASS = THREADS / SSSP_LANES_PER_NODE;
    while (true) {
        const int nf = sh.n_front;
        const int tag = sh.tag + 1;
        if (nf == 0) break;
        for (int base = 0; base < nf; base += NODES_PER_PASS) {
            const int qi = base + slot;
            if (qi < nf) {
                const int u = __ldcg(&front[qi]);
                const double du = __longlong_as_double((long long)__ldcg(&sc.dist[u]));
                if (du < radius) {
                    const int e1 = row_ptr[u + 1];
                    for (int e = row_ptr[u] + sub; e < e1; e += SSSP_LANES_PER_NODE) {
                        const int v = col[e];
                        const unsigned long long nd = (unsigned long long)__double_as_longlong(__dadd_rn(du, w[e]));
                        const unsigned long long old = atomicMin(&sc.dist[v], nd);
                        if (nd < old) {
                            if (old == INF_BITS) sc.touched[atomicAdd(&sh.n_touched, 1)] = v;
                            if (atomicExch(&sc.stamp[v], tag) != tag) next[atomicAdd(&sh.n_next, 1)] = v;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            sh.n_front = sh.n_next;
            sh.n_next = 0;
            sh.tag = tag;
        }
        int32_t *tmp = front;
        front = next;
        next = tmp;
        __syncthreads();
    }
}

__device__ __forceinline__ SsspScratch scratch_of(unsigned long long *dist, int32_t *stamp, int32_t *q0, int32_t *q1,
                                                  int32_t *touched, size_t n) {
    SsspScratch sc;
    const size_t o = (size_t)blockIdx.x * n;
    sc.dist = dist + o;
    sc.stamp = stamp + o;
    sc.q0 = q0 + o;
    sc.q1 = q1 + o;
    sc.touched = touched + o;
    return sc;
}

}  // namespace adsel
