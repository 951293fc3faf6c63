// comm.cu -- pool sharding across the GPUs of one NVSwitch box: the peer-visible arena, its exchange
// between ranks (CUDA IPC across processes, plain pointers inside one process) and the device-side
// barrier.  SURVEY.md §8e; the reference has no counterpart (every rank recomputes everything,
// det3d/selectors/base_selector.py:57).
//
// Each rank allocates ONE arena with a fixed layout that is identical on every rank (arena_layout below):
//     [0, 4 KB)   mailbox: greedy exchange slots, barrier counters, error flag   (comm.cuh)
//     kNN index table    int32 [n, k+1]    every rank writes its rows into every rank's copy
//     kNN distance table f64   [n, k+1]    (the all-gather is fused into the kNN kernel's stores)
//     descriptor levels  bf16  [n, dim]    each rank's BEV descriptors, pushed to every peer
//     descriptor norms   int32 [n]
//     neighbour lists    cursor | off int64 [n] | len int32 [n] | idx int32 [list_cap] | val f64 [list_cap]
//                        entries whose COLUMN this rank owns; every rank's list build scatters into them
//     d2 column slice    int32 [rows_pad, per]   dense path, world > 1: peers store the mirror images of their tiles here
// and maps every peer's arena into its own address space.  Kernels then address peer memory directly
// over NVLink; the only synchronisation primitives are the mailbox exchange and comm_barrier_kernel.
#include "comm.cuh"
#include "common.cuh"

namespace adsel {

constexpr size_t COMM_BARRIER_WORD = 2 * COMM_MAX_WORLD * COMM_SLOT_WORDS;   // [world] counters
constexpr size_t COMM_ERROR_WORD = COMM_BARRIER_WORD + COMM_MAX_WORLD;       // non-zero: a wait timed out

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct ArenaLayout {
    size_t knn_idx, knn_dist, levels, norm2, list_cursor, list_off, list_len, list_idx, list_val, d2, total;
};
static ArenaLayout arena_layout(int64_t n, int64_t k1, int64_t dim, int64_t list_cap, int64_t world) {
    ArenaLayout l;
    l.knn_idx = COMM_MAILBOX_BYTES;
    l.knn_dist = l.knn_idx + align256((size_t)n * k1 * sizeof(int32_t));
    l.levels = l.knn_dist + align256((size_t)n * k1 * sizeof(double));
    l.norm2 = l.levels + align256((size_t)n * dim * sizeof(uint16_t));
    l.list_cursor = l.norm2 + align256((size_t)n * sizeof(int32_t));
    l.list_off = l.list_cursor + 256;
    l.list_len = l.list_off + align256((size_t)n * sizeof(int64_t));
    l.list_idx = l.list_len + align256((size_t)n * sizeof(int32_t));
    l.list_val = l.list_idx + align256((size_t)list_cap * sizeof(int32_t));
    l.d2 = l.list_val + align256((size_t)list_cap * sizeof(double));
    // column slice of the descriptor distance matrix (dense path on several ranks): [rows_pad, per] int32, where
    // per = columns per rank; it lives in the arena because peers store the mirror images of their tiles into it
    size_t d2_bytes = 0;
    if (dim > 0 && world > 1) {
        const int64_t rows = (n + 127) / 128 * 128;
        int64_t per = (n + world - 1) / world;
        per = (per + 255) / 256 * 256;
        d2_bytes = (size_t)rows * per * sizeof(int32_t);
    }
    l.total = l.d2 + align256(d2_bytes);
    return l;
}
static ArenaLayout layout_of(const adsel_handle *h) { return arena_layout(h->arena_n, h->arena_k1, h->arena_dim, h->arena_list_cap, h->world); }

void update_shard(adsel_handle *h) {
    if (h->world <= 1 || h->n <= 0) {
        h->shard_lo = 0;
        h->shard_hi = h->n;
        return;
    }
    // contiguous slices, 256-aligned (block maxima of the greedy loop and the 256-wide distance tiles)
    long long per = ((long long)h->n + h->world - 1) / h->world;
    per = (per + 255) / 256 * 256;
    const long long lo = (long long)h->rank * per, hi = lo + per;
    h->shard_lo = (int32_t)(lo < h->n ? lo : h->n);
    h->shard_hi = (int32_t)(hi < h->n ? hi : h->n);
}

__global__ void comm_barrier_kernel(const CommView c, unsigned long long number) {
    const int lane = threadIdx.x;
    __threadfence_system();   // everything this rank wrote to peer memory earlier in the stream is visible first
    if (lane < c.world) st_release_sys(c.mailbox[lane] + COMM_BARRIER_WORD + c.rank, number);
    if (lane < c.world) {
        const unsigned long long *p = c.mailbox[c.rank] + COMM_BARRIER_WORD + lane;
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(p) < number) {
            if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > c.timeout_ns) {
                c.mailbox[c.rank][COMM_ERROR_WORD] = number;
                break;
            }
        }
    }
}

int comm_view(adsel_handle *h, CommView *view) {
    ADSEL_REQUIRE(h->comm_ready, "sharded call before adsel_comm_attach");
    view->rank = h->rank;
    view->world = h->world;
    view->seq_base = h->comm_seq;
    view->timeout_ns = (unsigned long long)(h->comm_timeout_s * 1e9);
    for (int w = 0; w < COMM_MAX_WORLD; ++w)
        view->mailbox[w] = w < h->world ? reinterpret_cast<unsigned long long *>(h->peer_arena[w]) : nullptr;
    return ADSEL_OK;
}

int comm_barrier(adsel_handle *h) {
    CommView view;
    ADSEL_TRY(comm_view(h, &view));
    h->comm_barrier_seq += 1;
    comm_barrier_kernel<<<1, 32, 0, h->stream>>>(view, h->comm_barrier_seq);
    h->launches += 1;
    ADSEL_CUDA(cudaGetLastError());
    return ADSEL_OK;
}

int comm_preload() {
    cudaFuncAttributes attr;
    ADSEL_CUDA(cudaFuncGetAttributes(&attr, comm_barrier_kernel));
    return ADSEL_OK;
}

static void comm_detach(adsel_handle *h) {
    for (int w = 0; w < COMM_MAX_WORLD; ++w) {
        if (h->peer_arena[w] && h->peer_ipc[w]) cudaIpcCloseMemHandle(h->peer_arena[w]);
        h->peer_arena[w] = nullptr;
        h->peer_ipc[w] = false;
    }
    h->comm_ready = false;
}

static void comm_release(adsel_handle *h) {
    comm_detach(h);
    if (h->arena) cudaFree(h->arena);
    h->arena = nullptr;
    h->arena_bytes = 0;
    h->comm_ready = false;
}

void comm_destroy(adsel_handle *h) {
    comm_release(h);
    if (h->pending_ev[0]) cudaEventDestroy(h->pending_ev[0]);
    if (h->pending_ev[1]) cudaEventDestroy(h->pending_ev[1]);
    h->pending_ev[0] = h->pending_ev[1] = nullptr;
    if (h->knn_ev[0]) cudaEventDestroy(h->knn_ev[0]);
    if (h->knn_ev[1]) cudaEventDestroy(h->knn_ev[1]);
    h->knn_ev[0] = h->knn_ev[1] = nullptr;
}

int greedy_preload();
int geodesic_preload();
int knn_preload();
int dense_preload();

}  // namespace adsel

using namespace adsel;

extern "C" {

int adsel_comm_init(adsel_handle *h, int32_t rank, int32_t world, int32_t n, int32_t k, int32_t dim, int64_t list_cap,
                    void *ipc_handle_out) {
    ADSEL_REQUIRE(h, "adsel_comm_init: null handle");
    ADSEL_REQUIRE(world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world, "adsel_comm_init: need 0 <= rank < world <= 8");
    ADSEL_REQUIRE(n > 0 && k >= 0 && k < 64 && dim >= 0, "adsel_comm_init: bad arena shape");
    ADSEL_REQUIRE(!h->pending, "adsel_comm_init: a sharded selection is pending");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    comm_release(h);
    // neighbour-list region: generous default (entries), a caller that saw an overflow passes what was needed
    if (list_cap <= 0) list_cap = k > 0 ? (long long)n * 256 / world + (long long)n * 16 : 0;
    const ArenaLayout l = arena_layout(n, k + 1, dim, list_cap, world);
    ADSEL_CUDA(cudaMalloc(&h->arena, l.total));
    ADSEL_CUDA(cudaMemset(h->arena, 0, COMM_MAILBOX_BYTES));
    h->arena_bytes = l.total;
    h->arena_n = n;
    h->arena_k1 = k + 1;
    h->arena_dim = dim;
    h->arena_list_cap = list_cap;
    h->rank = rank;
    h->world = world;
    h->comm_seq = 0;
    h->comm_barrier_seq = 0;
    h->peer_arena[rank] = h->arena;
    h->graph_ready = false;
    h->lists_mode = -1;
    update_shard(h);
    if (!h->pending_ev[0]) {
        ADSEL_CUDA(cudaEventCreate(&h->pending_ev[0]));
        ADSEL_CUDA(cudaEventCreate(&h->pending_ev[1]));
        ADSEL_CUDA(cudaEventCreate(&h->knn_ev[0]));
        ADSEL_CUDA(cudaEventCreate(&h->knn_ev[1]));
    }
    // every kernel of the sharded path is loaded now (see greedy_preload)
    ADSEL_TRY(greedy_preload());
    ADSEL_TRY(geodesic_preload());
    ADSEL_TRY(knn_preload());
    ADSEL_TRY(dense_preload());
    ADSEL_TRY(comm_preload());
    if (ipc_handle_out) {
        cudaIpcMemHandle_t ipc;
        ADSEL_CUDA(cudaIpcGetMemHandle(&ipc, h->arena));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
        memcpy(ipc_handle_out, &ipc, sizeof(ipc));
    }
    if (world == 1) h->comm_ready = true;
    return ADSEL_OK;
}

int adsel_comm_attach(adsel_handle *h, const void *ipc_handles) {
    ADSEL_REQUIRE(h && h->arena && ipc_handles, "adsel_comm_attach: call adsel_comm_init first");
    ADSEL_CUDA(cudaSetDevice(h->device));
    const unsigned char *src = static_cast<const unsigned char *>(ipc_handles);
    for (int w = 0; w < h->world; ++w) {
        if (w == h->rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, src + (size_t)w * sizeof(ipc), sizeof(ipc));
        void *p = nullptr;
        ADSEL_CUDA(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
        h->peer_arena[w] = static_cast<unsigned char *>(p);
        h->peer_ipc[w] = true;
    }
    h->comm_ready = true;
    return ADSEL_OK;
}

int adsel_comm_detach(adsel_handle *h) {
    ADSEL_REQUIRE(h, "adsel_comm_detach: null handle");
    ADSEL_REQUIRE(!h->pending, "adsel_comm_detach: a sharded selection is pending");
    ADSEL_CUDA(cudaSetDevice(h->device));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    comm_detach(h);
    if (h->arena) h->peer_arena[h->rank] = h->arena;
    return ADSEL_OK;
}

int adsel_comm_attach_local(adsel_handle *h, adsel_handle *const *handles) {
    ADSEL_REQUIRE(h && h->arena && handles, "adsel_comm_attach_local: call adsel_comm_init first");
    ADSEL_CUDA(cudaSetDevice(h->device));
    for (int w = 0; w < h->world; ++w) {
        adsel_handle *peer = handles[w];
        ADSEL_REQUIRE(peer && peer->arena && peer->world == h->world && peer->rank == w && peer->arena_bytes == h->arena_bytes,
                      "adsel_comm_attach_local: handles must be the world's handles in rank order, same arena shape");
        if (peer->device != h->device) {
            int can = 0;
            ADSEL_CUDA(cudaDeviceCanAccessPeer(&can, h->device, peer->device));
            ADSEL_REQUIRE(can, "adsel_comm_attach_local: no peer access between the two devices");
            const cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ADSEL_CUDA(e);
            cudaGetLastError();
        }
        h->peer_arena[w] = peer->arena;
        h->peer_ipc[w] = false;
    }
    h->comm_ready = true;
    return ADSEL_OK;
}

int adsel_comm_set_timeout(adsel_handle *h, double seconds) {
    ADSEL_REQUIRE(h && seconds > 0.0, "adsel_comm_set_timeout: bad argument");
    h->comm_timeout_s = seconds;
    return ADSEL_OK;
}

int adsel_comm_barrier(adsel_handle *h) {
    ADSEL_REQUIRE(h, "adsel_comm_barrier: null handle");
    ADSEL_CUDA(cudaSetDevice(h->device));
    return comm_barrier(h);
}

int adsel_comm_check(adsel_handle *h) {
    ADSEL_REQUIRE(h && h->arena, "adsel_comm_check: call adsel_comm_init first");
    ADSEL_CUDA(cudaSetDevice(h->device));
    unsigned long long flag = 0;
    ADSEL_CUDA(cudaMemcpyAsync(&flag, reinterpret_cast<unsigned long long *>(h->arena) + COMM_ERROR_WORD, sizeof(flag),
                               cudaMemcpyDeviceToHost, h->stream));
    ADSEL_CUDA(cudaStreamSynchronize(h->stream));
    if (flag != 0) {
        // reported once: the word is cleared so that a later, healthy call is not blamed for this one
        ADSEL_CUDA(cudaMemsetAsync(reinterpret_cast<unsigned long long *>(h->arena) + COMM_ERROR_WORD, 0, sizeof(flag), h->stream));
        ADSEL_CUDA(cudaStreamSynchronize(h->stream));
        set_error("device barrier " + std::to_string(flag) + " timed out waiting for a peer rank");
        return ADSEL_ERR_COMM;
    }
    return ADSEL_OK;
}

int adsel_shard_range(adsel_handle *h, int32_t *lo, int32_t *hi) {
    ADSEL_REQUIRE(h && lo && hi, "adsel_shard_range: null argument");
    *lo = h->shard_lo;
    *hi = h->shard_hi;
    return ADSEL_OK;
}

}  // extern "C"

namespace adsel {

// arena sub-buffers (same offsets on every rank)
int32_t *arena_knn_idx(adsel_handle *h, int w) { return reinterpret_cast<int32_t *>(h->peer_arena[w] + layout_of(h).knn_idx); }
double *arena_knn_dist(adsel_handle *h, int w) { return reinterpret_cast<double *>(h->peer_arena[w] + layout_of(h).knn_dist); }
uint16_t *arena_levels(adsel_handle *h, int w) { return reinterpret_cast<uint16_t *>(h->peer_arena[w] + layout_of(h).levels); }
int32_t *arena_norm2(adsel_handle *h, int w) { return reinterpret_cast<int32_t *>(h->peer_arena[w] + layout_of(h).norm2); }
int32_t *arena_d2(adsel_handle *h, int w) { return reinterpret_cast<int32_t *>(h->peer_arena[w] + layout_of(h).d2); }
ListStore arena_lists(adsel_handle *h, int w) {
    const ArenaLayout l = layout_of(h);
    unsigned char *a = h->peer_arena[w];
    ListStore st;
    st.off = reinterpret_cast<int64_t *>(a + l.list_off);
    st.len = reinterpret_cast<int32_t *>(a + l.list_len);
    st.idx = reinterpret_cast<int32_t *>(a + l.list_idx);
    st.val = reinterpret_cast<double *>(a + l.list_val);
    st.cursor = reinterpret_cast<unsigned long long *>(a + l.list_cursor);
    st.capacity = h->arena_list_cap;
    return st;
}

}  // namespace adsel
