// comm.cuh -- device-side exchange over peer memory (NVLink / NVSwitch) for the pool-sharded path.
//
// The reference issues no collective on this path: under torch.distributed.launch every rank redoes
// the whole selection (det3d/selectors/base_selector.py:57 gates only the dump).  Here the pool is
// sharded by column: every rank owns a contiguous slice of frames, and each greedy
// step ends with one exchange of the ranks' local (value, index) maxima.  The exchange is done by the
// persistent greedy kernel itself: plain 8-byte stores into every peer's mailbox (peer-mapped device
// memory) followed by a poll of its own mailbox -- no host round trip, no kernel boundary, no NCCL
// launch per step.
//
// Mailbox protocol.  A record is three 64-bit words, each carrying (payload32 << 32) | tag32 where
// tag = sequence number of the exchange: hi half of the value key, lo half of the key, frame index.
// Every word is self-validating (an 8-byte store is single-copy atomic), so no fence orders the
// words against each other and a reader simply polls until all three show the expected tag.  Slots
// are double-buffered by tag parity: a rank can post exchange s+2 only after it completed s+1, which
// needs every rank to have posted s+1, which they do only after reading all of s.
#pragma once
#include "common.cuh"

namespace adsel {

constexpr int COMM_MAX_WORLD = 8;                 // one NVSwitch box
constexpr int COMM_SLOT_WORDS = 4;                // 3 used + 1 pad (32-byte slots)
constexpr size_t COMM_MAILBOX_WORDS = 2 * COMM_MAX_WORLD * COMM_SLOT_WORDS + COMM_MAX_WORLD;   // greedy slots + barrier counters
constexpr size_t COMM_MAILBOX_BYTES = 4096;       // head of the arena, zero-initialised
constexpr unsigned COMM_ABORT_INDEX = 0xffffffffu;

struct CommView {
    int rank = 0, world = 1;
    unsigned seq_base = 0;                        // tag of this launch's first exchange minus one
    unsigned long long timeout_ns = 0;
    unsigned long long *mailbox[COMM_MAX_WORLD] = {nullptr};   // every rank's mailbox as seen from this device
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct ExchangeResult {
    unsigned long long key;   // winning value key (larger wins)
    int index;                // winning frame (lower wins among bit-equal keys)
    int status;               // 0 ok, 1 a rank posted an abort, 2 timed out waiting for a peer
};

// Per-lane mailbox addresses, resolved once per kernel (indexing the pointer table of a kernel parameter
// with a run-time lane id would otherwise go through local memory on every exchange).
struct ExchangeLane {
    unsigned long long *post;          // parity-0 slot word this lane writes in its peer's mailbox
    const unsigned long long *poll;    // parity-0 slot word this lane reads in the own mailbox
    bool active;
};
__device__ __forceinline__ ExchangeLane make_exchange_lane(const CommView &c) {
    const int lane = threadIdx.x & 31;
    const int peer = lane / 3, word = lane - 3 * peer;
    ExchangeLane e;
    e.active = lane < 3 * c.world;
    e.post = nullptr;
    e.poll = nullptr;
    if (e.active) {
        e.post = c.mailbox[peer] + ((size_t)c.rank * COMM_SLOT_WORDS + word);
        e.poll = c.mailbox[c.rank] + ((size_t)peer * COMM_SLOT_WORDS + word);
    }
    return e;
}

// Called by ALL 32 lanes of one warp with the same (key, index, tag).  Posts this rank's record to
// every rank (itself included) and returns the global winner, identical on every rank:
// larger key first, then lower index -- np.argmax's first-maximum rule across shards.
// index == COMM_ABORT_INDEX posts an abort that every rank observes in the same exchange.
// post == false: only wait for the records (another CTA of this rank's cluster posts the rank's record).
__device__ __forceinline__ ExchangeResult exchange_best(const CommView &c, const ExchangeLane &e, unsigned tag,
                                                        unsigned long long key, unsigned index, bool post = true) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t parity_off = (size_t)(tag & 1u) * COMM_MAX_WORLD * COMM_SLOT_WORDS;
    const int word = lane % 3;
    const bool active = e.active;
    if (active && post) {
        const unsigned long long payload = word == 0 ? (key >> 32) : word == 1 ? (key & 0xffffffffull) : (unsigned long long)index;
        st_relaxed_sys(e.post + parity_off, (payload << 32) | tag);
    }
    unsigned long long got = 0;
    bool ok = true;
    if (active) {
        const unsigned long long *p = e.poll + parity_off;
        const unsigned long long t0 = global_timer_ns();
        unsigned spins = 0;
        while (true) {
            got = ld_relaxed_sys(p);
            if ((unsigned)got == tag) break;
            if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > c.timeout_ns) {
                ok = false;
                break;
            }
        }
    }
    ok = __all_sync(full, ok);
    // lane w < world assembles rank w's record from lanes 3w, 3w+1, 3w+2
    const int src = lane < c.world ? 3 * lane : 0;
    const unsigned long long hi = __shfl_sync(full, got, src) >> 32;
    const unsigned long long lo = __shfl_sync(full, got, src + 1) >> 32;
    const unsigned idx = (unsigned)(__shfl_sync(full, got, src + 2) >> 32);
    unsigned long long k = 0ull;
    int i = 0x7fffffff;
    bool abort_seen = false;
    if (lane < c.world) {
        abort_seen = idx == COMM_ABORT_INDEX;
        if (!abort_seen) {
            k = (hi << 32) | lo;
            i = (int)idx;
        }
    }
    abort_seen = __any_sync(full, abort_seen);
    {   // global winner: three REDUX operations (see warp_best in greedy.cu)
        const unsigned khi = (unsigned)(k >> 32), klo = (unsigned)k;
        const unsigned mhi = __reduce_max_sync(full, khi);
        const unsigned mlo = __reduce_max_sync(full, khi == mhi ? klo : 0u);
        i = __reduce_min_sync(full, (khi == mhi && klo == mlo) ? i : 0x7fffffff);
        k = ((unsigned long long)mhi << 32) | mlo;
    }
    ExchangeResult r;
    r.key = k;
    r.index = i;
    r.status = !ok ? 2 : (abort_seen ? 1 : 0);
    return r;
}

}  // namespace adsel
