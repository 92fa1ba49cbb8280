// Shard merge fused with its exchange over NVLink peer memory ("local fused top-k, one
// exchange of N_q * k keys per rank, merge").  Instead of ncclAllGather + layout copy + merge kernel, ONE kernel per
// rank pushes its (n_q, k) sorted keys straight into every peer's mailbox (plain stores to peer-mapped memory,
// cudaIpc), publishes a per-(source rank, query) sequence flag with system-scope release, waits for the world's flags
// of its own query, and merges the world * k keys in shared memory.  No collective library call on the step.
//
// Mailbox (one per rank, cudaMalloc + cudaIpc, zero-initialised), all offsets in bytes:
//   [0, 64)                         header: seq counter, exit counter, status (local use only)
//   flags  [4][world][max_q] u32    sequence number of the last push per buffer / source rank / query
//   data   [4][world][max_q * max_k] u64
// Buffer = sequence & 3.  Modes: SYNC pushes step s and merges step s; DEFERRED pushes step s and merges step s-1
// (whose keys arrived a whole step ago: the exchange latency and the skew between ranks disappear behind the next
// scan); FLUSH merges the last pushed step without pushing.  A peer's push of step s+4 (the buffer this rank reads
// for step s) is ordered after the peer's step-(s+3) kernel, which waited for this rank's step-(s+2) flag, which
// this rank's stream issues only after its step-(s+1) kernel -- the last reader of step s -- completed: 4 buffers
// make the reuse safe in every mode.
#include "common.cuh"

namespace mdir {

constexpr int kMaxWorld = 16;
constexpr int kMailBufs = 4;
enum { kExchangeSync = 0, kExchangeDeferred = 1, kExchangeFlush = 2 };

struct Mailboxes {
    uint8_t* base[kMaxWorld];
};

__device__ __forceinline__ uint32_t* mb_flags(uint8_t* mb, int world, int max_q, int parity, int src) {
    return reinterpret_cast<uint32_t*>(mb + 64) + ((int64_t)parity * world + src) * max_q;
}
__device__ __forceinline__ uint64_t* mb_data(uint8_t* mb, int world, int max_q, int max_k, int parity, int src) {
    const int64_t flag_bytes = ((int64_t)kMailBufs * world * max_q * 4 + 63) & ~(int64_t)63;
    return reinterpret_cast<uint64_t*>(mb + 64 + flag_bytes) + ((int64_t)parity * world + src) * ((int64_t)max_q * max_k);
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// grid = n_q CTAs of 256 threads; dynamic smem = world * k * 8 bytes
__global__ void __launch_bounds__(256) shard_exchange_merge_kernel(const uint64_t* __restrict__ local_keys, int n_q, int k, int rank, int world,
                                                                   int max_q, int max_k, int mode, Mailboxes mbs,
                                                                   float* __restrict__ out_scores, int32_t* __restrict__ out_idx,
                                                                   const int32_t* __restrict__ local_status, int32_t* __restrict__ out_status) {
    extern __shared__ uint64_t sk[];
    __shared__ int s_fail;
    __shared__ uint32_t s_status;
    const int q = blockIdx.x;
    uint8_t* mine = mbs.base[rank];
    uint32_t* hdr = reinterpret_cast<uint32_t*>(mine);
    const uint32_t last = *reinterpret_cast<volatile uint32_t*>(hdr);          // same for every CTA: bumped by the last one out
    const uint32_t seq = mode == kExchangeFlush ? last : last + 1u;            // step pushed by this launch (none when flushing)
    const uint32_t mseq = mode == kExchangeDeferred ? seq - 1u : seq;          // step merged by this launch
    const int parity = (int)(seq & (kMailBufs - 1));
    if (threadIdx.x == 0) { s_fail = 0; s_status = 0u; }
    // The flag word carries the sequence number AND this rank's 2-bit status of the query (candidate overflow /
    // failed shortlist certificate): every rank ORs the world's statuses, so all of them agree on which queries of
    // the merged step need the (collective) recovery -- a rank never silently merges a peer's incomplete keys.
    const uint32_t my_status = (local_status && mode != kExchangeFlush) ? ((uint32_t)local_status[q] & 3u) : 0u;

    // 1. push this rank's keys of query q into every mailbox (own included), then publish the flag
    if (mode != kExchangeFlush) {
        for (int r = 0; r < world; ++r) {
            uint64_t* dst = mb_data(mbs.base[r], world, max_q, max_k, parity, rank) + (int64_t)q * max_k;
            for (int j = threadIdx.x; j < k; j += blockDim.x) dst[j] = local_keys[(int64_t)q * k + j];
        }
    }
    __syncthreads();                      // CTA-scope: every thread's stores happen-before the flag writers below
    if (threadIdx.x < world) {
        if (mode != kExchangeFlush) {
            __threadfence_system();       // one cumulative system-scope fence per flag writer, not one per thread
            st_release_sys(mb_flags(mbs.base[threadIdx.x], world, max_q, parity, rank) + q, (seq << 2) | my_status);
        }
        // 2. wait for source rank threadIdx.x's keys of query q of the step being merged (bounded: ~2 s, then report
        //    instead of hanging)
        if (mseq != 0u) {
            const uint32_t* f = mb_flags(mine, world, max_q, (int)(mseq & (kMailBufs - 1)), threadIdx.x) + q;
            bool ok = false;
            for (int it = 0; it < (1 << 24); ++it) {
                const uint32_t v = ld_acquire_sys(f);
                if ((v >> 2) == (mseq & 0x3fffffffu)) { ok = true; if (v & 3u) atomicOr(&s_status, v & 3u); break; }
                __nanosleep(100);
            }
            if (!ok) s_fail = 1;
        }
    }
    __syncthreads();
    const int mpar = (int)(mseq & (kMailBufs - 1));
    if (threadIdx.x == 0 && out_status) out_status[q] = (int32_t)(s_status | (s_fail ? 4u : 0u));
    if (s_fail || mseq == 0u) {
        if (threadIdx.x == 0 && s_fail) hdr[2] = 1u;                            // status: a peer never arrived
        for (int j = threadIdx.x; j < k; j += blockDim.x) {
            out_scores[(int64_t)q * k + j] = -INFINITY;
            out_idx[(int64_t)q * k + j] = -1;
        }
    } else {
        // 3. merge world sorted lists of k unique keys without sorting: the output position of a key is the number of
        //    keys smaller than it = its position in its own list + a lower_bound in each other list (shared memory)
        const int n = world * k;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int r = i / k, j = i - r * k;
            sk[i] = __ldcg(mb_data(mine, world, max_q, max_k, mpar, r) + (int64_t)q * max_k + j);   // written by a peer: bypass L1
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t key = sk[i];
            if (key == ~0ull) continue;                       // padding of a short shard list
            const int r = i / k;
            int pos = i - r * k;
            for (int o = 0; o < world && pos < k; ++o) {
                if (o == r) continue;
                const uint64_t* lst = sk + o * k;
                int lo = 0, hi = k;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (lst[mid] < key) lo = mid + 1; else hi = mid;
                }
                pos += lo;
            }
            if (pos < k) {
                out_scores[(int64_t)q * k + pos] = key_score(key);
                out_idx[(int64_t)q * k + pos] = (int32_t)(uint32_t)key;
            }
        }
        // fewer than k real keys in the whole world: pad the tail
        {
            int real = 0;
            for (int o = 0; o < world; ++o) {
                const uint64_t* lst = sk + o * k;
                int lo = 0, hi = k;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (lst[mid] != ~0ull) lo = mid + 1; else hi = mid;
                }
                real += lo;
            }
            for (int j = real + threadIdx.x; j < k; j += blockDim.x) {
                out_scores[(int64_t)q * k + j] = -INFINITY;
                out_idx[(int64_t)q * k + j] = -1;
            }
        }
    }
    // the last CTA out advances the sequence for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&hdr[1], 1u) == gridDim.x - 1) {
            hdr[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(hdr) = seq;                  // unchanged by a flush
        }
    }
}

}  // namespace mdir

using namespace mdir;

extern "C" size_t mdir_shard_mailbox_bytes(int world, int max_q, int max_k) {
    if (world < 1 || max_q < 1 || max_k < 1) return 0;
    const size_t flag_bytes = ((size_t)kMailBufs * world * max_q * 4 + 63) & ~(size_t)63;
    return 64 + flag_bytes + (size_t)kMailBufs * world * max_q * max_k * 8;
}

// cudaMalloc'ed (not from a caching allocator: the IPC handle must describe the whole allocation), zeroed.
// handle64 receives the 64-byte cudaIpcMemHandle_t to hand to the peers.
extern "C" int mdir_p2p_alloc(size_t bytes, void** ptr, void* handle64) {
    MDIR_CHECK_ARG(ptr && handle64 && bytes > 0);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    MDIR_CUDA(cudaMalloc(ptr, bytes));
    MDIR_CUDA(cudaMemset(*ptr, 0, bytes));
    MDIR_CUDA(cudaDeviceSynchronize());
    MDIR_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), *ptr));
    return 0;
}

extern "C" int mdir_p2p_open(const void* handle64, void** ptr) {
    MDIR_CHECK_ARG(ptr && handle64);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    MDIR_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int mdir_p2p_close(void* ptr) {
    if (ptr) MDIR_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int mdir_p2p_free(void* ptr) {
    if (ptr) MDIR_CUDA(cudaFree(ptr));
    return 0;
}

// mailboxes: host array of `world` device pointers (entry `rank` = this rank's own allocation, the others peer-mapped
// with mdir_p2p_open).  Every rank must call this with the same (n_q, k, mode) in the same order.  mode 0: exchange
// and merge this step; 1: push this step, merge the previous one (same n_q, k) into out_*; 2: merge the last pushed
// step without pushing (the drain after a run of mode-1 calls).  status = word 2 of the
// own mailbox header (mdir_shard_status): non-zero after a peer failed to arrive within the bounded wait.
extern "C" int mdir_shard_exchange_merge(const uint64_t* local_keys, int n_q, int k, int rank, int world, int max_q, int max_k, int mode,
                                         void* const* mailboxes, float* out_scores, int32_t* out_idx, const int32_t* local_status,
                                         int32_t* out_status, void* stream) {
    MDIR_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && n_q >= 0 && n_q <= max_q && k >= 1 && k <= max_k);
    MDIR_CHECK_ARG(world * (int64_t)k <= 16384 && mode >= kExchangeSync && mode <= kExchangeFlush);
    if (n_q == 0) return 0;
    MDIR_CHECK_ARG((local_keys || mode == kExchangeFlush) && mailboxes && out_scores && out_idx);
    Mailboxes mbs;
    for (int r = 0; r < kMaxWorld; ++r) mbs.base[r] = r < world ? static_cast<uint8_t*>(mailboxes[r]) : nullptr;
    for (int r = 0; r < world; ++r) MDIR_CHECK_ARG(mbs.base[r] != nullptr);
    const size_t smem = (size_t)world * k * 8;
    static PerDeviceOnce once;
    if (once.first() != 0) {
        MDIR_CUDA(cudaFuncSetAttribute(shard_exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        // the same L1 / shared-memory split as the scan kernel: CTAs of kernels with different carve-outs cannot share an
        // SM, and this kernel is meant to run beside the next step's scan (SearchPipeline overlap mode)
        MDIR_CUDA(cudaFuncSetAttribute(shard_exchange_merge_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    shard_exchange_merge_kernel<<<n_q, 256, smem, (cudaStream_t)stream>>>(local_keys, n_q, k, rank, world, max_q, max_k, mode, mbs, out_scores, out_idx,
                                                                          local_status, out_status);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_shard_status(const void* own_mailbox, int* status) {
    MDIR_CHECK_ARG(own_mailbox && status);
    uint32_t v = 0;
    MDIR_CUDA(cudaMemcpy(&v, static_cast<const uint8_t*>(own_mailbox) + 8, 4, cudaMemcpyDeviceToHost));
    *status = (int)v;
    return 0;
}
