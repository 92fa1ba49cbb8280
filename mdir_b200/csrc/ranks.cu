// Full per-query ranking: segmented, stable LSD radix sort on (score desc, index asc).
// Replaces np.argsort(-scores, axis=0) (cirscore.py:70) for the drop-in (n_db, n_q) ranks array.
//
// Layout: everything is query-major (n_q segments of n_db keys) on the device so all
// streams are coalesced; the reference's (n_db, n_q) C-order appears only in the first
// (score transpose -> u32 keys) and last (u32 ranks -> int64 transpose) kernels.
// 4 passes x {digit histogram per 2048-key chunk, per-segment scan, stable scatter staged through
// shared memory so each digit's keys leave as one coalesced run}.
#include "common.cuh"

namespace mdir {

constexpr int kItems = 8;             // keys per thread in the scatter kernel
constexpr int kScatterThreads = 256;
constexpr int kChunk = kScatterThreads * kItems;   // 2048 keys per CTA
constexpr int kScatterWarps = kScatterThreads / 32;

__device__ __forceinline__ uint32_t rank_key(float s) { return desc_key(s); }   // ascending key == descending score, NaN last

// scores (n_db, n_q) -> keys (n_q, n_db)
__global__ void __launch_bounds__(256) keys_transpose_kernel(const float* __restrict__ scores, int64_t n_db, int n_q,
                                                             uint32_t* __restrict__ keys) {
    __shared__ float tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int q0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int64_t row = r0 + r;
        const int q = q0 + tx;
        tile[r][tx] = (row < n_db && q < n_q) ? scores[row * n_q + q] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int q = q0 + r;
        const int64_t row = r0 + tx;
        if (q < n_q && row < n_db) keys[(int64_t)q * n_db + row] = rank_key(tile[tx][r]);
    }
}

__global__ void __launch_bounds__(256) keys_direct_kernel(const float* __restrict__ scores, int64_t total, uint32_t* __restrict__ keys) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < total) {
        const float4 v = *reinterpret_cast<const float4*>(scores + i);
        *reinterpret_cast<uint4*>(keys + i) = make_uint4(rank_key(v.x), rank_key(v.y), rank_key(v.z), rank_key(v.w));
    } else {
        for (int64_t j = i; j < total; ++j) keys[j] = rank_key(scores[j]);
    }
}

// counts[(q * 256 + digit) * n_chunks + chunk]
__global__ void __launch_bounds__(256) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n_db, int n_chunks, int shift,
                                                         uint32_t* __restrict__ counts) {
    __shared__ uint32_t h[256];
    const int chunk = blockIdx.x, q = blockIdx.y;
    h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t* src = keys + (int64_t)q * n_db;
    const int64_t base = (int64_t)chunk * kChunk;
    const int64_t end = min(base + kChunk, n_db);
    for (int64_t i = base + threadIdx.x; i < end; i += 256) atomicAdd(&h[(src[i] >> shift) & 0xffu], 1u);
    __syncthreads();
    counts[((int64_t)q * 256 + threadIdx.x) * n_chunks + chunk] = h[threadIdx.x];
}

// exclusive scan of counts[q] in (digit-major, chunk-minor) order, in place.  One CTA per q.
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ counts, int n_chunks) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const int q = blockIdx.x;
    uint32_t* c = counts + (int64_t)q * 256 * n_chunks;
    const int total = 256 * n_chunks;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (int base = 0; base < total; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < total ? c[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) wsum[w] = s;
        __syncthreads();
        if (w == 0) {
            uint32_t t = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += u;
            }
            wsum[lane] = t;      // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t excl = carry + (w ? wsum[w - 1] : 0u) + s - v;
        if (i < total) c[i] = excl;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + wsum[31];
        __syncthreads();
    }
}

// Stable scatter of one chunk.  vals_in == nullptr means "value = position" (first pass).
// Two phases: (1) every key gets its position inside the CHUNK sorted by digit (per-warp counters +
// match.any ranking, then a scan over warps and digits) and is staged there in shared memory;
// (2) the staged chunk is written out in order, so the keys of one digit go to consecutive global
// addresses (runs of ~32 keys = full 128-byte lines instead of 32 scattered 4-byte stores).
__global__ void __launch_bounds__(kScatterThreads, 4) radix_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                        const uint32_t* __restrict__ vals_in, int64_t n_db, int n_chunks,
                                                                        int shift, const uint32_t* __restrict__ offsets,
                                                                        uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    extern __shared__ uint32_t sm[];
    uint32_t* skey = sm;                               // kChunk
    uint32_t* sval = sm + kChunk;                      // kChunk
    uint32_t* cnt = sm + 2 * kChunk;                   // kScatterWarps x 256
    uint32_t* delta = cnt + kScatterWarps * 256;       // 256: global offset - local base of each digit
    uint32_t* wsum = delta + 256;                      // 8 (digit scan over 256 threads)
    const int chunk = blockIdx.x, q = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kScatterWarps * 256; i += kScatterThreads) cnt[i] = 0u;
    __syncthreads();
    const int64_t seg = (int64_t)q * n_db;
    const int64_t cbase = (int64_t)chunk * kChunk;
    const int64_t base = cbase + (int64_t)w * (kItems * 32);
    const int n_valid = (int)min((int64_t)kChunk, n_db - cbase);
    uint32_t key[kItems], rank[kItems];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        const int64_t i = base + it * 32 + lane;
        key[it] = i < n_db ? keys_in[seg + i] : 0u;
    }
    uint32_t* mycnt = cnt + w * 256;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        const int64_t i = base + it * 32 + lane;
        const bool valid = i < n_db;
        const int d = (int)((key[it] >> shift) & 0xffu);
        const unsigned peers = match_digit8(d, valid);
        uint32_t r = 0;
        if (valid) r = mycnt[d] + __popc(peers & lt);
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) mycnt[d] += __popc(peers);
        __syncwarp();
        rank[it] = r;
    }
    __syncthreads();
    // per digit: exclusive scan over the warps, then an exclusive scan of the digit totals over the 256 digits
    if (threadIdx.x < 256) {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < kScatterWarps; ++k) {
            const uint32_t t = cnt[k * 256 + d];
            cnt[k * 256 + d] = run;
            run += t;
        }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[w] = incl;
        // (threads 256..511 skip this block; the barrier below is outside it)
        delta[d] = incl - run;                          // exclusive within the warp for now
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        const int d = threadIdx.x;
        uint32_t pre = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < w) pre += wsum[k];
        const uint32_t local_base = delta[d] + pre;
#pragma unroll
        for (int k = 0; k < kScatterWarps; ++k) cnt[k * 256 + d] += local_base;      // warp offsets become chunk positions
        delta[d] = offsets[((int64_t)q * 256 + d) * n_chunks + chunk] - local_base;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        const int64_t i = base + it * 32 + lane;
        if (i < n_db) {
            const int d = (int)((key[it] >> shift) & 0xffu);
            const uint32_t l = mycnt[d] + rank[it];
            skey[l] = key[it];
            sval[l] = vals_in ? vals_in[seg + i] : (uint32_t)i;
        }
    }
    __syncthreads();
    for (int l = threadIdx.x; l < n_valid; l += kScatterThreads) {
        const uint32_t k = skey[l];
        const uint32_t pos = (uint32_t)l + delta[(k >> shift) & 0xffu];
        keys_out[seg + pos] = k;
        vals_out[seg + pos] = sval[l];
    }
}

// vals (n_q, n_db) u32 -> ranks (n_db, n_q) int64
__global__ void __launch_bounds__(256) ranks_transpose_kernel(const uint32_t* __restrict__ vals, int64_t n_db, int n_q,
                                                              int64_t* __restrict__ ranks, int64_t ranks_ld) {
    __shared__ uint32_t tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int q0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int q = q0 + r;
        const int64_t row = r0 + tx;
        tile[r][tx] = (q < n_q && row < n_db) ? vals[(int64_t)q * n_db + row] : 0u;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t row = r0 + r;
        const int q = q0 + tx;
        if (row < n_db && q < n_q) ranks[row * ranks_ld + q] = (int64_t)tile[tx][r];
    }
}

}  // namespace mdir

using namespace mdir;

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" size_t mdir_rank_workspace_bytes(int64_t n_db, int n_q) {
    if (n_db <= 0 || n_q <= 0) return 0;
    const size_t arr = align256((size_t)n_db * n_q * 4);
    const int64_t n_chunks = (n_db + kChunk - 1) / kChunk;
    return 4 * arr + align256((size_t)n_q * 256 * n_chunks * 4);
}

extern "C" int mdir_rank_scores(const float* scores, int64_t n_db, int n_q, int query_major, int64_t* ranks, int64_t ranks_ld,
                                void* ws, void* stream) {
    MDIR_CHECK_ARG(scores && ranks && ws && n_db >= 1 && n_q >= 1 && ranks_ld >= n_q);
    MDIR_CHECK_ARG(n_db < ((int64_t)1 << 32) && n_q <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t arr = align256((size_t)n_db * n_q * 4);
    uint8_t* w = (uint8_t*)ws;
    uint32_t* kA = (uint32_t*)w;
    uint32_t* kB = (uint32_t*)(w + arr);
    uint32_t* vA = (uint32_t*)(w + 2 * arr);
    uint32_t* vB = (uint32_t*)(w + 3 * arr);
    uint32_t* counts = (uint32_t*)(w + 4 * arr);
    const int n_chunks = (int)((n_db + kChunk - 1) / kChunk);
    const unsigned gx = (unsigned)((n_db + 31) / 32), gy = (unsigned)((n_q + 31) / 32);
    if (query_major) {
        const int64_t total = n_db * n_q;
        MDIR_CHECK_ARG(((uintptr_t)scores & 15) == 0);
        keys_direct_kernel<<<(unsigned)((total / 4 + 256) / 256), 256, 0, st>>>(scores, total, kA);
    } else {
        keys_transpose_kernel<<<dim3(gx, gy), 256, 0, st>>>(scores, n_db, n_q, kA);
    }
    MDIR_LAUNCH_CHECK();
    constexpr size_t kScatterSmem = (size_t)(2 * kChunk + kScatterWarps * 256 + 256 + 8) * 4;
    static PerDeviceOnce once;
    if (once.first() != 0)
        MDIR_CUDA(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem));
    uint32_t *kin = kA, *kout = kB, *vin = nullptr, *vout = vA;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        radix_hist_kernel<<<dim3(n_chunks, n_q), 256, 0, st>>>(kin, n_db, n_chunks, shift, counts);
        MDIR_LAUNCH_CHECK();
        radix_scan_kernel<<<n_q, 1024, 0, st>>>(counts, n_chunks);
        MDIR_LAUNCH_CHECK();
        radix_scatter_kernel<<<dim3(n_chunks, n_q), kScatterThreads, kScatterSmem, st>>>(kin, vin, n_db, n_chunks, shift, counts, kout, vout);
        MDIR_LAUNCH_CHECK();
        uint32_t* t = kin; kin = kout; kout = t;
        vin = vout;
        vout = (vout == vA) ? vB : vA;
    }
    ranks_transpose_kernel<<<dim3(gx, gy), 256, 0, st>>>(vin, n_db, n_q, ranks, ranks_ld);
    MDIR_LAUNCH_CHECK();
    return 0;
}
