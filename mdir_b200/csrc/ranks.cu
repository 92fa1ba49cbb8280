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

// exclusive scan of counts[q] in (bin-major, chunk-minor) order, in place.  One CTA per q.  n_bins = 256 radix digits,
// or the number of buckets of the sample-sort path below.
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ counts, int n_chunks, int n_bins) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const int q = blockIdx.x;
    uint32_t* c = counts + (int64_t)q * n_bins * n_chunks;
    const int total = n_bins * n_chunks;
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

// vals (n_q, n_db) u32 -> ranks (n_db, n_q) int64, on 64 x 64 tiles: a warp reads 64 consecutive rows of one query (256 B) and writes one row's 64
// consecutive queries (512 B, 16-byte streaming stores).
// vec != 0: n_db even, vals 8-byte aligned, ranks 16-byte aligned, ranks_ld even.
__global__ void __launch_bounds__(256) ranks_transpose64_kernel(const uint32_t* __restrict__ vals, int64_t n_db, int n_q,
                                                                int64_t* __restrict__ ranks, int64_t ranks_ld, int vec) {
    __shared__ uint32_t tile[64][65];                 // [query][row]
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const int q0 = blockIdx.y * 64;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (vec) {
        const int64_t row = r0 + 2 * lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ql = w + 8 * i, q = q0 + ql;
            uint2 v = make_uint2(0u, 0u);
            if (q < n_q && row < n_db) v = *reinterpret_cast<const uint2*>(vals + (int64_t)q * n_db + row);      // n_db even: row + 1 < n_db too
            tile[ql][2 * lane] = v.x;
            tile[ql][2 * lane + 1] = v.y;
        }
        __syncthreads();
        const int q = q0 + 2 * lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = w + 8 * i;
            const int64_t r = r0 + rl;
            if (r < n_db) {
                int64_t* dst = ranks + r * ranks_ld + q;
                if (q + 1 < n_q) {
                    const longlong2 o = make_longlong2((long long)tile[2 * lane][rl], (long long)tile[2 * lane + 1][rl]);
                    __stcs(reinterpret_cast<longlong2*>(dst), o);
                } else if (q < n_q) {
                    __stcs(dst, (int64_t)tile[2 * lane][rl]);
                }
            }
        }
    } else {
        for (int i = w; i < 64; i += 8) {
            const int q = q0 + i;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t row = r0 + lane + 32 * h;
                tile[i][lane + 32 * h] = (q < n_q && row < n_db) ? vals[(int64_t)q * n_db + row] : 0u;
            }
        }
        __syncthreads();
        for (int i = w; i < 64; i += 8) {
            const int64_t r = r0 + i;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = q0 + lane + 32 * h;
                if (r < n_db && q < n_q) ranks[r * ranks_ld + q] = (int64_t)tile[lane + 32 * h][i];
            }
        }
    }
}

static int launch_ranks_transpose(const uint32_t* vals, int64_t n_db, int n_q, int64_t* ranks, int64_t ranks_ld, cudaStream_t st) {
    const int vec = ((n_db & 1) == 0 && (ranks_ld & 1) == 0 && (((uintptr_t)vals) & 7) == 0 && (((uintptr_t)ranks) & 15) == 0) ? 1 : 0;
    ranks_transpose64_kernel<<<dim3((unsigned)((n_db + 63) / 64), (unsigned)((n_q + 63) / 64)), 256, 0, st>>>(vals, n_db, n_q, ranks, ranks_ld, vec);
    MDIR_LAUNCH_CHECK();
    return 0;
}


// ============================================================================================================
// Sample-sort path (mdir_rank_scores_fast): ONE partition pass + one shared-memory sort per bucket.
//
// The LSD radix sort above moves every (key, index) pair through HBM four times.  A segment here is one query's
// n_db scores (100,000 at BASELINE config C3): with data-adaptive splitters it can be cut, in one pass, into
// buckets small enough to be sorted entirely in shared memory:
//   1. ss_splitters_kernel   per query: a systematic sample of 32 x B composite keys (score key << 32 | row) is
//                            sorted in shared memory; every 32nd is a splitter.  Composite keys are unique, so ties
//                            in the scores cannot unbalance the buckets.
//   2. ss_scatter_kernel     per (query, 4096-key chunk): bucket of every key (binary search over the splitters in
//                            shared memory), pairs staged in shared memory grouped by bucket, one global atomicAdd per
//                            (chunk, bucket) reserves the run's slots in the bucket's region (2048 slots, ~2.7x the
//                            expected fill: no counting pass), runs written out coalesced
//   3. ss_offsets_kernel     per query: exclusive scan of the B bucket fills -> rank offset of every bucket
//   4. ss_bucket_sort_kernel per (query, bucket): interpolation counting sort in shared memory (fine bin = linear map
//                            of the composite between the bucket's min and max, histogram, scan, place, exact rank
//                            inside the ~0.5-element bins by comparison): O(n) for the smooth key distributions of
//                            similarity scores, O(n^2 / 256) per CTA at worst (never wrong); row indices land at their
//                            final ranks (query-major u32)
//   5. ranks_transpose64_kernel (n_q, n_db) u32 -> (n_db, n_q) int64, the reference's layout
// Traffic per pair: scores 4 + pairs 8 written + 8 read + ranks 4 + transpose 4 + 8 = 36 B, against ~100 B for the four
// radix passes.  A bucket that overflows its region (probability ~1e-9 per bucket for a random row order) raises
// *status and the caller re-runs mdir_rank_scores.
constexpr int kSsChunk = 2048;                         // keys per CTA in the scatter kernel
constexpr int kSsThreads = 256;
constexpr int kSsItems = kSsChunk / kSsThreads;        // 8
constexpr int kBucketTarget = 400;                     // average bucket size aimed for
constexpr int kBucketCap = 1024;                       // slots per bucket region = pairs the bucket sort stages
constexpr int kOversample = 32;
constexpr int kMaxBuckets = 256;
constexpr int kMaxSample = kMaxBuckets * kOversample;  // 8192
static_assert(kMaxBuckets <= kSsThreads, "one bucket per thread in the scatter scan");
constexpr int64_t kSsMaxRows = (int64_t)kMaxBuckets * kBucketTarget;      // 102,400: longer segments take the LSD path

constexpr int kTab = 1024;              // cells of the score -> bucket-range lookup table (per query)
struct SsTable {                        // written by ss_splitters_kernel, read by ss_scatter_kernel
    float hi, scale;                    // cell(score) = clamp(int((hi - score) * scale)); scale == 0: one cell (plain binary search)
    uint16_t lower[kTab + 2];           // lower[j] = number of splitters in cells < j; a key of cell j is in bucket [lower[j], lower[j + 1]]
};
// Monotone (non-decreasing) in the composite key: keys ascend as scores descend; NaN (last) -> last cell.
__device__ __forceinline__ int ss_cell(float score, float hi, float scale) {
    if (scale == 0.f) return 0;          // one cell = the full splitter range; (hi - inf) * 0 would be NaN and send +-inf to the last cell
    const float v = (hi - score) * scale;
    return (v == v) ? min(kTab - 1, max(0, (int)v)) : kTab - 1;
}

__device__ __forceinline__ uint32_t ss_load_key(const void* src, int is_key, int64_t i) {
    return is_key ? static_cast<const uint32_t*>(src)[i] : desc_key(static_cast<const float*>(src)[i]);
}
__device__ __forceinline__ uint64_t ss_composite(uint32_t key, uint32_t idx) { return ((uint64_t)key << 32) | (uint64_t)idx; }

// number of splitters <= x  (splitters ascending, n_spl = B - 1 of them)  ->  bucket in [0, B - 1].
// The search runs on the 32-bit score keys; the row half of the composite only matters on an exact key tie.
__device__ __forceinline__ int ss_bucket_of(const uint32_t* spl_key, const uint32_t* spl_idx, int lo, int hi, uint32_t key, uint32_t idx) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t sk = spl_key[mid];
        const bool le = sk < key || (sk == key && spl_idx[mid] <= idx);
        if (le) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint64_t ss_shfl_xor64(uint64_t v, int m) {
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

// Interpolation counting sort of the s unique 64-bit keys a[0, s) (shared memory).  tmp: s keys of scratch, cur: F
// counters (F a power of two >= s), red: 2 * 32 keys.  Calls emit(rank, key) exactly once per key with its rank.
// Every thread of the block must call it.
// by_score: bins linear in the SCORE between the extreme scores (the whole-segment sample: its keys span dozens of
// float octaves, and bins linear in the key bits would put a third of a Gaussian sample into a few dozen bins);
// otherwise (one bucket: a sub-octave key range, or a run of tied scores) linear in the composite itself.
// bounds != nullptr: every key lies in [bounds[0], bounds[1]] (a bucket between two splitters): no min / max pass.
template <int kRegKeys, typename Emit>
__device__ __forceinline__ void ss_interp_sort(const uint64_t* a, uint64_t* tmp, uint32_t* cur, uint64_t* red, int s, int F, bool by_score,
                                               const uint64_t* bounds, Emit emit) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint64_t mn = ~0ull, mx = 0ull;
    if (bounds == nullptr) {
        for (int i = threadIdx.x; i < s; i += blockDim.x) {
            const uint64_t x = a[i];
            mn = x < mn ? x : mn;
            mx = x > mx ? x : mx;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t a_ = ss_shfl_xor64(mn, o), b_ = ss_shfl_xor64(mx, o);
            mn = a_ < mn ? a_ : mn;
            mx = b_ > mx ? b_ : mx;
        }
        if (lane == 0) { red[w] = mn; red[32 + w] = mx; }
    }
    for (int f = threadIdx.x; f < F; f += blockDim.x) cur[f] = 0u;
    __syncthreads();
    if (bounds == nullptr) {
        mn = red[0];
        mx = red[32];
        for (int k = 1; k < nw; ++k) {
            mn = red[k] < mn ? red[k] : mn;
            mx = red[32 + k] > mx ? red[32 + k] : mx;
        }
    } else {
        mn = bounds[0];
        mx = bounds[1];
    }
    // fine bin = floor(position(x) * F / (position(mx) + 1)) with position monotone (non-decreasing) in x, so the bins
    // are ordered like the keys: position = x - mn, or score(mn) - score(x) (keys ascend as scores descend; NaN last)
    const float s_hi = key_score(mn), s_lo = key_score(mx);
    const bool use_score = by_score && (s_hi - s_lo) > 0.f && (s_hi - s_lo) < INFINITY;
    const float scale = use_score ? (float)F / ((s_hi - s_lo) * 1.0001f) : (float)F / (__ull2float_rz(mx - mn) + 1.0f);
    auto bin_of = [&](uint64_t x) {
        const float v = use_score ? (s_hi - key_score(x)) : __ull2float_rz(x - mn);
        return (v == v) ? min(F - 1, max(0, (int)(v * scale))) : F - 1;
    };
    // each thread keeps its first kRegKeys keys and their bins in registers across the phases (sized for the usual
    // list: 2 per thread for a ~400-key bucket on 256 threads, 8 for the 8,000-key sample on 1024); the rest is re-read
    uint64_t xr[kRegKeys];
    int fr[kRegKeys];
#pragma unroll
    for (int k = 0; k < kRegKeys; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        fr[k] = -1;
        if (i < s) {
            xr[k] = a[i];
            fr[k] = bin_of(xr[k]);
            atomicAdd(&cur[fr[k]], 1u);
        }
    }
    for (int i = threadIdx.x + kRegKeys * blockDim.x; i < s; i += blockDim.x) atomicAdd(&cur[bin_of(a[i])], 1u);
    __syncthreads();
    {   // exclusive scan of cur[0, F) in place: each thread owns F / blockDim consecutive counters
        const int per = (F + (int)blockDim.x - 1) / (int)blockDim.x;
        const int f0 = threadIdx.x * per;
        uint32_t sum = 0;
        for (int k = 0; k < per; ++k)
            if (f0 + k < F) sum += cur[f0 + k];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t* wsum = reinterpret_cast<uint32_t*>(red + 64);
        __syncthreads();                                   // red[0..63] fully read
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t pre = incl - sum;
        for (int k = 0; k < w; ++k) pre += wsum[k];
        for (int k = 0; k < per; ++k)
            if (f0 + k < F) {
                const uint32_t c = cur[f0 + k];
                cur[f0 + k] = pre;
                pre += c;
            }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRegKeys; ++k)
        if (fr[k] >= 0) tmp[atomicAdd(&cur[fr[k]], 1u)] = xr[k];       // afterwards cur[f] = END of bin f
    for (int i = threadIdx.x + kRegKeys * blockDim.x; i < s; i += blockDim.x) {
        const uint64_t x = a[i];
        tmp[atomicAdd(&cur[bin_of(x)], 1u)] = x;
    }
    __syncthreads();
    // exact rank = first slot of the key's bin + the number of smaller keys in that bin (~0.5 keys per bin)
    auto rank_of = [&](uint64_t x, int f) {
        const int begin = f ? (int)cur[f - 1] : 0, end = (int)cur[f];
        int r = begin;
        for (int j = begin; j < end; ++j) r += tmp[j] < x ? 1 : 0;
        return r;
    };
#pragma unroll
    for (int k = 0; k < kRegKeys; ++k)
        if (fr[k] >= 0) emit(rank_of(xr[k], fr[k]), xr[k]);
    for (int i = threadIdx.x + kRegKeys * blockDim.x; i < s; i += blockDim.x) {
        const uint64_t x = a[i];
        emit(rank_of(x, bin_of(x)), x);
    }
}

// dynamic smem: 2 * m keys + F counters + 64 keys + 32 counters
__global__ void __launch_bounds__(1024) ss_splitters_kernel(const void* __restrict__ src, int is_key, int64_t n_db, int B, int m, int F,
                                                            uint64_t* __restrict__ splitters, SsTable* __restrict__ table) {
    extern __shared__ uint64_t ss_smem[];
    uint64_t* a = ss_smem;
    uint64_t* tmp = a + m;
    uint64_t* red = tmp + m;                              // 64 keys + 32 counters (= 16 keys)
    uint32_t* cur = reinterpret_cast<uint32_t*>(red + 80);
    const int q = blockIdx.x;
    const void* seg = static_cast<const uint8_t*>(src) + (int64_t)q * n_db * 4;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const int64_t i = (int64_t)j * n_db / m;
        a[j] = ss_composite(ss_load_key(seg, is_key, i), (uint32_t)i);
    }
    __syncthreads();
    uint64_t* out = splitters + (int64_t)q * B;
    __shared__ uint64_t s_ext[2];
    __shared__ float s_tab[2];
    __shared__ uint32_t s_hist[kTab + 2];
    // splitter b = the sample's order statistic (b + 1) * m / B; ranks are unique, so each is written exactly once
    ss_interp_sort<8>(a, tmp, cur, red, m, F, true, nullptr, [&](int r, uint64_t x) {
        const int b = (int)(((int64_t)r * B + m - 1) / m);                 // the only b with (b * m) / B == r, if any (m >= B)
        if (b >= 1 && b <= B - 1 && (int)(((int64_t)b * m) / B) == r) out[b - 1] = x;
        if (r == 0) s_ext[0] = x;
        if (r == m - 1) s_ext[1] = x;
    });
    // score -> bucket-range table for the scatter kernel: cells linear in the score between the sample's extremes;
    // lower[j] = splitters in cells below j (they are smaller than every key of cell j, the map being monotone)
    for (int j = threadIdx.x; j < kTab + 2; j += blockDim.x) s_hist[j] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float hi = key_score(s_ext[0]), range = hi - key_score(s_ext[1]);
        s_tab[0] = hi;
        s_tab[1] = (range > 0.f && range < INFINITY) ? (float)kTab / (range * 1.0001f) : 0.f;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B - 1; b += blockDim.x) atomicAdd(&s_hist[ss_cell(key_score(out[b]), s_tab[0], s_tab[1]) + 1], 1u);
    __syncthreads();
    {   // inclusive scan of the kTab + 2 counters: 32 per lane of warp 0
        constexpr int kPer = (kTab + 2 + 31) / 32;
        if (threadIdx.x < 32) {
            const int j0 = threadIdx.x * kPer;
            uint32_t sum = 0;
            for (int k = 0; k < kPer; ++k)
                if (j0 + k < kTab + 2) sum += s_hist[j0 + k];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)threadIdx.x >= o) incl += u;
            }
            uint32_t run = incl - sum;
            for (int k = 0; k < kPer; ++k)
                if (j0 + k < kTab + 2) {
                    run += s_hist[j0 + k];
                    s_hist[j0 + k] = run;
                }
        }
    }
    __syncthreads();
    SsTable* t = table + q;
    if (threadIdx.x == 0) { t->hi = s_tab[0]; t->scale = s_tab[1]; }
    for (int j = threadIdx.x; j < kTab + 2; j += blockDim.x) t->lower[j] = (uint16_t)s_hist[j];
}

// dynamic smem: kSsChunk pairs + B x {splitter key, splitter row, counter, delta, fits} + kSsChunk bucket ids
__global__ void __launch_bounds__(kSsThreads, 8) ss_scatter_kernel(const void* __restrict__ src, int is_key, int64_t n_db, int B,
                                                                const uint64_t* __restrict__ splitters, const SsTable* __restrict__ table,
                                                                uint32_t* __restrict__ fill, uint64_t* __restrict__ pairs,
                                                                int32_t* __restrict__ status) {
    extern __shared__ uint64_t ss_smem[];
    uint64_t* stage = ss_smem;                                             // kSsChunk
    uint32_t* spl_key = reinterpret_cast<uint32_t*>(stage + kSsChunk);     // B
    uint32_t* spl_idx = spl_key + B;                                       // B
    uint32_t* cnt = spl_idx + B;                                           // B: counts, then chunk-local bases
    uint32_t* delta = cnt + B;                                             // B: slot in the bucket region - local base (mod 2^32)
    uint32_t* fits = delta + B;                                            // B: 0 when the run overflowed the bucket region
    uint16_t* sbucket = reinterpret_cast<uint16_t*>(fits + B);             // kSsChunk
    __shared__ uint16_t lower[kTab + 2];
    const int chunk = blockIdx.x, q = blockIdx.y;
    const int lane = threadIdx.x & 31;
    for (int b = threadIdx.x; b < B; b += kSsThreads) {
        const uint64_t sp = b < B - 1 ? splitters[(int64_t)q * B + b] : ~0ull;
        spl_key[b] = (uint32_t)(sp >> 32);
        spl_idx[b] = (uint32_t)sp;
        cnt[b] = 0u;
    }
    const SsTable* t = table + q;
    for (int j = threadIdx.x; j < kTab + 2; j += kSsThreads) lower[j] = t->lower[j];
    const float t_hi = t->hi, t_scale = t->scale;
    __syncthreads();
    const void* seg = static_cast<const uint8_t*>(src) + (int64_t)q * n_db * 4;
    const int64_t base = (int64_t)chunk * kSsChunk;
    uint32_t key[kSsItems], where[kSsItems];                               // where = bucket << 16 | rank inside (chunk, bucket)
#pragma unroll
    for (int it = 0; it < kSsItems; ++it) {
        const int64_t i = base + it * kSsThreads + threadIdx.x;
        key[it] = i < n_db ? ss_load_key(seg, is_key, i) : 0u;
    }
#pragma unroll
    for (int it = 0; it < kSsItems; ++it) {
        const int64_t i = base + it * kSsThreads + threadIdx.x;
        where[it] = 0u;
        if (i < n_db) {
            const int cell = ss_cell(key_score((uint64_t)key[it] << 32), t_hi, t_scale);
            const int b = ss_bucket_of(spl_key, spl_idx, (int)lower[cell], (int)lower[cell + 1], key[it], (uint32_t)i);
            where[it] = ((uint32_t)b << 16) | atomicAdd(&cnt[b], 1u);      // order inside a bucket is free: the bucket gets sorted
        }
    }
    __syncthreads();
    // one thread per bucket reserves this chunk's run in the bucket's region (all the global atomics in flight at
    // once); a run that does not fit is dropped and reported.  delta keeps (region slot - local base) mod 2^32.
    for (int b = threadIdx.x; b < B; b += kSsThreads) {
        const uint32_t c = cnt[b];
        const uint32_t slot = c ? atomicAdd(&fill[(int64_t)q * B + b], c) : 0u;
        const bool ok = slot + c <= (uint32_t)kBucketCap;
        if (!ok) atomicOr(status, 1);
        fits[b] = ok ? 1u : 0u;
        delta[b] = (uint32_t)b * kBucketCap + slot;
    }
    __syncthreads();
    {   // exclusive scan of the B <= kSsThreads counts, one bucket per thread
        __shared__ uint32_t wsum[kSsThreads / 32];
        const int b = threadIdx.x;
        const uint32_t c = b < B ? cnt[b] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t lbase = incl - c;
        for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) lbase += wsum[k];
        if (b < B) {
            cnt[b] = lbase;
            delta[b] -= lbase;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSsItems; ++it) {
        const int64_t i = base + it * kSsThreads + threadIdx.x;
        if (i < n_db) {
            const uint32_t b = where[it] >> 16;
            const uint32_t l = cnt[b] + (where[it] & 0xffffu);
            stage[l] = ss_composite(key[it], (uint32_t)i);
            sbucket[l] = (uint16_t)b;
        }
    }
    __syncthreads();
    const int n_valid = (int)min((int64_t)kSsChunk, n_db - base);
    uint64_t* out = pairs + (int64_t)q * B * kBucketCap;
    for (int l = threadIdx.x; l < n_valid; l += kSsThreads) {
        const int b = sbucket[l];
        if (fits[b]) out[(uint32_t)l + delta[b]] = stage[l];
    }
}

// fill (n_q, B) -> offsets (n_q, B): rank of the first element of every bucket
__global__ void __launch_bounds__(256) ss_offsets_kernel(const uint32_t* __restrict__ fill, int B, uint32_t* __restrict__ offsets) {
    __shared__ uint32_t sh[kMaxBuckets];
    const int q = blockIdx.x;
    for (int b = threadIdx.x; b < B; b += blockDim.x) sh[b] = min(fill[(int64_t)q * B + b], (uint32_t)kBucketCap);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < B; ++b) {
            const uint32_t c = sh[b];
            sh[b] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) offsets[(int64_t)q * B + b] = sh[b];
}

// direct != 0: B == 1, the pairs are built from the scores / keys themselves (short segments: no partition pass) and
// staged in shared memory; otherwise the bucket is read straight from its (L2-resident) region, three times.
// dynamic smem: kBucketCap keys (+ kBucketCap more when direct) + 2 * kBucketCap counters + 80 keys
constexpr int kSortThreads = 256;
__global__ void __launch_bounds__(kSortThreads, 8) ss_bucket_sort_kernel(const uint64_t* __restrict__ pairs, const void* __restrict__ src, int is_key, int direct,
                                                             int64_t n_db, int B, const uint32_t* __restrict__ fill,
                                                             const uint32_t* __restrict__ offsets, const uint64_t* __restrict__ splitters,
                                                             uint32_t* __restrict__ vals) {
    extern __shared__ uint64_t ss_smem[];
    uint64_t* tmp = ss_smem;
    uint64_t* red = tmp + kBucketCap;
    uint32_t* cur = reinterpret_cast<uint32_t*>(red + 80);
    const int b = blockIdx.x, q = blockIdx.y;
    int s;
    uint32_t off = 0u;
    const uint64_t* a;
    if (direct) {
        s = (int)n_db;
        uint64_t* stage = reinterpret_cast<uint64_t*>(cur + 2 * kBucketCap);
        const void* seg = static_cast<const uint8_t*>(src) + (int64_t)q * n_db * 4;
        for (int i = threadIdx.x; i < s; i += blockDim.x) stage[i] = ss_composite(ss_load_key(seg, is_key, i), (uint32_t)i);
        a = stage;
    } else {
        s = (int)min(fill[(int64_t)q * B + b], (uint32_t)kBucketCap);
        off = offsets[(int64_t)q * B + b];
        a = pairs + ((int64_t)q * B + b) * kBucketCap;
    }
    if (s <= 0) return;
    __syncthreads();
    int F = 64;
    while (F < 2 * s) F <<= 1;
    uint32_t* out = vals + (int64_t)q * n_db + off;
    // an inner bucket lies between two splitters: its key range is known without looking at the keys
    __shared__ uint64_t bnd[2];
    const bool inner = !direct && b >= 1 && b + 1 < B;
    if (inner && threadIdx.x < 2) bnd[threadIdx.x] = splitters[(int64_t)q * B + b - 1 + threadIdx.x] - threadIdx.x;
    ss_interp_sort<2>(a, tmp, cur, red, s, F, false, inner ? bnd : nullptr, [&](int r, uint64_t x) { out[r] = (uint32_t)x; });
}

struct SsPlan {
    int B, m, F_sample;
};
static SsPlan ss_plan(int64_t n_db) {
    SsPlan p;
    p.B = n_db <= kBucketCap ? 1 : (int)((n_db + kBucketTarget - 1) / kBucketTarget);      // B == 1: direct in-smem sort
    if (p.B > kMaxBuckets) p.B = kMaxBuckets;
    int64_t m = (int64_t)p.B * kOversample;
    if (m > n_db) m = n_db;
    p.m = (int)m;
    p.F_sample = 64;
    while (p.F_sample < 2 * p.m) p.F_sample <<= 1;
    return p;
}


// ============================================================================================================
// Histogram-sort path (mdir_rank_scores_hist): the same two-level idea as the sample sort, with every per-key step cut
// to a table lookup.  Similarity scores of one query are a smooth, near-Gaussian population: a 4096-cell histogram that
// is LINEAR IN THE SCORE over mean +- 4 sigma (both from a strided sample; the end cells take the tails) separates
// 100,000 rows into cells of <= ~80 rows.  Per query:
//   1. hs_plan_kernel        sample statistics -> (hi, scale); exact cell histogram of the whole row in shared memory;
//                            prefix sums; cell -> bucket table with bucket = floor(first rank of the cell / 2048), so a
//                            bucket is a run of whole cells holding < 2048 + (largest cell) rows; exact start / size /
//                            cell range of every bucket.  No sorting, no sampling error: the layout is exact.
//   2. hs_scatter_kernel     per (query, 4096-row chunk): bucket = table[cell(score)] (one LDS, no search), pairs staged
//                            in shared memory grouped by bucket, one global atomicAdd per (chunk, bucket) on a cursor
//                            that starts at the bucket's exact offset: the pair array is compact (8 B per row) and the
//                            runs are ~80 pairs long (full lines)
//   3. hs_bucket_sort_kernel one CTA per bucket (<= 4096 rows, 8 per thread in registers): interpolation counting sort
//                            with 4096 fine bins linear in the score over the bucket's cell range, exact rank inside a
//                            fine bin by comparing (score key, row) composites; ranks leave coalesced
//   4. ranks_transpose64_kernel
// Everything is monotone in the score key, so the result is bit-identical to the stable argsort.  A query whose
// histogram has a bucket of more than 4096 rows (massive ties, a spike, heavy tails) is flagged in *status (bit 1) and
// the caller re-runs the call through the sample sort, which splits ties by row.
constexpr int kHsCells = 4096;
constexpr int kHsT = 2048;                   // bucket = floor(first rank of the cell / kHsT)
constexpr int kHsCap = 4096;                 // rows one CTA sorts
constexpr int kHsF = 4096;                   // fine bins of the bucket sort
constexpr int kHsFPad = kHsF + kHsF / 8;     // one pad word per 8 counters: a thread's 8 consecutive counters are conflict-free
constexpr int kHsSortThreads = 512;
constexpr int kHsPer = kHsCap / kHsSortThreads;            // 8 rows per thread
constexpr int kHsPlanThreads = 256;          // 8 CTAs per SM: 1,024 queries are one wave
constexpr int kHsChunk = 2048;
constexpr int kHsScatterThreads = 256;
constexpr int kHsItems = kHsChunk / kHsScatterThreads;     // 8
constexpr int kHsMinRows = 1025, kHsMaxRows = 131072;
constexpr int kHsMaxBuckets = kHsMaxRows / kHsT + 2;       // 66
constexpr int kHsSample = 4096;
static_assert(kHsMaxBuckets <= 256, "bucket ids are bytes");

struct HsRange {
    float hi, scale;          // cell(score) = clamp(int((hi - score) * scale), 0, kHsCells - 1); NaN -> last cell
    uint32_t heavy;           // 1: some bucket exceeds kHsCap rows -> this query needs the sample sort
    uint32_t pad;
};
struct HsBucket {
    uint32_t start, size;     // first rank and number of rows
    uint32_t cells;           // first non-empty cell | last non-empty cell << 16
    uint32_t cursor;          // next free slot of the bucket in the pair array (scatter kernel)
};

static inline int hs_buckets(int64_t n_db) { return (int)(n_db / kHsT) + 2; }

// Position of a score on the cell axis.  Monotone non-increasing in the score (every operation is a correctly rounded,
// monotone fp32 operation), equal for +0 / -0, NaN for NaN: consistent with the order of the score keys.
__device__ __forceinline__ float hs_pos(float score, float hi, float scale) { return __fmul_rn(__fsub_rn(hi, score), scale); }
__device__ __forceinline__ int hs_cell(float score, float hi, float scale) {
    const float v = hs_pos(score, hi, scale);
    return (v == v) ? min(kHsCells - 1, max(0, (int)v)) : kHsCells - 1;
}

// scores (n_db, n_q) -> (n_q, n_db), values untouched
__global__ void __launch_bounds__(256) scores_transpose_kernel(const float* __restrict__ scores, int64_t n_db, int n_q, float* __restrict__ out) {
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
        if (q < n_q && row < n_db) out[(int64_t)q * n_db + row] = tile[tx][r];
    }
}

// dynamic smem: kHsCells counters + 4 * B words
__global__ void __launch_bounds__(kHsPlanThreads, 8) hs_plan_kernel(const float* __restrict__ src, int n_db, int B,
                                                                    HsRange* __restrict__ range, uint8_t* __restrict__ table,
                                                                    HsBucket* __restrict__ buckets, int32_t* __restrict__ status) {
    extern __shared__ uint32_t hs_smem[];
    uint32_t* hist = hs_smem;                 // kHsCells
    uint32_t* st = hist + kHsCells;           // B: first rank of the bucket
    uint32_t* sz = st + B;                    // B: rows
    uint32_t* cf = sz + B;                    // B: first non-empty cell
    uint32_t* cl = cf + B;                    // B: last non-empty cell
    __shared__ float red[3][kHsPlanThreads / 32];
    __shared__ uint32_t wsum[kHsPlanThreads / 32];
    __shared__ float s_rng[2];
    __shared__ uint32_t s_heavy;
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* seg = src + (int64_t)q * n_db;
    {   // ---- range from a strided sample: mean +- 4 sigma of its finite scores (all 8 loads of a thread in flight)
        constexpr int kS = kHsSample / kHsPlanThreads;
        const int m = min(kHsSample, n_db);
        float xs[kS];
#pragma unroll
        for (int k = 0; k < kS; ++k) {
            const int j = threadIdx.x + k * kHsPlanThreads;
            xs[k] = j < m ? seg[(int64_t)j * n_db / m] : INFINITY;
        }
        float sum = 0.f, sq = 0.f, cnt = 0.f;
#pragma unroll
        for (int k = 0; k < kS; ++k)
            if (fabsf(xs[k]) < INFINITY) { sum += xs[k]; sq = fmaf(xs[k], xs[k], sq); cnt += 1.f; }
        sum = warp_sum(sum); sq = warp_sum(sq); cnt = warp_sum(cnt);
        if (lane == 0) { red[0][w] = sum; red[1][w] = sq; red[2][w] = cnt; }
        for (int c = threadIdx.x; c < kHsCells / 4; c += kHsPlanThreads) reinterpret_cast<uint4*>(hist)[c] = make_uint4(0u, 0u, 0u, 0u);
        for (int b = threadIdx.x; b < B; b += kHsPlanThreads) { st[b] = 0xffffffffu; sz[b] = 0u; cf[b] = 0xffffu; cl[b] = 0u; }
        if (threadIdx.x == 0) s_heavy = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b2 = 0.f, n = 0.f;
            for (int k = 0; k < kHsPlanThreads / 32; ++k) { a += red[0][k]; b2 += red[1][k]; n += red[2][k]; }
            const float mean = n > 0.f ? a / n : 0.f;
            const float var = n > 0.f ? fmaxf(b2 / n - mean * mean, 0.f) : 0.f;
            const float sd = sqrtf(var);
            float hi = mean + 4.f * sd, scale = (float)kHsCells / (8.f * sd);
            if (!(sd > 0.f) || !(scale < INFINITY) || !(fabsf(hi) < INFINITY)) { hi = mean; scale = 0.f; }     // one cell: flagged below
            s_rng[0] = hi;
            s_rng[1] = scale;
        }
        __syncthreads();
    }
    const float hi = s_rng[0], scale = s_rng[1];
    // ---- exact cell histogram of the whole row (four 128-bit loads in flight per thread)
    if ((n_db & 3) == 0 && (((uintptr_t)seg) & 15) == 0) {
        const float4* p4 = reinterpret_cast<const float4*>(seg);
        const int n4 = n_db >> 2;
        auto add4 = [&](const float4& v) {
            atomicAdd(&hist[hs_cell(v.x, hi, scale)], 1u);
            atomicAdd(&hist[hs_cell(v.y, hi, scale)], 1u);
            atomicAdd(&hist[hs_cell(v.z, hi, scale)], 1u);
            atomicAdd(&hist[hs_cell(v.w, hi, scale)], 1u);
        };
        int i = threadIdx.x;
        for (; i + 3 * kHsPlanThreads < n4; i += 4 * kHsPlanThreads) {
            const float4 v0 = p4[i], v1 = p4[i + kHsPlanThreads], v2 = p4[i + 2 * kHsPlanThreads], v3 = p4[i + 3 * kHsPlanThreads];
            add4(v0);
            add4(v1);
            add4(v2);
            add4(v3);
        }
        for (; i < n4; i += kHsPlanThreads) add4(p4[i]);
    } else {
        for (int i = threadIdx.x; i < n_db; i += kHsPlanThreads) atomicAdd(&hist[hs_cell(seg[i], hi, scale)], 1u);
    }
    __syncthreads();
    // ---- exclusive prefix over the cells: thread t owns cells [16t, 16t + 16)
    constexpr int kPer = kHsCells / kHsPlanThreads;
    static_assert(kPer == 16, "16 table bytes per thread");
    uint32_t c[kPer], sum = 0;
#pragma unroll
    for (int k4 = 0; k4 < kPer / 4; ++k4) {
        const uint4 v = reinterpret_cast<const uint4*>(hist)[(kPer / 4) * threadIdx.x + k4];
        c[4 * k4] = v.x; c[4 * k4 + 1] = v.y; c[4 * k4 + 2] = v.z; c[4 * k4 + 3] = v.w;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) sum += c[k];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    uint32_t run = incl - sum;
    {
        const uint32_t mine = lane < kHsPlanThreads / 32 ? wsum[lane] : 0u;      // sum of the warps before this one
        run += warp_sum_int((int)(lane < w ? mine : 0u));
    }
    uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int cell = threadIdx.x * kPer + k;
        const uint32_t b = run / kHsT;                // < B by construction (run <= n_db)
        packed[k >> 2] |= b << (8 * (k & 3));
        if (c[k]) {
            atomicMin(&st[b], run);
            atomicAdd(&sz[b], c[k]);
            atomicMin(&cf[b], (uint32_t)cell);
            atomicMax(&cl[b], (uint32_t)cell);
        }
        run += c[k];
    }
    *reinterpret_cast<uint4*>(table + (int64_t)q * kHsCells + threadIdx.x * kPer) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    __syncthreads();
    HsBucket* bk = buckets + (int64_t)q * B;
    for (int b = threadIdx.x; b < B; b += kHsPlanThreads) {
        HsBucket o;
        o.size = sz[b];
        o.start = o.size ? st[b] : 0u;
        o.cells = o.size ? (cf[b] | (cl[b] << 16)) : 0u;
        o.cursor = o.start;
        bk[b] = o;
        if (o.size > (uint32_t)kHsCap) s_heavy = 1u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // scale == 0 (a constant row, or no finite spread): (hi - s) * 0 is NaN for an infinite s, which would send +inf to
        // the LAST cell; such a row goes to the sample sort whatever its size
        if (!(scale > 0.f)) s_heavy = 1u;
        HsRange r;
        r.hi = hi; r.scale = scale; r.heavy = s_heavy; r.pad = 0u;
        range[q] = r;
        if (s_heavy) atomicOr(status, 2);
    }
}

// static smem: 2048 pairs + the 4096-entry table: 21 KB, eight CTAs per SM.
// pairs: (row, raw fp32 score bits); the sort kernel turns the score into its key.
__global__ void __launch_bounds__(kHsScatterThreads, 8) hs_scatter_kernel(const float* __restrict__ src, int n_db, int B,
                                                                        const HsRange* __restrict__ range, const uint8_t* __restrict__ table,
                                                                        HsBucket* __restrict__ buckets, uint2* __restrict__ pairs) {
    __shared__ uint2 stage[kHsChunk];
    __shared__ __align__(16) uint8_t tab[kHsCells];
    __shared__ uint32_t cnt[kHsMaxBuckets + 2];        // counts, then chunk-local bases
    __shared__ uint32_t delta[kHsMaxBuckets + 2];      // slot in the pair array - local base (mod 2^32)
    const int chunk = blockIdx.x, q = blockIdx.y;
    const HsRange rg = range[q];
    if (rg.heavy) return;
    const int lane = threadIdx.x & 31;
    const float* seg = src + (int64_t)q * n_db;
    const int base = chunk * kHsChunk;
    float sc[kHsItems];
#pragma unroll
    for (int it = 0; it < kHsItems; ++it) {
        const int i = base + it * kHsScatterThreads + threadIdx.x;
        sc[it] = i < n_db ? seg[i] : 0.f;
    }
    {
        static_assert(kHsCells / 16 == kHsScatterThreads, "one 16-byte table load per thread");
        reinterpret_cast<uint4*>(tab)[threadIdx.x] = reinterpret_cast<const uint4*>(table + (int64_t)q * kHsCells)[threadIdx.x];
        if (threadIdx.x < kHsMaxBuckets + 2) cnt[threadIdx.x] = 0u;
    }
    __syncthreads();
    uint32_t where[kHsItems];                                                    // bucket << 16 | rank inside (chunk, bucket)
#pragma unroll
    for (int it = 0; it < kHsItems; ++it) {
        const int i = base + it * kHsScatterThreads + threadIdx.x;
        where[it] = 0xffffffffu;
        if (i < n_db) {
            const uint32_t b = tab[hs_cell(sc[it], rg.hi, rg.scale)];
            where[it] = (b << 16) | atomicAdd(&cnt[b], 1u);                      // order inside a bucket is free: the bucket gets sorted
        }
    }
    __syncthreads();
    // exclusive scan of the B <= 66 counts by warp 0 (three per lane) + one global atomicAdd per non-empty (chunk, bucket)
    if (threadIdx.x < 32) {
        uint32_t c[3], g[3], sum = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int b = 3 * lane + k;
            c[k] = b < B ? cnt[b] : 0u;
            g[k] = c[k] ? atomicAdd(&buckets[(int64_t)q * B + b].cursor, c[k]) : 0u;
            sum += c[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t lbase = incl - sum;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int b = 3 * lane + k;
            if (b < B) { cnt[b] = lbase; delta[b] = g[k] - lbase; }
            lbase += c[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kHsItems; ++it) {
        if (where[it] != 0xffffffffu) {
            const uint32_t b = where[it] >> 16;
            const uint32_t l = cnt[b] + (where[it] & 0xffffu);
            stage[l] = make_uint2((uint32_t)(base + it * kHsScatterThreads + threadIdx.x), __float_as_uint(sc[it]));
        }
    }
    __syncthreads();
    // copy-out by run: warp w takes buckets w, w + 8, ...; the staged pairs of a bucket are contiguous and go to
    // contiguous slots, so a run costs three broadcast loads and then one LDS.64 + one coalesced store per 32 pairs (a
    // per-pair bucket id / offset lookup cost a third of the kernel's shared-memory wavefronts)
    const int n_valid = min(kHsChunk, n_db - base);
    uint2* out = pairs + (int64_t)q * n_db;
    for (int b = (int)(threadIdx.x >> 5); b < B; b += kHsScatterThreads / 32) {
        const int l0 = (int)cnt[b], l1 = b + 1 < B ? (int)cnt[b + 1] : n_valid;
        const uint32_t d = delta[b];
        for (int l = l0 + lane; l < l1; l += 32) out[(uint32_t)l + d] = stage[l];
    }
}

// One CTA per bucket.  dynamic smem: kHsCap keys + kHsFPad counters
__global__ void __launch_bounds__(kHsSortThreads, 3) hs_bucket_sort_kernel(const uint2* __restrict__ pairs, int n_db, int B,
                                                                          const HsRange* __restrict__ range,
                                                                          const HsBucket* __restrict__ buckets, uint32_t* __restrict__ vals) {
    extern __shared__ uint64_t hs_smem64[];
    __shared__ uint32_t wsum[kHsSortThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x, q = blockIdx.y;
    const HsRange rg = range[q];
    if (rg.heavy) return;
    const HsBucket bk = buckets[(int64_t)q * B + b];
    const int n = (int)bk.size;
    if (n == 0) return;
    uint64_t* tmp = hs_smem64;
    uint32_t* cur = reinterpret_cast<uint32_t*>(hs_smem64 + kHsCap);
    const float c0 = (float)(bk.cells & 0xffffu);
    const float fscale = (float)kHsF / (float)((bk.cells >> 16) - (bk.cells & 0xffffu) + 1u);
    const uint2* in = pairs + (int64_t)q * n_db + bk.start;
    uint2 x[kHsPer];
#pragma unroll
    for (int k = 0; k < kHsPer; ++k) {
        const int i = threadIdx.x + kHsSortThreads * k;
        if (i < n) x[k] = in[i];
    }
    static_assert(kHsFPad % 4 == 0, "128-bit zeroing");
    for (int j = threadIdx.x; j < kHsFPad / 4; j += kHsSortThreads) reinterpret_cast<uint4*>(cur)[j] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    int fr[kHsPer];              // fine bin, then the row's position in tmp, then its rank
#pragma unroll
    for (int k = 0; k < kHsPer; ++k) {
        const int i = threadIdx.x + kHsSortThreads * k;
        fr[k] = -1;
        if (i < n) {
            const float sc = __uint_as_float(x[k].y);
            const float v = hs_pos(sc, rg.hi, rg.scale);
            // the same clamped cell position the plan used, refined: monotone non-decreasing in the key
            const float vc = (v == v) ? fminf(fmaxf(v, 0.f), (float)(kHsCells - 1)) : (float)(kHsCells - 1);
            fr[k] = min(kHsF - 1, max(0, (int)(__fmul_rn(__fsub_rn(vc, c0), fscale))));
            x[k].y = desc_key(sc);
            atomicAdd(&cur[fr[k] + (fr[k] >> 3)], 1u);
        }
    }
    __syncthreads();
    {   // exclusive scan: thread t owns fine bins [8t, 8t + 8)
        uint32_t cc[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { cc[k] = cur[9 * threadIdx.x + k]; sum += cc[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        uint32_t run = incl - sum;
        {
            const uint32_t mine = lane < kHsSortThreads / 32 ? wsum[lane] : 0u;
            run += (uint32_t)warp_sum_int((int)(lane < w ? mine : 0u));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { cur[9 * threadIdx.x + k] = run; run += cc[k]; }
    }
    __syncthreads();
    uint32_t pos[kHsPer];
#pragma unroll
    for (int k = 0; k < kHsPer; ++k)
        if (fr[k] >= 0) {
            pos[k] = atomicAdd(&cur[fr[k] + (fr[k] >> 3)], 1u);                  // afterwards cur[f] = END of fine bin f
            tmp[pos[k]] = ((uint64_t)x[k].y << 32) | x[k].x;
        }
    __syncthreads();
    // exact rank = first slot of the fine bin + the number of smaller composites in it: alone in the bin (the usual case)
    // -> its own slot; two rows -> one comparison with the other slot; more (ties) -> a loop over the bin
#pragma unroll
    for (int k = 0; k < kHsPer; ++k) {
        if (fr[k] >= 0) {
            const int f = fr[k];
            const int end = (int)cur[f + (f >> 3)];
            const int begin = f ? (int)cur[f - 1 + ((f - 1) >> 3)] : 0;
            const int c = end - begin;
            int r = (int)pos[k];
            if (c == 2) {
                const uint64_t me = ((uint64_t)x[k].y << 32) | x[k].x;
                r = begin + (tmp[2 * begin + 1 - r] < me ? 1 : 0);
            } else if (c > 2) {
                const uint64_t me = ((uint64_t)x[k].y << 32) | x[k].x;
                r = begin;
                for (int j = begin; j < end; ++j) r += tmp[j] < me ? 1 : 0;
            }
            fr[k] = r;
        }
    }
    __syncthreads();
    // rows in rank order through the (now free) counter area, then out as one coalesced run
#pragma unroll
    for (int k = 0; k < kHsPer; ++k)
        if (fr[k] >= 0) cur[fr[k]] = x[k].x;
    __syncthreads();
    uint32_t* out = vals + (int64_t)q * n_db + bk.start;
#pragma unroll
    for (int k = 0; k < kHsPer; ++k) {
        const int i = threadIdx.x + kHsSortThreads * k;
        if (i < n) out[i] = cur[i];
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
    // the always-complete path: segmented stable LSD radix sort (4 x 8 bits)
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
        radix_scan_kernel<<<n_q, 1024, 0, st>>>(counts, n_chunks, 256);
        MDIR_LAUNCH_CHECK();
        radix_scatter_kernel<<<dim3(n_chunks, n_q), kScatterThreads, kScatterSmem, st>>>(kin, vin, n_db, n_chunks, shift, counts, kout, vout);
        MDIR_LAUNCH_CHECK();
        uint32_t* t = kin; kin = kout; kout = t;
        vin = vout;
        vout = (vout == vA) ? vB : vA;
    }
    { const int rc = launch_ranks_transpose(vin, n_db, n_q, ranks, ranks_ld, st); if (rc) return rc; }
    return 0;
}


// ---- sample-sort path -------------------------------------------------------------------------------------------
extern "C" size_t mdir_rank_fast_workspace_bytes(int64_t n_db, int n_q) {
    if (n_db <= 0 || n_q <= 0) return 0;
    if (n_db > kSsMaxRows) return mdir_rank_workspace_bytes(n_db, n_q);
    const SsPlan p = ss_plan(n_db);
    const size_t n_pairs = (size_t)n_db * n_q;
    return align256((size_t)n_q * p.B * 8) + 2 * align256((size_t)n_q * p.B * 4) + align256((size_t)n_q * sizeof(SsTable)) +
           align256((size_t)n_q * p.B * kBucketCap * 8) + 2 * align256(n_pairs * 4);
}

extern "C" int mdir_rank_scores_fast(const float* scores, int64_t n_db, int n_q, int query_major, int64_t* ranks, int64_t ranks_ld,
                                     void* ws, int32_t* status, void* stream) {
    MDIR_CHECK_ARG(scores && ranks && ws && status && n_db >= 1 && n_q >= 1 && ranks_ld >= n_q);
    MDIR_CHECK_ARG(n_db < ((int64_t)1 << 32) && n_q <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    MDIR_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    if (n_db > kSsMaxRows) return mdir_rank_scores(scores, n_db, n_q, query_major, ranks, ranks_ld, ws, stream);
    const SsPlan p = ss_plan(n_db);
    const size_t n_pairs = (size_t)n_db * n_q;
    uint8_t* w = (uint8_t*)ws;
    uint64_t* splitters = (uint64_t*)w;          w += align256((size_t)n_q * p.B * 8);
    uint32_t* fill = (uint32_t*)w;               w += align256((size_t)n_q * p.B * 4);
    uint32_t* offsets = (uint32_t*)w;            w += align256((size_t)n_q * p.B * 4);
    SsTable* table = (SsTable*)w;                w += align256((size_t)n_q * sizeof(SsTable));
    uint64_t* pairs = (uint64_t*)w;              w += align256((size_t)n_q * p.B * kBucketCap * 8);
    uint32_t* vals = (uint32_t*)w;               w += align256(n_pairs * 4);
    uint32_t* keys_t = (uint32_t*)w;             // only for the (n_db, n_q) input layout
    const unsigned gx = (unsigned)((n_db + 31) / 32), gy = (unsigned)((n_q + 31) / 32);
    const void* src = scores;
    int is_key = 0;
    if (!query_major) {
        keys_transpose_kernel<<<dim3(gx, gy), 256, 0, st>>>(scores, n_db, n_q, keys_t);
        MDIR_LAUNCH_CHECK();
        src = keys_t;
        is_key = 1;
    }
    const size_t scatter_smem = (size_t)kSsChunk * 8 + (size_t)p.B * 20 + (size_t)kSsChunk * 2;
    const size_t sort_smem = (size_t)kBucketCap * 8 + 80 * 8 + (size_t)2 * kBucketCap * 4 + (p.B == 1 ? (size_t)kBucketCap * 8 : 0);
    const size_t spl_smem = (size_t)p.m * 16 + 80 * 8 + (size_t)p.F_sample * 4;
    static PerDeviceOnce once;
    if (once.first() != 0) {
        MDIR_CUDA(cudaFuncSetAttribute(ss_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)kSsChunk * 10 + (size_t)kMaxBuckets * 20)));
        MDIR_CUDA(cudaFuncSetAttribute(ss_splitters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)kMaxSample * 16 + 80 * 8 + (size_t)2 * kMaxSample * 4)));
        MDIR_CUDA(cudaFuncSetAttribute(ss_bucket_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)kBucketCap * 16 + 80 * 8 + (size_t)2 * kBucketCap * 4)));
    }
    if (p.B > 1) {
        MDIR_CUDA(cudaMemsetAsync(fill, 0, (size_t)n_q * p.B * 4, st));
        ss_splitters_kernel<<<n_q, 1024, spl_smem, st>>>(src, is_key, n_db, p.B, p.m, p.F_sample, splitters, table);
        MDIR_LAUNCH_CHECK();
        const int n_chunks = (int)((n_db + kSsChunk - 1) / kSsChunk);
        ss_scatter_kernel<<<dim3(n_chunks, n_q), kSsThreads, scatter_smem, st>>>(src, is_key, n_db, p.B, splitters, table, fill, pairs, status);
        MDIR_LAUNCH_CHECK();
        ss_offsets_kernel<<<n_q, 256, 0, st>>>(fill, p.B, offsets);
        MDIR_LAUNCH_CHECK();
    }
    ss_bucket_sort_kernel<<<dim3(p.B, n_q), kSortThreads, sort_smem, st>>>(pairs, src, is_key, p.B == 1 ? 1 : 0, n_db, p.B, fill, offsets, splitters,
                                                                                       vals);
    MDIR_LAUNCH_CHECK();
    { const int rc = launch_ranks_transpose(vals, n_db, n_q, ranks, ranks_ld, st); if (rc) return rc; }
    return 0;
}


// ---- histogram-sort path ----------------------------------------------------------------------------------------
static inline bool hs_eligible(int64_t n_db) { return n_db >= kHsMinRows && n_db <= kHsMaxRows; }

extern "C" size_t mdir_rank_hist_workspace_bytes(int64_t n_db, int n_q) {
    if (n_db <= 0 || n_q <= 0) return 0;
    if (!hs_eligible(n_db)) return mdir_rank_fast_workspace_bytes(n_db, n_q);
    const size_t n_pairs = (size_t)n_db * n_q;
    const int B = hs_buckets(n_db);
    return align256((size_t)n_q * sizeof(HsRange)) + align256((size_t)n_q * kHsCells) + align256((size_t)n_q * B * sizeof(HsBucket)) +
           align256(n_pairs * 8) + 2 * align256(n_pairs * 4);
}

extern "C" int mdir_rank_scores_hist(const float* scores, int64_t n_db, int n_q, int query_major, int64_t* ranks, int64_t ranks_ld,
                                     void* ws, int32_t* status, void* stream) {
    MDIR_CHECK_ARG(scores && ranks && ws && status && n_db >= 1 && n_q >= 1 && ranks_ld >= n_q);
    MDIR_CHECK_ARG(n_db < ((int64_t)1 << 32) && n_q <= 65535);
    if (!hs_eligible(n_db)) return mdir_rank_scores_fast(scores, n_db, n_q, query_major, ranks, ranks_ld, ws, status, stream);
    cudaStream_t st = (cudaStream_t)stream;
    MDIR_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    const size_t n_pairs = (size_t)n_db * n_q;
    const int B = hs_buckets(n_db);
    uint8_t* w = (uint8_t*)ws;
    HsRange* range = (HsRange*)w;                w += align256((size_t)n_q * sizeof(HsRange));
    uint8_t* table = w;                          w += align256((size_t)n_q * kHsCells);
    HsBucket* buckets = (HsBucket*)w;            w += align256((size_t)n_q * B * sizeof(HsBucket));
    uint2* pairs = (uint2*)w;                    w += align256(n_pairs * 8);
    uint32_t* vals = (uint32_t*)w;               w += align256(n_pairs * 4);
    uint32_t* keys_t = (uint32_t*)w;             // only for the (n_db, n_q) input layout
    const unsigned gx = (unsigned)((n_db + 31) / 32), gy = (unsigned)((n_q + 31) / 32);
    const float* src = scores;
    if (!query_major) {
        scores_transpose_kernel<<<dim3(gx, gy), 256, 0, st>>>(scores, n_db, n_q, reinterpret_cast<float*>(keys_t));
        MDIR_LAUNCH_CHECK();
        src = reinterpret_cast<const float*>(keys_t);
    }
    const size_t plan_smem = (size_t)(kHsCells + 4 * B) * 4;
    const size_t sort_smem = (size_t)kHsCap * 8 + (size_t)kHsFPad * 4;
    static PerDeviceOnce once;
    if (once.first() != 0)
        MDIR_CUDA(cudaFuncSetAttribute(hs_bucket_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
    hs_plan_kernel<<<n_q, kHsPlanThreads, plan_smem, st>>>(src, (int)n_db, B, range, table, buckets, status);
    MDIR_LAUNCH_CHECK();
    const int n_chunks = (int)((n_db + kHsChunk - 1) / kHsChunk);
    hs_scatter_kernel<<<dim3(n_chunks, n_q), kHsScatterThreads, 0, st>>>(src, (int)n_db, B, range, table, buckets, pairs);
    MDIR_LAUNCH_CHECK();
    hs_bucket_sort_kernel<<<dim3(B, n_q), kHsSortThreads, sort_smem, st>>>(pairs, (int)n_db, B, range, buckets, vals);
    MDIR_LAUNCH_CHECK();
    { const int rc = launch_ranks_transpose(vals, n_db, n_q, ranks, ranks_ld, st); if (rc) return rc; }
    return 0;
}
