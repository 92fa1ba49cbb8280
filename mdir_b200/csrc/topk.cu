// Selection machinery around the similarity scan: radix-select of the k-th key, candidate
// finalisation (bitonic sort in shared memory), fp32 shortlist re-scoring, fp32 -> bf16 packing.
// All orderings are on the 64-bit key (score descending, index ascending) of common.cuh, so
// results equal np.argsort(-scores, kind='stable')[:k] bit for bit.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mdir {

__device__ __forceinline__ uint32_t sample_pos_to_idx(int64_t i, int sample_stride, uint32_t idx_base) {
    if (sample_stride <= 1) return idx_base + (uint32_t)i;
    const int64_t j = i / MDIR_SCAN_TILE_ROWS, r = i - j * MDIR_SCAN_TILE_ROWS;
    return idx_base + (uint32_t)(j * sample_stride * MDIR_SCAN_TILE_ROWS + r);
}

// One CTA (1024 threads) per query: MSB-first 8-bit radix select over 64-bit keys.
__global__ void __launch_bounds__(1024) select_kth_kernel(const float* __restrict__ scores, int64_t ld, int64_t n, int kth,
                                                          int sample_stride, uint32_t idx_base, uint64_t* __restrict__ tau,
                                                          uint64_t* __restrict__ cand, uint32_t* __restrict__ cand_count, int cap) {
    __shared__ uint32_t hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_k;
    const int q = blockIdx.x;
    const float* sc = scores + (int64_t)q * ld;
    uint64_t result;
    if ((int64_t)kth > n) {
        result = ~0ull;
    } else {
        if (threadIdx.x == 0) { s_prefix = 0ull; s_k = (uint32_t)kth; }
        for (int pass = 0; pass < 8; ++pass) {
            const int shift = 56 - 8 * pass;
            if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
            __syncthreads();
            const uint64_t prefix = s_prefix;
            const uint64_t himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
            for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint64_t key = make_key(sc[i], sample_pos_to_idx(i, sample_stride, idx_base));
                if ((key & himask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 0xffu], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t k = s_k, acc = 0;
                int b = 0;
                for (; b < 256; ++b) {
                    if (acc + hist[b] >= k) break;
                    acc += hist[b];
                }
                s_k = k - acc;
                s_prefix = prefix | ((uint64_t)b << shift);
            }
            __syncthreads();
        }
        result = s_prefix;
    }
    if (threadIdx.x == 0) tau[q] = result;
    if (cand) {
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint64_t key = make_key(sc[i], sample_pos_to_idx(i, sample_stride, idx_base));
            if (key <= result) {
                const uint32_t pos = atomicAdd(&cand_count[q], 1u);
                if (pos < (uint32_t)cap) cand[(int64_t)q * cap + pos] = key;
            }
        }
    }
}

// One CTA (1024 threads) per query; dynamic smem = npow2 * 8 bytes.
__global__ void __launch_bounds__(1024) topk_finalize_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cand_count,
                                                             int cap, int k, int npow2_max, float* __restrict__ out_scores,
                                                             int32_t* __restrict__ out_idx, uint64_t* __restrict__ out_keys,
                                                             uint64_t* __restrict__ tau, int32_t* __restrict__ overflow) {
    extern __shared__ uint64_t skeys[];
    const int q = blockIdx.x;
    const uint32_t count = cand_count[q];
    const int cnt = (int)min(count, (uint32_t)cap);
    int n = 32;
    while (n < cnt) n <<= 1;
    n = min(n, npow2_max);
    const uint64_t* src = cand + (int64_t)q * cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) skeys[i] = i < cnt ? src[i] : ~0ull;
    __syncthreads();
    for (int kk = 2; kk <= n; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = skeys[i], b = skeys[ixj];
                    const bool asc = (i & kk) == 0;
                    if ((a > b) == asc) { skeys[i] = b; skeys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const uint64_t key = j < cnt ? skeys[j] : ~0ull;
        const bool ok = j < cnt && key != ~0ull;
        if (out_scores) out_scores[(int64_t)q * k + j] = ok ? key_score(key) : -INFINITY;
        if (out_idx) out_idx[(int64_t)q * k + j] = ok ? (int32_t)(uint32_t)key : -1;
        if (out_keys) out_keys[(int64_t)q * k + j] = ok ? key : ~0ull;
    }
    if (threadIdx.x == 0 && overflow) {
        const bool ovf = count > (uint32_t)cap;
        overflow[q] = ovf ? 1 : 0;
        if (ovf && tau) tau[q] = skeys[min(k, cnt) - 1];
    }
}

// One warp per (query, shortlist entry): exact fp32 dot product against the fp32 master copy.
__global__ void __launch_bounds__(256) rescore_f32_kernel(const float* __restrict__ db32, int64_t n_db, uint32_t idx_base,
                                                          const float* __restrict__ q32, int n_q, int D,
                                                          const int32_t* __restrict__ idx, int kk, uint64_t* __restrict__ out_keys) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= (int64_t)n_q * kk) return;
    const int q = (int)(item / kk);
    const int32_t gi = idx[item];
    const int64_t row = (int64_t)(uint32_t)gi - (int64_t)idx_base;
    if (gi < 0 || row < 0 || row >= n_db) {
        if (lane == 0) out_keys[item] = ~0ull;
        return;
    }
    const float* a = db32 + row * D;
    const float* b = q32 + (int64_t)q * D;
    float acc = 0.f;
    if ((D & 3) == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        for (int i = lane; i < (D >> 2); i += 32) {
            const float4 x = a4[i], y = b4[i];
            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
        }
    } else {
        for (int i = lane; i < D; i += 32) acc = fmaf(a[i], b[i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out_keys[item] = make_key(acc, (uint32_t)gi);
}

// fp32 -> bf16 (round-to-nearest-even) packing into the row-major (n, D) layout the scan wants.
// src_is_Dxn: src is the reference's (D, n) column-per-image matrix -> tiled transpose.
__global__ void __launch_bounds__(256) pack_bf16_rows_kernel(const float* __restrict__ src, int64_t total, __nv_bfloat16* __restrict__ dst) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < total) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (int64_t j = i; j < total; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
}

__global__ void __launch_bounds__(256) pack_bf16_transpose_kernel(const float* __restrict__ src, int64_t n, int D,
                                                                  __nv_bfloat16* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int64_t n0 = (int64_t)blockIdx.x * 32;
    const int d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r;
        const int64_t i = n0 + tx;
        tile[r][tx] = (d < D && i < n) ? src[(int64_t)d * n + i] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t i = n0 + r;
        const int d = d0 + tx;
        if (i < n && d < D) dst[i * D + d] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// alpha-QE / DBA accumulation (NOT in the reference).  One CTA per query:
// acc[q, :] = sum_j max(s_j, 0)^alpha * db32[idx_j - idx_base, :] over the entries this shard owns.
__global__ void __launch_bounds__(256) qe_accumulate_kernel(const float* __restrict__ db32, int64_t n_db, uint32_t idx_base, int D,
                                                            const int32_t* __restrict__ idx, const float* __restrict__ scores,
                                                            int n_qe, float alpha, float* __restrict__ acc) {
    __shared__ float w_s[256];
    __shared__ int64_t row_s[256];
    const int q = blockIdx.x;
    for (int j = threadIdx.x; j < n_qe; j += blockDim.x) {
        const int32_t gi = idx[(int64_t)q * n_qe + j];
        const int64_t row = (int64_t)(uint32_t)gi - (int64_t)idx_base;
        const bool own = gi >= 0 && row >= 0 && row < n_db;
        row_s[j] = own ? row : -1;
        w_s[j] = own ? powf(fmaxf(scores[(int64_t)q * n_qe + j], 0.f), alpha) : 0.f;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < n_qe; ++j)
            if (row_s[j] >= 0) a = fmaf(w_s[j], db32[row_s[j] * D + d], a);
        acc[(int64_t)q * D + d] = a;
    }
}

// out[n, :] = (a[n, :] + b[n, :]) / ||a[n, :] + b[n, :]||   (b may be NULL)
__global__ void __launch_bounds__(256) add_l2n_kernel(const float* a, const float* b, int D, float* out) {
    __shared__ float red[32];
    const int n = blockIdx.x;
    float s = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float v = a[(int64_t)n * D + d] + (b ? b[(int64_t)n * D + d] : 0.f);
        s += v * v;
    }
    s = block_sum(s, red);
    const float inv = 1.0f / sqrtf(s);
    for (int d = threadIdx.x; d < D; d += blockDim.x)
        out[(int64_t)n * D + d] = (a[(int64_t)n * D + d] + (b ? b[(int64_t)n * D + d] : 0.f)) * inv;
}

}  // namespace mdir

using namespace mdir;

extern "C" uint64_t mdir_make_key(float score, uint32_t index) { return make_key(score, index); }
extern "C" float mdir_key_score(uint64_t key) { return key_score(key); }

extern "C" int mdir_select_kth(const float* scores, int64_t ld, int64_t n, int n_q, int kth, int sample_stride,
                               uint32_t idx_base, uint64_t* tau, uint64_t* cand, uint32_t* cand_count, int cap, void* stream) {
    MDIR_CHECK_ARG(scores && tau && n >= 0 && n_q >= 0 && kth >= 1 && ld >= n);
    MDIR_CHECK_ARG(cand == nullptr || (cand_count != nullptr && cap >= 1));
    if (n_q == 0) return 0;
    select_kth_kernel<<<n_q, 1024, 0, (cudaStream_t)stream>>>(scores, ld, n, kth, sample_stride, idx_base, tau, cand,
                                                              cand_count, cap);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_topk_finalize(const uint64_t* cand, const uint32_t* cand_count, int cap, int n_q, int k, float* out_scores,
                                  int32_t* out_idx, uint64_t* out_keys, uint64_t* tau, int32_t* overflow, void* stream) {
    MDIR_CHECK_ARG(cand && cand_count && cap >= 1 && cap <= 16384 && n_q >= 0 && k >= 1);
    if (n_q == 0) return 0;
    int npow2 = 32;
    while (npow2 < cap) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * 8;
    static bool attr_set = false;
    if (!attr_set) {
        MDIR_CUDA(cudaFuncSetAttribute(topk_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        attr_set = true;
    }
    topk_finalize_kernel<<<n_q, 1024, smem, (cudaStream_t)stream>>>(cand, cand_count, cap, k, npow2, out_scores, out_idx,
                                                                     out_keys, tau, overflow);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_rescore_f32(const float* db32, int64_t n_db, uint32_t idx_base, const float* q32, int n_q, int D,
                                const int32_t* idx, int kk, uint64_t* out_keys, void* stream) {
    MDIR_CHECK_ARG(db32 && q32 && idx && out_keys && n_db >= 0 && n_q >= 0 && D > 0 && kk >= 1);
    MDIR_CHECK_ARG((((uintptr_t)db32 | (uintptr_t)q32) & 15) == 0);
    const int64_t items = (int64_t)n_q * kk;
    if (items == 0) return 0;
    rescore_f32_kernel<<<(unsigned)((items + 7) / 8), 256, 0, (cudaStream_t)stream>>>(db32, n_db, idx_base, q32, n_q, D, idx, kk,
                                                                                      out_keys);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_pack_bf16(const float* src, int64_t n, int D, int src_is_Dxn, uint16_t* dst, void* stream) {
    MDIR_CHECK_ARG(src && dst && n >= 0 && D > 0);
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!src_is_Dxn) {
        MDIR_CHECK_ARG((((uintptr_t)src & 15) | ((uintptr_t)dst & 7)) == 0);
        const int64_t total = n * D;
        const int64_t threads = (total + 3) / 4;
        pack_bf16_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(src, total, (__nv_bfloat16*)dst);
    } else {
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((D + 31) / 32));
        MDIR_CHECK_ARG(grid.y <= 65535);
        pack_bf16_transpose_kernel<<<grid, 256, 0, st>>>(src, n, D, (__nv_bfloat16*)dst);
    }
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_qe_accumulate(const float* db32, int64_t n_db, uint32_t idx_base, int D, const int32_t* idx, const float* scores,
                                  int n_q, int n_qe, float alpha, float* acc, void* stream) {
    MDIR_CHECK_ARG(db32 && idx && scores && acc && n_db >= 0 && D > 0 && n_q >= 0 && n_qe >= 1 && n_qe <= 256);
    if (n_q == 0) return 0;
    qe_accumulate_kernel<<<n_q, 256, 0, (cudaStream_t)stream>>>(db32, n_db, idx_base, D, idx, scores, n_qe, alpha, acc);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_add_l2n(const float* a, const float* b, int n, int D, float* out, void* stream) {
    MDIR_CHECK_ARG(a && out && n >= 0 && D > 0);
    if (n == 0) return 0;
    add_l2n_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(a, b, D, out);
    MDIR_LAUNCH_CHECK();
    return 0;
}
