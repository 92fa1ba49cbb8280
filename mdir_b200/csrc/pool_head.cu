// GeM / MAC / SPoC pooling, L2N, multi-scale aggregation and Lw whitening projection.
// HBM-bound: the feature maps are read exactly once with 128-bit loads, one warp per
// (map, channel) plane, warp-shuffle reduction.  See DESIGN.md "pool_planes".
#include "common.cuh"

namespace mdir {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float lg2_ftz(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// x^p for the multi-scale power mean (x = a normalised descriptor entry): MUFU.LG2 + MUFU.EX2, ~1e-6 relative;
// 0 -> 0, negative -> NaN like torch.pow with a fractional exponent.  The libm powf here was 60 % of the kernel.
__device__ __forceinline__ float fast_pow(float x, float p) {
    return ex2_ftz(p * lg2_ftz(x));
}

template <int KIND, bool P3>
__device__ __forceinline__ float pool_elem(float x, float p, float eps) {
    if (KIND == MDIR_POOL_GEM) {
        float v = fmaxf(x, eps);
        if (P3) return v * v * v;
        return ex2_ftz(p * lg2_ftz(v));        // MUFU.LG2 + MUFU.EX2; v >= eps > 0 so no denormal handling needed
    }
    return x;
}
template <int KIND>
__device__ __forceinline__ float pool_comb(float a, float b) {
    return KIND == MDIR_POOL_MAC ? fmaxf(a, b) : a + b;
}
template <int KIND, bool P3>
__device__ __forceinline__ float pool_elem4(float4 a, float p, float eps) {
    return pool_comb<KIND>(pool_comb<KIND>(pool_elem<KIND, P3>(a.x, p, eps), pool_elem<KIND, P3>(a.y, p, eps)),
                           pool_comb<KIND>(pool_elem<KIND, P3>(a.z, p, eps), pool_elem<KIND, P3>(a.w, p, eps)));
}

// One warp per UNIT of G consecutive planes of one map (G = 4 when C % 4 == 0, else 1).  The G
// planes are contiguous in memory, so the warp streams them as ONE run of G*hw floats: a scalar
// head up to the first 16-byte boundary, aligned float4 chunks (up to 8 per lane in flight = 4 KB
// per warp), a scalar tail.  A chunk almost always lies inside a single plane (its plane id costs
// three compares); the <= 3 chunks per unit that straddle a plane boundary take a per-element path.
// Per-plane fixed costs (address math, reduction, final pow) are shared by the G planes, which is
// what keeps the instruction stream below the HBM time for small planes (16x12 = 192 floats).
template <int KIND, int G>
__device__ __forceinline__ void acc_add(float (&acc)[G], int pid, float v) {
#pragma unroll
    for (int g = 0; g < G; ++g)
        if (G == 1 || pid == g) acc[g] = pool_comb<KIND>(acc[g], v);
}

template <int KIND, bool P3, int G>
__global__ void __launch_bounds__(256) pool_planes_kernel(const float* __restrict__ x, const int64_t* __restrict__ off,
                                                          const int32_t* __restrict__ hws, int n_maps, int C,
                                                          int hw_uniform, float p, float eps, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t units_per_map = (uint32_t)C / G;
    const uint32_t n_units = (uint32_t)n_maps * units_per_map;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const float ident = KIND == MDIR_POOL_MAC ? -INFINITY : 0.f;
    const float inv_p = 1.0f / p;
    for (uint32_t unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < n_units; unit += warps_total) {
        int hw;
        const float* src;
        if (off) {
            const uint32_t map = unit / units_per_map;
            const uint32_t c0 = (unit - map * units_per_map) * G;
            hw = __ldg(hws + map);
            src = x + __ldg(off + map) + (int64_t)c0 * hw;
        } else {
            hw = hw_uniform;
            src = x + (int64_t)unit * G * hw;
        }
        const int T = G * hw;                                   // floats in this unit
        const int b1 = hw, b2 = 2 * hw, b3 = 3 * hw;            // plane boundaries (element offsets)
        int head = (int)(((16u - ((uint32_t)(uintptr_t)src & 15u)) & 15u) >> 2);
        head = min(head, T);
        const int nbody = (T - head) >> 2;
        const int tail0 = head + (nbody << 2);
        const float4* body = reinterpret_cast<const float4*>(src + head);
        float acc[G];
#pragma unroll
        for (int g = 0; g < G; ++g) acc[g] = ident;
        // scalar head / tail elements (each < 4)
        {
            const int e = lane < head ? lane : tail0 + (lane - head);
            if (lane < head || e < T) {
                if (e < T) {
                    const int pid = G == 1 ? 0 : (e >= b1) + (e >= b2) + (e >= b3);
                    acc_add<KIND, G>(acc, pid, pool_elem<KIND, P3>(__ldg(src + e), p, eps));
                }
            }
        }
        for (int j0 = 0; j0 < nbody; j0 += 256) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = j0 + u * 32 + lane;
                if (i < nbody) v[u] = ldg_stream(body + i);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = j0 + u * 32 + lane;
                if (i < nbody) {
                    const int e = head + 4 * i;
                    const int pid = G == 1 ? 0 : (e >= b1) + (e >= b2) + (e >= b3);
                    const int pid3 = G == 1 ? 0 : (e + 3 >= b1) + (e + 3 >= b2) + (e + 3 >= b3);
                    if (pid == pid3) {
                        acc_add<KIND, G>(acc, pid, pool_elem4<KIND, P3>(v[u], p, eps));
                    } else {                                     // chunk straddles a plane boundary (rare)
                        const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int pe = (e + t >= b1) + (e + t >= b2) + (e + t >= b3);
                            acc_add<KIND, G>(acc, pe, pool_elem<KIND, P3>(vv[t], p, eps));
                        }
                    }
                }
            }
        }
        // reduce: G == 4 uses a butterfly that leaves plane (lane >> 3)'s total in lanes with (lane & 7) == 0
        float r;
        if (G == 4) {
            const bool h = (lane & 16) != 0;
            float k0 = h ? acc[2 % G] : acc[0], k1 = h ? acc[3 % G] : acc[1 % G];
            const float s0 = h ? acc[0] : acc[2 % G], s1 = h ? acc[1 % G] : acc[3 % G];
            k0 = pool_comb<KIND>(k0, __shfl_xor_sync(0xffffffffu, s0, 16));
            k1 = pool_comb<KIND>(k1, __shfl_xor_sync(0xffffffffu, s1, 16));
            const bool q = (lane & 8) != 0;
            r = q ? k1 : k0;
            const float snd = q ? k0 : k1;
            r = pool_comb<KIND>(r, __shfl_xor_sync(0xffffffffu, snd, 8));
            r = pool_comb<KIND>(r, __shfl_xor_sync(0xffffffffu, r, 4));
            r = pool_comb<KIND>(r, __shfl_xor_sync(0xffffffffu, r, 2));
            r = pool_comb<KIND>(r, __shfl_xor_sync(0xffffffffu, r, 1));
        } else {
            r = KIND == MDIR_POOL_MAC ? warp_max(acc[0]) : warp_sum(acc[0]);
        }
        if (KIND != MDIR_POOL_MAC) r = r / (float)hw;
        if (KIND == MDIR_POOL_GEM) r = powf(r, inv_p);           // all lanes: uniform, no divergence
        if (G == 4) {
            if ((lane & 7) == 0) out[(int64_t)unit * 4 + (lane >> 3)] = r;
        } else {
            if (lane == 0) out[unit] = r;
        }
    }
}

// x (N, C, inner): out = x / (||x||_2 over C + eps).  One CTA per (n, inner-chunk of 32).
__global__ void l2n_kernel(const float* x, int N, int C, int inner, float eps, float* out) {   // may run in place
    __shared__ float red[32];
    if (inner == 1) {
        const int n = blockIdx.x;
        const float* src = x + (int64_t)n * C;
        float s = 0.f;
        for (int c = threadIdx.x; c < C; c += blockDim.x) { float v = src[c]; s += v * v; }
        s = block_sum(s, red);
        const float inv = 1.0f / (sqrtf(s) + eps);
        for (int c = threadIdx.x; c < C; c += blockDim.x) out[(int64_t)n * C + c] = src[c] * inv;
        return;
    }
    // generic: thread per inner position, loop over C (coalesced across inner)
    const int chunks = (inner + blockDim.x - 1) / blockDim.x;
    const int n = blockIdx.x / chunks;
    const int j = (blockIdx.x % chunks) * blockDim.x + threadIdx.x;
    if (j >= inner) return;
    const float* src = x + (int64_t)n * C * inner + j;
    float s = 0.f;
    for (int c = 0; c < C; ++c) { float v = src[(int64_t)c * inner]; s += v * v; }
    const float inv = 1.0f / (sqrtf(s) + eps);
    float* dst = out + (int64_t)n * C * inner + j;
    for (int c = 0; c < C; ++c) dst[(int64_t)c * inner] = src[(int64_t)c * inner] * inv;
}

// One CTA per image.  pooled (n_img, S, C) -> out (n_img, C).
__global__ void __launch_bounds__(256) ms_aggregate_kernel(const float* __restrict__ pooled, int S, int C, float l2n_eps,
                                                           float msp, const float* __restrict__ m, float* __restrict__ out) {
    __shared__ float red[32];
    __shared__ float inv_norm[16];
    const int img = blockIdx.x;
    const float* src = pooled + (int64_t)img * S * C;
    if (l2n_eps < 0.f) {
        if (threadIdx.x < S) inv_norm[threadIdx.x] = 1.0f;
    } else {
        for (int s0 = 0; s0 < S; s0 += 4) {          // the loads of up to four scales are issued together
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (s0 + u < S) { const float v = src[(s0 + u) * C + c]; a[u] += v * v; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (s0 + u < S) {
                    const float t = block_sum(a[u], red);
                    if (threadIdx.x == 0) inv_norm[s0 + u] = 1.0f / (sqrtf(t) + l2n_eps);
                }
            }
        }
    }
    __syncthreads();
    const bool one = (msp == 1.0f);
    const float inv_msp = 1.0f / msp;
    float nn = 0.f;
    float* dst = out + (int64_t)img * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = 0.f;
        for (int s = 0; s < S; ++s) {
            float o = src[s * C + c] * inv_norm[s];
            v += one ? o : fast_pow(o, msp);
        }
        v = v / (float)S;
        if (!one) v = fast_pow(v, inv_msp);
        dst[c] = v;
        nn += v * v;
    }
    nn = block_sum(nn, red);
    const float inv = 1.0f / sqrtf(nn);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = dst[c] * inv;
        if (m) v -= m[c];
        dst[c] = v;
    }
}

// GEMV-style projection for small batches: one warp per output row j, all n (<= 8) vectors.
template <int NV>
__global__ void __launch_bounds__(256) whiten_gemv_kernel(const float* __restrict__ v, const float* __restrict__ m, int D,
                                                          const float* __restrict__ P, int dims, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= dims) return;
    const float* row = P + (int64_t)j * D;
    float acc[NV];
#pragma unroll
    for (int n = 0; n < NV; ++n) acc[n] = 0.f;
    if ((D & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        for (int i = lane; i < (D >> 2); i += 32) {
            float4 pr = ldg_stream(r4 + i);
            float4 mm = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m) mm = reinterpret_cast<const float4*>(m)[i];
#pragma unroll
            for (int n = 0; n < NV; ++n) {
                float4 vv = reinterpret_cast<const float4*>(v + (int64_t)n * D)[i];
                acc[n] += pr.x * (vv.x - mm.x) + pr.y * (vv.y - mm.y) + pr.z * (vv.z - mm.z) + pr.w * (vv.w - mm.w);
            }
        }
    } else {
        for (int i = lane; i < D; i += 32) {
            float pr = row[i];
            const float mm = m ? m[i] : 0.f;
#pragma unroll
            for (int n = 0; n < NV; ++n) acc[n] += pr * (v[(int64_t)n * D + i] - mm);
        }
    }
#pragma unroll
    for (int n = 0; n < NV; ++n) {
        float s = warp_sum(acc[n]);
        if (lane == 0) out[(int64_t)n * dims + j] = s;
    }
}

// Batched projection out(n, dims) = (V(n, D) - m) * P(dims, D)^T, fp32 CUDA-core tiles:
// BM (images) x 64 (output dims) x 16 (k) per CTA, 4x4 outputs per thread, 128-bit global and
// shared-memory accesses.  BM = 32 for small batches so the grid still covers the 148 SMs.
// P is re-read once per BM images: negligible next to the feature-map stream that produced V.
template <int BM>
__global__ void __launch_bounds__(BM * 4) whiten_sgemm_kernel(const float* __restrict__ V, const float* __restrict__ m, int n, int D,
                                                              const float* __restrict__ P, int dims, float* __restrict__ out) {
    __shared__ __align__(16) float sV[16][BM + 4];
    __shared__ __align__(16) float sP[16][64 + 4];
    constexpr int kThreads = BM * 4;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // tx -> 4 output dims, ty -> 4 images
    const int n0 = blockIdx.y * BM, j0 = blockIdx.x * 64;
    const bool vec = (D & 3) == 0;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < D; k0 += 16) {
        // tile loads: one float4 along k per (row, k-quad); BM*4 quads for V, 256 quads for P
        for (int t = threadIdx.x; t < BM * 4; t += kThreads) {
            const int r = t >> 2, kk = (t & 3) << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < n) {
                const float* src = V + (int64_t)(n0 + r) * D + k0 + kk;
                if (vec && k0 + kk + 3 < D) {
                    v = *reinterpret_cast<const float4*>(src);
                    if (m) { const float4 mm = *reinterpret_cast<const float4*>(m + k0 + kk); v.x -= mm.x; v.y -= mm.y; v.z -= mm.z; v.w -= mm.w; }
                } else {
                    float* pv = &v.x;
                    for (int e = 0; e < 4; ++e)
                        if (k0 + kk + e < D) pv[e] = src[e] - (m ? m[k0 + kk + e] : 0.f);
                }
            }
            sV[kk + 0][r] = v.x; sV[kk + 1][r] = v.y; sV[kk + 2][r] = v.z; sV[kk + 3][r] = v.w;
        }
        for (int t = threadIdx.x; t < 64 * 4; t += kThreads) {
            const int r = t >> 2, kk = (t & 3) << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + r < dims) {
                const float* src = P + (int64_t)(j0 + r) * D + k0 + kk;
                if (vec && k0 + kk + 3 < D) {
                    v = ldg_stream(reinterpret_cast<const float4*>(src));
                } else {
                    float* pv = &v.x;
                    for (int e = 0; e < 4; ++e)
                        if (k0 + kk + e < D) pv[e] = src[e];
                }
            }
            sP[kk + 0][r] = v.x; sP[kk + 1][r] = v.y; sP[kk + 2][r] = v.z; sP[kk + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&sV[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&sP[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + ty * 4 + i, jj = j0 + tx * 4 + j;
            if (nn < n && jj < dims) out[(int64_t)nn * dims + jj] = acc[i][j];
        }
}

}  // namespace mdir

using namespace mdir;

extern "C" int mdir_pool(int kind, const float* x, const int64_t* off, const int32_t* hw, int n_maps, int C,
                         int hw_uniform, float p, float eps, float* out, void* stream) {
    MDIR_CHECK_ARG(kind >= 0 && kind <= 2);
    MDIR_CHECK_ARG(x && out && n_maps >= 0 && C > 0);
    MDIR_CHECK_ARG((off == nullptr) == (hw == nullptr));
    MDIR_CHECK_ARG(off != nullptr || hw_uniform > 0);
    MDIR_CHECK_ARG(((uintptr_t)x & 3) == 0);
    MDIR_CHECK_ARG((int64_t)n_maps * C < ((int64_t)1 << 31));
    if (n_maps == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool quad = (C % 4) == 0;
    const int64_t n_units = (int64_t)n_maps * (quad ? C / 4 : C);
    const int64_t blocks_needed = (n_units + 7) / 8;
    const int grid = (int)(blocks_needed < (int64_t)kNumSMs * 32 ? blocks_needed : (int64_t)kNumSMs * 32);
#define MDIR_POOL_LAUNCH(KIND, P3)                                                                                          \
    do {                                                                                                                    \
        if (quad) pool_planes_kernel<KIND, P3, 4><<<grid, 256, 0, st>>>(x, off, hw, n_maps, C, hw_uniform, p, eps, out);   \
        else pool_planes_kernel<KIND, P3, 1><<<grid, 256, 0, st>>>(x, off, hw, n_maps, C, hw_uniform, p, eps, out);        \
    } while (0)
    if (kind == MDIR_POOL_GEM) {
        MDIR_CHECK_ARG(p > 0.f || p < 0.f);
        if (p == 3.0f) MDIR_POOL_LAUNCH(MDIR_POOL_GEM, true);
        else MDIR_POOL_LAUNCH(MDIR_POOL_GEM, false);
    } else if (kind == MDIR_POOL_MAC) {
        MDIR_POOL_LAUNCH(MDIR_POOL_MAC, false);
    } else {
        MDIR_POOL_LAUNCH(MDIR_POOL_SPOC, false);
    }
#undef MDIR_POOL_LAUNCH
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_l2n(const float* x, int N, int C, int inner, float eps, float* out, void* stream) {
    MDIR_CHECK_ARG(x && out && N >= 0 && C > 0 && inner > 0);
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (inner == 1) {
        l2n_kernel<<<N, 256, 0, st>>>(x, N, C, inner, eps, out);
    } else {
        const int chunks = (inner + 127) / 128;
        l2n_kernel<<<N * chunks, 128, 0, st>>>(x, N, C, inner, eps, out);
    }
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_ms_aggregate(const float* pooled, int n_img, int S, int C, float l2n_eps, float msp, const float* m,
                                 float* out, void* stream) {
    MDIR_CHECK_ARG(pooled && out && n_img >= 0 && S >= 1 && S <= 16 && C > 0);
    if (n_img == 0) return 0;
    ms_aggregate_kernel<<<n_img, 256, 0, (cudaStream_t)stream>>>(pooled, S, C, l2n_eps, msp, m, out);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_whiten_project(const float* v, const float* m, int n, int D, const float* P, int dims, float renorm_eps,
                                   float* out, void* stream) {
    MDIR_CHECK_ARG(v && P && out && n >= 0 && D > 0 && dims > 0);
    MDIR_CHECK_ARG(((uintptr_t)v & 15) == 0 && ((uintptr_t)P & 15) == 0 && ((uintptr_t)m & 15) == 0);
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 4) {
        const int grid = (dims + 7) / 8;
        switch (n) {
            case 1: whiten_gemv_kernel<1><<<grid, 256, 0, st>>>(v, m, D, P, dims, out); break;
            case 2: whiten_gemv_kernel<2><<<grid, 256, 0, st>>>(v, m, D, P, dims, out); break;
            case 3: whiten_gemv_kernel<3><<<grid, 256, 0, st>>>(v, m, D, P, dims, out); break;
            default: whiten_gemv_kernel<4><<<grid, 256, 0, st>>>(v, m, D, P, dims, out); break;
        }
    } else {
        const int jt = (dims + 63) / 64;
        if ((int64_t)jt * ((n + 63) / 64) >= 2 * kNumSMs) {
            whiten_sgemm_kernel<64><<<dim3(jt, (n + 63) / 64), 256, 0, st>>>(v, m, n, D, P, dims, out);
        } else {
            whiten_sgemm_kernel<32><<<dim3(jt, (n + 31) / 32), 128, 0, st>>>(v, m, n, D, P, dims, out);
        }
    }
    MDIR_LAUNCH_CHECK();
    if (renorm_eps >= 0.f) {
        l2n_kernel<<<n, 256, 0, st>>>(out, n, dims, 1, renorm_eps, out);
        MDIR_LAUNCH_CHECK();
    }
    return 0;
}
