// mAP / precision@k from the (N_db, N_q) ranks array on the device.
// Replaces compute_ap / compute_map (mdir/external/cirtorch/utils/evaluate.py:3-111): for every
// query, the rank positions of its positives ("ok") are shifted up by the number of junk images
// ranked before them, and the AP is the trapezoid sum over the precision/recall curve.
// SURVEY.md section 8f row f3 ("next"): lets the large configurations report mAP without shipping
// N_db x N_q int64 ranks to the host.
#include "common.cuh"

namespace mdir {

// membership bitmaps: bit (q * words + idx / 32, idx % 32); class 0 = ok, class 1 = junk
__global__ void __launch_bounds__(256) gnd_bitmap_kernel(const int64_t* __restrict__ items, const int32_t* __restrict__ item_query,
                                                         const int32_t* __restrict__ item_class, int64_t n_items, int64_t n_db,
                                                         int64_t words, int n_q, uint32_t* __restrict__ bitmap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const int64_t idx = items[i];
    if (idx < 0 || idx >= n_db) return;
    const int q = item_query[i], c = item_class[i];
    atomicOr(&bitmap[((int64_t)c * n_q + q) * words + (idx >> 5)], 1u << (idx & 31));
}

__device__ __forceinline__ int block_excl_scan_1024(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    *total = warp_tot[31];
    return incl - v + (w ? warp_tot[w - 1] : 0);
}

// One CTA (1024 threads) per query.  ranks: (n_db, n_q) int64 C-order (ld = n_q).
__global__ void __launch_bounds__(1024) compute_ap_kernel(const int64_t* __restrict__ ranks, int64_t n_db, int n_q,
                                                          const uint32_t* __restrict__ bitmap, int64_t words,
                                                          const int32_t* __restrict__ n_pos, const int32_t* __restrict__ kappas, int n_kappa,
                                                          double* __restrict__ aps, double* __restrict__ prs) {
    __shared__ int warp_tot[32];
    __shared__ double red_d[32];
    __shared__ int kcnt[16];
    __shared__ int s_maxpos;
    const int q = blockIdx.x;
    const int nres = n_pos[q];
    if (nres == 0) {                                   // evaluate.py:68-72: excluded from the average
        if (threadIdx.x == 0) aps[q] = __longlong_as_double(0x7ff8000000000000ll);
        for (int j = threadIdx.x; j < n_kappa; j += blockDim.x) prs[(int64_t)q * n_kappa + j] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    const uint32_t* okmap = bitmap + (int64_t)q * words;
    const uint32_t* junkmap = bitmap + ((int64_t)n_q + q) * words;
    if (threadIdx.x < 16) kcnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_maxpos = 0;
    const double recall_step = 1.0 / (double)nres;
    int base_pos = 0, base_junk = 0;                   // positives / junk seen in earlier chunks
    double ap = 0.0;                                   // accumulated by every thread identically (chunk order)
    __syncthreads();
    for (int64_t r0 = 0; r0 < n_db && base_pos < nres; r0 += 1024) {
        const int64_t r = r0 + threadIdx.x;
        int f_ok = 0, f_junk = 0;
        if (r < n_db) {
            const int64_t idx = ranks[r * n_q + q];
            if (idx >= 0 && idx < n_db) {
                f_ok = (okmap[idx >> 5] >> (idx & 31)) & 1u;
                f_junk = (junkmap[idx >> 5] >> (idx & 31)) & 1u;
            }
        }
        int tot_ok, tot_junk;
        const int i_loc = block_excl_scan_1024(f_ok, warp_tot, &tot_ok);
        const int j_loc = block_excl_scan_1024(f_junk, warp_tot, &tot_junk);
        double contrib = 0.0;
        if (f_ok) {
            const int i = base_pos + i_loc;                       // index among positives (0-based)
            const int64_t rank = r - (base_junk + j_loc);         // position after removing earlier junk
            const double p0 = rank == 0 ? 1.0 : (double)i / (double)rank;
            const double p1 = (double)(i + 1) / (double)(rank + 1);
            contrib = (p0 + p1) * recall_step / 2.0;
            const int pos1 = (int)rank + 1;                       // 1-based, for precision@k
            for (int j = 0; j < n_kappa; ++j)
                if (pos1 <= kappas[j]) atomicAdd(&kcnt[j], 1);
            atomicMax(&s_maxpos, pos1);
        }
        // deterministic chunk sum: warp shuffle tree + fixed-order sum over warps
        double s = contrib;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0) red_d[threadIdx.x >> 5] = s;
        __syncthreads();
        double chunk = 0.0;
        for (int w = 0; w < 32; ++w) chunk += red_d[w];
        ap += chunk;
        base_pos += tot_ok;
        base_junk += tot_junk;
        __syncthreads();
    }
    if (threadIdx.x == 0) aps[q] = ap;
    __syncthreads();
    for (int j = threadIdx.x; j < n_kappa; j += blockDim.x) {
        const int maxpos = s_maxpos;                              // evaluate.py:101-104
        double v = __longlong_as_double(0x7ff8000000000000ll);
        if (maxpos > 0) {
            const int kq = min(maxpos, kappas[j]);
            const int cnt = kappas[j] <= maxpos ? kcnt[j] : base_pos;
            v = (double)cnt / (double)kq;
        }
        prs[(int64_t)q * n_kappa + j] = v;
    }
}

}  // namespace mdir

using namespace mdir;

extern "C" size_t mdir_map_workspace_bytes(int64_t n_db, int n_q) {
    if (n_db <= 0 || n_q <= 0) return 0;
    const int64_t words = (n_db + 31) / 32;
    return (size_t)2 * n_q * words * 4;
}

extern "C" int mdir_compute_ap(const int64_t* ranks, int64_t n_db, int n_q, const int64_t* items, const int32_t* item_query,
                               const int32_t* item_class, int64_t n_items, const int32_t* n_pos, const int32_t* kappas, int n_kappa,
                               double* aps, double* prs, void* ws, void* stream) {
    MDIR_CHECK_ARG(ranks && n_pos && aps && ws && n_db >= 1 && n_q >= 1 && n_kappa >= 0 && n_kappa <= 16 && n_items >= 0);
    MDIR_CHECK_ARG(n_kappa == 0 || (kappas && prs));
    MDIR_CHECK_ARG(n_items == 0 || (items && item_query && item_class));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t words = (n_db + 31) / 32;
    MDIR_CUDA(cudaMemsetAsync(ws, 0, (size_t)2 * n_q * words * 4, st));
    if (n_items) {
        gnd_bitmap_kernel<<<(unsigned)((n_items + 255) / 256), 256, 0, st>>>(items, item_query, item_class, n_items, n_db, words, n_q,
                                                                             (uint32_t*)ws);
        MDIR_LAUNCH_CHECK();
    }
    compute_ap_kernel<<<n_q, 1024, 0, st>>>(ranks, n_db, n_q, (const uint32_t*)ws, words, n_pos, kappas, n_kappa, aps, prs);
    MDIR_LAUNCH_CHECK();
    return 0;
}
