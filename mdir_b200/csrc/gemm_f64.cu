// fp64 contractions for learning the whitening.
//   reference: mdir/external/cirtorch/utils/whiten.py:14-53 -- np.dot(df, df.T), np.dot(P, X - m), np.dot(Xc, Xc.T)
//   on (D, N) matrices whose columns are images: O(D^2 N) work, the rest of whitenlearn is O(D^3).
// C (M, N) = alpha * (A - a_sub) * op(B - b_sub): A (M, K) row-major; B either (N, K) row-major ("NT": both
// operands contiguous along K, the covariance shape) or (K, N) row-major ("NN": the projection shape).  a_sub / b_sub
// are optional per-ROW constants of A / of B in its own storage (the mean subtraction X - m fused into the loads).
// SIMT double-precision FMAs (the DFMA pipe is the fp64 peak on this part): 128x128 block tile, 16-deep K steps
// staged through shared memory, 8x8 accumulators per thread, optional split-K into partial planes that a
// second kernel adds in a fixed order (deterministic).
#include "common.cuh"

namespace mdir {

constexpr int kGT = 128;      // block tile (rows and columns)
constexpr int kGK = 16;       // K step

template <bool B_KMAJOR>
__global__ void __launch_bounds__(256) gemm_f64_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ a_sub,
                                                       const double* __restrict__ B, int64_t ldb, const double* __restrict__ b_sub, int M,
                                                       int N, int64_t K, int64_t k_chunk, double alpha, double* __restrict__ C, int64_t ldc,
                                                       int64_t plane_stride) {
    __shared__ __align__(16) double As[kGK][kGT];
    __shared__ __align__(16) double Bs[kGK][kGT];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * kGT, n0 = blockIdx.x * kGT;
    const int64_t k_begin = (int64_t)blockIdx.z * k_chunk;
    const int64_t k_end = min(K, k_begin + k_chunk);
    const int lr = tid & 127, lk = (tid >> 7) * 8;          // loader role: one tile row (or column), 8 consecutive k
    const int tx = tid & 15, ty = tid >> 4;                 // compute role: rows ty*4 + {0..3, 64..67}, cols tx*4 + {0..3, 64..67}

    const bool a_ok = m0 + lr < M;
    const double* a_ptr = A + (int64_t)(m0 + lr) * lda;
    const double a_off = (a_ok && a_sub) ? a_sub[m0 + lr] : 0.0;
    const bool b_ok = n0 + lr < N;
    const double b_off_k = (B_KMAJOR && b_ok && b_sub) ? b_sub[n0 + lr] : 0.0;

    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    double ra[8], rb[8];
    auto fetch = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t k = k0 + lk + i;
            const bool kin = k < k_end;
            ra[i] = (a_ok && kin) ? __ldg(a_ptr + k) - a_off : 0.0;
            if (B_KMAJOR) {
                rb[i] = (b_ok && kin) ? __ldg(B + (int64_t)(n0 + lr) * ldb + k) - b_off_k : 0.0;
            } else {
                rb[i] = (b_ok && kin) ? __ldg(B + k * ldb + (n0 + lr)) - (b_sub ? b_sub[k] : 0.0) : 0.0;
            }
        }
    };

    if (k_begin < k_end) fetch(k_begin);
    for (int64_t k0 = k_begin; k0 < k_end; k0 += kGK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            As[lk + i][lr] = ra[i];
            Bs[lk + i][lr] = rb[i];
        }
        __syncthreads();
        if (k0 + kGK < k_end) fetch(k0 + kGK);          // next step's global loads overlap this step's FMAs
#pragma unroll
        for (int k = 0; k < kGK; ++k) {
            double a[8], b[8];
            const double2 a0 = *reinterpret_cast<const double2*>(&As[k][ty * 4]);
            const double2 a1 = *reinterpret_cast<const double2*>(&As[k][ty * 4 + 2]);
            const double2 a2 = *reinterpret_cast<const double2*>(&As[k][64 + ty * 4]);
            const double2 a3 = *reinterpret_cast<const double2*>(&As[k][64 + ty * 4 + 2]);
            const double2 b0 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4]);
            const double2 b1 = *reinterpret_cast<const double2*>(&Bs[k][tx * 4 + 2]);
            const double2 b2 = *reinterpret_cast<const double2*>(&Bs[k][64 + tx * 4]);
            const double2 b3 = *reinterpret_cast<const double2*>(&Bs[k][64 + tx * 4 + 2]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y; a[6] = a3.x; a[7] = a3.y;
            b[0] = b0.x; b[1] = b0.y; b[2] = b1.x; b[3] = b1.y; b[4] = b2.x; b[5] = b2.y; b[6] = b3.x; b[7] = b3.y;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    double* out = C + (int64_t)blockIdx.z * plane_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + ty * 4 + (i & 3) + (i >> 2) * 64;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = n0 + tx * 4 + (j & 3) + (j >> 2) * 64;
            if (c < N) out[(int64_t)r * ldc + c] = alpha * acc[i][j];
        }
    }
}

// C[i] = alpha * (P_0[i] + P_1[i] + ... ) in plane order (deterministic split-K reduction); planes are dense (M*N)
__global__ void __launch_bounds__(256) sum_planes_f64_kernel(const double* __restrict__ planes, int64_t plane_stride, int n_planes, int M,
                                                             int N, double alpha, double* __restrict__ C, int64_t ldc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    double s = 0.0;
    for (int p = 0; p < n_planes; ++p) s += planes[p * plane_stride + i];
    const int64_t r = i / N;
    C[r * ldc + (i - r * N)] = alpha * s;
}

// out (D, n_pairs) row-major: out[d][j] = X[d][qidx[j]] - X[d][pidx[j]]     (whiten.py:41, df)
__global__ void __launch_bounds__(256) pair_diff_kernel(const double* __restrict__ X, int64_t ldx, int D, int64_t n_cols,
                                                        const int64_t* __restrict__ qidx, const int64_t* __restrict__ pidx, int64_t n_pairs,
                                                        double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int d = blockIdx.y;
    if (j >= n_pairs) return;
    const int64_t a = qidx[j], b = pidx[j];
    const bool ok = a >= 0 && a < n_cols && b >= 0 && b < n_cols;
    out[(int64_t)d * n_pairs + j] = ok ? X[(int64_t)d * ldx + a] - X[(int64_t)d * ldx + b] : 0.0;
}

// mean[d] = mean_j X[d][idx[j]] (idx == nullptr: all n_idx columns), sequential-order fp64 sum per block-strided lane
__global__ void __launch_bounds__(256) cols_mean_kernel(const double* __restrict__ X, int64_t ldx, const int64_t* __restrict__ idx,
                                                        int64_t n_idx, int64_t n_cols, double* __restrict__ mean) {
    __shared__ double red[8];
    const int d = blockIdx.x;
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < n_idx; j += blockDim.x) {
        const int64_t c = idx ? idx[j] : j;
        if (c >= 0 && c < n_cols) s += X[(int64_t)d * ldx + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        mean[d] = t / (double)n_idx;
    }
}

__global__ void __launch_bounds__(256) f32_to_f64_kernel(const float* __restrict__ src, int64_t n, double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (double)src[i];
}

static int pick_splits(int M, int N, int64_t K) {
    const int64_t tiles = (int64_t)((M + kGT - 1) / kGT) * ((N + kGT - 1) / kGT);
    int64_t s = (2 * kNumSMs + tiles - 1) / tiles;
    const int64_t max_by_k = (K + 16 * kGK - 1) / (16 * kGK);       // at least 256 k per split
    if (s > max_by_k) s = max_by_k;
    if (s > 64) s = 64;
    return s < 1 ? 1 : (int)s;
}

}  // namespace mdir

using namespace mdir;

extern "C" size_t mdir_gemm_f64_workspace_bytes(int M, int N, int64_t K) {
    const int s = pick_splits(M, N, K);
    return s > 1 ? (size_t)s * M * N * sizeof(double) : 0;
}

extern "C" int mdir_gemm_f64(const double* A, int64_t lda, const double* a_sub, const double* B, int64_t ldb, const double* b_sub,
                             int b_is_kxn, int M, int N, int64_t K, double alpha, double* C, int64_t ldc, void* ws, void* stream) {
    MDIR_CHECK_ARG(M >= 0 && N >= 0 && K >= 0);
    if (M == 0 || N == 0) return 0;
    MDIR_CHECK_ARG(A && B && C && lda >= K && ldc >= N && (b_is_kxn ? ldb >= N : ldb >= K));
    const int splits = pick_splits(M, N, K);
    MDIR_CHECK_ARG(splits == 1 || ws);
    const int64_t k_chunk = ((K + splits - 1) / splits + kGK - 1) / kGK * kGK;
    dim3 grid((N + kGT - 1) / kGT, (M + kGT - 1) / kGT, splits);
    double* out = splits > 1 ? static_cast<double*>(ws) : C;
    const int64_t out_ld = splits > 1 ? N : ldc;
    const int64_t plane = splits > 1 ? (int64_t)M * N : 0;
    const double a1 = splits > 1 ? 1.0 : alpha;
    if (b_is_kxn)
        gemm_f64_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, a_sub, B, ldb, b_sub, M, N, K, k_chunk, a1, out, out_ld, plane);
    else
        gemm_f64_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, a_sub, B, ldb, b_sub, M, N, K, k_chunk, a1, out, out_ld, plane);
    MDIR_LAUNCH_CHECK();
    if (splits > 1) {
        const int64_t n = (int64_t)M * N;
        sum_planes_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, plane, splits, M, N, alpha, C, ldc);
        MDIR_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int mdir_pair_diff_f64(const double* X, int64_t ldx, int D, int64_t n_cols, const int64_t* qidx, const int64_t* pidx,
                                  int64_t n_pairs, double* out, void* stream) {
    MDIR_CHECK_ARG(D >= 0 && n_pairs >= 0 && n_cols >= 0);
    if (D == 0 || n_pairs == 0) return 0;
    MDIR_CHECK_ARG(X && qidx && pidx && out && ldx >= n_cols && D <= 65535);
    dim3 grid((unsigned)((n_pairs + 255) / 256), D);
    pair_diff_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, D, n_cols, qidx, pidx, n_pairs, out);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_cols_mean_f64(const double* X, int64_t ldx, int D, int64_t n_cols, const int64_t* idx, int64_t n_idx,
                                  double* mean, void* stream) {
    MDIR_CHECK_ARG(D >= 0 && n_idx >= 1 && n_cols >= 0);
    if (D == 0) return 0;
    MDIR_CHECK_ARG(X && mean && ldx >= n_cols && (idx || n_idx <= n_cols));
    cols_mean_kernel<<<D, 256, 0, (cudaStream_t)stream>>>(X, ldx, idx, n_idx, n_cols, mean);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_f32_to_f64(const float* src, int64_t n, double* dst, void* stream) {
    MDIR_CHECK_ARG(n >= 0);
    if (n == 0) return 0;
    MDIR_CHECK_ARG(src && dst);
    f32_to_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, dst);
    MDIR_LAUNCH_CHECK();
    return 0;
}
