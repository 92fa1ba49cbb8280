// Hard-negative selection for training tuples: the consumer of the full
// ranking inside cirtorch's TuplesDataset.create_epoch_tuples.
//   reference: mdir/external/cirtorch/datasets/traindataset.py:250-267
//     walk ranks[:, q] from the best score down; skip images whose cluster (3D model) is the query's or
//     was already used; the first nnum survivors are the negatives; their L2 distance to the query is
//     torch.pow(q - p + 1e-6, 2).sum().sqrt().
// One warp per query: 32 ranks are fetched (and their clusters gathered) at a time, then accepted in rank
// order with a ballot loop, so the choice equals the reference's sequential walk exactly.
#include "common.cuh"

namespace mdir {

constexpr int kMaxNeg = 63;      // negatives per query (the "seen clusters" list lives in shared memory)

__global__ void __launch_bounds__(128) mine_negatives_kernel(const int64_t* __restrict__ ranks, int64_t ranks_ld, int64_t n_pool,
                                                             int n_q, const int32_t* __restrict__ pool_cluster,
                                                             const int32_t* __restrict__ q_cluster, int nnum,
                                                             int64_t* __restrict__ out_pos, int32_t* __restrict__ out_found) {
    __shared__ int32_t seen_s[4][kMaxNeg + 1];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + w;
    if (q >= n_q) return;
    int32_t* seen = seen_s[w];
    if (lane == 0) seen[0] = q_cluster[q];
    __syncwarp();
    int ns = 1, found = 0;
    for (int64_t r0 = 0; r0 < n_pool && found < nnum; r0 += 32) {
        const int64_t r = r0 + lane;
        bool pending = r < n_pool;
        int64_t pos = -1;
        int32_t c = -1;
        if (pending) {
            pos = ranks[r * ranks_ld + q];
            pending = pos >= 0 && pos < n_pool;          // defensive: a corrupt rank is skipped, never dereferenced
            if (pending) c = pool_cluster[pos];
        }
        while (found < nnum) {
            bool fresh = pending;
            for (int i = 0; fresh && i < ns; ++i) fresh = seen[i] != c;
            const unsigned mask = __ballot_sync(0xffffffffu, fresh);
            if (!mask) break;
            const int leader = __ffs(mask) - 1;
            const int32_t lc = __shfl_sync(0xffffffffu, c, leader);
            const int64_t lpos = __shfl_sync(0xffffffffu, pos, leader);
            if (lane == 0) {
                out_pos[(int64_t)q * nnum + found] = lpos;
                seen[ns] = lc;
            }
            __syncwarp();
            ++ns;
            ++found;
            pending = pending && lane > leader;          // lanes up to the leader are settled for good
        }
    }
    if (lane == 0) {
        out_found[q] = found;
        for (int j = found; j < nnum; ++j) out_pos[(int64_t)q * nnum + j] = -1;
    }
}

// out[q, j] = sqrt(sum_d (qv[q, d] - pool[pos[q, j], d] + eps)^2); one warp per pair, fp32 like the reference.
__global__ void __launch_bounds__(256) pair_l2dist_kernel(const float* __restrict__ qv, const float* __restrict__ pool,
                                                          const int64_t* __restrict__ pos, int64_t n_pairs, int nnum, int D, float eps,
                                                          float* __restrict__ out) {
    const int64_t pair = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pair >= n_pairs) return;
    const int64_t p = pos[pair];
    if (p < 0) {
        if (lane == 0) out[pair] = __int_as_float(0x7fc00000);      // no negative in this slot
        return;
    }
    const float* a = qv + (pair / nnum) * D;
    const float* b = pool + p * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float t = (a[d] - b[d]) + eps;
        acc = fmaf(t, t, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[pair] = sqrtf(acc);
}

}  // namespace mdir

using namespace mdir;

extern "C" int mdir_mine_negatives(const int64_t* ranks, int64_t ranks_ld, int64_t n_pool, int n_q, const int32_t* pool_cluster,
                                   const int32_t* q_cluster, int nnum, int64_t* out_pos, int32_t* out_found, void* stream) {
    MDIR_CHECK_ARG(n_q >= 0 && n_pool >= 0 && nnum >= 0 && nnum <= kMaxNeg);
    if (n_q == 0 || nnum == 0) return 0;
    MDIR_CHECK_ARG(ranks && pool_cluster && q_cluster && out_pos && out_found && ranks_ld >= n_q);
    mine_negatives_kernel<<<(n_q + 3) / 4, 128, 0, (cudaStream_t)stream>>>(ranks, ranks_ld, n_pool, n_q, pool_cluster, q_cluster, nnum,
                                                                           out_pos, out_found);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_pair_l2dist(const float* q, const float* pool, const int64_t* pos, int n_q, int nnum, int D, float eps, float* out,
                                void* stream) {
    MDIR_CHECK_ARG(n_q >= 0 && nnum >= 0 && D >= 1);
    const int64_t n_pairs = (int64_t)n_q * nnum;
    if (n_pairs == 0) return 0;
    MDIR_CHECK_ARG(q && pool && pos && out);
    pair_l2dist_kernel<<<(unsigned)((n_pairs + 7) / 8), 256, 0, (cudaStream_t)stream>>>(q, pool, pos, n_pairs, nnum, D, eps, out);
    MDIR_LAUNCH_CHECK();
    return 0;
}
