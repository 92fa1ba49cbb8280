// Torch-free GPU bring-up test for libmdir_b200.so (run under gpurun: tools/selftest [big]).
// Checks the C ABI against straightforward host computations and prints timings.
// Not part of the product; the parity tests proper are tests/ (-m gpu) against oracle/.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../include/mdir_b200.h"

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)
#define MD(x)                                                                          \
    do {                                                                               \
        int r_ = (x);                                                                  \
        if (r_ != 0) {                                                                 \
            printf("mdir error %d (%s) at %s:%d\n", r_, mdir_last_error(), __FILE__, __LINE__); \
            exit(3);                                                                   \
        }                                                                              \
    } while (0)

static uint32_t rng_state = 12345;
static inline uint32_t xr() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5; return rng_state; }
static inline float frand() { return (xr() >> 8) * (1.0f / 16777216.0f); }
static inline float nrand() { float u = frand() + 1e-7f, v = frand(); return sqrtf(-2.f * logf(u)) * cosf(6.2831853f * v); }

static float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

__global__ void fill_bf16(__nv_bfloat16* p, int64_t n, uint32_t seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    p[i] = __float2bfloat16_rn(((int)(h & 0xffff) - 32768) * (1.0f / 32768.0f) * 0.03f);
}

static int failures = 0;
static void report(const char* name, bool ok, double val) {
    printf("[%s] %-34s %.3e\n", ok ? " ok " : "FAIL", name, val);
    if (!ok) ++failures;
}

static void test_pool() {
    const int N = 2, C = 64, h = 23, w = 17, hw = h * w;
    std::vector<float> x((size_t)N * C * hw);
    for (auto& v : x) v = std::max(nrand(), 0.f) * 2.f;
    float *dx, *dout;
    CK(cudaMalloc(&dx, x.size() * 4));
    CK(cudaMalloc(&dout, N * C * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    for (float p : {3.0f, 2.9137f}) {
        MD(mdir_pool(MDIR_POOL_GEM, dx, nullptr, nullptr, N, C, hw, p, 1e-6f, dout, 0));
        std::vector<float> out(N * C);
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0;
        for (int pl = 0; pl < N * C; ++pl) {
            double s = 0;
            for (int i = 0; i < hw; ++i) s += pow(std::max((double)x[(size_t)pl * hw + i], 1e-6), (double)p);
            double ref = pow(s / hw, 1.0 / p);
            worst = std::max(worst, fabs(out[pl] - ref) / ref);
        }
        report(p == 3.0f ? "gem p=3 max rel err" : "gem p=2.9137 max rel err", worst < 1e-5, worst);
    }
    cudaFree(dx); cudaFree(dout);
}

struct Scan {
    int64_t n_db; int n_q, D;
    std::vector<float> db, q;      // bf16-rounded values
    __nv_bfloat16 *d_db, *d_q;
};

static Scan make_scan(int64_t n_db, int n_q, int D) {
    Scan s; s.n_db = n_db; s.n_q = n_q; s.D = D;
    s.db.resize((size_t)n_db * D); s.q.resize((size_t)n_q * D);
    std::vector<__nv_bfloat16> hdb(s.db.size()), hq(s.q.size());
    for (size_t i = 0; i < s.db.size(); ++i) { s.db[i] = bf16_round(nrand() * 0.05f); hdb[i] = __float2bfloat16_rn(s.db[i]); }
    for (size_t i = 0; i < s.q.size(); ++i) { s.q[i] = bf16_round(nrand() * 0.05f); hq[i] = __float2bfloat16_rn(s.q[i]); }
    CK(cudaMalloc(&s.d_db, hdb.size() * 2)); CK(cudaMalloc(&s.d_q, hq.size() * 2));
    CK(cudaMemcpy(s.d_db, hdb.data(), hdb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s.d_q, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice));
    return s;
}

static void test_dense(int64_t n_db, int n_q, int D) {
    Scan s = make_scan(n_db, n_q, D);
    float* d_out;
    CK(cudaMalloc(&d_out, (size_t)n_q * n_db * 4));
    CK(cudaMemset(d_out, 0xff, (size_t)n_q * n_db * 4));
    MD(mdir_sim_scan_bf16((uint16_t*)s.d_db, n_db, (uint16_t*)s.d_q, n_q, D, MDIR_SCAN_DENSE, 0, 0, d_out, n_db, nullptr, 0, nullptr,
                          nullptr, 0, 0, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)n_q * n_db);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0; int64_t bad = 0;
    for (int qi = 0; qi < n_q; ++qi)
        for (int64_t r = 0; r < n_db; ++r) {
            double ref = 0;
            for (int d = 0; d < D; ++d) ref += (double)s.db[(size_t)r * D + d] * s.q[(size_t)qi * D + d];
            double e = fabs(out[(size_t)qi * n_db + r] - ref);
            if (!(e < 1e-4)) { if (bad < 5) printf("   mismatch q=%d r=%lld got %g ref %g\n", qi, (long long)r, out[(size_t)qi * n_db + r], ref); ++bad; }
            if (e == e) worst = std::max(worst, e);
        }
    char name[96];
    snprintf(name, sizeof name, "dense scan %lldx%dx%d max abs err", (long long)n_db, n_q, D);
    report(name, bad == 0, worst);
    cudaFree(d_out); cudaFree(s.d_db); cudaFree(s.d_q);
}

// sample -> select -> filter -> finalize, compared with an exact sort of the GPU's own dense scores
static void test_topk(int64_t n_db, int n_q, int D, int k, int stride, int n_sample, int cap_s, bool fused = false) {
    const int cap_l = 96; const int64_t cap = cap_s + 148 * (int64_t)cap_l; const int NS = MDIR_CAND_SEGS;
    Scan s = make_scan(n_db, n_q, D);
    float *d_dense, *d_sample, *d_os; int32_t* d_oi; uint64_t *d_tau, *d_cand; uint32_t* d_cnt; int32_t* d_ovf;
    const int64_t n_samp_rows = (int64_t)n_sample * MDIR_SCAN_TILE_ROWS;
    CK(cudaMalloc(&d_dense, (size_t)n_q * n_db * 4));
    CK(cudaMalloc(&d_sample, (size_t)n_q * n_samp_rows * 4));
    CK(cudaMalloc(&d_os, (size_t)n_q * k * 4)); CK(cudaMalloc(&d_oi, (size_t)n_q * k * 4));
    CK(cudaMalloc(&d_tau, n_q * 8)); CK(cudaMalloc(&d_cand, (size_t)n_q * cap * 8));
    CK(cudaMalloc(&d_cnt, n_q * NS * 4)); CK(cudaMalloc(&d_ovf, n_q * 4));
    MD(mdir_sim_scan_bf16((uint16_t*)s.d_db, n_db, (uint16_t*)s.d_q, n_q, D, MDIR_SCAN_DENSE, 0, 0, d_dense, n_db, nullptr, 0, nullptr,
                          nullptr, 0, 0, 0));
    CK(cudaMemset(d_cnt, fused ? 0xff : 0, n_q * NS * 4));
    void* d_ws = nullptr;
    if (fused) {
        const size_t wsb = mdir_sim_scan_fused_workspace_bytes(n_q);
        CK(cudaMalloc(&d_ws, wsb));
        CK(cudaMemset(d_ws, 0, wsb));
        for (int rep = 0; rep < 3; ++rep)      // repeated launches: the kernel re-arms its own arrival counters
            MD(mdir_sim_scan_fused_bf16((uint16_t*)s.d_db, n_db, (uint16_t*)s.d_q, n_q, D, k, d_tau, 0, d_cand, d_cnt, cap_s, cap_l, d_ws, 0));
    } else {
    MD(mdir_sim_scan_bf16((uint16_t*)s.d_db, n_db, (uint16_t*)s.d_q, n_q, D, MDIR_SCAN_SAMPLE, stride, n_sample, d_sample, n_samp_rows,
                          nullptr, 0, nullptr, nullptr, 0, 0, 0));
    MD(mdir_select_kth(d_sample, n_samp_rows, n_samp_rows, n_q, k, stride, 0, d_tau, d_cand, cap, d_cnt, NS, cap_s, 1, 0));
    MD(mdir_sim_scan_bf16((uint16_t*)s.d_db, n_db, (uint16_t*)s.d_q, n_q, D, MDIR_SCAN_FILTER, stride, n_sample, nullptr, 0, d_tau, 0,
                          d_cand, d_cnt, cap_s, cap_l, 0));
    }
    MD(mdir_topk_finalize(d_cand, cap, d_cnt, NS, cap_s, cap_l, n_q, k, d_os, d_oi, nullptr, d_tau, d_ovf, 0));
    CK(cudaDeviceSynchronize());
    std::vector<float> dense((size_t)n_q * n_db), os((size_t)n_q * k);
    std::vector<int32_t> oi((size_t)n_q * k), ovf(n_q);
    std::vector<uint32_t> cnt_all((size_t)n_q * NS), cnt(n_q);
    CK(cudaMemcpy(dense.data(), d_dense, dense.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(os.data(), d_os, os.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(oi.data(), d_oi, oi.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ovf.data(), d_ovf, n_q * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cnt_all.data(), d_cnt, (size_t)n_q * NS * 4, cudaMemcpyDeviceToHost));
    for (int qi = 0; qi < n_q; ++qi) { cnt[qi] = 0; for (int sg = 0; sg < NS; ++sg) cnt[qi] += cnt_all[(size_t)qi * NS + sg]; }
    int64_t bad = 0; uint32_t maxcnt = 0; int novf = 0;
    for (int qi = 0; qi < n_q; ++qi) {
        maxcnt = std::max(maxcnt, cnt[qi]); novf += ovf[qi];
        std::vector<int32_t> idx(n_db);
        std::iota(idx.begin(), idx.end(), 0);
        const float* sc = &dense[(size_t)qi * n_db];
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return sc[a] > sc[b]; });
        for (int j = 0; j < k; ++j)
            if (oi[(size_t)qi * k + j] != idx[j] || os[(size_t)qi * k + j] != sc[idx[j]]) {
                if (bad < 5) printf("   topk mismatch q=%d j=%d got (%d,%g) want (%d,%g)\n", qi, j, oi[(size_t)qi * k + j], os[(size_t)qi * k + j], idx[j], sc[idx[j]]);
                ++bad;
            }
    }
    char name[128];
    snprintf(name, sizeof name, "%s %lldx%dx%d k=%d (maxcand %u, ovf %d)", fused ? "fused" : "topk", (long long)n_db, n_q, D, k, maxcnt, novf);
    if (d_ws) cudaFree(d_ws);
    report(name, bad == 0 && novf == 0, (double)bad);
    cudaFree(d_dense); cudaFree(d_sample); cudaFree(d_os); cudaFree(d_oi); cudaFree(d_tau); cudaFree(d_cand); cudaFree(d_cnt); cudaFree(d_ovf);
    cudaFree(s.d_db); cudaFree(s.d_q);
}

static void test_ranks(int64_t n_db, int n_q, bool ties) {
    std::vector<float> sc((size_t)n_db * n_q);
    for (auto& v : sc) { v = nrand(); if (ties) v = roundf(v * 4) / 4; }
    if (ties) { sc[3] = -0.0f; sc[7 * n_q] = 0.0f; }
    float* d_sc; int64_t* d_r; void* ws;
    CK(cudaMalloc(&d_sc, sc.size() * 4)); CK(cudaMalloc(&d_r, sc.size() * 8));
    CK(cudaMalloc(&ws, mdir_rank_workspace_bytes(n_db, n_q)));
    CK(cudaMemcpy(d_sc, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice));
    MD(mdir_rank_scores(d_sc, n_db, n_q, 0, d_r, n_q, ws, 0));
    CK(cudaDeviceSynchronize());
    std::vector<int64_t> r(sc.size());
    CK(cudaMemcpy(r.data(), d_r, r.size() * 8, cudaMemcpyDeviceToHost));
    int64_t bad = 0;
    for (int qi = 0; qi < n_q; ++qi) {
        std::vector<int64_t> idx(n_db);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return sc[a * n_q + qi] > sc[b * n_q + qi]; });
        for (int64_t j = 0; j < n_db; ++j)
            if (r[j * n_q + qi] != idx[j]) { if (bad < 5) printf("   rank mismatch q=%d j=%lld got %lld want %lld\n", qi, (long long)j, (long long)r[j * n_q + qi], (long long)idx[j]); ++bad; }
    }
    char name[96];
    snprintf(name, sizeof name, "rank_scores %lldx%d %s", (long long)n_db, n_q, ties ? "ties" : "random");
    report(name, bad == 0, (double)bad);
    cudaFree(d_sc); cudaFree(d_r); cudaFree(ws);
}

static void bench_big(int64_t n_db, int n_q, int n_sample) {
    const int D = 2048, k = 200, cap_s = 8192, cap_l = 96, NS = MDIR_CAND_SEGS; const int64_t cap = cap_s + 148 * (int64_t)cap_l;
    const int n_tiles = (int)((n_db + 255) / 256);
    const int stride = n_tiles / n_sample;
    printf("bench_big n_db=%lld n_q=%d n_sample=%d stride=%d\n", (long long)n_db, n_q, n_sample, stride);
    __nv_bfloat16 *d_db, *d_q;
    CK(cudaMalloc(&d_db, (size_t)n_db * D * 2)); CK(cudaMalloc(&d_q, (size_t)n_q * D * 2));
    fill_bf16<<<(unsigned)(((int64_t)n_db * D + 255) / 256), 256>>>(d_db, (int64_t)n_db * D, 1u);
    fill_bf16<<<(n_q * D + 255) / 256, 256>>>(d_q, (int64_t)n_q * D, 7u);
    CK(cudaDeviceSynchronize());
    float *d_sample, *d_os; int32_t* d_oi; uint64_t *d_tau, *d_cand; uint32_t* d_cnt; int32_t* d_ovf;
    const int64_t n_samp_rows = (int64_t)n_sample * 256;
    CK(cudaMalloc(&d_sample, (size_t)n_q * n_samp_rows * 4));
    CK(cudaMalloc(&d_os, (size_t)n_q * k * 4)); CK(cudaMalloc(&d_oi, (size_t)n_q * k * 4));
    CK(cudaMalloc(&d_tau, n_q * 8)); CK(cudaMalloc(&d_cand, (size_t)n_q * cap * 8));
    CK(cudaMalloc(&d_cnt, n_q * NS * 4)); CK(cudaMalloc(&d_ovf, n_q * 4));
    cudaEvent_t e[6];
    for (auto& ev : e) CK(cudaEventCreate(&ev));
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemsetAsync(d_cnt, 0, n_q * NS * 4));
        CK(cudaEventRecord(e[0]));
        MD(mdir_sim_scan_bf16((uint16_t*)d_db, n_db, (uint16_t*)d_q, n_q, D, MDIR_SCAN_SAMPLE, stride, n_sample, d_sample, n_samp_rows,
                              nullptr, 0, nullptr, nullptr, 0, 0, 0));
        CK(cudaEventRecord(e[1]));
        MD(mdir_select_kth(d_sample, n_samp_rows, n_samp_rows, n_q, k, stride, 0, d_tau, d_cand, cap, d_cnt, NS, cap_s, 1, 0));
        CK(cudaEventRecord(e[2]));
        MD(mdir_sim_scan_bf16((uint16_t*)d_db, n_db, (uint16_t*)d_q, n_q, D, MDIR_SCAN_FILTER, stride, n_sample, nullptr, 0, d_tau, 0,
                              d_cand, d_cnt, cap_s, cap_l, 0));
        CK(cudaEventRecord(e[3]));
        MD(mdir_topk_finalize(d_cand, cap, d_cnt, NS, cap_s, cap_l, n_q, k, d_os, d_oi, nullptr, d_tau, d_ovf, 0));
        CK(cudaEventRecord(e[4]));
        CK(cudaDeviceSynchronize());
        float t[4];
        for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&t[i], e[i], e[i + 1]));
        std::vector<uint32_t> cnt((size_t)n_q * NS);
        CK(cudaMemcpy(cnt.data(), d_cnt, (size_t)n_q * NS * 4, cudaMemcpyDeviceToHost));
        uint32_t mx = 0, mxseg = 0;
        for (int qi = 0; qi < n_q; ++qi) { uint32_t tot = 0; for (int sg = 0; sg < NS; ++sg) { tot += cnt[(size_t)qi * NS + sg]; if (sg) mxseg = std::max(mxseg, cnt[(size_t)qi * NS + sg]); } mx = std::max(mx, tot); }
        std::vector<int32_t> ovf(n_q); CK(cudaMemcpy(ovf.data(), d_ovf, n_q * 4, cudaMemcpyDeviceToHost));
        int novf = 0; for (auto o : ovf) novf += o;
        if (rep == 1) printf("   max segment fill %u of %d, overflow flags %d\n", mxseg, cap_l, novf);
        const double gb = (double)(n_db - n_samp_rows) * D * 2 / 1e9;
        if (rep) printf("rep %d: sample %.3f ms  select %.3f ms  filter %.3f ms (%.0f GB/s)  finalize %.3f ms  total %.3f ms  maxcand %u\n", rep,
               t[0], t[1], t[2], gb / (t[2] * 1e-3), t[3], t[0] + t[1] + t[2] + t[3], mx);
    }
    {   // the one-launch route
        void* d_ws; const size_t wsb = mdir_sim_scan_fused_workspace_bytes(n_q);
        CK(cudaMalloc(&d_ws, wsb)); CK(cudaMemset(d_ws, 0, wsb));
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e[0]));
            MD(mdir_sim_scan_fused_bf16((uint16_t*)d_db, n_db, (uint16_t*)d_q, n_q, D, k, d_tau, 0, d_cand, d_cnt, cap_s, cap_l, d_ws, 0));
            CK(cudaEventRecord(e[1]));
            MD(mdir_topk_finalize(d_cand, cap, d_cnt, NS, cap_s, cap_l, n_q, k, d_os, d_oi, nullptr, d_tau, d_ovf, 0));
            CK(cudaEventRecord(e[2]));
            CK(cudaDeviceSynchronize());
            float t0, t1;
            CK(cudaEventElapsedTime(&t0, e[0], e[1])); CK(cudaEventElapsedTime(&t1, e[1], e[2]));
            std::vector<uint32_t> cnt((size_t)n_q * NS);
            CK(cudaMemcpy(cnt.data(), d_cnt, (size_t)n_q * NS * 4, cudaMemcpyDeviceToHost));
            uint32_t mx = 0, mxseg = 0;
            for (int qi = 0; qi < n_q; ++qi) { uint32_t tot = 0; for (int sg = 0; sg < NS; ++sg) { tot += cnt[(size_t)qi * NS + sg]; mxseg = std::max(mxseg, cnt[(size_t)qi * NS + sg]); } mx = std::max(mx, tot); }
            std::vector<int32_t> ovf(n_q); CK(cudaMemcpy(ovf.data(), d_ovf, n_q * 4, cudaMemcpyDeviceToHost));
            int novf = 0; for (auto o : ovf) novf += o;
            if (rep) printf("fused rep %d: scan %.3f ms (%.0f GB/s)  finalize %.3f ms  total %.3f ms  maxcand %u maxseg %u ovf %d\n", rep, t0,
                            (double)n_db * D * 2 / 1e9 / (t0 * 1e-3), t1, t0 + t1, mx, mxseg, novf);
        }
        cudaFree(d_ws);
    }
    cudaFree(d_db); cudaFree(d_q); cudaFree(d_sample); cudaFree(d_os); cudaFree(d_oi); cudaFree(d_tau); cudaFree(d_cand); cudaFree(d_cnt); cudaFree(d_ovf);
}

// how the persistent scan's time depends on the number of 256-row tiles (rounds of 148 CTAs, tail rounds)
static void bench_tiles() {
    const int D = 2048, n_q = 70;
    const int64_t max_rows = 256 * 1200;
    __nv_bfloat16 *d_db, *d_q; float* d_out;
    CK(cudaMalloc(&d_db, (size_t)max_rows * D * 2)); CK(cudaMalloc(&d_q, (size_t)n_q * D * 2));
    CK(cudaMalloc(&d_out, (size_t)n_q * max_rows * 4));
    fill_bf16<<<(unsigned)((max_rows * D + 255) / 256), 256>>>(d_db, max_rows * D, 1u);
    fill_bf16<<<(n_q * D + 255) / 256, 256>>>(d_q, (int64_t)n_q * D, 7u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int tiles[] = {1, 13, 74, 148, 161, 222, 296, 309, 444, 457, 592, 1184};
    for (int nt : tiles) {
        const int64_t n_db = (int64_t)nt * 256;
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            MD(mdir_sim_scan_bf16((uint16_t*)d_db, n_db, (uint16_t*)d_q, n_q, D, MDIR_SCAN_DENSE, 0, 0, d_out, n_db, nullptr, 0, nullptr,
                                  nullptr, 0, 0, 0));
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) best = std::min(best, ms);
        }
        printf("tiles %5d (%.2f rounds): %.1f us  -> %.0f GB/s\n", nt, nt / 148.0, best * 1e3, (double)n_db * D * 2 / 1e9 / (best * 1e-3));
    }
    cudaFree(d_db); cudaFree(d_q); cudaFree(d_out);
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "tiles")) { bench_tiles(); return 0; }
    MD(mdir_device_check());
    printf("abi %d\n", mdir_abi_version());
    test_pool();
    test_dense(256, 16, 64);
    test_dense(300, 70, 136);
    test_dense(1000, 70, 2048);
    test_dense(777, 128, 512);
    test_topk(20000, 70, 256, 100, 8, 8, 8192);
    test_topk(5000, 5, 64, 10, 2, 3, 1024);
    test_topk(100000, 128, 128, 200, 12, 32, 8192);
    test_topk(20000, 70, 256, 100, 0, 0, 0, true);
    test_topk(100000, 128, 128, 200, 0, 0, 0, true);
    test_topk(513, 5, 64, 10, 0, 0, 0, true);
    test_topk(300000, 70, 64, 132, 0, 0, 0, true);
    test_ranks(5000, 7, false);
    test_ranks(4993, 70, true);
    test_ranks(100, 3, true);
    if (argc > 1 && !strcmp(argv[1], "big")) {
        bench_big(1001001, 70, 176);
        bench_big(500501, 70, 87);
        bench_big(250251, 70, 44);
        bench_big(125126, 70, 32);
        bench_big(1001001, 128, 176);
        bench_big(1001001, 16, 176);
    }
    printf("%s (%d failures)\n", failures ? "SELFTEST FAILED" : "SELFTEST PASSED", failures);
    return failures ? 1 : 0;
}
