// CLAHE on batches of ragged 8UC1 images, bit-exact against cv2.createCLAHE(...).apply
// (algorithm: SURVEY.md App. A; oracle/oracle.py:clahe_u8).
//
//  kernel 1  clahe_lut_kernel    one CTA per (tile, image): the tile is read as aligned 16-byte
//            chunks (4 in flight per thread), per-warp private uint32 histograms in shared memory
//            updated with unit-increment atomics (ATOMS.POPC.INC; run-length aggregation and per-lane
//            byte counters were both tried and lost), clip-limit redistribution, block scan,
//            LUT = sat_u8(rint(cdf * 255/area)).
//  kernel 2  clahe_interp_kernel one CTA per interpolation cell (the rectangle between four
//            tile centres, where the four contributing LUTs are fixed): the four LUTs are
//            interleaved into one float4[256] table in shared memory so each pixel costs a
//            single LDS.128; the bilinear blend uses individually rounded fp32 mul/add in
//            OpenCV's association (no FMA contraction) and round-half-even; a thread owns 8
//            consecutive pixels (64-bit loads/stores, 4 rows in flight).
//  The batch is processed in chunks of <= 48 MB of pixels so that pass 2 re-reads from L2.
#include <type_traits>

#include "common.cuh"

namespace mdir {

struct ClaheGeom {
    int tw, th, ext_w, ext_h;
};

__host__ __device__ __forceinline__ ClaheGeom clahe_geom(int H, int W, int tiles_x, int tiles_y) {
    ClaheGeom g;
    if (W % tiles_x == 0 && H % tiles_y == 0) {
        g.ext_w = W;
        g.ext_h = H;
    } else {
        // NB: when only one dimension is indivisible the other still gets a full extra pad
        g.ext_w = W + (tiles_x - (W % tiles_x));
        g.ext_h = H + (tiles_y - (H % tiles_y));
    }
    g.tw = g.ext_w / tiles_x;
    g.th = g.ext_h / tiles_y;
    return g;
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

// Per-image constants, computed once per call by clahe_prep_kernel: the runtime integer divisions and the
// double-precision clip limit cost ~500 instructions per warp when every CTA of the two big kernels redoes them
// (a third of all instructions the interpolation kernel executed).
struct ClahePrep {
    int tw, th, ext_w, ext_h;
    float inv_tw, inv_th, lut_scale;
    int clip_limit;
};

__global__ void __launch_bounds__(256) clahe_prep_kernel(const mdir_image_desc* __restrict__ descs, int n_img, double clip, int tiles_x,
                                                         int tiles_y, ClahePrep* __restrict__ prep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_img) return;
    const ClaheGeom g = clahe_geom(descs[i].H, descs[i].W, tiles_x, tiles_y);
    ClahePrep p;
    p.tw = g.tw; p.th = g.th; p.ext_w = g.ext_w; p.ext_h = g.ext_h;
    p.inv_tw = __fdiv_rn(1.0f, (float)g.tw);
    p.inv_th = __fdiv_rn(1.0f, (float)g.th);
    const int area = g.tw * g.th;
    p.lut_scale = __fdiv_rn(255.0f, (float)area);
    p.clip_limit = 0;
    if (clip > 0.0) {
        p.clip_limit = (int)(clip * (double)area / 256.0);
        p.clip_limit = max(p.clip_limit, 1);
    }
    prep[i] = p;
}

// histogram update of the bytes [jlo, jhi) of a 16-byte chunk (row ends only)
__device__ __forceinline__ void hist_chunk_partial(uint32_t* h, const uint32_t (&w)[4], int jlo, int jhi) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (j >= jlo && j < jhi) atomicAdd(&h[(w[j >> 2] >> (8 * (j & 3))) & 0xffu], 1u);
}

// all 16 bytes of a chunk: one unit-increment shared-memory atomic per pixel.  The unit increment matters: it
// compiles to ATOMS.POPC.INC, which folds lanes that hit the same bin into one update, so dark images (a quarter
// of a warp in bin 0) do not serialise the way a variable-increment ATOMS.ADD does.
__device__ __forceinline__ void hist_chunk_full(uint32_t* h, const uint32_t (&w)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        atomicAdd(&h[w[k] & 0xffu], 1u);
        atomicAdd(&h[(w[k] >> 8) & 0xffu], 1u);
        atomicAdd(&h[(w[k] >> 16) & 0xffu], 1u);
        atomicAdd(&h[w[k] >> 24], 1u);
    }
}

__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ src, const mdir_image_desc* __restrict__ descs,
                                                        const ClahePrep* __restrict__ prep, uint8_t* __restrict__ luts) {
    __shared__ uint32_t whist[8][256];
    __shared__ int red_i[8];
    __shared__ int scan_w[8];
    const int img = blockIdx.z;
    const int ty = blockIdx.y, tx = blockIdx.x;
    const int tiles_x = gridDim.x, tiles_y = gridDim.y;
    const mdir_image_desc d = descs[img];
    const ClahePrep g = prep[img];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&whist[0][0])[i] = 0u;
    __syncthreads();

    const uint8_t* base = src + d.src_off;
    const int x0 = tx * g.tw, y0 = ty * g.th;
    uint32_t* myh = whist[w];
    // columns [x0, x0 + vw) of the tile lie inside the image; the rest (right border tiles of
    // images whose width is not a multiple of tiles_x) are BORDER_REFLECT_101 copies
    const int vw = max(0, min(g.tw, d.W - x0));
    // interior tile whose rows are whole aligned 16-byte chunks (every tile of e.g. 768 x 1024 with 8 x 8 tiles): no
    // reflection, no partial chunks, one multiply-add per address
    const bool fast = vw == g.tw && y0 + g.th <= d.H && (g.tw & 15) == 0 && ((d.src_pitch | (int64_t)(uintptr_t)(base + x0)) & 15) == 0;
    if (fast) {
        const int cpr = g.tw >> 4;
        const int total = g.th * cpr;
        const uint8_t* tb = base + (int64_t)y0 * d.src_pitch + x0;
        const int step_r = 256 / cpr, step_c = 256 - step_r * cpr;
        int r = (int)threadIdx.x / cpr, c = (int)threadIdx.x - r * cpr;
        for (int c0 = threadIdx.x; c0 < total; c0 += 4 * 256) {
            uint32_t wv[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (c0 + u * 256 < total) {
                    const uint4 q = *reinterpret_cast<const uint4*>(tb + (int64_t)r * d.src_pitch + 16 * c);
                    wv[u][0] = q.x; wv[u][1] = q.y; wv[u][2] = q.z; wv[u][3] = q.w;
                }
                r += step_r;
                c += step_c;
                if (c >= cpr) { c -= cpr; ++r; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c0 + u * 256 < total) hist_chunk_full(myh, wv[u]);
        }
    } else if (vw > 0) {
        // 16-byte chunks: cpr per tile row (rows start at arbitrary alignment), 4 loads in flight per thread
        const int cpr = (vw + 15) / 16 + 1;
        const int total = g.th * cpr;
        // chunk ci = (row r, chunk-in-row c); consecutive chunks of a thread are 256 apart: advance (r, c) instead of dividing
        const int step_r = 256 / cpr, step_c = 256 - step_r * cpr;
        int r = (int)threadIdx.x / cpr, c = (int)threadIdx.x - r * cpr;
        for (int c0 = threadIdx.x; c0 < total; c0 += 4 * 256) {
            uint32_t wv[4][4];
            int jlo[4], jhi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ci = c0 + u * 256;
                jlo[u] = 0; jhi[u] = 0;
                if (ci < total) {
                    const uint8_t* rowp = base + (int64_t)reflect101(y0 + r, d.H) * d.src_pitch;
                    const uint8_t* seg_lo = rowp + x0;
                    const uint8_t* seg_hi = seg_lo + vw;
                    const uint8_t* chunk = (const uint8_t*)((uintptr_t)seg_lo & ~(uintptr_t)15) + 16 * c;
                    const uint8_t* lo = chunk > seg_lo ? chunk : seg_lo;
                    const uint8_t* hi = chunk + 16 < seg_hi ? chunk + 16 : seg_hi;
                    if (lo < hi) {
                        jlo[u] = (int)(lo - chunk);
                        jhi[u] = (int)(hi - chunk);
                        if (chunk >= rowp && chunk + 16 <= rowp + d.W) {          // whole chunk inside this image row
                            const uint4 q = *reinterpret_cast<const uint4*>(chunk);
                            wv[u][0] = q.x; wv[u][1] = q.y; wv[u][2] = q.z; wv[u][3] = q.w;
                        } else {
                            wv[u][0] = wv[u][1] = wv[u][2] = wv[u][3] = 0u;
#pragma unroll
                            for (int j = 0; j < 16; ++j)          // compile-time indices keep wv in registers
                                if (j >= jlo[u] && j < jhi[u]) wv[u][j >> 2] |= (uint32_t)chunk[j] << (8 * (j & 3));
                        }
                    }
                }
                r += step_r;
                c += step_c;
                if (c >= cpr) { c -= cpr; ++r; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (jlo[u] == 0 && jhi[u] == 16) hist_chunk_full(myh, wv[u]);
                else if (jlo[u] < jhi[u]) hist_chunk_partial(myh, wv[u], jlo[u], jhi[u]);
            }
        }
    }
    const int rw = g.tw - vw;
    if (rw > 0) {
        const int total = g.th * rw;
        for (int i = threadIdx.x; i < total; i += 256) {
            const int r = i / rw, c = vw + (i - r * rw);
            const uint8_t* rowp = base + (int64_t)reflect101(y0 + r, d.H) * d.src_pitch;
            atomicAdd(&myh[rowp[reflect101(x0 + c, d.W)]], 1u);
        }
    }
    __syncthreads();

    const int i = threadIdx.x;
    int h = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) h += (int)whist[k][i];

    const int clip_limit = g.clip_limit;
    if (clip_limit > 0) {
        int excess = max(h - clip_limit, 0);
        h = min(h, clip_limit);
        int s = warp_sum_int(excess);
        if (lane == 0) red_i[w] = s;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) clipped += red_i[k];
        const int redist = clipped / 256;
        int residual = clipped - redist * 256;
        h += redist;
        if (residual != 0) {
            const int step = max(256 / residual, 1);
            if ((i % step) == 0 && (i / step) < residual) h += 1;
        }
    }
    // inclusive scan over the 256 bins
    int s = h;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) scan_w[w] = s;
    __syncthreads();
    int pre = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < w) pre += scan_w[k];
    s += pre;
    int q = __float2int_rn(__fmul_rn((float)s, g.lut_scale));
    q = min(max(q, 0), 255);
    luts[(((int64_t)img * tiles_y + ty) * tiles_x + tx) * 256 + i] = (uint8_t)q;
}

__global__ void __launch_bounds__(256) clahe_interp_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           const mdir_image_desc* __restrict__ descs, const ClahePrep* __restrict__ prep,
                                                           const uint8_t* __restrict__ luts) {
    __shared__ float4 lutf[256];          // the four contributing LUTs, pre-converted: one LDS.128 per pixel
    const int img = blockIdx.z;
    const int cy = blockIdx.y, cx = blockIdx.x;
    const int tiles_x = (int)gridDim.x - 1, tiles_y = (int)gridDim.y - 1;
    const mdir_image_desc d = descs[img];
    const ClahePrep g = prep[img];
    // nominal pixel ranges of this cell: raw tile index floor(x/tw - 0.5) == cx - 1.  The exact fp32
    // boundary can differ from the nominal one by a pixel, hence the margin + the per-pixel ownership test.
    const int margin = 1 + (max(g.tw, g.th) >> 9);
    int xs = (cx == 0) ? 0 : (((2 * cx - 1) * g.tw + 1) >> 1) - margin;
    int xe = (((2 * cx + 1) * g.tw + 1) >> 1) + margin;
    int ys = (cy == 0) ? 0 : (((2 * cy - 1) * g.th + 1) >> 1) - margin;
    int ye = (((2 * cy + 1) * g.th + 1) >> 1) + margin;
    xs = max(xs, 0); ys = max(ys, 0);
    xe = min(xe, d.W); ye = min(ye, d.H);
    if (xs >= xe || ys >= ye) return;

    {
        const int ty1 = max(cy - 1, 0), ty2 = min(cy, tiles_y - 1);
        const int tx1 = max(cx - 1, 0), tx2 = min(cx, tiles_x - 1);
        const uint8_t* L = luts + (int64_t)img * tiles_y * tiles_x * 256;
        const int v = threadIdx.x;
        lutf[v] = make_float4((float)L[(ty1 * tiles_x + tx1) * 256 + v], (float)L[(ty1 * tiles_x + tx2) * 256 + v],
                              (float)L[(ty2 * tiles_x + tx1) * 256 + v], (float)L[(ty2 * tiles_x + tx2) * 256 + v]);
    }
    __syncthreads();

    const float inv_tw = g.inv_tw;
    const float inv_th = g.inv_th;
    // shrink the nominal rectangle to the pixels this cell really owns (ownership is monotone in x and in y), so
    // that the thread mapping below wastes no lanes on the safety margin
    {
        auto own = [](int v, float inv, int c) { return (int)floorf(__fsub_rn(__fmul_rn((float)v, inv), 0.5f)) == c - 1; };
        while (xs < xe && !own(xs, inv_tw, cx)) ++xs;
        while (xe > xs && !own(xe - 1, inv_tw, cx)) --xe;
        while (ys < ye && !own(ys, inv_th, cy)) ++ys;
        while (ye > ys && !own(ye - 1, inv_th, cy)) --ye;
        if (xs >= xe || ys >= ye) return;
    }
    const uint8_t* sbase = src + d.src_off;
    uint8_t* dbase = dst + d.dst_off;
    const bool vec_ok = (((uintptr_t)sbase | (uintptr_t)dbase | (uintptr_t)d.src_pitch | (uintptr_t)d.dst_pitch) & 7) == 0;

    // thread -> (group of 8 consecutive pixels, row slot); a thread keeps its x-group for all rows so the
    // 8 column weights live in registers.  Groups are aligned to 8 in image coordinates.
    const int xs8 = xs & ~7;
    const int n_groups = (xe - xs8 + 7) >> 3;
    const int gpp = min(n_groups, 256);                 // groups per pass
    const int rows_pp = 256 / gpp;                      // rows per pass
    const int gslot = threadIdx.x % gpp, rslot = threadIdx.x / gpp;
    if (rslot >= rows_pp) return;
    // 4 or 6 rows in flight per thread, whichever leaves fewer idle row slots in the last sweep over the cell
    const int n_rows = ye - ys;
    const bool six_rows = ((n_rows + 6 * rows_pp - 1) / (6 * rows_pp)) * 6 <= ((n_rows + 4 * rows_pp - 1) / (4 * rows_pp)) * 4;
    for (int g0 = gslot; g0 < n_groups; g0 += gpp) {
        const int x8 = xs8 + g0 * 8;
        float xa[8], xa1[8];
        uint32_t mine = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int x = x8 + e;
            const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
            const float fl = floorf(txf);
            xa[e] = __fsub_rn(txf, fl);
            xa1[e] = __fsub_rn(1.0f, xa[e]);
            if (x >= xs && x < xe && (int)fl == cx - 1) mine |= 1u << e;
        }
        if (!mine) continue;
        const bool in_row = x8 + 7 < d.W;
        const bool full = (mine == 0xffu) && in_row;
        auto rows = [&](auto U_) {
        constexpr int U = decltype(U_)::value;
        for (int yb = ys + rslot; yb < ye; yb += U * rows_pp) {
            uint2 pix[U];
            // U row loads in flight
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = yb + u * rows_pp;
                pix[u] = make_uint2(0u, 0u);
                if (y < ye) {
                    const uint8_t* srow = sbase + (int64_t)y * d.src_pitch;
                    if (vec_ok && in_row) {
                        pix[u] = *reinterpret_cast<const uint2*>(srow + x8);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            if (x8 + e < d.W) {
                                const uint32_t b = (uint32_t)srow[x8 + e] << (8 * (e & 3));
                                if (e < 4) pix[u].x |= b; else pix[u].y |= b;
                            }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = yb + u * rows_pp;
                if (y >= ye) continue;
                const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
                const float fly = floorf(tyf);
                if ((int)fly != cy - 1) continue;
                const float ya = __fsub_rn(tyf, fly);
                const float ya1 = __fsub_rn(1.0f, ya);
                uint32_t outw[2] = {0u, 0u};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint32_t pw = e < 4 ? pix[u].x : pix[u].y;
                    const float4 l = lutf[(pw >> (8 * (e & 3))) & 0xffu];
                    const float top = __fadd_rn(__fmul_rn(l.x, xa1[e]), __fmul_rn(l.y, xa[e]));
                    const float bot = __fadd_rn(__fmul_rn(l.z, xa1[e]), __fmul_rn(l.w, xa[e]));
                    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
                    // round-half-even via the 1.5*2^23 magic constant; 0 <= res < 255.5 so the byte cannot wrap
                    const uint32_t q = __float_as_uint(__fadd_rn(res, 12582912.0f)) & 0xffu;
                    outw[e >> 2] |= (uint32_t)q << (8 * (e & 3));
                }
                uint8_t* drow = dbase + (int64_t)y * d.dst_pitch;
                if (vec_ok && full) {
                    *reinterpret_cast<uint2*>(drow + x8) = make_uint2(outw[0], outw[1]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (mine & (1u << e)) drow[x8 + e] = (uint8_t)(outw[e >> 2] >> (8 * (e & 3)));
                }
            }
        }
        };
        if (six_rows) rows(std::integral_constant<int, 6>{}); else rows(std::integral_constant<int, 4>{});
    }
}

}  // namespace mdir

using namespace mdir;

extern "C" size_t mdir_clahe_workspace_bytes(int n_img, int tiles_x, int tiles_y) {
    if (n_img < 0 || tiles_x <= 0 || tiles_y <= 0) return 0;
    return (size_t)n_img * tiles_x * tiles_y * 256 + (size_t)n_img * sizeof(ClahePrep);
}

extern "C" int mdir_clahe_u8(const uint8_t* src, uint8_t* dst, const mdir_image_desc* descs, int n_img, int max_H, int max_W,
                             double clip, int tiles_x, int tiles_y, void* ws, void* stream) {
    MDIR_CHECK_ARG(src && dst && descs && ws);
    MDIR_CHECK_ARG(n_img >= 0 && n_img <= 65535);
    MDIR_CHECK_ARG(tiles_x >= 1 && tiles_y >= 1 && tiles_x * tiles_y <= 4096 && (((uintptr_t)ws) & 15) == 0);
    MDIR_CHECK_ARG(max_H >= 1 && max_W >= 1);
    if (n_img == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* luts = (uint8_t*)ws;
    ClahePrep* prep = reinterpret_cast<ClahePrep*>(luts + (size_t)n_img * tiles_x * tiles_y * 256);
    clahe_prep_kernel<<<(n_img + 255) / 256, 256, 0, st>>>(descs, n_img, clip, tiles_x, tiles_y, prep);
    MDIR_LAUNCH_CHECK();

    // Both passes read the source; processing the batch in chunks of <= ~48 MB of pixels lets the
    // interpolation pass of a chunk hit L2 (126 MB) for the pixels its LUT pass just streamed.
    const int64_t px = (int64_t)max_H * max_W;
    int chunk = (int)((int64_t)48 * 1024 * 1024 / (px > 0 ? px : 1));
    if (chunk < 1) chunk = 1;
    for (int i0 = 0; i0 < n_img; i0 += chunk) {
        const int n = (n_img - i0) < chunk ? (n_img - i0) : chunk;
        uint8_t* l = luts + (size_t)i0 * tiles_x * tiles_y * 256;
        clahe_lut_kernel<<<dim3(tiles_x, tiles_y, n), 256, 0, st>>>(src, descs + i0, prep + i0, l);
        MDIR_LAUNCH_CHECK();
        clahe_interp_kernel<<<dim3(tiles_x + 1, tiles_y + 1, n), 256, 0, st>>>(src, dst, descs + i0, prep + i0, l);
        MDIR_LAUNCH_CHECK();
    }
    return 0;
}
