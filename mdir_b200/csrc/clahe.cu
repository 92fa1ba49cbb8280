// CLAHE on batches of ragged 8UC1 images, bit-exact against cv2.createCLAHE(...).apply
// (algorithm: SURVEY.md App. A; oracle/oracle.py:clahe_u8).
//
//  kernel 1  clahe_lut_kernel    one CTA per (tile, image): per-warp private uint32
//            histograms in shared memory filled with warp-aggregated (match.any) updates,
//            clip-limit redistribution, block scan, LUT = sat_u8(rint(cdf * 255/area)).
//  kernel 2  clahe_interp_kernel one CTA per interpolation cell (the rectangle between four
//            tile centres, where the four contributing LUTs are fixed): the four LUTs are
//            interleaved into one uint32[256] table in shared memory so each pixel costs a
//            single LDS; the bilinear blend uses individually rounded fp32 mul/add in
//            OpenCV's association (no FMA contraction) and round-half-even.
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

__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t* __restrict__ src, const mdir_image_desc* __restrict__ descs,
                                                        double clip, int tiles_x, int tiles_y, uint8_t* __restrict__ luts) {
    __shared__ uint32_t whist[8][256];
    __shared__ int red_i[8];
    __shared__ int scan_w[8];
    const int img = blockIdx.y;
    const int tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const mdir_image_desc d = descs[img];
    const ClaheGeom g = clahe_geom(d.H, d.W, tiles_x, tiles_y);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&whist[0][0])[i] = 0u;
    __syncthreads();

    const uint8_t* base = src + d.src_off;
    const int x0 = tx * g.tw, y0 = ty * g.th;
    uint32_t* myh = whist[w];
    for (int r = w; r < g.th; r += 8) {
        const int sy = reflect101(y0 + r, d.H);
        const uint8_t* row = base + (int64_t)sy * d.src_pitch;
        for (int c0 = 0; c0 < g.tw; c0 += 32) {
            const int c = c0 + lane;
            const bool valid = c < g.tw;
            int v = 256 + lane;                       // unique sentinel: never groups with a pixel value
            if (valid) v = row[reflect101(x0 + c, d.W)];
            const unsigned peers = __match_any_sync(0xffffffffu, v);
            if (valid && lane == (__ffs(peers) - 1)) myh[v] += __popc(peers);   // leaders hold distinct bins
            __syncwarp();
        }
    }
    __syncthreads();

    const int i = threadIdx.x;
    int h = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) h += (int)whist[k][i];

    const int area = g.tw * g.th;
    int clip_limit = 0;
    if (clip > 0.0) {
        clip_limit = (int)(clip * (double)area / 256.0);
        clip_limit = max(clip_limit, 1);
    }
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
    const float lut_scale = __fdiv_rn(255.0f, (float)area);
    int q = __float2int_rn(__fmul_rn((float)s, lut_scale));
    q = min(max(q, 0), 255);
    luts[(((int64_t)img * tiles_y + ty) * tiles_x + tx) * 256 + i] = (uint8_t)q;
}

__device__ __forceinline__ float u8_to_f32(uint32_t b) {   // exact, avoids the quarter-rate I2F
    return __fadd_rn(__uint_as_float(0x4B000000u | b), -8388608.0f);
}

__global__ void __launch_bounds__(256) clahe_interp_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                           const mdir_image_desc* __restrict__ descs, int tiles_x, int tiles_y,
                                                           const uint8_t* __restrict__ luts) {
    __shared__ uint32_t lut4[256];
    const int img = blockIdx.y;
    const int cell = blockIdx.x;
    const int cy = cell / (tiles_x + 1), cx = cell - cy * (tiles_x + 1);
    const mdir_image_desc d = descs[img];
    const ClaheGeom g = clahe_geom(d.H, d.W, tiles_x, tiles_y);
    // nominal pixel ranges of this cell: raw tile index floor(x/tw - 0.5) == cx - 1
    const int margin = 2 + (max(g.tw, g.th) >> 8);
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
        lut4[v] = (uint32_t)L[(ty1 * tiles_x + tx1) * 256 + v] | ((uint32_t)L[(ty1 * tiles_x + tx2) * 256 + v] << 8) |
                  ((uint32_t)L[(ty2 * tiles_x + tx1) * 256 + v] << 16) | ((uint32_t)L[(ty2 * tiles_x + tx2) * 256 + v] << 24);
    }
    __syncthreads();

    const float inv_tw = __fdiv_rn(1.0f, (float)g.tw);
    const float inv_th = __fdiv_rn(1.0f, (float)g.th);
    const uint8_t* sbase = src + d.src_off;
    uint8_t* dbase = dst + d.dst_off;
    const bool vec_ok = (((uintptr_t)sbase | (uintptr_t)dbase | (uintptr_t)d.src_pitch | (uintptr_t)d.dst_pitch) & 3) == 0;

    const int gx = threadIdx.x & 31, gy = threadIdx.x >> 5;   // 32 x-groups of 4 pixels, 8 rows per pass
    const int xs4 = xs & ~3;
    for (int x4 = xs4 + gx * 4; x4 < xe; x4 += 128) {
        // per-column weights for the 4 pixels of this group (reused over all rows)
        float xa[4], xa1[4];
        bool mine_x[4];
        bool any_x = false, all_x = true;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int x = x4 + e;
            const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
            const float fl = floorf(txf);
            xa[e] = __fsub_rn(txf, fl);
            xa1[e] = __fsub_rn(1.0f, xa[e]);
            mine_x[e] = (x >= xs) && (x < xe) && ((int)fl == cx - 1);
            any_x |= mine_x[e];
            all_x &= mine_x[e];
        }
        if (!any_x) continue;
        for (int y = ys + gy; y < ye; y += 8) {
            const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
            const float fly = floorf(tyf);
            if ((int)fly != cy - 1) continue;
            const float ya = __fsub_rn(tyf, fly);
            const float ya1 = __fsub_rn(1.0f, ya);
            const uint8_t* srow = sbase + (int64_t)y * d.src_pitch;
            uint8_t* drow = dbase + (int64_t)y * d.dst_pitch;
            uint32_t pix;
            const bool full = all_x && (x4 + 3 < d.W);
            if (vec_ok && x4 + 3 < d.W) {
                pix = *reinterpret_cast<const uint32_t*>(srow + x4);
            } else {
                pix = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (x4 + e < d.W) pix |= (uint32_t)srow[x4 + e] << (8 * e);
            }
            uint32_t outw = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t l = lut4[(pix >> (8 * e)) & 0xffu];
                const float l11 = u8_to_f32(l & 0xffu), l12 = u8_to_f32((l >> 8) & 0xffu);
                const float l21 = u8_to_f32((l >> 16) & 0xffu), l22 = u8_to_f32(l >> 24);
                const float top = __fadd_rn(__fmul_rn(l11, xa1[e]), __fmul_rn(l12, xa[e]));
                const float bot = __fadd_rn(__fmul_rn(l21, xa1[e]), __fmul_rn(l22, xa[e]));
                const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
                // round-half-even via the 1.5*2^23 magic constant (0 <= res < 2^22)
                int q = (int)(__float_as_uint(__fadd_rn(res, 12582912.0f)) & 0x3ffu);
                q = min(q, 255);
                outw |= (uint32_t)q << (8 * e);
            }
            if (vec_ok && full) {
                *reinterpret_cast<uint32_t*>(drow + x4) = outw;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (mine_x[e]) drow[x4 + e] = (uint8_t)(outw >> (8 * e));
            }
        }
    }
}

}  // namespace mdir

using namespace mdir;

extern "C" size_t mdir_clahe_workspace_bytes(int n_img, int tiles_x, int tiles_y) {
    if (n_img < 0 || tiles_x <= 0 || tiles_y <= 0) return 0;
    return (size_t)n_img * tiles_x * tiles_y * 256;
}

extern "C" int mdir_clahe_u8(const uint8_t* src, uint8_t* dst, const mdir_image_desc* descs, int n_img, int max_H, int max_W,
                             double clip, int tiles_x, int tiles_y, void* ws, void* stream) {
    MDIR_CHECK_ARG(src && dst && descs && ws);
    MDIR_CHECK_ARG(n_img >= 0 && n_img <= 65535);
    MDIR_CHECK_ARG(tiles_x >= 1 && tiles_y >= 1 && tiles_x * tiles_y <= 4096);
    MDIR_CHECK_ARG(max_H >= 1 && max_W >= 1);
    (void)max_H; (void)max_W;
    if (n_img == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* luts = (uint8_t*)ws;
    clahe_lut_kernel<<<dim3(tiles_x * tiles_y, n_img), 256, 0, st>>>(src, descs, clip, tiles_x, tiles_y, luts);
    MDIR_LAUNCH_CHECK();
    clahe_interp_kernel<<<dim3((tiles_x + 1) * (tiles_y + 1), n_img), 256, 0, st>>>(src, dst, descs, tiles_x, tiles_y, luts);
    MDIR_LAUNCH_CHECK();
    return 0;
}
