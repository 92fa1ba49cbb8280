// RGB <-> Lab around the CLAHE kernel, following OpenCV's FLOAT code path as called by the reference
// (rgb2normspace / normspace2rgb, mdir/components/data/transform/functional.py:24-48).
//
//  * RGB -> Lab (cv2.cvtColor(float32, COLOR_RGB2LAB)) is NOT the closed-form Lab formula: OpenCV
//    interpolates a 33^3 fixed-point table trilinearly with 4-bit weights (color_lab.cpp RGB2Lab_f,
//    trilinearInterpolate; LAB_BASE = 2^14).  The table (mdir_b200/data/rgb2lab_lut_s16.npy, recovered from
//    the wheel by tools/make_lab_lut.py) is passed in; with it the integer arithmetic here is bit-exact.
//  * Lab -> RGB (COLOR_LAB2RGB, float) is closed form + a 1024-interval natural cubic spline for the sRGB
//    gamma (Lab2RGBfloat, splineInterpolate); the spline table is built on the host and passed in.
//
//  kernel 1  rgb_to_l_u8_kernel     RGB f32 HWC -> the uint8 L plane CLAHE consumes:
//                                    u8 = trunc(((L + 0) / 100) * 255)          (functional.py:27,117)
//  kernel 2  lab_clahe_to_rgb_kernel RGB f32 HWC (for a, b) + CLAHE'd u8 L plane -> RGB f32 HWC
//                                    (functional.py:128-129,41); a, b are recomputed rather than stored.
// Both are elementwise and HBM-bound: 13 B/pixel and 25 B/pixel.
#include "common.cuh"

namespace mdir {

constexpr int kLabDim = 33;

struct LabInterp {
    int base;          // lattice index of the cell origin
    int w[8];          // trilinear weights, sum = 4096
    int off[8];        // lattice offsets of the 8 corners
};

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// cv2: iR = cvRound(clip(R) * LAB_BASE); cell = iR >> 9; 4-bit fraction = (iR & 511) >> 5
__device__ __forceinline__ void lab_cell(float R, float G, float B, int& tx, int& ty, int& tz, int& fx, int& fy, int& fz) {
    const int iR = __float2int_rn(__fmul_rn(clip01(R), 16384.0f));
    const int iG = __float2int_rn(__fmul_rn(clip01(G), 16384.0f));
    const int iB = __float2int_rn(__fmul_rn(clip01(B), 16384.0f));
    tx = iR >> 9; ty = iG >> 9; tz = iB >> 9;
    fx = (iR & 511) >> 5; fy = (iG & 511) >> 5; fz = (iB & 511) >> 5;
}

// lut: (33,33,33) entries of short4 {L, a, b, 0}.  CH_MASK selects which channels to interpolate.
template <bool WANT_L, bool WANT_AB>
__device__ __forceinline__ void lab_interp(const short4* __restrict__ lut, float R, float G, float B, int& iL, int& ia, int& ib) {
    int tx, ty, tz, fx, fy, fz;
    lab_cell(R, G, B, tx, ty, tz, fx, fy, fz);
    const int tx1 = min(tx + 1, kLabDim - 1), ty1 = min(ty + 1, kLabDim - 1), tz1 = min(tz + 1, kLabDim - 1);
    int aL = 0, aa = 0, ab = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int dx = c & 1, dy = (c >> 1) & 1, dz = c >> 2;
        const int w = (dx ? fx : 16 - fx) * (dy ? fy : 16 - fy) * (dz ? fz : 16 - fz);
        const short4 v = __ldg(&lut[((dx ? tx1 : tx) * kLabDim + (dy ? ty1 : ty)) * kLabDim + (dz ? tz1 : tz)]);
        if (WANT_L) aL += (int)v.x * w;
        if (WANT_AB) { aa += (int)v.y * w; ab += (int)v.z * w; }
    }
    iL = (aL + 2048) >> 12;             // CV_DESCALE(x, 12)
    ia = (aa + 2048) >> 12;
    ib = (ab + 2048) >> 12;
}

__global__ void __launch_bounds__(256) rgb_to_l_u8_kernel(const float* __restrict__ rgb, const mdir_rgb_desc* __restrict__ descs,
                                                          const short4* __restrict__ lut, uint8_t* __restrict__ l_out) {
    const mdir_rgb_desc d = descs[blockIdx.y];
    const int64_t npx = (int64_t)d.H * d.W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x * blockDim.x) {
        const float* p = rgb + d.rgb_off + i * 3;
        int iL, ia, ib;
        lab_interp<true, false>(lut, p[0], p[1], p[2], iL, ia, ib);
        const float L = __fmul_rn(__fmul_rn((float)iL, 1.0f / 16384.0f), 100.0f);     // cv2: L*1.0f/LAB_BASE, then *100
        const float chan = __fdiv_rn(L, 100.0f);                                        // rgb2normspace: (L + 0) / 100
        l_out[d.l_off + i] = (uint8_t)(int)__fmul_rn(chan, 255.0f);                      // (chan*255).astype(uint8): truncation
    }
}

__device__ __forceinline__ float spline_eval(const float4* __restrict__ tab, float x) {
    // splineInterpolate(x * 1024, gammaTab, 1024)
    const float xs = __fmul_rn(x, 1024.0f);
    int ix = (int)xs;
    ix = min(max(ix, 0), 1023);
    const float t = xs - (float)ix;
    const float4 c = __ldg(&tab[ix]);
    return ((c.w * t + c.z) * t + c.y) * t + c.x;
}

__global__ void __launch_bounds__(256) lab_clahe_to_rgb_kernel(const float* __restrict__ rgb, const mdir_rgb_desc* __restrict__ descs,
                                                               const short4* __restrict__ lut, const float4* __restrict__ gamma_tab,
                                                               const uint8_t* __restrict__ l_in, float* __restrict__ out) {
    const mdir_rgb_desc d = descs[blockIdx.y];
    const int64_t npx = (int64_t)d.H * d.W;
    // XYZ -> sRGB (D65) with the white point folded in, as Lab2RGBfloat builds them
    const float C0 = 3.240479f * 0.950456f, C1 = -1.53715f, C2 = -0.498535f * 1.088754f;
    const float C3 = -0.969256f * 0.950456f, C4 = 1.875991f, C5 = 0.041556f * 1.088754f;
    const float C6 = 0.055648f * 0.950456f, C7 = -0.204043f, C8 = 1.057311f * 1.088754f;
    const float lThresh = 0.008856f * 903.3f;
    const float fThresh = 7.787f * 0.008856f + 16.0f / 116.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x * blockDim.x) {
        const float* p = rgb + d.rgb_off + i * 3;
        int iL, ia, ib;
        lab_interp<false, true>(lut, p[0], p[1], p[2], iL, ia, ib);
        // cv2 output a, b; then the reference's normspace round trip (functional.py:27,41)
        float a = __fsub_rn(__fmul_rn(__fmul_rn((float)ia, 1.0f / 16384.0f), 256.0f), 128.0f);
        float b = __fsub_rn(__fmul_rn(__fmul_rn((float)ib, 1.0f / 16384.0f), 256.0f), 128.0f);
        a = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(a, 128.0f), 255.0f), 255.0f), 128.0f);
        b = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(b, 128.0f), 255.0f), 255.0f), 128.0f);
        const float li = __fmul_rn(__fdiv_rn((float)l_in[d.l_off + i], 255.0f), 100.0f);   // u8 / 255.0, then * 100 - 0
        // Lab2RGBfloat
        float y, fy;
        if (li <= lThresh) {
            y = li / 903.3f;
            fy = 7.787f * y + 16.0f / 116.0f;
        } else {
            fy = (li + 16.0f) / 116.0f;
            y = fy * fy * fy;
        }
        float fx = a / 500.0f + fy, fz = fy - b / 200.0f;
        fx = fx <= fThresh ? (fx - 16.0f / 116.0f) / 7.787f : fx * fx * fx;
        fz = fz <= fThresh ? (fz - 16.0f / 116.0f) / 7.787f : fz * fz * fz;
        const float ro = clip01(C0 * fx + C1 * y + C2 * fz);
        const float go = clip01(C3 * fx + C4 * y + C5 * fz);
        const float bo = clip01(C6 * fx + C7 * y + C8 * fz);
        float* o = out + d.out_off + i * 3;
        o[0] = spline_eval(gamma_tab, ro);
        o[1] = spline_eval(gamma_tab, go);
        o[2] = spline_eval(gamma_tab, bo);
    }
}

}  // namespace mdir

using namespace mdir;

extern "C" int mdir_rgb_to_l_u8(const float* rgb, const mdir_rgb_desc* descs, int n_img, int64_t max_pixels, const int16_t* lut,
                                uint8_t* l_out, void* stream) {
    MDIR_CHECK_ARG(rgb && descs && lut && l_out && n_img >= 0 && max_pixels >= 1);
    MDIR_CHECK_ARG(((uintptr_t)lut & 7) == 0 && n_img <= 65535);
    if (n_img == 0) return 0;
    int64_t gx = (max_pixels + 255) / 256;
    if (gx > 4 * kNumSMs * 8) gx = 4 * kNumSMs * 8;
    rgb_to_l_u8_kernel<<<dim3((unsigned)gx, n_img), 256, 0, (cudaStream_t)stream>>>(rgb, descs, (const short4*)lut, l_out);
    MDIR_LAUNCH_CHECK();
    return 0;
}

extern "C" int mdir_lab_clahe_to_rgb(const float* rgb, const mdir_rgb_desc* descs, int n_img, int64_t max_pixels, const int16_t* lut,
                                     const float* gamma_tab, const uint8_t* l_in, float* out, void* stream) {
    MDIR_CHECK_ARG(rgb && descs && lut && gamma_tab && l_in && out && n_img >= 0 && max_pixels >= 1);
    MDIR_CHECK_ARG(((uintptr_t)lut & 7) == 0 && ((uintptr_t)gamma_tab & 15) == 0 && n_img <= 65535);
    if (n_img == 0) return 0;
    int64_t gx = (max_pixels + 255) / 256;
    if (gx > 4 * kNumSMs * 8) gx = 4 * kNumSMs * 8;
    lab_clahe_to_rgb_kernel<<<dim3((unsigned)gx, n_img), 256, 0, (cudaStream_t)stream>>>(rgb, descs, (const short4*)lut,
                                                                                          (const float4*)gamma_tab, l_in, out);
    MDIR_LAUNCH_CHECK();
    return 0;
}
