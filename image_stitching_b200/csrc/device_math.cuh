// device_math.cuh - device-side arithmetic shared by all kernels (bit-exact contract of SURVEY.md Appendix A).
#pragma once
#include <climits>
#include <cstdint>
#include <type_traits>

#include "device_types.hpp"

namespace isb {

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cv_round(float v)
{  // cvRound on x86 (cvtss2si): half-to-even, INT_MIN when out of range or NaN
    return (fabsf(v) < 2147483648.f) ? __float2int_rn(v) : INT_MIN;
}
__device__ __forceinline__ int sat_s16(int v) { return min(max(v, -32768), 32767); }
__device__ __forceinline__ int sat_u8(int v) { return min(max(v, 0), 255); }
// static_cast<short>(float): cvttss2si then the low 16 bits
__device__ __forceinline__ int trunc_s16(float v) { return (int)(short)__float2int_rz(v); }

// cv::BORDER_REFLECT  fedcba|abcdefgh|hgfedcb.  One reflection covers -n <= p < 2n; farther indices (sentinel
// coordinates, saturated shorts) take the out-of-line modulo path so the hot path carries no integer division.
static __device__ __noinline__ int reflect_far(int p, int n)
{
    if (n == 1) return 0;
    if (p < 0) p = -p - 1;
    const int m = 2 * n;
    p %= m;
    return p < n ? p : m - 1 - p;
}
__device__ __forceinline__ int reflect(int p, int n)
{
    if ((unsigned)p < (unsigned)n) return p;
    const int q = p < 0 ? -p - 1 : 2 * n - 1 - p;
    if ((unsigned)q < (unsigned)n) return q;
    return reflect_far(p, n);
}
__device__ __forceinline__ int reflect101(int p, int n)
{  // cv::BORDER_REFLECT_101 for |overshoot| < n (pyramid taps overshoot by <= 2)
    if (n == 1) return 0;
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return max(p, 0);
}

struct XY { float x, y; };

// mapBackward with the transcendentals taken from the separable host tables (A.2)
__device__ __forceinline__ XY inverse_map(const float* __restrict__ kr, F2 c, F2 r)
{
    const float x_ = __fmul_rn(r.a, c.a), y_ = r.b, z_ = __fmul_rn(r.a, c.b);
    float x = __fadd_rn(__fadd_rn(__fmul_rn(kr[0], x_), __fmul_rn(kr[1], y_)), __fmul_rn(kr[2], z_));
    float y = __fadd_rn(__fadd_rn(__fmul_rn(kr[3], x_), __fmul_rn(kr[4], y_)), __fmul_rn(kr[5], z_));
    const float z = __fadd_rn(__fadd_rn(__fmul_rn(kr[6], x_), __fmul_rn(kr[7], y_)), __fmul_rn(kr[8], z_));
    if (z > 0.f) {
        x = __fdiv_rn(x, z);
        y = __fdiv_rn(y, z);
    } else {
        x = y = -1.f;
    }
    return XY{x, y};
}

// cv::remap INTER_LINEAR on 8U: fixed-point coordinates (1/32 px) and 15-bit weights (A.3)
struct BilinearTaps {
    int x0, y0, w00, w01, w10, w11;
};
__device__ __forceinline__ BilinearTaps bilinear_taps(XY m)
{
    const int sx = cv_round(__fmul_rn(m.x, 32.f)), sy = cv_round(__fmul_rn(m.y, 32.f));
    BilinearTaps t;
    t.x0 = sat_s16(sx >> 5);
    t.y0 = sat_s16(sy >> 5);
    const int a = sx & 31, b = sy & 31;
    t.w00 = (32 - a) * (32 - b) * 32;
    t.w01 = a * (32 - b) * 32;
    t.w10 = (32 - a) * b * 32;
    t.w11 = a * b * 32;
    return t;
}

template <int CH, bool REFLECT>
__device__ __forceinline__ void sample_linear(const ImageDev& I, XY m, int out[CH])
{
    const BilinearTaps t = bilinear_taps(m);
    int xa = t.x0, xb = t.x0 + 1, ya = t.y0, yb = t.y0 + 1;
    bool vxa = true, vxb = true, vya = true, vyb = true;
    if (REFLECT) {
        xa = reflect(xa, I.sw); xb = reflect(xb, I.sw); ya = reflect(ya, I.sh); yb = reflect(yb, I.sh);
    } else {
        vxa = (unsigned)xa < (unsigned)I.sw; vxb = (unsigned)xb < (unsigned)I.sw;
        vya = (unsigned)ya < (unsigned)I.sh; vyb = (unsigned)yb < (unsigned)I.sh;
        xa = vxa ? xa : 0; xb = vxb ? xb : 0; ya = vya ? ya : 0; yb = vyb ? yb : 0;
    }
    const uint8_t* ra = I.src + (long long)ya * I.spitch;
    const uint8_t* rb = I.src + (long long)yb * I.spitch;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int p00 = (vxa && vya) ? __ldg(ra + xa * CH + c) : 0;
        const int p01 = (vxb && vya) ? __ldg(ra + xb * CH + c) : 0;
        const int p10 = (vxa && vyb) ? __ldg(rb + xa * CH + c) : 0;
        const int p11 = (vxb && vyb) ? __ldg(rb + xb * CH + c) : 0;
        out[c] = sat_u8((p00 * t.w00 + p01 * t.w01 + p10 * t.w10 + p11 * t.w11 + (1 << 14)) >> 15);
    }
}

// validity of the nearest/constant mask warp: 255 iff round-half-even(x), (y) fall inside the source
__device__ __forceinline__ bool nearest_inside(const ImageDev& I, XY m, int& ix, int& iy)
{
    ix = sat_s16(cv_round(m.x));
    iy = sat_s16(cv_round(m.y));
    return (unsigned)ix < (unsigned)I.sw && (unsigned)iy < (unsigned)I.sh;
}

// cv::resize(f32, INTER_LINEAR) of the gain grid evaluated at one ROI pixel (A.7)
__device__ __forceinline__ float gain_at(const ImageDev& I, LinCoefDev cx, LinCoefDev cy)
{
    const int c0 = cx.ofs, c1 = min(cx.ofs + 1, I.gw - 1);
    const int r0 = min(max(cy.ofs, 0), I.gh - 1), r1 = min(max(cy.ofs + 1, 0), I.gh - 1);
    const float a1 = cx.frac, a0 = __fsub_rn(1.f, a1), b1 = cy.frac, b0 = __fsub_rn(1.f, b1);
    const float* g0 = I.gain + r0 * I.gw;
    const float* g1 = I.gain + r1 * I.gw;
    const float h0 = __fadd_rn(__fmul_rn(__ldg(g0 + c0), a0), __fmul_rn(__ldg(g0 + c1), a1));
    const float h1 = __fadd_rn(__fmul_rn(__ldg(g1 + c0), a0), __fmul_rn(__ldg(g1 + c1), a1));
    return __fadd_rn(__fmul_rn(h0, b0), __fmul_rn(h1, b1));
}

// cv::resize(8U, INTER_LINEAR_EXACT) of the dilated seam mask at one ROI pixel (A.4)
__device__ __forceinline__ int seam_at(const uint8_t* __restrict__ dil, int mw, int mh, uint32_t tx, uint32_t ty)
{
    const int c0 = tx >> 16, ax = tx & 0xffff, r0 = ty >> 16, ay = ty & 0xffff;
    const int c1 = min(c0 + 1, mw - 1), r1 = min(r0 + 1, mh - 1);
    const uint8_t* p0 = dil + r0 * mw;
    const uint8_t* p1 = dil + r1 * mw;
    const int h0 = __ldg(p0 + c0) * (256 - ax) + __ldg(p0 + c1) * ax;
    const int h1 = __ldg(p1 + c0) * (256 - ax) + __ldg(p1 + c1) * ax;
    return (h0 * (256 - ay) + h1 * ay + 32768) >> 16;
}

// weight taps in OpenCV's operation order (A.5): `simd` = the 4-lane SIMD formulation, else the scalar one
__device__ __forceinline__ float wdown_h(float r0, float r1, float r2, float r3, float r4, bool simd)
{
    if (simd) return __fadd_rn(__fmul_rn(r2, 6.f), __fadd_rn(__fmul_rn(__fadd_rn(r1, r3), 4.f), __fadd_rn(r0, r4)));
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2, 6.f), __fmul_rn(__fadd_rn(r1, r3), 4.f)), r0), r4);
}
__device__ __forceinline__ float wdown_v(float r0, float r1, float r2, float r3, float r4, bool simd)
{
    float v;
    if (simd)
        v = __fadd_rn(__fmul_rn(__fadd_rn(__fadd_rn(r1, r3), r2), 4.f), __fadd_rn(__fadd_rn(r0, r4), __fadd_rn(r2, r2)));
    else
        v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2, 6.f), __fmul_rn(__fadd_rn(r1, r3), 4.f)), r0), r4);
    return __fmul_rn(v, 1.f / 256.f);
}

// level-l sample of a tile's Gaussian / weight pyramid, whichever storage the tile uses
__device__ __forceinline__ int tile_g(const TileDev& T, int l, int p, int x, int y)
{
    if (T.packed) return (int)((T.P[l][y * T.ppitch[l] + x] >> (8 * p)) & 0xffu);
    return T.G[l][p * T.gplane[l] + (long long)y * T.gpitch[l] + x];
}
__device__ __forceinline__ float tile_w(const TileDev& T, int l, int x, int y)
{
    if (l == 0 && T.packed) return __fmul_rn((float)(T.P[0][y * T.ppitch[0] + x] >> 24), (float)(1. / 255.));
    return T.W[l][(long long)y * T.wpitch[l] + x];
}

// collapsed destination level: 16S x4 interleaved
__device__ __forceinline__ void c_unpack(uint2 v, int& b, int& g, int& r)
{
    b = (short)(v.x & 0xffffu);
    g = (short)(v.x >> 16);
    r = (short)(v.y & 0xffffu);
}
__device__ __forceinline__ uint2 c_pack(int b, int g, int r)
{
    return make_uint2(((uint32_t)b & 0xffffu) | ((uint32_t)g << 16), (uint32_t)r & 0xffffu);
}

// cv::pyrUp (to exactly 2x) of a collapsed level at one fine position; edge rule s[-1] := s[1], s[n] := s[n-1]
__device__ __forceinline__ void pyrup_c_at(const uint2* __restrict__ c, int pitch, int wc, int hc, int fx, int fy, int out[3])
{
    const int cx = fx >> 1, cy = fy >> 1;
    const int xi[3] = {cx == 0 ? (wc > 1 ? 1 : 0) : cx - 1, cx, cx == wc - 1 ? cx : cx + 1};
    const int yi[3] = {cy == 0 ? (hc > 1 ? 1 : 0) : cy - 1, cy, cy == hc - 1 ? cy : cy + 1};
    const int wx[3] = {(fx & 1) ? 0 : 1, (fx & 1) ? 4 : 6, (fx & 1) ? 4 : 1};
    const int wy[3] = {(fy & 1) ? 0 : 1, (fy & 1) ? 4 : 6, (fy & 1) ? 4 : 1};
    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int h[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int b, g, r;
            c_unpack(c[yi[j] * pitch + xi[i]], b, g, r);
            h[0] += wx[i] * b; h[1] += wx[i] * g; h[2] += wx[i] * r;
        }
        acc[0] += wy[j] * h[0]; acc[1] += wy[j] * h[1]; acc[2] += wy[j] * h[2];
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) out[p] = sat_s16((acc[p] + 32) >> 6);
}

// the same for one channel of a tile's (l+1) Gaussian level
__device__ __forceinline__ int pyrup_tile_at(const TileDev& T, int lc, int p, int fx, int fy)
{
    const int wc = T.w >> lc, hc = T.h >> lc;
    const int cx = fx >> 1, cy = fy >> 1;
    const int xi[3] = {cx == 0 ? (wc > 1 ? 1 : 0) : cx - 1, cx, cx == wc - 1 ? cx : cx + 1};
    const int yi[3] = {cy == 0 ? (hc > 1 ? 1 : 0) : cy - 1, cy, cy == hc - 1 ? cy : cy + 1};
    const int wx[3] = {(fx & 1) ? 0 : 1, (fx & 1) ? 4 : 6, (fx & 1) ? 4 : 1};
    const int wy[3] = {(fy & 1) ? 0 : 1, (fy & 1) ? 4 : 6, (fy & 1) ? 4 : 1};
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        int h = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) h += wx[i] * tile_g(T, lc, p, xi[i], yi[j]);
        acc += wy[j] * h;
    }
    return sat_s16((acc + 32) >> 6);
}

}  // namespace isb
