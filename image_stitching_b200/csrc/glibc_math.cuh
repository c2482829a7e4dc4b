// glibc_math.cuh - atan2f / acosf with the results of glibc's float routines (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c,
// e_acosf.c: the fdlibm single-precision algorithms), restated in plain binary32 operations so that the device produces the
// bits the reference's host code gets from libm.  Needed where a transcendental is NOT separable: the forward map of
// RotationWarper::warpBackward (u = scale * atan2f(x_, z_), v = scale * (pi - acosf(w)) per pixel).  Every operation rounds
// on its own (the file is compiled with -fmad=false; divisions and square roots are the IEEE ones).  Pinned against the
// glibc of this image (2.39) over 3 x 10^8 random arguments per function, bit for bit: tools/check_glibc_math.c.
#pragma once
#include <cstdint>

namespace isb {
namespace gm {

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

__device__ inline float atanf_(float x)
{
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
                          -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f,
                          1.6285819933e-02f};
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25
        if (ix > 0x7f800000) return __fadd_rn(x, x);
        const float r = __fadd_rn(atanhi[3], atanlo[3]);
        return hx > 0 ? r : -r;
    }
    if (ix < 0x3ee00000) {  // |x| < 0.4375
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = u2f((uint32_t)ix);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x)); }
            else { id = 1; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x))); }
            else { id = 3; x = __fdiv_rn(-1.0f, x); }
        }
    }
    const float z = __fmul_rn(x, x), w = __fmul_rn(z, z);
    float s1 = __fadd_rn(aT[8], __fmul_rn(w, aT[10]));
    s1 = __fadd_rn(aT[6], __fmul_rn(w, s1));
    s1 = __fadd_rn(aT[4], __fmul_rn(w, s1));
    s1 = __fadd_rn(aT[2], __fmul_rn(w, s1));
    s1 = __fmul_rn(z, __fadd_rn(aT[0], __fmul_rn(w, s1)));
    float s2 = __fadd_rn(aT[7], __fmul_rn(w, aT[9]));
    s2 = __fadd_rn(aT[5], __fmul_rn(w, s2));
    s2 = __fadd_rn(aT[3], __fmul_rn(w, s2));
    s2 = __fmul_rn(w, __fadd_rn(aT[1], __fmul_rn(w, s2)));
    const float xs = __fmul_rn(x, __fadd_rn(s1, s2));
    if (id < 0) return __fsub_rn(x, xs);
    const float r = __fsub_rn(atanhi[id], __fsub_rn(__fsub_rn(xs, atanlo[id]), x));
    return hx < 0 ? -r : r;
}

__device__ inline float atan2f_(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_2 = u2f(0x3fc90fdbu), pi_o_4 = u2f(0x3f490fdbu), pi = u2f(0x40490fdbu), pi_lo = u2f(0xb3bbbd2eu);
    const int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return __fadd_rn(x, y);
    if (hx == 0x3f800000) return atanf_(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);  // 2 * sign(x) + sign(y)
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? __fadd_rn(pi, tiny) : __fsub_rn(-pi, tiny);
    }
    if (ix == 0) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return __fadd_rn(pi_o_4, tiny);
            case 1: return __fsub_rn(-pi_o_4, tiny);
            case 2: return __fadd_rn(__fmul_rn(3.0f, pi_o_4), tiny);
            default: return __fsub_rn(__fmul_rn(-3.0f, pi_o_4), tiny);
            }
        }
        switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return __fadd_rn(pi, tiny);
        default: return __fsub_rn(-pi, tiny);
        }
    }
    if (iy == 0x7f800000) return hy < 0 ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = atanf_(u2f(f2u(__fdiv_rn(y, x)) & 0x7fffffffu));
    switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return __fsub_rn(pi, __fsub_rn(z, pi_lo));
    default: return __fsub_rn(__fsub_rn(z, pi_lo), pi);
    }
}

__device__ inline float acosf_(float x)
{
    const float pi = u2f(0x40490fdau), pio2_hi = u2f(0x3fc90fdau), pio2_lo = u2f(0x33a22168u);
    const float pS0 = u2f(0x3e2aaaabu), pS1 = u2f(0xbea6b090u), pS2 = u2f(0x3e4e0aa8u), pS3 = u2f(0xbd241146u), pS4 = u2f(0x3a4f7f04u),
                pS5 = u2f(0x3811ef08u), qS1 = u2f(0xc019d139u), qS2 = u2f(0x4001572du), qS3 = u2f(0xbf303361u), qS4 = u2f(0x3d9dc62eu);
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : __fadd_rn(pi, __fmul_rn(2.0f, pio2_lo));
    if (ix > 0x3f800000) return __fdiv_rn(__fsub_rn(x, x), __fsub_rn(x, x));
    float z;
    if (ix < 0x3f000000) {
        if (ix <= 0x32800000) return __fadd_rn(pio2_hi, pio2_lo);
        z = __fmul_rn(x, x);
    } else if (hx < 0) z = __fmul_rn(__fadd_rn(1.0f, x), 0.5f);
    else z = __fmul_rn(__fsub_rn(1.0f, x), 0.5f);
    float p = __fadd_rn(pS4, __fmul_rn(z, pS5));
    p = __fadd_rn(pS3, __fmul_rn(z, p));
    p = __fadd_rn(pS2, __fmul_rn(z, p));
    p = __fadd_rn(pS1, __fmul_rn(z, p));
    p = __fmul_rn(z, __fadd_rn(pS0, __fmul_rn(z, p)));
    float q = __fadd_rn(qS3, __fmul_rn(z, qS4));
    q = __fadd_rn(qS2, __fmul_rn(z, q));
    q = __fadd_rn(qS1, __fmul_rn(z, q));
    q = __fadd_rn(1.0f, __fmul_rn(z, q));
    const float r = __fdiv_rn(p, q);
    if (ix < 0x3f000000) return __fsub_rn(pio2_hi, __fsub_rn(x, __fsub_rn(pio2_lo, __fmul_rn(x, r))));
    const float s = __fsqrt_rn(z);
    if (hx < 0) {
        const float w = __fsub_rn(__fmul_rn(r, s), pio2_lo);
        return __fsub_rn(pi, __fmul_rn(2.0f, __fadd_rn(s, w)));
    }
    const float df = u2f(f2u(s) & 0xfffff000u);
    const float c = __fdiv_rn(__fsub_rn(z, __fmul_rn(df, df)), __fadd_rn(s, df));
    const float w = __fadd_rn(__fmul_rn(r, s), c);
    return __fmul_rn(2.0f, __fadd_rn(df, w));
}

}  // namespace gm
}  // namespace isb
