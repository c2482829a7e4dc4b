/* C twin of image_stitching_b200/csrc/glibc_math.cuh (dev tool: tools/check_glibc_math.c compares it with libm). */
#include <stdint.h>
#include <string.h>
#ifndef FM_FN
#define FM_FN static inline
#endif
FM_FN float fm_u2f(uint32_t u){ float f; memcpy(&f,&u,4); return f; }
FM_FN uint32_t fm_f2u(float f){ uint32_t u; memcpy(&u,&f,4); return u; }
FM_FN float fm_atanf(float x)
{
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f,
                          -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f,
                          1.6285819933e-02f};
    const int32_t hx = (int32_t)fm_f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    if (ix >= FM_ATAN_BIG) {
        if (ix > 0x7f800000) return x + x;
        if (hx > 0) return atanhi[3] + atanlo[3];
        return -atanhi[3] - atanlo[3];
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) return x;
        id = -1;
    } else {
        x = fm_u2f((uint32_t)ix);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x, w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return hx < 0 ? -r : r;
}
FM_FN float fm_atan2f(float y, float x)
{
    const float tiny = 1.0e-30f, pi_o_2 = fm_u2f(0x3fc90fdbu), pi = fm_u2f(0x40490fdbu), pi_lo = fm_u2f(0xb3bbbd2eu);
    const int32_t hx = (int32_t)fm_f2u(x), hy = (int32_t)fm_f2u(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return fm_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
        case 0: case 1: return y;
        case 2: return pi + tiny;
        default: return -pi - tiny;
        }
    }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            switch (m) {
            case 0: return fm_u2f(0x3f490fdbu) + tiny;
            case 1: return -fm_u2f(0x3f490fdbu) - tiny;
            case 2: return 3.0f * fm_u2f(0x3f490fdbu) + tiny;
            default: return -3.0f * fm_u2f(0x3f490fdbu) - tiny;
            }
        } else {
            switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
            }
        }
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = fm_atanf(fm_u2f(fm_f2u(y / x) & 0x7fffffffu));
    switch (m) {
    case 0: return z;
    case 1: return fm_u2f(fm_f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}
FM_FN float fm_acosf(float x)
{
    const float pi = fm_u2f(0x40490fdau), pio2_hi = fm_u2f(0x3fc90fdau), pio2_lo = fm_u2f(0x33a22168u);
    const float pS0 = fm_u2f(0x3e2aaaabu), pS1 = fm_u2f(0xbea6b090u), pS2 = fm_u2f(0x3e4e0aa8u), pS3 = fm_u2f(0xbd241146u),
                pS4 = fm_u2f(0x3a4f7f04u), pS5 = fm_u2f(0x3811ef08u), qS1 = fm_u2f(0xc019d139u), qS2 = fm_u2f(0x4001572du),
                qS3 = fm_u2f(0xbf303361u), qS4 = fm_u2f(0x3d9dc62eu);
    const int32_t hx = (int32_t)fm_f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    if (ix > 0x3f800000) return (x - x) / (x - x);
    if (ix < 0x3f000000) {
        if (ix <= FM_ACOS_TINY) return pio2_hi + pio2_lo;
        const float z = x * x;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (hx < 0) {
        const float z = (1.0f + x) * 0.5f;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float s = FM_SQRTF(z);
        const float r = p / q;
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    } else {
        const float z = (1.0f - x) * 0.5f;
        const float s = FM_SQRTF(z);
        const float df = fm_u2f(fm_f2u(s) & 0xfffff000u);
        const float c = (z - df * df) / (s + df);
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = p / q;
        const float w = r * s + c;
        return 2.0f * (df + w);
    }
}
