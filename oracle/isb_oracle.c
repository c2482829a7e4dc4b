/*
 * isb_oracle.c - TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, scalar, single
 * thread) of the arithmetic the reference's compositing hot path executes.
 *
 * Reference path: /root/reference/image_stitching/image_stitching.cpp:1086-1229.  Every
 * flop of that path runs inside a third-party dependency that is NOT vendored under
 * /root/reference: OpenCV (vcpkg `opencv4[world]`, builtin-baseline 7bc5b8cd..., see
 * vcpkg.json:4-11).  This file restates OpenCV's published algorithms for
 *   cv::detail::RotationWarper::{warpRoi,buildMaps,warp}   (call sites :1138,:1154,:1159)
 *   cv::remap 8U INTER_LINEAR/BORDER_REFLECT, INTER_NEAREST/BORDER_CONSTANT
 *   cv::detail::BlocksGainCompensator::apply               (:1162)
 *   cv::dilate(3x3), cv::resize(INTER_LINEAR_EXACT), &     (:1169-1171)
 *   cv::detail::MultiBandBlender::{prepare,feed,blend}     (:1173-1225)
 *   saturate 16S->8U of imwrite                            (:1228)
 * following SURVEY.md Appendix A item by item.
 *
 * PARITY PINNING: the reference ships no tests / golden vectors for this path.  The
 * restatement is pinned instead against outputs of the dependency itself (cv2 4.13.0
 * wheel, IPP off) run in the build container: tests/test_oracle_vs_cv2.py compares every
 * function here with cv2 bit for bit, and tests/golden/ holds cv2-generated vectors (with
 * the generating script) that are re-checked wherever cv2 is absent.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may load this library.
 * The product (image_stitching_b200/) never links, loads or calls it.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * All float expressions are written so that every operation rounds to binary32
 * separately (no FMA contraction) - that is what the SSE-baseline OpenCV build does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define ORC_API __attribute__((visibility("default")))

enum { ORC_SPHERICAL = 0, ORC_CYLINDRICAL = 1 };
enum { ORC_NEAREST = 0, ORC_LINEAR = 1 };

typedef struct {
    float scale;
    float k[9], rinv[9], r_kinv[9], k_rinv[9];
} orc_projector;

/* ---- A.1 ProjectorBase::setCameraParams ------------------------------------------- */
static void mat3_mul_f32(const float* a, const float* b, float* c)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = a[i * 3 + 0] * b[0 * 3 + j];
            s = s + a[i * 3 + 1] * b[1 * 3 + j];
            s = s + a[i * 3 + 2] * b[2 * 3 + j];
            c[i * 3 + j] = s;
        }
}

/* cv::invert(3x3 f32, DECOMP_LU): closed-form adjugate evaluated in double, rounded to f32 */
static void mat3_inv_f32(const float* m, float* inv)
{
    double a00 = m[0], a01 = m[1], a02 = m[2], a10 = m[3], a11 = m[4], a12 = m[5], a20 = m[6], a21 = m[7],
           a22 = m[8];
    double d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
    if (d == 0) {
        memset(inv, 0, 9 * sizeof(float));
        return;
    }
    d = 1. / d;
    double t[9];
    t[0] = (a11 * a22 - a12 * a21) * d;
    t[1] = (a02 * a21 - a01 * a22) * d;
    t[2] = (a01 * a12 - a02 * a11) * d;
    t[3] = (a12 * a20 - a10 * a22) * d;
    t[4] = (a00 * a22 - a02 * a20) * d;
    t[5] = (a02 * a10 - a00 * a12) * d;
    t[6] = (a10 * a21 - a11 * a20) * d;
    t[7] = (a01 * a20 - a00 * a21) * d;
    t[8] = (a00 * a11 - a01 * a10) * d;
    for (int i = 0; i < 9; ++i) inv[i] = (float)t[i];
}

ORC_API void orc_projector_setup(orc_projector* p, float scale, const float* K, const float* R)
{
    float kinv[9], rt[9];
    p->scale = scale;
    memcpy(p->k, K, sizeof(p->k));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rt[i * 3 + j] = R[j * 3 + i];
    memcpy(p->rinv, rt, sizeof(rt));
    mat3_inv_f32(K, kinv);
    mat3_mul_f32(R, kinv, p->r_kinv);
    mat3_mul_f32(K, rt, p->k_rinv);
}

/* ---- A.2 forward / backward maps ---------------------------------------------------- */
static const float PI_F = (float)3.1415926535897932384626433832795;

static void map_forward(const orc_projector* p, int kind, float x, float y, float* u, float* v)
{
    const float* r = p->r_kinv;
    float x_ = r[0] * x + r[1] * y + r[2];
    float y_ = r[3] * x + r[4] * y + r[5];
    float z_ = r[6] * x + r[7] * y + r[8];
    if (kind == ORC_SPHERICAL) {
        *u = p->scale * atan2f(x_, z_);
        float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        *v = p->scale * (PI_F - acosf(w == w ? w : 0));
    } else {
        *u = p->scale * atan2f(x_, z_);
        *v = p->scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    }
}

static void map_backward(const orc_projector* p, int kind, float u, float v, float* x, float* y)
{
    const float* k = p->k_rinv;
    float x_, y_, z_;
    u /= p->scale;
    v /= p->scale;
    if (kind == ORC_SPHERICAL) {
        float sinv = sinf(PI_F - v);
        x_ = sinv * sinf(u);
        y_ = cosf(PI_F - v);
        z_ = sinv * cosf(u);
    } else {
        x_ = sinf(u);
        y_ = v;
        z_ = cosf(u);
    }
    float z;
    *x = k[0] * x_ + k[1] * y_ + k[2] * z_;
    *y = k[3] * x_ + k[4] * y_ + k[5] * z_;
    z = k[6] * x_ + k[7] * y_ + k[8] * z_;
    if (z > 0) {
        *x /= z;
        *y /= z;
    } else
        *x = *y = -1;
}

ORC_API void orc_map_forward(int kind, float scale, const float* K, const float* R, float x, float y, float* uv)
{
    orc_projector p;
    orc_projector_setup(&p, scale, K, R);
    map_forward(&p, kind, x, y, uv, uv + 1);
}

ORC_API void orc_map_backward(int kind, float scale, const float* K, const float* R, float u, float v, float* xy)
{
    orc_projector p;
    orc_projector_setup(&p, scale, K, R);
    map_backward(&p, kind, u, v, xy, xy + 1);
}

/* detectResultRoiByBorder (+ the spherical pole rule); returns tl,br INCLUSIVE as buildMaps uses them */
static void detect_roi(const orc_projector* p, int kind, int w, int h, int* tl, int* br)
{
    float tl_uf = INFINITY, tl_vf = INFINITY, br_uf = -INFINITY, br_vf = -INFINITY;
    float u, v;
    for (int x = 0; x < w; ++x) {
        map_forward(p, kind, (float)x, 0.f, &u, &v);
        tl_uf = fminf(tl_uf, u); tl_vf = fminf(tl_vf, v); br_uf = fmaxf(br_uf, u); br_vf = fmaxf(br_vf, v);
        map_forward(p, kind, (float)x, (float)(h - 1), &u, &v);
        tl_uf = fminf(tl_uf, u); tl_vf = fminf(tl_vf, v); br_uf = fmaxf(br_uf, u); br_vf = fmaxf(br_vf, v);
    }
    for (int y = 0; y < h; ++y) {
        map_forward(p, kind, 0.f, (float)y, &u, &v);
        tl_uf = fminf(tl_uf, u); tl_vf = fminf(tl_vf, v); br_uf = fmaxf(br_uf, u); br_vf = fmaxf(br_vf, v);
        map_forward(p, kind, (float)(w - 1), (float)y, &u, &v);
        tl_uf = fminf(tl_uf, u); tl_vf = fminf(tl_vf, v); br_uf = fmaxf(br_uf, u); br_vf = fmaxf(br_vf, v);
    }
    tl[0] = (int)tl_uf; tl[1] = (int)tl_vf; br[0] = (int)br_uf; br[1] = (int)br_vf;
    if (kind == ORC_SPHERICAL) {
        int tl_u = tl[0], tl_v = tl[1], br_u = br[0], br_v = br[1];
        const float* ri = p->rinv;
        const float* k = p->k;
        float x = ri[1], y = ri[4], z = ri[7];
        if (y > 0.f) {
            float x_ = (k[0] * x + k[1] * y) / z + k[2];
            float y_ = k[4] * y / z + k[5];
            if (x_ > 0.f && x_ < w && y_ > 0.f && y_ < h) {
                int pu = 0, pv = (int)(float)(3.1415926535897932384626433832795 * (double)p->scale);
                if (pu < tl_u) tl_u = pu; if (pv < tl_v) tl_v = pv;
                if (pu > br_u) br_u = pu; if (pv > br_v) br_v = pv;
            }
        }
        x = ri[1]; y = -ri[4]; z = ri[7];
        if (y > 0.f) {
            float x_ = (k[0] * x + k[1] * y) / z + k[2];
            float y_ = k[4] * y / z + k[5];
            if (x_ > 0.f && x_ < w && y_ > 0.f && y_ < h) {
                int pu = 0, pv = 0;
                if (pu < tl_u) tl_u = pu; if (pv < tl_v) tl_v = pv;
                if (pu > br_u) br_u = pu; if (pv > br_v) br_v = pv;
            }
        }
        tl[0] = tl_u; tl[1] = tl_v; br[0] = br_u; br[1] = br_v;
    }
}

/* warpRoi = Rect(tl, br + 1) ; buildMaps/warp use Rect(tl, br) with (h+1)x(w+1) maps */
ORC_API void orc_warp_roi(int kind, float scale, int w, int h, const float* K, const float* R, int* rect_xywh)
{
    orc_projector p;
    int tl[2], br[2];
    orc_projector_setup(&p, scale, K, R);
    detect_roi(&p, kind, w, h, tl, br);
    rect_xywh[0] = tl[0];
    rect_xywh[1] = tl[1];
    rect_xywh[2] = br[0] - tl[0] + 1;
    rect_xywh[3] = br[1] - tl[1] + 1;
}

/* xmap/ymap are (rect_h x rect_w) with rect = orc_warp_roi() */
ORC_API void orc_build_maps(int kind, float scale, int w, int h, const float* K, const float* R, float* xmap,
                            float* ymap)
{
    orc_projector p;
    int tl[2], br[2];
    orc_projector_setup(&p, scale, K, R);
    detect_roi(&p, kind, w, h, tl, br);
    int mw = br[0] - tl[0] + 1;
    for (int v = tl[1]; v <= br[1]; ++v)
        for (int u = tl[0]; u <= br[0]; ++u) {
            float x, y;
            map_backward(&p, kind, (float)u, (float)v, &x, &y);
            xmap[(size_t)(v - tl[1]) * mw + (u - tl[0])] = x;
            ymap[(size_t)(v - tl[1]) * mw + (u - tl[0])] = y;
        }
}

/* ---- A.3 cv::remap, 8-bit ----------------------------------------------------------- */
static inline int border_reflect(int i, int n)
{ /* BORDER_REFLECT: fedcba|abcdefgh|hgfedcb */
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i - 1;
        else i = 2 * n - 1 - i;
    }
    return i;
}
static inline int border_reflect101(int i, int n)
{ /* gfedcb|abcdefgh|gfedcba */
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}
static inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
static inline int sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
/* cvRound(float) on x86 = cvtss2si: round-half-even, INT_MIN on overflow/NaN */
static inline int cv_round_f(float v)
{
    if (!(v > -2147483648.f && v < 2147483648.f)) return INT_MIN;
    return (int)lrintf(v);
}

/* border: 0 = BORDER_CONSTANT(0), 1 = BORDER_REFLECT */
ORC_API void orc_remap_linear_8u(const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* xmap,
                                 const float* ymap, int dw, int dh, uint8_t* dst, size_t dpitch, int border)
{
    for (int dy = 0; dy < dh; ++dy)
        for (int dx = 0; dx < dw; ++dx) {
            float fx = xmap[(size_t)dy * dw + dx], fy = ymap[(size_t)dy * dw + dx];
            int sx = cv_round_f(fx * 32.f), sy = cv_round_f(fy * 32.f);
            int x0 = sat_short(sx >> 5), y0 = sat_short(sy >> 5);
            int a = sx & 31, b = sy & 31;
            int w00 = (32 - a) * (32 - b) * 32, w01 = a * (32 - b) * 32, w10 = (32 - a) * b * 32, w11 = a * b * 32;
            for (int c = 0; c < ch; ++c) {
                int p[4];
                for (int k = 0; k < 4; ++k) {
                    int xx = x0 + (k & 1), yy = y0 + (k >> 1);
                    if (border == 1) {
                        xx = border_reflect(xx, sw);
                        yy = border_reflect(yy, sh);
                        p[k] = src[(size_t)yy * spitch + (size_t)xx * ch + c];
                    } else
                        p[k] = (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * spitch + (size_t)xx * ch + c] : 0;
                }
                int v = (p[0] * w00 + p[1] * w01 + p[2] * w10 + p[3] * w11 + (1 << 14)) >> 15;
                dst[(size_t)dy * dpitch + (size_t)dx * ch + c] = (uint8_t)sat_u8(v);
            }
        }
}

ORC_API void orc_remap_nearest_8u(const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* xmap,
                                  const float* ymap, int dw, int dh, uint8_t* dst, size_t dpitch, int border)
{
    for (int dy = 0; dy < dh; ++dy)
        for (int dx = 0; dx < dw; ++dx) {
            int ix = sat_short(cv_round_f(xmap[(size_t)dy * dw + dx]));
            int iy = sat_short(cv_round_f(ymap[(size_t)dy * dw + dx]));
            for (int c = 0; c < ch; ++c) {
                int v;
                if (border == 1)
                    v = src[(size_t)border_reflect(iy, sh) * spitch + (size_t)border_reflect(ix, sw) * ch + c];
                else
                    v = (ix >= 0 && ix < sw && iy >= 0 && iy < sh) ? src[(size_t)iy * spitch + (size_t)ix * ch + c] : 0;
                dst[(size_t)dy * dpitch + (size_t)dx * ch + c] = (uint8_t)v;
            }
        }
}

/* RotationWarper::warp: dst is rect_h x rect_w (rect = orc_warp_roi); returns corner = rect.tl */
ORC_API void orc_warp(int kind, float scale, const uint8_t* src, int sw, int sh, int ch, size_t spitch,
                      const float* K, const float* R, int interp, int border, uint8_t* dst, size_t dpitch,
                      int* corner_xy)
{
    int rect[4];
    orc_warp_roi(kind, scale, sw, sh, K, R, rect);
    size_t n = (size_t)rect[2] * rect[3];
    float* xmap = (float*)malloc(n * sizeof(float));
    float* ymap = (float*)malloc(n * sizeof(float));
    orc_build_maps(kind, scale, sw, sh, K, R, xmap, ymap);
    if (interp == ORC_LINEAR)
        orc_remap_linear_8u(src, sw, sh, ch, spitch, xmap, ymap, rect[2], rect[3], dst, dpitch, border);
    else
        orc_remap_nearest_8u(src, sw, sh, ch, spitch, xmap, ymap, rect[2], rect[3], dst, dpitch, border);
    corner_xy[0] = rect[0];
    corner_xy[1] = rect[1];
    free(xmap);
    free(ymap);
}

/* RotationWarperBase::warpBackward: forward map of every pixel of the original frame (libm atan2f / acosf), minus the warped
 * ROI's top-left, then cv::remap of the warped image.  Returns -1 when src is not warpRoi(dst size) large (CV_Assert). */
ORC_API int orc_warp_backward(int kind, float scale, const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* K,
                              const float* R, int interp, int border, int dw, int dh, uint8_t* dst, size_t dpitch)
{
    orc_projector p;
    int tl[2], br[2];
    orc_projector_setup(&p, scale, K, R);
    detect_roi(&p, kind, dw, dh, tl, br);
    if (br[0] - tl[0] + 1 != sw || br[1] - tl[1] + 1 != sh) return -1;
    size_t n = (size_t)dw * dh;
    float* xmap = (float*)malloc(n * sizeof(float));
    float* ymap = (float*)malloc(n * sizeof(float));
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            float u, v;
            map_forward(&p, kind, (float)x, (float)y, &u, &v);
            xmap[(size_t)y * dw + x] = u - (float)tl[0];
            ymap[(size_t)y * dw + x] = v - (float)tl[1];
        }
    if (interp == ORC_LINEAR) orc_remap_linear_8u(src, sw, sh, ch, spitch, xmap, ymap, dw, dh, dst, dpitch, border);
    else orc_remap_nearest_8u(src, sw, sh, ch, spitch, xmap, ymap, dw, dh, dst, dpitch, border);
    free(xmap);
    free(ymap);
    return 0;
}

/* ---- A.4 dilate 3x3 + resize INTER_LINEAR_EXACT (8UC1) ------------------------------ */
ORC_API void orc_dilate3x3_8u(const uint8_t* src, int w, int h, uint8_t* dst)
{ /* cv::dilate(src, dst, Mat()): 3x3 rect max; the default border never wins a max */
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int m = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                    int v = src[(size_t)yy * w + xx];
                    if (v > m) m = v;
                }
            dst[(size_t)y * w + x] = (uint8_t)m;
        }
}

/* inv_scale = fx (the fx/fy form of cv::resize) or 0 for the dsize form, where OpenCV sets inv_scale = dn / sn;
 * either way the bit-exact resize works with scale = 1 / inv_scale (resize.cpp, interpolationLinear). */
static void linear_exact_coeffs(int sn, int dn, double inv_scale, int* ofs, int* alpha)
{
    if (inv_scale <= 0) inv_scale = (double)dn / sn;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dn; ++d) {
        double f = (d + 0.5) * scale - 0.5;
        int s = (int)floor(f);
        int a = (int)lrint((f - s) * 256.0);
        if (s < 0) { s = 0; a = 0; }
        if (s >= sn - 1) { s = sn - 1; a = 0; }
        ofs[d] = s;
        alpha[d] = a;
    }
}

/* cv::resize(INTER_LINEAR_EXACT) on 8U with `ch` interleaved channels; fx = fy = 0 selects the dsize form */
ORC_API void orc_resize_linear_exact_8u_ex(const uint8_t* src, int sw, int sh, int ch, uint8_t* dst, int dw, int dh,
                                           double fx, double fy)
{
    int* xo = (int*)malloc(sizeof(int) * dw * 2);
    int* xa = xo + dw;
    int* yo = (int*)malloc(sizeof(int) * dh * 2);
    int* ya = yo + dh;
    linear_exact_coeffs(sw, dw, fx, xo, xa);
    linear_exact_coeffs(sh, dh, fy, yo, ya);
    for (int y = 0; y < dh; ++y) {
        int s0 = yo[y], s1 = s0 + 1 < sh ? s0 + 1 : s0;
        for (int x = 0; x < dw; ++x) {
            int c0 = xo[x], c1 = c0 + 1 < sw ? c0 + 1 : c0;
            for (int c = 0; c < ch; ++c) {
                int r0 = src[((size_t)s0 * sw + c0) * ch + c] * (256 - xa[x]) + src[((size_t)s0 * sw + c1) * ch + c] * xa[x];
                int r1 = src[((size_t)s1 * sw + c0) * ch + c] * (256 - xa[x]) + src[((size_t)s1 * sw + c1) * ch + c] * xa[x];
                dst[((size_t)y * dw + x) * ch + c] = (uint8_t)((r0 * (256 - ya[y]) + r1 * ya[y] + 32768) >> 16);
            }
        }
    }
    free(xo);
    free(yo);
}

ORC_API void orc_resize_linear_exact_8u(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh)
{
    orc_resize_linear_exact_8u_ex(src, sw, sh, 1, dst, dw, dh, 0, 0);
}

/* cv::rotate: code 0 = ROTATE_90_CLOCKWISE, 1 = ROTATE_180 (image_stitching.cpp:1093-1103).  dst is h x w for code 0. */
ORC_API void orc_rotate_8u(const uint8_t* src, int w, int h, int ch, int code, uint8_t* dst)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < ch; ++c) {
                uint8_t v = src[((size_t)y * w + x) * ch + c];
                if (code == 0) dst[((size_t)x * h + (h - 1 - y)) * ch + c] = v;       /* (x, y) -> (h-1-y, x) */
                else dst[((size_t)(h - 1 - y) * w + (w - 1 - x)) * ch + c] = v;
            }
}

/* ---- A.7 gain map: cv::resize(f32, INTER_LINEAR) + multiply -------------------------- */
ORC_API void orc_resize_linear_f32(const float* src, int sw, int sh, float* dst, int dw, int dh)
{
    double sx_ = (double)sw / dw, sy_ = (double)sh / dh;
    int* xo = (int*)malloc(sizeof(int) * dw);
    float* xa = (float*)malloc(sizeof(float) * dw);
    for (int d = 0; d < dw; ++d) {
        float fx = (float)((d + 0.5) * sx_ - 0.5);
        int s = (int)floorf(fx);
        fx -= s;
        if (s < 0) { fx = 0; s = 0; }
        if (s >= sw - 1) { fx = 0; s = sw - 1; }
        xo[d] = s;
        xa[d] = fx;
    }
    for (int y = 0; y < dh; ++y) {
        float fy = (float)((y + 0.5) * sy_ - 0.5);
        int s = (int)floorf(fy);
        fy -= s;
        int s0 = s < 0 ? 0 : (s > sh - 1 ? sh - 1 : s);
        int s1 = s + 1 < 0 ? 0 : (s + 1 > sh - 1 ? sh - 1 : s + 1);
        float b0 = 1.f - fy, b1 = fy;
        for (int x = 0; x < dw; ++x) {
            int c0 = xo[x], c1 = c0 + 1 < sw ? c0 + 1 : c0;
            float a0 = 1.f - xa[x], a1 = xa[x];
            float r0 = src[s0 * sw + c0] * a0 + src[s0 * sw + c1] * a1;
            float r1 = src[s1 * sw + c0] * a0 + src[s1 * sw + c1] * a1;
            dst[(size_t)y * dw + x] = r0 * b0 + r1 * b1;
        }
    }
    free(xo);
    free(xa);
}

/* BlocksGainCompensator::apply: image(8UC3) = sat_u8(rint(float(p) * G)) */
ORC_API void orc_gain_apply_8uc3(uint8_t* img, int w, int h, size_t pitch, const float* gain, int gw, int gh)
{
    float* G = (float*)malloc(sizeof(float) * (size_t)w * h);
    if (gw == w && gh == h) memcpy(G, gain, sizeof(float) * (size_t)w * h);
    else orc_resize_linear_f32(gain, gw, gh, G, w, h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float g = G[(size_t)y * w + x];
            for (int c = 0; c < 3; ++c) {
                uint8_t* p = img + (size_t)y * pitch + (size_t)x * 3 + c;
                *p = (uint8_t)sat_u8(cv_round_f((float)*p * g));
            }
        }
    free(G);
}

/* ---- A.5 pyramids --------------------------------------------------------------------- */
ORC_API void orc_pyrdown_16s(const int16_t* src, int w, int h, int ch, int16_t* dst)
{
    int ow = (w + 1) / 2, oh = (h + 1) / 2;
    static const int k5[5] = {1, 4, 6, 4, 1};
    for (int y = 0; y < oh; ++y)
        for (int x = 0; x < ow; ++x)
            for (int c = 0; c < ch; ++c) {
                int s = 0;
                for (int j = 0; j < 5; ++j) {
                    int yy = border_reflect101(2 * y - 2 + j, h);
                    int r = 0;
                    for (int i = 0; i < 5; ++i) {
                        int xx = border_reflect101(2 * x - 2 + i, w);
                        r += k5[i] * src[((size_t)yy * w + xx) * ch + c];
                    }
                    s += k5[j] * r;
                }
                dst[((size_t)y * ow + x) * ch + c] = (int16_t)sat_short((s + 128) >> 8);
            }
}

/* pyrUp to exactly (2w x 2h) */
ORC_API void orc_pyrup_16s(const int16_t* src, int w, int h, int ch, int16_t* dst)
{
    int ow = 2 * w;
    int* rows = (int*)malloc(sizeof(int) * (size_t)ow * ch * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < ch; ++c) {
                const int16_t* s = src + (size_t)y * w * ch + c;
                int xm = x == 0 ? (w > 1 ? 1 : 0) : x - 1; /* s[-1] := s[1] */
                int xp = x == w - 1 ? w - 1 : x + 1;       /* s[n]  := s[n-1] */
                rows[((size_t)y * ow + 2 * x) * ch + c] = s[xm * ch] + 6 * s[x * ch] + s[xp * ch];
                rows[((size_t)y * ow + 2 * x + 1) * ch + c] = 4 * (s[x * ch] + s[xp * ch]);
            }
    for (int y = 0; y < h; ++y) {
        int ym = y == 0 ? (h > 1 ? 1 : 0) : y - 1;
        int yp = y == h - 1 ? h - 1 : y + 1;
        for (int i = 0; i < ow * ch; ++i) {
            int r0 = rows[(size_t)ym * ow * ch + i], r1 = rows[(size_t)y * ow * ch + i], r2 = rows[(size_t)yp * ow * ch + i];
            dst[(size_t)(2 * y) * ow * ch + i] = (int16_t)sat_short((r0 + 6 * r1 + r2 + 32) >> 6);
            dst[(size_t)(2 * y + 1) * ow * ch + i] = (int16_t)sat_short((4 * (r1 + r2) + 32) >> 6);
        }
    }
    free(rows);
}

/* pyrDown 32FC1 with OpenCV's operation order (SSE-baseline universal intrinsics, 4 lanes) */
ORC_API void orc_pyrdown_32f(const float* src, int w, int h, float* dst)
{
    int ow = (w + 1) / 2, oh = (h + 1) / 2;
    int width0 = (w - 3) / 2 + 1;
    if (width0 > ow) width0 = ow;
    int simd_h_end = width0 >= 1 ? 1 + 4 * ((width0 - 1) / 4) : 0; /* columns [1, simd_h_end) use the SIMD order */
    int simd_v_end = 4 * (ow / 4);
    float* hrows = (float*)malloc(sizeof(float) * (size_t)ow * h);
    for (int y = 0; y < h; ++y) {
        const float* s = src + (size_t)y * w;
        float* r = hrows + (size_t)y * ow;
        for (int x = 0; x < ow; ++x) {
            float r0 = s[border_reflect101(2 * x - 2, w)], r1 = s[border_reflect101(2 * x - 1, w)],
                  r2 = s[border_reflect101(2 * x, w)], r3 = s[border_reflect101(2 * x + 1, w)],
                  r4 = s[border_reflect101(2 * x + 2, w)];
            if (x >= 1 && x < simd_h_end) {
                float t = (r1 + r3) * 4.f;
                float q = r0 + r4;
                t = t + q;
                r[x] = r2 * 6.f + t;
            } else {
                float t = r2 * 6.f;
                float q = (r1 + r3) * 4.f;
                t = t + q;
                t = t + r0;
                r[x] = t + r4;
            }
        }
    }
    for (int y = 0; y < oh; ++y) {
        const float* p0 = hrows + (size_t)border_reflect101(2 * y - 2, h) * ow;
        const float* p1 = hrows + (size_t)border_reflect101(2 * y - 1, h) * ow;
        const float* p2 = hrows + (size_t)border_reflect101(2 * y, h) * ow;
        const float* p3 = hrows + (size_t)border_reflect101(2 * y + 1, h) * ow;
        const float* p4 = hrows + (size_t)border_reflect101(2 * y + 2, h) * ow;
        for (int x = 0; x < ow; ++x) {
            float v;
            if (x < simd_v_end) {
                float a = (p1[x] + p3[x]) + p2[x];
                float b = (p0[x] + p4[x]) + (p2[x] + p2[x]);
                a = a * 4.f;
                v = a + b;
            } else {
                float t = p2[x] * 6.f;
                float q = (p1[x] + p3[x]) * 4.f;
                t = t + q;
                t = t + p0[x];
                v = t + p4[x];
            }
            dst[(size_t)y * ow + x] = v * (1.f / 256.f);
        }
    }
    free(hrows);
}

/* ---- A.6 MultiBandBlender ------------------------------------------------------------ */
#define ORC_MAX_LEVELS 24
typedef struct {
    int nb_requested, nb;
    int roi[4];       /* dst_roi_ (padded) x,y,w,h */
    int roi_final[4]; /* dst_roi_final_ */
    int lw[ORC_MAX_LEVELS], lh[ORC_MAX_LEVELS];
    int16_t* lap[ORC_MAX_LEVELS];
    float* wt[ORC_MAX_LEVELS];
} orc_blender;

ORC_API void orc_result_roi(const int* corners_xy, const int* sizes_wh, int n, int* rect_xywh)
{ /* cv::detail::resultRoi (image_stitching.cpp:1176) */
    int tlx = INT_MAX, tly = INT_MAX, brx = INT_MIN, bry = INT_MIN;
    for (int i = 0; i < n; ++i) {
        if (corners_xy[2 * i] < tlx) tlx = corners_xy[2 * i];
        if (corners_xy[2 * i + 1] < tly) tly = corners_xy[2 * i + 1];
        if (corners_xy[2 * i] + sizes_wh[2 * i] > brx) brx = corners_xy[2 * i] + sizes_wh[2 * i];
        if (corners_xy[2 * i + 1] + sizes_wh[2 * i + 1] > bry) bry = corners_xy[2 * i + 1] + sizes_wh[2 * i + 1];
    }
    rect_xywh[0] = tlx; rect_xywh[1] = tly; rect_xywh[2] = brx - tlx; rect_xywh[3] = bry - tly;
}

ORC_API orc_blender* orc_blender_create(int num_bands)
{
    orc_blender* b = (orc_blender*)calloc(1, sizeof(orc_blender));
    b->nb_requested = num_bands;
    return b;
}

static void blender_free_pyr(orc_blender* b)
{
    for (int i = 0; i < ORC_MAX_LEVELS; ++i) {
        free(b->lap[i]); b->lap[i] = NULL;
        free(b->wt[i]); b->wt[i] = NULL;
    }
}

ORC_API void orc_blender_destroy(orc_blender* b)
{
    if (!b) return;
    blender_free_pyr(b);
    free(b);
}

/* MultiBandBlender::prepare(Rect) */
ORC_API void orc_blender_prepare(orc_blender* b, const int* roi_xywh)
{
    blender_free_pyr(b);
    memcpy(b->roi_final, roi_xywh, 4 * sizeof(int));
    memcpy(b->roi, roi_xywh, 4 * sizeof(int));
    double max_len = (double)(roi_xywh[2] > roi_xywh[3] ? roi_xywh[2] : roi_xywh[3]);
    int lim = (int)ceil(log(max_len) / log(2.0));
    b->nb = b->nb_requested < lim ? b->nb_requested : lim;
    int m = 1 << b->nb;
    b->roi[2] += (m - b->roi[2] % m) % m;
    b->roi[3] += (m - b->roi[3] % m) % m;
    b->lw[0] = b->roi[2]; b->lh[0] = b->roi[3];
    for (int i = 0; i <= b->nb; ++i) {
        if (i > 0) { b->lw[i] = (b->lw[i - 1] + 1) / 2; b->lh[i] = (b->lh[i - 1] + 1) / 2; }
        b->lap[i] = (int16_t*)calloc((size_t)b->lw[i] * b->lh[i] * 3, sizeof(int16_t));
        b->wt[i] = (float*)calloc((size_t)b->lw[i] * b->lh[i], sizeof(float));
    }
}

ORC_API int orc_blender_num_bands(const orc_blender* b) { return b->nb; }
ORC_API void orc_blender_get_rois(const orc_blender* b, int* roi_padded, int* roi_final)
{
    memcpy(roi_padded, b->roi, 4 * sizeof(int));
    memcpy(roi_final, b->roi_final, 4 * sizeof(int));
}

/* the tile rect feed() works on; also exported so the product's planner can be checked */
ORC_API void orc_blender_tile_rect(const orc_blender* b, int w, int h, int tlx, int tly, int* tl_new, int* br_new)
{
    int nb = b->nb, gap = 3 * (1 << nb), m = 1 << nb;
    int rx = b->roi[0], ry = b->roi[1], rbx = rx + b->roi[2], rby = ry + b->roi[3];
    int tx = tlx - gap > rx ? tlx - gap : rx, ty = tly - gap > ry ? tly - gap : ry;
    int bx = tlx + w + gap < rbx ? tlx + w + gap : rbx, by = tly + h + gap < rby ? tly + h + gap : rby;
    tx = rx + (((tx - rx) >> nb) << nb);
    ty = ry + (((ty - ry) >> nb) << nb);
    int width = bx - tx, height = by - ty;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    bx = tx + width; by = ty + height;
    int dy = by - rby > 0 ? by - rby : 0, dx = bx - rbx > 0 ? bx - rbx : 0;
    tx -= dx; bx -= dx; ty -= dy; by -= dy;
    tl_new[0] = tx; tl_new[1] = ty; br_new[0] = bx; br_new[1] = by;
}

/* MultiBandBlender::feed(img 16SC3, mask 8U, tl) */
ORC_API void orc_blender_feed(orc_blender* b, const int16_t* img, const uint8_t* mask, int w, int h, int tlx, int tly)
{
    int nb = b->nb, tl_new[2], br_new[2];
    orc_blender_tile_rect(b, w, h, tlx, tly, tl_new, br_new);
    int top = tly - tl_new[1], left = tlx - tl_new[0];
    int tw = br_new[0] - tl_new[0], th = br_new[1] - tl_new[1];
    int16_t* gp[ORC_MAX_LEVELS];
    float* wp[ORC_MAX_LEVELS];
    int pw[ORC_MAX_LEVELS], ph[ORC_MAX_LEVELS];
    pw[0] = tw; ph[0] = th;
    gp[0] = (int16_t*)malloc(sizeof(int16_t) * 3 * (size_t)tw * th);
    wp[0] = (float*)calloc((size_t)tw * th, sizeof(float));
    /* copyMakeBorder(BORDER_REFLECT) for the image, BORDER_CONSTANT(0) for the weight */
    const float inv255 = (float)(1. / 255.);
    for (int y = 0; y < th; ++y) {
        int sy = border_reflect(y - top, h);
        for (int x = 0; x < tw; ++x) {
            int sx = border_reflect(x - left, w);
            for (int c = 0; c < 3; ++c) gp[0][((size_t)y * tw + x) * 3 + c] = img[((size_t)sy * w + sx) * 3 + c];
            if (y - top >= 0 && y - top < h && x - left >= 0 && x - left < w)
                wp[0][(size_t)y * tw + x] = (float)mask[(size_t)(y - top) * w + (x - left)] * inv255;
        }
    }
    for (int i = 0; i < nb; ++i) {
        pw[i + 1] = (pw[i] + 1) / 2; ph[i + 1] = (ph[i] + 1) / 2;
        gp[i + 1] = (int16_t*)malloc(sizeof(int16_t) * 3 * (size_t)pw[i + 1] * ph[i + 1]);
        wp[i + 1] = (float*)malloc(sizeof(float) * (size_t)pw[i + 1] * ph[i + 1]);
        orc_pyrdown_16s(gp[i], pw[i], ph[i], 3, gp[i + 1]);
        orc_pyrdown_32f(wp[i], pw[i], ph[i], wp[i + 1]);
    }
    /* createLaplacePyr: pyr[i] = sat(pyr[i] - pyrUp(pyr[i+1])) */
    for (int i = 0; i < nb; ++i) {
        int16_t* up = (int16_t*)malloc(sizeof(int16_t) * 3 * (size_t)pw[i] * ph[i]);
        orc_pyrup_16s(gp[i + 1], pw[i + 1], ph[i + 1], 3, up);
        size_t cnt = (size_t)pw[i] * ph[i] * 3;
        for (size_t k = 0; k < cnt; ++k) gp[i][k] = (int16_t)sat_short((int)gp[i][k] - (int)up[k]);
        free(up);
    }
    int y_tl = tl_new[1] - b->roi[1], y_br = br_new[1] - b->roi[1];
    int x_tl = tl_new[0] - b->roi[0], x_br = br_new[0] - b->roi[0];
    for (int i = 0; i <= nb; ++i) {
        int rw = x_br - x_tl, rh = y_br - y_tl;
        for (int y = 0; y < rh; ++y)
            for (int x = 0; x < rw; ++x) {
                float wv = wp[i][(size_t)y * pw[i] + x];
                size_t d = (size_t)(y + y_tl) * b->lw[i] + (x + x_tl);
                for (int c = 0; c < 3; ++c) {
                    int16_t add = (int16_t)(int)((float)gp[i][((size_t)y * pw[i] + x) * 3 + c] * wv);
                    b->lap[i][d * 3 + c] = (int16_t)(b->lap[i][d * 3 + c] + add);
                }
                b->wt[i][d] = b->wt[i][d] + wv;
            }
        x_tl /= 2; y_tl /= 2; x_br /= 2; y_br /= 2;
    }
    for (int i = 0; i <= nb; ++i) { free(gp[i]); free(wp[i]); }
}

/* MultiBandBlender::blend: dst 16SC3 + dst_mask 8U of size roi_final (w x h) */
ORC_API void orc_blender_blend(orc_blender* b, int16_t* dst, uint8_t* dst_mask)
{
    const float WEIGHT_EPS = 1e-5f;
    int nb = b->nb;
    for (int i = 0; i <= nb; ++i) {
        size_t n = (size_t)b->lw[i] * b->lh[i];
        for (size_t k = 0; k < n; ++k) {
            float den = b->wt[i][k] + WEIGHT_EPS;
            for (int c = 0; c < 3; ++c) b->lap[i][k * 3 + c] = (int16_t)(int)((float)b->lap[i][k * 3 + c] / den);
        }
    }
    for (int i = nb; i > 0; --i) {
        int16_t* up = (int16_t*)malloc(sizeof(int16_t) * 3 * (size_t)b->lw[i - 1] * b->lh[i - 1]);
        orc_pyrup_16s(b->lap[i], b->lw[i], b->lh[i], 3, up);
        size_t cnt = (size_t)b->lw[i - 1] * b->lh[i - 1] * 3;
        for (size_t k = 0; k < cnt; ++k) b->lap[i - 1][k] = (int16_t)sat_short((int)up[k] + (int)b->lap[i - 1][k]);
        free(up);
    }
    int fw = b->roi_final[2], fh = b->roi_final[3];
    for (int y = 0; y < fh; ++y)
        for (int x = 0; x < fw; ++x) {
            size_t s = (size_t)y * b->lw[0] + x, d = (size_t)y * fw + x;
            int on = b->wt[0][s] > WEIGHT_EPS;
            dst_mask[d] = on ? 255 : 0;
            for (int c = 0; c < 3; ++c) dst[d * 3 + c] = on ? b->lap[0][s * 3 + c] : 0;
        }
    blender_free_pyr(b);
}

/* ---- the whole compositing loop (image_stitching.cpp:1086-1229) ----------------------- */
/* imgs[i]: 8UC3 tightly packed w_i x h_i ; gains[i]: gh x gw f32 or NULL ; seam[i]: 8UC1 or NULL.
 * out8 (roi_w*roi_h*3) and out_mask (roi_w*roi_h) must be sized from orc_compose_roi(). */
ORC_API void orc_compose_roi(int kind, float scale, int n, const int* wh, const float* Ks, const float* Rs,
                             int* corners_xy, int* sizes_wh, int* dst_roi)
{
    for (int i = 0; i < n; ++i) {
        int r[4];
        orc_warp_roi(kind, scale, wh[2 * i], wh[2 * i + 1], Ks + 9 * i, Rs + 9 * i, r);
        corners_xy[2 * i] = r[0]; corners_xy[2 * i + 1] = r[1];
        sizes_wh[2 * i] = r[2]; sizes_wh[2 * i + 1] = r[3];
    }
    orc_result_roi(corners_xy, sizes_wh, n, dst_roi);
}

ORC_API void orc_compose(int kind, float scale, int n, const uint8_t* const* imgs, const int* wh, const float* Ks,
                         const float* Rs, const float* const* gains, const int* gain_wh,
                         const uint8_t* const* seam, const int* seam_wh, int num_bands, int16_t* out16,
                         uint8_t* out8, uint8_t* out_mask)
{
    int* corners = (int*)malloc(sizeof(int) * 4 * n);
    int* sizes = corners + 2 * n;
    int roi[4];
    orc_compose_roi(kind, scale, n, wh, Ks, Rs, corners, sizes, roi);
    orc_blender* b = orc_blender_create(num_bands);
    orc_blender_prepare(b, roi);
    for (int i = 0; i < n; ++i) {
        int w = sizes[2 * i], h = sizes[2 * i + 1], sw = wh[2 * i], sh = wh[2 * i + 1], c[2];
        uint8_t* iw = (uint8_t*)malloc((size_t)w * h * 3);
        uint8_t* mw = (uint8_t*)malloc((size_t)w * h);
        uint8_t* ones = (uint8_t*)malloc((size_t)sw * sh);
        memset(ones, 255, (size_t)sw * sh);
        orc_warp(kind, scale, imgs[i], sw, sh, 3, (size_t)sw * 3, Ks + 9 * i, Rs + 9 * i, ORC_LINEAR, 1, iw,
                 (size_t)w * 3, c);
        orc_warp(kind, scale, ones, sw, sh, 1, (size_t)sw, Ks + 9 * i, Rs + 9 * i, ORC_NEAREST, 0, mw, (size_t)w, c);
        free(ones);
        if (gains && gains[i]) orc_gain_apply_8uc3(iw, w, h, (size_t)w * 3, gains[i], gain_wh[2 * i], gain_wh[2 * i + 1]);
        int16_t* is = (int16_t*)malloc(sizeof(int16_t) * (size_t)w * h * 3);
        for (size_t k = 0; k < (size_t)w * h * 3; ++k) is[k] = iw[k];
        if (seam && seam[i]) {
            int mw_ = seam_wh[2 * i], mh_ = seam_wh[2 * i + 1];
            uint8_t* dil = (uint8_t*)malloc((size_t)mw_ * mh_);
            uint8_t* up = (uint8_t*)malloc((size_t)w * h);
            orc_dilate3x3_8u(seam[i], mw_, mh_, dil);
            orc_resize_linear_exact_8u(dil, mw_, mh_, up, w, h);
            for (size_t k = 0; k < (size_t)w * h; ++k) mw[k] &= up[k];
            free(dil);
            free(up);
        }
        orc_blender_feed(b, is, mw, w, h, corners[2 * i], corners[2 * i + 1]);
        free(iw); free(mw); free(is);
    }
    size_t npx = (size_t)roi[2] * roi[3];
    int16_t* r16 = out16 ? out16 : (int16_t*)malloc(sizeof(int16_t) * npx * 3);
    orc_blender_blend(b, r16, out_mask);
    if (out8)
        for (size_t k = 0; k < npx * 3; ++k) out8[k] = (uint8_t)sat_u8(r16[k]);
    if (!out16) free(r16);
    orc_blender_destroy(b);
    free(corners);
}

/* ---- Blender::NO and FeatherBlender (image_stitching.cpp:1179, 1186-1191) ------------ */
/* cv::detail::createWeightMap: distanceTransform(mask, DIST_L1, 3) * sharpness, truncated at 1.
 * The 3x3 L1 chamfer is the exact city-block distance to the nearest zero pixel of the mask; with no zero pixel
 * the non-IPP build yields 65534 (the IPP build FLT_MAX) - both clamp to 1 for any sharpness >= 1.6e-5. */
ORC_API void orc_create_weight_map(const uint8_t* mask, int w, int h, float sharpness, float* weight)
{
    const int INF = 1 << 28;
    int* d = (int*)malloc(sizeof(int) * (size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) d[i] = mask[i] ? INF : 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int* p = d + (size_t)y * w + x;
            if (x > 0 && p[-1] + 1 < *p) *p = p[-1] + 1;
            if (y > 0 && p[-w] + 1 < *p) *p = p[-w] + 1;
        }
    for (int y = h - 1; y >= 0; --y)
        for (int x = w - 1; x >= 0; --x) {
            int* p = d + (size_t)y * w + x;
            if (x < w - 1 && p[1] + 1 < *p) *p = p[1] + 1;
            if (y < h - 1 && p[w] + 1 < *p) *p = p[w] + 1;
        }
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        float dist = d[i] >= INF ? 65534.f : (float)d[i];
        float v = dist * sharpness;
        weight[i] = v > 1.f ? 1.f : v;
    }
    free(d);
}

typedef struct {
    int type; /* 0 = Blender::NO, 1 = FeatherBlender */
    float sharpness;
    int roi[4];
    int16_t* dst;
    uint8_t* dst_mask;
    float* dst_weight;
} orc_simple_blender;

ORC_API orc_simple_blender* orc_simple_blender_create(int type, float sharpness)
{
    orc_simple_blender* b = (orc_simple_blender*)calloc(1, sizeof(orc_simple_blender));
    b->type = type;
    b->sharpness = sharpness;
    return b;
}
ORC_API void orc_simple_blender_destroy(orc_simple_blender* b)
{
    if (!b) return;
    free(b->dst); free(b->dst_mask); free(b->dst_weight); free(b);
}
ORC_API void orc_simple_blender_prepare(orc_simple_blender* b, const int* roi)
{
    free(b->dst); free(b->dst_mask); free(b->dst_weight);
    memcpy(b->roi, roi, 4 * sizeof(int));
    size_t n = (size_t)roi[2] * roi[3];
    b->dst = (int16_t*)calloc(n * 3, sizeof(int16_t));
    b->dst_mask = (uint8_t*)calloc(n, 1);
    b->dst_weight = (float*)calloc(n, sizeof(float));
}
ORC_API void orc_simple_blender_feed(orc_simple_blender* b, const int16_t* img, const uint8_t* mask, int w, int h, int tlx, int tly)
{
    int dx = tlx - b->roi[0], dy = tly - b->roi[1], dw = b->roi[2];
    if (b->type == 0) { /* Blender::feed */
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                size_t s = (size_t)y * w + x, d = (size_t)(dy + y) * dw + dx + x;
                if (mask[s])
                    for (int c = 0; c < 3; ++c) b->dst[d * 3 + c] = img[s * 3 + c];
                b->dst_mask[d] |= mask[s];
            }
        return;
    }
    float* wm = (float*)malloc(sizeof(float) * (size_t)w * h);
    orc_create_weight_map(mask, w, h, b->sharpness, wm);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            size_t s = (size_t)y * w + x, d = (size_t)(dy + y) * dw + dx + x;
            for (int c = 0; c < 3; ++c)
                b->dst[d * 3 + c] = (int16_t)(b->dst[d * 3 + c] + (int16_t)(int)((float)img[s * 3 + c] * wm[s]));
            b->dst_weight[d] = b->dst_weight[d] + wm[s];
        }
    free(wm);
}
ORC_API void orc_simple_blender_blend(orc_simple_blender* b, int16_t* dst, uint8_t* dst_mask)
{
    size_t n = (size_t)b->roi[2] * b->roi[3];
    for (size_t i = 0; i < n; ++i) {
        if (b->type == 1) {
            float den = b->dst_weight[i] + 1e-5f;
            for (int c = 0; c < 3; ++c) b->dst[i * 3 + c] = (int16_t)(int)((float)b->dst[i * 3 + c] / den);
            b->dst_mask[i] = b->dst_weight[i] > 1e-5f ? 255 : 0;
        }
        dst_mask[i] = b->dst_mask[i];
        for (int c = 0; c < 3; ++c) dst[i * 3 + c] = b->dst_mask[i] ? b->dst[i * 3 + c] : 0;
    }
}
