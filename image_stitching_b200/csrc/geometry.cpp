#include "geometry.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>

#include "../../include/image_stitching.h"

namespace isb {

namespace {
const float kPiF = static_cast<float>(3.1415926535897932384626433832795);

// 3x3 product with every entry accumulated sequentially in float32: ((a0*b0)+(a1*b1))+(a2*b2)
void mul3(const float* a, const float* b, float* c)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = a[3 * i] * b[j];
            acc += a[3 * i + 1] * b[3 + j];
            acc += a[3 * i + 2] * b[6 + j];
            c[3 * i + j] = acc;
        }
}

// cv::invert on a 3x3 float32: adjugate / determinant evaluated in double, rounded once to float
void inv3(const float* m, float* out)
{
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0) { std::memset(out, 0, 9 * sizeof(float)); return; }
    det = 1. / det;
    const double adj[9] = {(e * i - f * h) * det, (c * h - b * i) * det, (b * f - c * e) * det,
                           (f * g - d * i) * det, (a * i - c * g) * det, (c * d - a * f) * det,
                           (d * h - e * g) * det, (b * g - a * h) * det, (a * e - b * d) * det};
    for (int k = 0; k < 9; ++k) out[k] = (float)adj[k];
}
}  // namespace

void Projector::set(int kind_, float scale_, const float K[9], const float R[9])
{
    kind = kind_;
    scale = scale_;
    std::memcpy(k, K, sizeof(k));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) rinv[3 * r + c] = R[3 * c + r];
    float kinv[9];
    inv3(K, kinv);
    mul3(R, kinv, r_kinv);
    mul3(K, rinv, k_rinv);
}

void Projector::forward(float x, float y, float& u, float& v) const
{
    const float x_ = r_kinv[0] * x + r_kinv[1] * y + r_kinv[2];
    const float y_ = r_kinv[3] * x + r_kinv[4] * y + r_kinv[5];
    const float z_ = r_kinv[6] * x + r_kinv[7] * y + r_kinv[8];
    u = scale * atan2f(x_, z_);
    if (kind == ISB_WARP_SPHERICAL) {
        const float w = y_ / sqrtf(x_ * x_ + y_ * y_ + z_ * z_);
        v = scale * (kPiF - acosf(w == w ? w : 0));
    } else {
        v = scale * y_ / sqrtf(x_ * x_ + z_ * z_);
    }
}

void Projector::backward(float u, float v, float& x, float& y) const
{
    u /= scale;
    v /= scale;
    float x_, y_, z_;
    if (kind == ISB_WARP_SPHERICAL) {
        const float sinv = sinf(kPiF - v);
        x_ = sinv * sinf(u);
        y_ = cosf(kPiF - v);
        z_ = sinv * cosf(u);
    } else {
        x_ = sinf(u);
        y_ = v;
        z_ = cosf(u);
    }
    x = k_rinv[0] * x_ + k_rinv[1] * y_ + k_rinv[2] * z_;
    y = k_rinv[3] * x_ + k_rinv[4] * y_ + k_rinv[5] * z_;
    const float z = k_rinv[6] * x_ + k_rinv[7] * y_ + k_rinv[8] * z_;
    if (z > 0) { x /= z; y /= z; }
    else x = y = -1;
}

void Projector::detect_roi(int src_w, int src_h, int tl[2], int br[2]) const
{
    float lo_u = std::numeric_limits<float>::max(), lo_v = lo_u, hi_u = -lo_u, hi_v = -lo_u;
    auto take = [&](float px, float py) {
        float u, v;
        forward(px, py, u, v);
        lo_u = std::min(lo_u, u); lo_v = std::min(lo_v, v);
        hi_u = std::max(hi_u, u); hi_v = std::max(hi_v, v);
    };
    for (int x = 0; x < src_w; ++x) { take((float)x, 0.f); take((float)x, (float)(src_h - 1)); }
    for (int y = 0; y < src_h; ++y) { take(0.f, (float)y); take((float)(src_w - 1), (float)y); }
    tl[0] = (int)lo_u; tl[1] = (int)lo_v; br[0] = (int)hi_u; br[1] = (int)hi_v;

    if (kind != ISB_WARP_SPHERICAL) return;
    // the poles: if one projects strictly inside the source image the ROI must reach v = pi*scale (or 0) at u = 0
    float tl_uf = (float)tl[0], tl_vf = (float)tl[1], br_uf = (float)br[0], br_vf = (float)br[1];
    for (int pole = 0; pole < 2; ++pole) {
        const float x = rinv[1], y = pole == 0 ? rinv[4] : -rinv[4], z = rinv[7];
        if (!(y > 0.f)) continue;
        const float x_ = (k[0] * x + k[1] * y) / z + k[2];
        const float y_ = k[4] * y / z + k[5];
        if (x_ > 0.f && x_ < src_w && y_ > 0.f && y_ < src_h) {
            const float pv = pole == 0 ? static_cast<float>(3.1415926535897932384626433832795 * scale) : 0.f;
            tl_uf = std::min(tl_uf, 0.f); tl_vf = std::min(tl_vf, pv);
            br_uf = std::max(br_uf, 0.f); br_vf = std::max(br_vf, pv);
        }
    }
    tl[0] = (int)tl_uf; tl[1] = (int)tl_vf; br[0] = (int)br_uf; br[1] = (int)br_vf;
}

Rect Projector::warp_roi(int src_w, int src_h) const
{
    int tl[2], br[2];
    detect_roi(src_w, src_h, tl, br);
    return Rect{tl[0], tl[1], br[0] - tl[0] + 1, br[1] - tl[1] + 1};
}

void build_trig_tables(const Projector& p, const Rect& roi, std::vector<Float2>& col, std::vector<Float2>& row)
{
    col.resize(roi.w);
    row.resize(roi.h);
    for (int i = 0; i < roi.w; ++i) {
        const float u = (float)(roi.x + i) / p.scale;
        col[i] = Float2{sinf(u), cosf(u)};
    }
    for (int j = 0; j < roi.h; ++j) {
        const float v = (float)(roi.y + j) / p.scale;
        if (p.kind == ISB_WARP_SPHERICAL) row[j] = Float2{sinf(kPiF - v), cosf(kPiF - v)};
        else row[j] = Float2{1.f, v};
    }
}

void build_linear_exact_table(int src_n, int dst_n, std::vector<uint32_t>& tab, double inv_scale)
{
    tab.resize(dst_n);
    // cv::resize: inv_scale = fx when given, else dst/src; the bit-exact path then uses scale = 1 / inv_scale
    if (inv_scale <= 0) inv_scale = (double)dst_n / src_n;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst_n; ++d) {
        const double f = (d + 0.5) * scale - 0.5;
        int s = (int)std::floor(f);
        int a = (int)std::lrint((f - s) * 256.0);
        if (s < 0) { s = 0; a = 0; }
        if (s >= src_n - 1) { s = src_n - 1; a = 0; }
        tab[d] = ((uint32_t)s << 16) | (uint32_t)a;
    }
}

void build_linear_f32_table(int src_n, int dst_n, bool horizontal, std::vector<LinCoef>& tab)
{
    tab.resize(dst_n);
    const double scale = (double)src_n / dst_n;
    for (int d = 0; d < dst_n; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (horizontal) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= src_n - 1) { f = 0; s = src_n - 1; }
        }
        tab[d] = LinCoef{s, f};  // vertical: rows s, s+1 are clamped by the consumer, the fraction is kept
    }
}

Rect result_roi(const int* corners_xy, const int* sizes_wh, int n)
{
    int tlx = INT_MAX, tly = INT_MAX, brx = INT_MIN, bry = INT_MIN;
    for (int i = 0; i < n; ++i) {
        tlx = std::min(tlx, corners_xy[2 * i]);
        tly = std::min(tly, corners_xy[2 * i + 1]);
        brx = std::max(brx, corners_xy[2 * i] + sizes_wh[2 * i]);
        bry = std::max(bry, corners_xy[2 * i + 1] + sizes_wh[2 * i + 1]);
    }
    return Rect{tlx, tly, brx - tlx, bry - tly};
}

void BlendGeometry::prepare(const Rect& dst_roi, int requested_bands)
{
    roi_final = dst_roi;
    roi = dst_roi;
    const double max_len = (double)std::max(dst_roi.w, dst_roi.h);
    nb = std::min(requested_bands, (int)std::ceil(std::log(max_len) / std::log(2.0)));
    const int m = 1 << nb;
    roi.w += (m - roi.w % m) % m;
    roi.h += (m - roi.h % m) % m;
}

void BlendGeometry::tile_rect(int w, int h, int tlx, int tly, int tl_new[2], int br_new[2]) const
{
    const int gap = 3 * (1 << nb), m = 1 << nb;
    const int rbx = roi.x + roi.w, rby = roi.y + roi.h;
    int tx = std::max(roi.x, tlx - gap), ty = std::max(roi.y, tly - gap);
    int bx = std::min(rbx, tlx + w + gap), by = std::min(rby, tly + h + gap);
    tx = roi.x + (((tx - roi.x) >> nb) << nb);
    ty = roi.y + (((ty - roi.y) >> nb) << nb);
    int width = bx - tx, height = by - ty;
    width += (m - width % m) % m;
    height += (m - height % m) % m;
    bx = tx + width;
    by = ty + height;
    const int dy = std::max(by - rby, 0), dx = std::max(bx - rbx, 0);
    tl_new[0] = tx - dx; tl_new[1] = ty - dy;
    br_new[0] = bx - dx; br_new[1] = by - dy;
}

// Strip cuts on the 2^nb grid, balanced by WORK rather than by rows: a strip recomputes a halo of 4 cell rows above and 3
// below (Composer::plan), except at the panorama's own top / bottom.  With c cell rows per interior strip the first strip
// gets c + 4 and the last c + 3 (each saves one halo), so every strip warps about c + 7 cell rows; the remainder goes to
// the leading strips.  Pure arithmetic: every rank computes every rank's rows.
void strip_rows(int padded_h, int nb, int strip_index, int strip_count, int& y0, int& y1)
{
    const int cells = padded_h >> nb;  // padded_h is a multiple of 2^nb
    const int n = strip_count;
    auto cut = [&](int i) -> int {     // first cell row of strip i (i in [0, n])
        if (i <= 0) return 0;
        if (i >= n) return cells;
        if (n == 1) return cells;
        const int c = (cells - 7) / n;  // interior strips
        if (c < 1) return (int)((long long)cells * i / n);  // panorama too short for the scheme: plain equal cuts
        const int rem = (cells - 7) - c * n;
        // strips 0 .. i-1: the first one is 4 taller; `rem` extra rows go one each to the leading strips
        return 4 + c * i + (i < rem ? i : rem);
    };
    y0 = cut(strip_index) << nb;
    y1 = cut(strip_index + 1) << nb;
}

}  // namespace isb
