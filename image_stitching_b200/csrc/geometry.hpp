// geometry.hpp - host-side integer/float geometry of the compositing path: projector setup, ROI,
// separable trig tables, resultRoi, MultiBandBlender tile arithmetic and the strip planner.
//
// Everything here must be BIT-EXACT with OpenCV's cv::detail::RotationWarperBase<P> /
// MultiBandBlender (driven by image_stitching.cpp:1116-1141, 1173-1193), so this translation
// unit is compiled with -ffp-contract=off and uses glibc sinf/cosf/atan2f/acosf on the host.
#pragma once
#include <cstdint>
#include <vector>

namespace isb {

struct Rect { int x = 0, y = 0, w = 0, h = 0; };

// ProjectorBase::setCameraParams
struct Projector {
    int kind = 0;  // ISB_WARP_SPHERICAL / ISB_WARP_CYLINDRICAL
    float scale = 1.f;
    float k[9], rinv[9], r_kinv[9], k_rinv[9];

    void set(int kind, float scale, const float K[9], const float R[9]);
    void forward(float x, float y, float& u, float& v) const;   // mapForward
    void backward(float u, float v, float& x, float& y) const;  // mapBackward
    // detectResultRoi: tl, br inclusive (buildMaps' Rect(tl, br)); warpRoi is Rect(tl, br + 1)
    void detect_roi(int src_w, int src_h, int tl[2], int br[2]) const;
    Rect warp_roi(int src_w, int src_h) const;
};

// Separable inverse-map tables: the map's transcendentals depend on one
// coordinate each, so the device needs only O(w + h) host-computed values per image:
//   col[i] = (sinf(u'), cosf(u'))                     u' = float(roi.x + i) / scale
//   row[j] = (sinf(pi - v'), cosf(pi - v'))  spherical v' = float(roi.y + j) / scale
//          = (1, v')                         cylindrical
// and per pixel  x_ = row.a * col.s ; y_ = row.b ; z_ = row.a * col.c.
struct Float2 { float a, b; };
void build_trig_tables(const Projector& p, const Rect& roi, std::vector<Float2>& col, std::vector<Float2>& row);

// cv::resize(INTER_LINEAR_EXACT) 8U coefficients per destination index: (ofs << 16) | alpha(0..256)
void build_linear_exact_table(int src_n, int dst_n, std::vector<uint32_t>& tab, double inv_scale = 0.0);
// cv::resize(INTER_LINEAR) f32 coefficients per destination index: ofs and fraction.
// `clamp_frac`: horizontal pass zeroes the fraction at the borders, the vertical pass does not.
struct LinCoef { int ofs; float frac; };
void build_linear_f32_table(int src_n, int dst_n, bool horizontal, std::vector<LinCoef>& tab);

Rect result_roi(const int* corners_xy, const int* sizes_wh, int n);

// MultiBandBlender::prepare / feed integer arithmetic
struct BlendGeometry {
    int nb = 0;      // actual band count after the clamp
    Rect roi;        // dst_roi_ (padded to 2^nb)
    Rect roi_final;  // dst_roi_final_
    void prepare(const Rect& dst_roi, int requested_bands);
    // rect feed() builds pyramids on (pano coordinates, tl inclusive / br exclusive)
    void tile_rect(int w, int h, int tlx, int tly, int tl_new[2], int br_new[2]) const;
};

// strip i of n over the padded panorama rows, boundaries on the 2^nb grid; rows beyond final_h are
// still owned by the last strip (they are never output).
void strip_rows(int padded_h, int nb, int strip_index, int strip_count, int& y0, int& y1);

}  // namespace isb
