// capi.cu - the extern "C" boundary declared in include/image_stitching.h.
// Every entry point converts internal exceptions into cv::Error-style codes; nothing throws across it.
#include <cmath>
#include <cstring>
#include <string>

#include "cam_io.hpp"
#include "engine.hpp"
#include "kernels.cuh"
#include "pose_math.hpp"

using namespace isb;

struct isb_warper { Warper impl; isb_warper(int k, float s) : impl(k, s) {} };
struct isb_compensator { Compensator impl; isb_compensator(int w, int h) : impl(w, h) {} };
struct isb_blender { Blender impl; explicit isb_blender(int nb) : impl(nb) {} };
struct isb_simple_blender { SimpleBlender impl; isb_simple_blender(int t, float s) : impl(t, s) {} };
struct isb_timelapser { Timelapser impl; explicit isb_timelapser(int t) : impl(t) {} };
struct isb_composer { ComposerPool impl; explicit isb_composer(const isb_config& c) : impl(c) {} };

static thread_local std::string t_error;

template <typename F>
static int guarded(F&& f)
{
    try {
        f();
        return ISB_OK;
    } catch (const Error& e) {
        t_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        t_error = "host allocation failed";
        return ISB_ERR_NO_MEM;
    } catch (const std::exception& e) {
        t_error = e.what();
        return ISB_ERR_IO;
    } catch (...) {
        t_error = "unknown error";
        return ISB_ERR_IO;
    }
}
#define NOT_NULL(p)                                                            \
    do {                                                                       \
        if (!(p)) throw Error(ISB_ERR_NULL_PTR, std::string(#p) + " is null"); \
    } while (0)

extern "C" {

const char* isb_last_error(void) { return t_error.c_str(); }
const char* isb_version(void) { return "image_stitching_b200 0.1 (sm_100a)"; }
int isb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int isb_set_stream(void* s)
{
    set_current_stream(static_cast<cudaStream_t>(s));
    return ISB_OK;
}
void isb_reload_env(void) { reload_env_switches(); }
long long isb_launch_count(int reset) { return launch_count(reset != 0); }

// ---- cameras / pose math ------------------------------------------------------------------------
void isb_camera_K(const isb_camera* cam, float K[9])
{
    // CameraParams::K() is CV_64F; the loop converts it with convertTo(CV_32F) (image_stitching.cpp:1150-1151)
    const double k[9] = {cam->focal, 0, cam->ppx, 0, cam->focal * cam->aspect, cam->ppy, 0, 0, 1};
    for (int i = 0; i < 9; ++i) K[i] = (float)k[i];
}

void isb_quat_from_rotation_matrix(const double R[9], double q[4])
{
    const Quat<double> v = Quat<double>::from_rotation(R);
    q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
}
void isb_quat_to_rotation_matrix(const double q[4], double R[9])
{
    Quat<double> v;
    v.x = q[0]; v.y = q[1]; v.z = q[2]; v.w = q[3];
    v.to_rotation(R);
}
void isb_quat_from_euler(const double e[3], int order, double q[4])
{
    const Quat<double> v = Quat<double>::from_euler(e[0], e[1], e[2], static_cast<EulerOrder>(order));
    q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
}
void isb_quat_from_axis_angle(const double axis[3], double angle, double q[4])
{
    const Quat<double> v = Quat<double>::from_axis_angle(axis, angle);
    q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w;
}
void isb_quat_multiply(const double a[4], const double b[4], double out[4])
{
    Quat<double> qa, qb;
    qa.x = a[0]; qa.y = a[1]; qa.z = a[2]; qa.w = a[3];
    qb.x = b[0]; qb.y = b[1]; qb.z = b[2]; qb.w = b[3];
    const Quat<double> v = Quat<double>::multiply(qa, qb);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
void isb_quat_slerp(const double a[4], const double b[4], double t, double out[4])
{
    Quat<double> qa, qb;
    qa.x = a[0]; qa.y = a[1]; qa.z = a[2]; qa.w = a[3];
    qb.x = b[0]; qb.y = b[1]; qb.z = b[2]; qb.w = b[3];
    const Quat<double> v = qa.slerp(qb, t);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
void isb_pose_from_cam_transform(const double R_in[9], int is_portrait, double R_out[9])
{
    // image_stitching.cpp:485-517
    const Quat<double> q = Quat<double>::from_rotation(R_in);
    Quat<double> f;
    if (is_portrait) { f.x = q.y; f.y = q.x; f.z = -q.z; f.w = q.w; }
    else { f.x = -q.x; f.y = q.y; f.z = -q.z; f.w = q.w; }
    f.to_rotation(R_out);
}
int isb_rotation_matrix_to_euler(const double R[9], int order, double e[3])
{
    return guarded([&] {
        if (order < 0 || order > 5 || !rotation_to_euler<double>(R, static_cast<EulerOrder>(order), e))
            throw Error(ISB_ERR_BAD_ARG, "unknown euler order");
    });
}
int isb_euler_to_rotation_matrix(const double e[3], int order, double R[9])
{
    return guarded([&] {
        if (order < 0 || order > 5 || !euler_to_rotation<double>(e, static_cast<EulerOrder>(order), R))
            throw Error(ISB_ERR_BAD_ARG, "unknown euler order");
    });
}

// ---- serializer -----------------------------------------------------------------------------------
int isb_parse_matrix_str(const char* s, double* out, int capacity, int* side)
{
    return guarded([&] {
        NOT_NULL(s); NOT_NULL(out); NOT_NULL(side);
        std::vector<double> v;
        if (!parse_matrix_str(s, v, *side)) throw Error(ISB_ERR_IO, "malformed matrix string");
        if ((int)v.size() > capacity) throw Error(ISB_ERR_OUT_OF_RANGE, "output buffer too small");
        std::memcpy(out, v.data(), v.size() * sizeof(double));
    });
}
int isb_serialize_matrix(const double* m, int rows, int cols, int is_f32, char* buf, size_t cap)
{
    return guarded([&] {
        NOT_NULL(m); NOT_NULL(buf);
        const std::string s = serialize_matrix(m, rows, cols, is_f32 != 0);
        if (s.size() + 1 > cap) throw Error(ISB_ERR_OUT_OF_RANGE, "output buffer too small");
        std::memcpy(buf, s.c_str(), s.size() + 1);
    });
}
int isb_deserialize_matrix(const char* s, float* out, int capacity, int* rows, int* cols)
{
    return guarded([&] {
        NOT_NULL(s); NOT_NULL(out); NOT_NULL(rows); NOT_NULL(cols);
        std::vector<float> v;
        if (!deserialize_matrix(s, v, *rows, *cols)) throw Error(ISB_ERR_IO, "malformed matrix text");
        if ((int)v.size() > capacity) throw Error(ISB_ERR_OUT_OF_RANGE, "output buffer too small");
        std::memcpy(out, v.data(), v.size() * sizeof(float));
    });
}
int isb_save_cams(const char* path, const isb_camera* cams, int n)
{
    return guarded([&] {
        NOT_NULL(cams);
        if (!save_cams(path, cams, n)) throw Error(ISB_ERR_IO, "cannot write cams.data");
    });
}
int isb_load_cams(const char* path, isb_camera* cams, int capacity, int* n)
{
    return guarded([&] {
        NOT_NULL(n);
        std::vector<isb_camera> v;
        if (!load_cams(path, v)) throw Error(ISB_ERR_IO, "cannot read / parse cams.data");
        *n = (int)v.size();
        if (cams) {
            if ((int)v.size() > capacity) throw Error(ISB_ERR_OUT_OF_RANGE, "output buffer too small");
            std::memcpy(cams, v.data(), v.size() * sizeof(isb_camera));
        }
    });
}
int isb_save_indices(const char* path, const int* idx, int n)
{
    return guarded([&] {
        NOT_NULL(idx);
        if (!save_indices(path, idx, n)) throw Error(ISB_ERR_IO, "cannot write indices.data");
    });
}
int isb_load_indices(const char* path, int* idx, int capacity, int* n)
{
    return guarded([&] {
        NOT_NULL(n);
        std::vector<int> v;
        if (!load_indices(path, v)) throw Error(ISB_ERR_IO, "cannot read indices.data");
        *n = (int)v.size();
        if (idx) {
            if ((int)v.size() > capacity) throw Error(ISB_ERR_OUT_OF_RANGE, "output buffer too small");
            std::memcpy(idx, v.data(), v.size() * sizeof(int));
        }
    });
}

// ---- warper ---------------------------------------------------------------------------------------
isb_warper* isb_warper_create(int kind, float scale)
{
    if (kind != ISB_WARP_SPHERICAL && kind != ISB_WARP_CYLINDRICAL) {
        t_error = "unsupported warp kind (only spherical and cylindrical are on the hot path)";
        return nullptr;
    }
    return new (std::nothrow) isb_warper(kind, scale);
}
void isb_warper_destroy(isb_warper* w) { delete w; }
float isb_warper_get_scale(const isb_warper* w) { return w ? w->impl.scale() : 0.f; }
int isb_warper_set_scale(isb_warper* w, float s)
{
    return guarded([&] { NOT_NULL(w); w->impl.set_scale(s); });
}
int isb_warper_warp_roi(isb_warper* w, int sw, int sh, const float K[9], const float R[9], int rect[4])
{
    return guarded([&] {
        NOT_NULL(w); NOT_NULL(rect);
        const Rect r = w->impl.warp_roi(sw, sh, K, R);
        rect[0] = r.x; rect[1] = r.y; rect[2] = r.w; rect[3] = r.h;
    });
}
int isb_warper_warp_point(isb_warper* w, const float pt[2], const float K[9], const float R[9], float out[2])
{
    return guarded([&] { NOT_NULL(w); NOT_NULL(pt); NOT_NULL(out); w->impl.warp_point(pt, K, R, out, false); });
}
int isb_warper_warp_point_backward(isb_warper* w, const float pt[2], const float K[9], const float R[9], float out[2])
{
    return guarded([&] { NOT_NULL(w); NOT_NULL(pt); NOT_NULL(out); w->impl.warp_point(pt, K, R, out, true); });
}
int isb_warper_build_maps(isb_warper* w, int sw, int sh, const float K[9], const float R[9], float* xmap, float* ymap,
                          size_t pitch, int rect[4])
{
    return guarded([&] {
        NOT_NULL(w);
        const Rect r = w->impl.build_maps(sw, sh, K, R, xmap, ymap, pitch);
        if (rect) { rect[0] = r.x; rect[1] = r.y; rect[2] = r.w; rect[3] = r.h; }
    });
}
int isb_warper_warp(isb_warper* w, const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float K[9],
                    const float R[9], int interp, int border, uint8_t* dst, size_t dpitch, int corner[2])
{
    return guarded([&] { NOT_NULL(w); w->impl.warp(src, sw, sh, ch, spitch, K, R, interp, border, dst, dpitch, corner); });
}
int isb_warper_warp_backward(isb_warper* w, const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* K, const float* R,
                             int interp, int border, int dw, int dh, uint8_t* dst, size_t dpitch)
{
    return guarded([&] { NOT_NULL(w); w->impl.warp_backward(src, sw, sh, ch, spitch, K, R, interp, border, dw, dh, dst, dpitch); });
}

// ---- compensator / seam -----------------------------------------------------------------------------
isb_compensator* isb_compensator_create(int bw, int bh) { return new (std::nothrow) isb_compensator(bw, bh); }
void isb_compensator_destroy(isb_compensator* c) { delete c; }
int isb_compensator_set_mat_gains(isb_compensator* c, int n, const float* const* gains, const int* gw, const int* gh)
{
    return guarded([&] { NOT_NULL(c); NOT_NULL(gains); NOT_NULL(gw); NOT_NULL(gh); c->impl.set_gains(n, gains, gw, gh); });
}
int isb_compensator_get_mat_gain(const isb_compensator* c, int index, float* out, int capacity, int* gw, int* gh)
{
    return guarded([&] {
        NOT_NULL(c); NOT_NULL(gw); NOT_NULL(gh);
        if (index < 0 || index >= c->impl.count()) throw Error(ISB_ERR_OUT_OF_RANGE, "gain index out of range");
        const std::vector<float>& g = c->impl.gain(index, *gw, *gh);
        if (out) {
            if ((int)g.size() > capacity) throw Error(ISB_ERR_OUT_OF_RANGE, "output buffer too small");
            std::memcpy(out, g.data(), g.size() * sizeof(float));
        }
    });
}
int isb_compensator_apply(isb_compensator* c, int index, const int corner[2], uint8_t* image, int w, int h, size_t pitch,
                          const uint8_t* mask, size_t mask_pitch)
{
    (void)corner; (void)mask; (void)mask_pitch;  // BlocksCompensator::apply ignores both
    return guarded([&] { NOT_NULL(c); c->impl.apply(index, image, w, h, pitch); });
}
int isb_seam_mask_apply(const uint8_t* seam, int mw, int mh, size_t spitch, uint8_t* mask, int w, int h, size_t pitch)
{
    return guarded([&] { seam_mask_apply(seam, mw, mh, spitch, mask, w, h, pitch); });
}

int isb_rotate(const uint8_t* src, int w, int h, int ch, size_t spitch, int code, uint8_t* dst, size_t dpitch)
{
    return guarded([&] { rotate_image(src, w, h, ch, spitch, code, dst, dpitch); });
}
int isb_resize_linear_exact(const uint8_t* src, int sw, int sh, int ch, size_t spitch, uint8_t* dst, int dw, int dh,
                            size_t dpitch, double fx, double fy)
{
    return guarded([&] { resize_linear_exact(src, sw, sh, ch, spitch, dst, dw, dh, dpitch, fx, fy); });
}

// ---- blender --------------------------------------------------------------------------------------
int isb_result_roi(const int* corners, const int* sizes, int n, int rect[4])
{
    return guarded([&] {
        NOT_NULL(corners); NOT_NULL(sizes); NOT_NULL(rect);
        ISB_ASSERT(n > 0);
        const Rect r = result_roi(corners, sizes, n);
        rect[0] = r.x; rect[1] = r.y; rect[2] = r.w; rect[3] = r.h;
    });
}
int isb_num_bands_for(int dst_w, int dst_h, float blend_strength)
{
    // image_stitching.cpp:1177-1183
    const float blend_width = std::sqrt(static_cast<float>((long long)dst_w * dst_h)) * blend_strength / 100.f;
    if (blend_width < 1.f) return -1;
    return static_cast<int>(std::ceil(std::log(blend_width) / std::log(2.)) - 1.);
}
isb_blender* isb_blender_create(int nb) { return new (std::nothrow) isb_blender(nb); }
void isb_blender_destroy(isb_blender* b) { delete b; }
int isb_blender_set_num_bands(isb_blender* b, int nb)
{
    return guarded([&] { NOT_NULL(b); b->impl.set_num_bands(nb); });
}
int isb_blender_num_bands(const isb_blender* b) { return b ? b->impl.num_bands() : -1; }
int isb_blender_actual_num_bands(const isb_blender* b) { return b ? b->impl.actual_bands() : -1; }
int isb_blender_prepare(isb_blender* b, const int* corners, const int* sizes, int n)
{
    return guarded([&] {
        NOT_NULL(b); NOT_NULL(corners); NOT_NULL(sizes);
        ISB_ASSERT(n > 0);
        b->impl.prepare(result_roi(corners, sizes, n));
    });
}
int isb_blender_prepare_roi(isb_blender* b, const int rect[4])
{
    return guarded([&] { NOT_NULL(b); NOT_NULL(rect); b->impl.prepare(Rect{rect[0], rect[1], rect[2], rect[3]}); });
}
int isb_blender_get_rois(const isb_blender* b, int padded[4], int fin[4])
{
    return guarded([&] {
        NOT_NULL(b);
        const BlendGeometry& g = b->impl.geom();
        if (padded) { padded[0] = g.roi.x; padded[1] = g.roi.y; padded[2] = g.roi.w; padded[3] = g.roi.h; }
        if (fin) { fin[0] = g.roi_final.x; fin[1] = g.roi_final.y; fin[2] = g.roi_final.w; fin[3] = g.roi_final.h; }
    });
}
int isb_blender_tile_rect(const isb_blender* b, int w, int h, int tlx, int tly, int rect[4])
{
    return guarded([&] {
        NOT_NULL(b); NOT_NULL(rect);
        b->impl.geom().tile_rect(w, h, tlx, tly, rect, rect + 2);
    });
}
int isb_blender_feed(isb_blender* b, const int16_t* img, size_t ipitch, const uint8_t* mask, size_t mpitch, int w, int h,
                     int tlx, int tly)
{
    return guarded([&] { NOT_NULL(b); b->impl.feed(img, ipitch, mask, mpitch, w, h, tlx, tly); });
}
int isb_blender_blend(isb_blender* b, int16_t* dst, size_t dpitch, uint8_t* dmask, size_t mpitch)
{
    return guarded([&] { NOT_NULL(b); b->impl.blend(dst, dpitch, dmask, mpitch); });
}

// ---- Blender::NO / FeatherBlender -------------------------------------------------------------------
isb_simple_blender* isb_simple_blender_create(int type, float sharpness)
{
    if (type != ISB_BLENDER_NO && type != ISB_BLENDER_FEATHER) {
        t_error = "isb_simple_blender_create: type must be ISB_BLENDER_NO or ISB_BLENDER_FEATHER";
        return nullptr;
    }
    return new (std::nothrow) isb_simple_blender(type, sharpness);
}
void isb_simple_blender_destroy(isb_simple_blender* b) { delete b; }
int isb_simple_blender_set_sharpness(isb_simple_blender* b, float sharpness)
{
    return guarded([&] { NOT_NULL(b); b->impl.set_sharpness(sharpness); });
}
float isb_simple_blender_sharpness(const isb_simple_blender* b) { return b ? b->impl.sharpness() : -1.f; }
int isb_simple_blender_prepare(isb_simple_blender* b, const int* corners, const int* sizes, int n)
{
    return guarded([&] {
        NOT_NULL(b); NOT_NULL(corners); NOT_NULL(sizes);
        ISB_ASSERT(n > 0);
        b->impl.prepare(result_roi(corners, sizes, n));
    });
}
int isb_simple_blender_prepare_roi(isb_simple_blender* b, const int rect[4])
{
    return guarded([&] { NOT_NULL(b); NOT_NULL(rect); b->impl.prepare(Rect{rect[0], rect[1], rect[2], rect[3]}); });
}
int isb_simple_blender_feed(isb_simple_blender* b, const int16_t* img, size_t ipitch, const uint8_t* mask, size_t mpitch, int w,
                            int h, int tlx, int tly)
{
    return guarded([&] { NOT_NULL(b); b->impl.feed(img, ipitch, mask, mpitch, w, h, tlx, tly); });
}
int isb_simple_blender_blend(isb_simple_blender* b, int16_t* dst, size_t dpitch, uint8_t* dmask, size_t mpitch)
{
    return guarded([&] { NOT_NULL(b); b->impl.blend(dst, dpitch, dmask, mpitch); });
}
int isb_create_weight_map(const uint8_t* mask, size_t mpitch, int w, int h, float sharpness, float* weight, size_t wpitch)
{
    return guarded([&] { SimpleBlender::weight_map(mask, mpitch, w, h, sharpness, weight, wpitch); });
}

// ---- Timelapser -------------------------------------------------------------------------------------
isb_timelapser* isb_timelapser_create(int type)
{
    if (type != ISB_TIMELAPSER_AS_IS && type != ISB_TIMELAPSER_CROP) {
        t_error = "isb_timelapser_create: type must be ISB_TIMELAPSER_AS_IS or ISB_TIMELAPSER_CROP";
        return nullptr;
    }
    return new (std::nothrow) isb_timelapser(type);
}
void isb_timelapser_destroy(isb_timelapser* t) { delete t; }
int isb_timelapser_initialize(isb_timelapser* t, const int* corners, const int* sizes, int n, int roi[4])
{
    return guarded([&] {
        NOT_NULL(t); NOT_NULL(corners); NOT_NULL(sizes);
        t->impl.initialize(corners, sizes, n);
        if (roi) { roi[0] = t->impl.roi().x; roi[1] = t->impl.roi().y; roi[2] = t->impl.roi().w; roi[3] = t->impl.roi().h; }
    });
}
int isb_timelapser_process(isb_timelapser* t, const int16_t* img, size_t ipitch, int w, int h, int tlx, int tly)
{
    return guarded([&] { NOT_NULL(t); t->impl.process(img, ipitch, w, h, tlx, tly); });
}
int isb_timelapser_get_dst(isb_timelapser* t, int16_t* dst, size_t dpitch)
{
    return guarded([&] { NOT_NULL(t); t->impl.get_dst(dst, dpitch); });
}

// ---- crop -------------------------------------------------------------------------------------------
int isb_crop_rect(const uint8_t* mask, int w, int h, size_t pitch, int rect[4], int* n_points)
{
    return guarded([&] { crop_rect(mask, w, h, pitch, rect, n_points); });
}
int isb_crop_rect_image(const void* img, int w, int h, size_t pitch, int is_16s, int rect[4], int* n_points)
{
    return guarded([&] { crop_rect_image(img, w, h, pitch, is_16s, rect, n_points); });
}

// ---- imwrite(.jpg) ------------------------------------------------------------------------------------
int isb_jpeg_encode(const void* image, int w, int h, size_t pitch, int is_16s, int quality, uint8_t* out, size_t capacity, size_t* out_size)
{
    return guarded([&] { jpeg_encode(image, w, h, pitch, is_16s, quality, out, capacity, out_size); });
}
int isb_jpeg_release_workspace(void) { return guarded([&] { jpeg_release_workspace(); }); }

// ---- composer ---------------------------------------------------------------------------------------
isb_composer* isb_composer_create(const isb_config* cfg)
{
    if (!cfg) { t_error = "cfg is null"; return nullptr; }
    isb_config c = *cfg;
    if (c.strip_count <= 0) c.strip_count = 1;
    return new (std::nothrow) isb_composer(c);
}
void isb_composer_destroy(isb_composer* c) { delete c; }
int isb_composer_plan(isb_composer* c, const isb_camera* cams, const int* sizes_wh, int n, int* corners, int* sizes,
                      int dst_roi[4])
{
    return guarded([&] { NOT_NULL(c); c->impl.plan(cams, sizes_wh, n, corners, sizes, dst_roi); });
}
int isb_composer_run(isb_composer* c, const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seams, int n,
                     isb_pano* out)
{
    return guarded([&] { NOT_NULL(c); c->impl.run(imgs, gains, seams, n, out); });
}
int isb_composer_sync(isb_composer* c)
{
    return guarded([&] { NOT_NULL(c); c->impl.sync(); });
}
int isb_composer_join(isb_composer* c)
{
    return guarded([&] { NOT_NULL(c); c->impl.join(); });
}
int isb_composer_last_timings(isb_composer* c, float* ms, int cap)
{
    int n = 0;
    const int rc = guarded([&] { NOT_NULL(c); n = c->impl.last().timings(ms, cap); });
    return rc == ISB_OK ? n : rc;
}
const char* isb_composer_stage_name(int stage) { return Composer::stage_name(stage); }
int isb_composer_byte_model(isb_composer* c, double* S, double* M, double* Ap, double* B)
{
    return guarded([&] { NOT_NULL(c); c->impl.first().byte_model(S, M, Ap, B); });
}
int isb_composer_source_band(isb_composer* c, int index, int* lo, int* hi)
{
    return guarded([&] {
        NOT_NULL(c);
        const std::vector<int>& b = c->impl.first().src_band();
        if (index < 0 || 2 * (size_t)index + 1 >= b.size()) throw Error(ISB_ERR_OUT_OF_RANGE, "image index out of range (plan first)");
        if (lo) *lo = b[2 * index];
        if (hi) *hi = b[2 * index + 1];
    });
}
int isb_composer_strip_rows(isb_composer* c, int* y0, int* y1)
{
    return guarded([&] {
        NOT_NULL(c); NOT_NULL(y0); NOT_NULL(y1);
        c->impl.first().planned_rows(*y0, *y1);
    });
}
long long isb_composer_last_h2d_bytes(isb_composer* c) { return c ? (long long)c->impl.last().last_h2d_bytes() : 0; }
int isb_compose(const isb_image* imgs, const isb_camera* cams, const isb_gainmap* gains, const isb_mask* seams, int n,
                const isb_config* cfg, isb_pano* out)
{
    return guarded([&] {
        NOT_NULL(imgs); NOT_NULL(cams); NOT_NULL(cfg); NOT_NULL(out);
        isb_config c = *cfg;
        if (c.strip_count <= 0) c.strip_count = 1;
        Composer comp(c);
        std::vector<int> sz(2 * n);
        for (int i = 0; i < n; ++i) { sz[2 * i] = imgs[i].width; sz[2 * i + 1] = imgs[i].height; }
        comp.plan(cams, sz.data(), n, nullptr, nullptr, nullptr);
        comp.run(imgs, gains, seams, n, out);
        ISB_CUDA(cudaStreamSynchronize(current_stream()));
    });
}
int isb_device_malloc(size_t bytes, void** p)
{
    return guarded([&] { NOT_NULL(p); require_device(); ISB_CUDA(cudaMalloc(p, bytes ? bytes : 1)); });
}
int isb_device_free(void* p)
{
    return guarded([&] { if (p) ISB_CUDA(cudaFree(p)); });
}
int isb_ipc_get_handle(const void* p, unsigned char handle[64])
{
    return guarded([&] {
        NOT_NULL(p); NOT_NULL(handle);
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
        cudaIpcMemHandle_t h;
        ISB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(p)));
        std::memcpy(handle, &h, 64);
    });
}
int isb_ipc_open_handle(const unsigned char handle[64], void** p)
{
    return guarded([&] {
        NOT_NULL(p); NOT_NULL(handle);
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle, 64);
        ISB_CUDA(cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess));
    });
}
int isb_ipc_close_handle(void* p)
{
    return guarded([&] { if (p) ISB_CUDA(cudaIpcCloseMemHandle(p)); });
}
int isb_memcpy(void* dst, const void* src, size_t bytes, int synchronize)
{
    return guarded([&] {
        NOT_NULL(dst); NOT_NULL(src);
        ISB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, current_stream()));
        if (synchronize) ISB_CUDA(cudaStreamSynchronize(current_stream()));
    });
}
int isb_strip_rows(int padded_h, int final_h, int nb, int idx, int count, int* y0, int* y1)
{
    return guarded([&] {
        NOT_NULL(y0); NOT_NULL(y1);
        ISB_ASSERT(count >= 1 && idx >= 0 && idx < count && nb >= 0 && padded_h % (1 << nb) == 0);
        strip_rows(padded_h, nb, idx, count, *y0, *y1);
        *y0 = std::min(*y0, final_h);
        *y1 = std::min(*y1, final_h);
    });
}

}  // extern "C"
