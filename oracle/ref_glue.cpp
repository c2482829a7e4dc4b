// ref_glue.cpp - TEST INFRASTRUCTURE ONLY.  C entry points over the reference's OWN helper sources, compiled unmodified
// from /root/reference/image_stitching/{euler_order.h, quaternion.h, euler.h, serializer.cpp, cropper.cpp} against the
// cvshim/ stand-in for the few OpenCV core types they use (OpenCV's C++ headers are not in this image).  Built by
// `make -C oracle _ref` into oracle/_ref/libisb_ref.so; tests/test_ref_helpers.py and tests/test_crop.py compare the
// product's host helpers (isb_quat_*, isb_*euler*, isb_pose_from_cam_transform, isb_*serialize*, isb_crop_rect) with it.
// No reference source is copied: the files are #included / compiled from where they lie.
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <string_view>
#include <vector>

#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
#include <opencv2/stitching/detail/camera.hpp>

#include "euler_order.h"
#include "quaternion.h"
#include "euler.h"
#include "serializer.h"
#include "cropper.h"

extern "C" {
cv::cvshim_find_contours_fn cvshim_find_contours_cb = nullptr;
cv::cvshim_draw_contour_fn cvshim_draw_contour_cb = nullptr;
}

#define REF_API extern "C" __attribute__((visibility("default")))

static cv::Mat mat3(const double* R)
{
    cv::Mat m(cv::Size(3, 3), CV_64F);
    for (int i = 0; i < 9; ++i) m.at<double>(i / 3, i % 3) = R[i];
    return m;
}
static void out3(const cv::Mat& m, double* R)
{
    for (int i = 0; i < 9; ++i) R[i] = m.at<double>(i / 3, i % 3);
}
static void outq(const Quaternion<double>& q, double* o) { o[0] = q.x(); o[1] = q.y(); o[2] = q.z(); o[3] = q.w(); }

REF_API void ref_quat_from_rotation_matrix(const double* R, double* q)
{
    Quaternion<double> Q;
    Q.setFromRotationMatrix<double>(mat3(R));
    outq(Q, q);
}
REF_API void ref_quat_to_rotation_matrix(const double* q, double* R)
{
    Quaternion<double> Q(q[0], q[1], q[2], q[3]);
    out3(Q.toRotationMatrix(), R);
}
REF_API void ref_quat_from_euler(const double* e, int order, double* q)
{
    Quaternion<double> Q;
    Q.setFromEuler(cv::Vec<double, 3>(e[0], e[1], e[2]), (EulerOrder)order);
    outq(Q, q);
}
REF_API void ref_quat_from_axis_angle(const double* axis, double angle, double* q)
{
    Quaternion<double> Q;
    Q.setFromAxisAngle(cv::Vec<double, 3>(axis[0], axis[1], axis[2]), angle);
    outq(Q, q);
}
REF_API void ref_quat_multiply(const double* a, const double* b, double* o)
{
    Quaternion<double> A(a[0], a[1], a[2], a[3]), B(b[0], b[1], b[2], b[3]), Q;
    Q.multiplyQuaternions(A, B);
    outq(Q, o);
}
REF_API void ref_quat_slerp(const double* a, const double* b, double t, double* o)
{
    Quaternion<double> A(a[0], a[1], a[2], a[3]), B(b[0], b[1], b[2], b[3]);
    A.slerp(B, t);
    outq(A, o);
}
// the EXIF pose fix-up exactly as image_stitching.cpp:485-517 drives the class (setFromRotationMatrix<double>, the
// component flips through set(), toRotationMatrix)
REF_API void ref_pose_from_cam_transform(const double* R_in, int is_portrait, double* R_out)
{
    Quaternion<double> q, q2;
    q.setFromRotationMatrix<double>(mat3(R_in));
    if (is_portrait) q2.set(q.y(), q.x(), -q.z(), q.w());
    else q2.set(-q.x(), q.y(), -q.z(), q.w());
    q = q2;
    out3(q.toRotationMatrix(), R_out);
}
REF_API void ref_rotation_matrix_to_euler(const double* R, int order, double* e)
{
    const cv::Vec<double, 3> v = rotationMatrixToEulerAngles<double>(mat3(R), (EulerOrder)order);
    e[0] = v[0]; e[1] = v[1]; e[2] = v[2];
}
REF_API void ref_euler_to_rotation_matrix(const double* e, int order, double* R)
{
    out3(eulerAnglesToRotationMatrix<double>(cv::Vec<double, 3>(e[0], e[1], e[2]), (EulerOrder)order), R);
}

// ---- serializer.cpp ---------------------------------------------------------------------------------------------
REF_API int ref_parse_matrix_str(const char* s, double* out, int cap)
{
    const cv::Mat m = parseMatrixStr(s);
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c)
            if (r * m.cols + c < cap) out[r * m.cols + c] = m.at<double>(r, c);
    return m.rows;
}
REF_API int ref_serialize_matrix(const double* v, int rows, int cols, int is_f32, char* buf, int cap)
{
    cv::Mat m(rows, cols, is_f32 ? CV_32F : CV_64F);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            if (is_f32) m.at<float>(r, c) = (float)v[r * cols + c];
            else m.at<double>(r, c) = v[r * cols + c];
        }
    const std::string s = serializeMatrix(m);
    if ((int)s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
REF_API int ref_deserialize_matrix(const char* s, float* out, int cap, int* rows, int* cols)
{
    const cv::Mat m = deserializeMatrix(s);
    *rows = m.rows; *cols = m.cols;
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c)
            if (r * m.cols + c < cap) out[r * m.cols + c] = m.at<float>(r, c);
    return 0;
}
struct RefCam { double focal, aspect, ppx, ppy; float R[9]; float t[3]; };  // layout of isb_camera
// serializeCameraParams / deserializeCameraParams use the fixed path ./cams.data: the caller chdir()s into a scratch dir
REF_API void ref_save_cams(const RefCam* cams, int n)
{
    std::vector<cv::detail::CameraParams> v(n);
    for (int i = 0; i < n; ++i) {
        v[i].focal = cams[i].focal; v[i].aspect = cams[i].aspect; v[i].ppx = cams[i].ppx; v[i].ppy = cams[i].ppy;
        v[i].R = cv::Mat(3, 3, CV_32F);
        for (int k = 0; k < 9; ++k) v[i].R.at<float>(k / 3, k % 3) = cams[i].R[k];
        v[i].t = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) v[i].t.at<float>(k, 0) = cams[i].t[k];
    }
    serializeCameraParams(v);
}
REF_API int ref_load_cams(RefCam* cams, int cap)
{
    const std::vector<cv::detail::CameraParams> v = deserializeCameraParams();
    for (int i = 0; i < (int)v.size() && i < cap; ++i) {
        cams[i].focal = v[i].focal; cams[i].aspect = v[i].aspect; cams[i].ppx = v[i].ppx; cams[i].ppy = v[i].ppy;
        for (int k = 0; k < 9; ++k) cams[i].R[k] = (k / 3 < v[i].R.rows && k % 3 < v[i].R.cols) ? v[i].R.at<float>(k / 3, k % 3) : 0.f;
        for (int k = 0; k < 3; ++k) cams[i].t[k] = (k < v[i].t.rows && v[i].t.cols > 0) ? v[i].t.at<float>(k, 0) : 0.f;
    }
    return (int)v.size();
}
REF_API void ref_save_indices(const int* idx, int n) { serializeIndices(std::vector<int>(idx, idx + n)); }
REF_API int ref_load_indices(int* idx, int cap)
{
    const std::vector<int> v = deserializeIndices();
    for (int i = 0; i < (int)v.size() && i < cap; ++i) idx[i] = v[i];
    return (int)v.size();
}

// ---- cropper.cpp -------------------------------------------------------------------------------------------------
REF_API void ref_set_contour_callbacks(cv::cvshim_find_contours_fn f, cv::cvshim_draw_contour_fn d)
{
    cvshim_find_contours_cb = f;
    cvshim_draw_contour_cb = d;
}
// crop(source) on a 16SC3 (what blend() returns) or 8UC3 image; reports the rectangle it cropped to by locating the
// returned ROI view inside the buffer it was cut from (crop() only narrows `source` to source(croppingMask))
REF_API void ref_crop(const void* img, int w, int h, int is_16s, int* rect_xywh)
{
    cv::Mat m(h, w, is_16s ? CV_16SC3 : CV_8UC3);
    for (int r = 0; r < h; ++r) std::memcpy(m.ptr(r), (const char*)img + (size_t)r * w * m.elemSize(), (size_t)w * m.elemSize());
    crop(m);
    rect_xywh[2] = m.cols;
    rect_xywh[3] = m.rows;
    rect_xywh[0] = rect_xywh[1] = -1;
    // convertTo() inside crop() re-allocated the image as 8UC3; the view's origin is found in that buffer
    // through the offset the shim keeps between the view and its allocation
    const size_t off = m.view_offset();
    rect_xywh[1] = (int)(off / m.step());
    rect_xywh[0] = (int)((off % m.step()) / m.elemSize());
}
REF_API int ref_check_interior_exterior(const uint8_t* mask, int w, int h, const int* rect, int* tblr)
{
    cv::Mat m(h, w, CV_8UC1);
    for (int r = 0; r < h; ++r) std::memcpy(m.ptr(r), mask + (size_t)r * w, (size_t)w);
    return checkInteriorExterior(m, cv::Rect(rect[0], rect[1], rect[2], rect[3]), tblr[0], tblr[1], tblr[2], tblr[3]) ? 1 : 0;
}
