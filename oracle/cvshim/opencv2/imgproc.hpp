// TEST INFRASTRUCTURE ONLY - see core.hpp.  The three imgproc calls cropper.cpp makes.  cvtColor is restated (8-bit
// fixed-point RGB2GRAY, checked against cv2 in tests/test_crop.py); findContours / drawContours are NOT restated:
// they call back into the test process, which answers with the real OpenCV (cv2) - so crop() below them is the
// reference's own code and the contour primitives are OpenCV's own.
#pragma once
#include "core.hpp"
namespace cv {
enum { COLOR_RGB2GRAY = 7 };
enum { RETR_EXTERNAL = 0 };
enum { CHAIN_APPROX_NONE = 1 };

inline void cvtColor(const Mat& src, Mat& dst, int code)
{
    assert(code == COLOR_RGB2GRAY && src.type() == CV_8UC3);
    (void)code;
    Mat g(src.rows, src.cols, CV_8UC1);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) {
            const uint8_t* p = src.ptr(r) + 3 * c;  // R, G, B order for RGB2GRAY
            g.ptr(r)[c] = (uint8_t)((p[0] * 4899 + p[1] * 9617 + p[2] * 1868 + (1 << 13)) >> 14);
        }
    dst = g;
}

// callbacks installed by the test harness (ctypes): both receive / fill tightly packed buffers
extern "C" {
typedef int (*cvshim_find_contours_fn)(const uint8_t* mask, int w, int h, int** xy_out, int** lens_out, int* n_out);
typedef void (*cvshim_draw_contour_fn)(uint8_t* img, int w, int h, const int* xy, int npts);
extern cvshim_find_contours_fn cvshim_find_contours_cb;
extern cvshim_draw_contour_fn cvshim_draw_contour_cb;
}

inline void findContours(const Mat& image, std::vector<std::vector<Point>>& contours, std::vector<Vec4i>& hierarchy, int mode,
                         int method, Point offset = Point())
{
    assert(mode == RETR_EXTERNAL && method == CHAIN_APPROX_NONE && image.type() == CV_8UC1 && cvshim_find_contours_cb);
    (void)mode; (void)method; (void)offset;
    Mat tight = image.clone();
    int *xy = nullptr, *lens = nullptr, n = 0;
    cvshim_find_contours_cb(tight.ptr(), tight.cols, tight.rows, &xy, &lens, &n);
    contours.clear();
    hierarchy.clear();
    size_t k = 0;
    for (int i = 0; i < n; ++i) {
        std::vector<Point> c(lens[i]);
        for (int j = 0; j < lens[i]; ++j, ++k) c[j] = Point(xy[2 * k], xy[2 * k + 1]);
        contours.push_back(std::move(c));
        hierarchy.push_back(Vec4i(i + 1 < n ? i + 1 : -1, i - 1, -1, -1));
    }
}

inline void drawContours(Mat& image, const std::vector<std::vector<Point>>& contours, int idx, const Scalar& color, int thickness,
                         int lineType, const std::vector<Vec4i>& hierarchy, int maxLevel, Point offset = Point())
{
    assert(thickness == -1 && lineType == 8 && maxLevel == 0 && image.type() == CV_8UC1 && color[0] == 255 && cvshim_draw_contour_cb);
    (void)thickness; (void)lineType; (void)hierarchy; (void)maxLevel; (void)offset; (void)color;
    std::vector<int> xy;
    for (const Point& p : contours.at(idx)) { xy.push_back(p.x); xy.push_back(p.y); }
    Mat tight(image.rows, image.cols, CV_8UC1);
    cvshim_draw_contour_cb(tight.ptr(), tight.cols, tight.rows, xy.data(), (int)contours.at(idx).size());
    for (int r = 0; r < image.rows; ++r) std::memcpy(image.ptr(r), tight.ptr(r), (size_t)image.cols);
}
}  // namespace cv
