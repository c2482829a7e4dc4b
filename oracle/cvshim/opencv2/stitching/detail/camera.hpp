// TEST INFRASTRUCTURE ONLY - see ../../core.hpp.  cv::detail::CameraParams as serializer.cpp uses it.
#pragma once
#include "../../core.hpp"
namespace cv { namespace detail {
struct CameraParams {
    double focal = 1, aspect = 1, ppx = 0, ppy = 0;
    Mat R, t;
};
} }
