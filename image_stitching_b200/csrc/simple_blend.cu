// simple_blend.cu - cv::detail::Blender (Blender::NO) and cv::detail::FeatherBlender, the two other blenders the
// reference's compositing loop can select (image_stitching.cpp:1175-1191), plus
// cv::detail::createWeightMap.  Same prepare / feed / blend contract as the multi-band blender; all state lives in HBM.
//
//   createWeightMap(mask, sharpness) = min(1, sharpness * distanceTransform(mask, DIST_L1, 3))
// The 3x3 L1 chamfer is the exact city-block distance to the nearest zero pixel, which separates:
//   dv(x, y) = min_y' |y - y'| over zero pixels of column x                      (one thread per column, two sweeps)
//   d(x, y)  = min_x' dv(x', y) + |x - x'|
//            = min( x + prefmin_x'<=x (dv - x'),  -x + sufmin_x'>=x (dv + x') )    (one warp per row, shuffle scans)
// A mask without any zero pixel gives an "infinite" distance; OpenCV then yields 65534 (or FLT_MAX with IPP): both
// clamp to weight 1 for any sharpness >= 1.6e-5, which is asserted.
#include <algorithm>
#include <climits>

#include "device_math.cuh"
#include "engine.hpp"
#include "kernels.cuh"

namespace isb {

constexpr int kDistInf = 1 << 28;

// pass 1: vertical distances.  Thread = column; coalesced across the warp, sequential down / up the column.
__global__ void __launch_bounds__(128) dist_columns_kernel(const uint8_t* __restrict__ mask, long long mpitch, int w, int h,
                                                           int* __restrict__ dv)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    int d = kDistInf;
    for (int y = 0; y < h; ++y) {
        d = mask[y * mpitch + x] ? min(d + 1, kDistInf) : 0;
        dv[(long long)y * w + x] = d;
    }
    d = kDistInf;
    for (int y = h - 1; y >= 0; --y) {
        const int v = dv[(long long)y * w + x];
        d = v == 0 ? 0 : min(d + 1, kDistInf);
        dv[(long long)y * w + x] = min(v, d);
    }
}

// pass 2: horizontal min-plus combination + weight.  Warp = row.
__global__ void __launch_bounds__(256) dist_rows_kernel(const int* __restrict__ dv, int w, int h, float sharpness,
                                                        float* __restrict__ weight, long long wpitch)
{
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (y >= h) return;
    const int* __restrict__ r = dv + (long long)y * w;
    float* __restrict__ o = weight + y * wpitch;
    const int chunks = (w + 31) >> 5;
    // forward: prefix min of (dv - x); stash x + that in the output row (as int bits)
    int carry = INT_MAX / 2;
    for (int c = 0; c < chunks; ++c) {
        const int x = c * 32 + lane;
        int v = x < w ? r[x] - x : INT_MAX / 2;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, s);
            if (lane >= s) v = min(v, t);
        }
        v = min(v, carry);
        carry = __shfl_sync(0xffffffffu, v, 31);
        if (x < w) o[x] = __int_as_float(min(v + x, kDistInf));
    }
    // backward: suffix min of (dv + x)
    carry = INT_MAX / 2;
    for (int c = chunks - 1; c >= 0; --c) {
        const int x = c * 32 + lane;
        int v = x < w ? r[x] + x : INT_MAX / 2;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_down_sync(0xffffffffu, v, s);
            if (lane + s < 32) v = min(v, t);
        }
        v = min(v, carry);
        carry = __shfl_sync(0xffffffffu, v, 0);
        if (x < w) {
            const int d = min(__float_as_int(o[x]), min(v - x, kDistInf));
            // distanceTransform returns float; no zero pixel anywhere -> 65534 in the non-IPP build
            const float dist = d >= kDistInf ? 65534.f : (float)d;
            o[x] = fminf(__fmul_rn(dist, sharpness), 1.f);
        }
    }
}

// Blender::feed: dst(mask != 0) = img; dst_mask |= mask
__global__ void __launch_bounds__(256) feed_no_kernel(const int16_t* __restrict__ img, long long ipitch,
                                                      const uint8_t* __restrict__ mask, long long mpitch, int w, int h,
                                                      int16_t* __restrict__ dst, uint8_t* __restrict__ dmask, int dw, int dx, int dy)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const uint8_t m = mask[y * mpitch + x];
    const long long d = (long long)(dy + y) * dw + dx + x;
    if (m) {
        const int16_t* s = reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(img) + y * ipitch) + 3 * x;
        dst[3 * d] = s[0]; dst[3 * d + 1] = s[1]; dst[3 * d + 2] = s[2];
        dmask[d] |= m;
    }
}

// FeatherBlender::feed: dst += (short)(img * w) (int16 wrap-around as the reference's short arithmetic), dst_weight += w
__global__ void __launch_bounds__(256) feed_feather_kernel(const int16_t* __restrict__ img, long long ipitch,
                                                           const float* __restrict__ wmap, int w, int h,
                                                           int16_t* __restrict__ dst, float* __restrict__ dweight, int dw, int dx,
                                                           int dy)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const float wt = wmap[(long long)y * w + x];
    const long long d = (long long)(dy + y) * dw + dx + x;
    const int16_t* s = reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(img) + y * ipitch) + 3 * x;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[3 * d + c] = (int16_t)(dst[3 * d + c] + trunc_s16(__fmul_rn((float)s[c], wt)));
    dweight[d] = __fadd_rn(dweight[d], wt);
}

// FeatherBlender::blend: normalizeUsingWeightMap, dst_mask = weight > 1e-5, then Blender::blend zeroes outside the mask
__global__ void __launch_bounds__(256) blend_simple_kernel(const int16_t* __restrict__ acc, const uint8_t* __restrict__ amask,
                                                           const float* __restrict__ aweight, int feather, int w, int h,
                                                           int16_t* __restrict__ out, long long opitch, uint8_t* __restrict__ omask,
                                                           long long ompitch)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const long long i = (long long)y * w + x;
    int v[3] = {acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]};
    int m = amask[i];
    if (feather) {
        const float wt = aweight[i], den = __fadd_rn(wt, 1e-5f);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = trunc_s16(__fdiv_rn((float)v[c], den));
        m = wt > 1e-5f ? 255 : 0;
    }
    if (!m) v[0] = v[1] = v[2] = 0;
    if (out) {
        int16_t* o = reinterpret_cast<int16_t*>(reinterpret_cast<char*>(out) + y * opitch) + 3 * x;
        o[0] = (int16_t)v[0]; o[1] = (int16_t)v[1]; o[2] = (int16_t)v[2];
    }
    if (omask) omask[y * ompitch + x] = (uint8_t)m;
}

static void weight_map_device(const uint8_t* dmask, long long mpitch, int w, int h, float sharpness, float* dweight,
                              long long wpitch_elems, DevBuf& scratch, cudaStream_t st)
{
    int* dv = static_cast<int*>(scratch.ensure((size_t)w * h * sizeof(int)));
    dist_columns_kernel<<<(w + 127) / 128, 128, 0, st>>>(dmask, mpitch, w, h, dv);
    count_launch();
    dist_rows_kernel<<<(h + 7) / 8, 256, 0, st>>>(dv, w, h, sharpness, dweight, wpitch_elems);
    count_launch();
    ISB_CUDA(cudaGetLastError());
}

void SimpleBlender::weight_map(const uint8_t* mask, size_t mpitch, int w, int h, float sharpness, float* weight, size_t wpitch)
{
    require_device();
    if (!mask || !weight) throw Error(ISB_ERR_NULL_PTR, "mask/weight are null");
    ISB_ASSERT(w > 0 && h > 0 && mpitch >= (size_t)w && wpitch >= (size_t)w * sizeof(float) && wpitch % sizeof(float) == 0);
    ISB_ASSERT(sharpness >= 1.6e-5f);  // keeps the "no zero pixel" distance (65534 / FLT_MAX) clamped to weight 1
    cudaStream_t st = current_stream();
    DevBuf mbuf, wbuf, scratch;
    const uint8_t* dm = mask;
    long long dmp = (long long)mpitch;
    if (mem_kind(mask) != MemKind::Device) {
        dmp = w;
        copy2d(mbuf.ensure((size_t)w * h), w, mask, mpitch, w, h, st);
        dm = mbuf.as<uint8_t>();
    }
    const bool wdev = mem_kind(weight) == MemKind::Device;
    float* dw = wdev ? weight : static_cast<float*>(wbuf.ensure((size_t)w * h * sizeof(float)));
    weight_map_device(dm, dmp, w, h, sharpness, dw, wdev ? (long long)(wpitch / sizeof(float)) : w, scratch, st);
    if (!wdev) copy2d(weight, wpitch, dw, (size_t)w * sizeof(float), (size_t)w * sizeof(float), h, st);
    ISB_CUDA(cudaStreamSynchronize(st));
}

void SimpleBlender::prepare(const Rect& roi)
{
    require_device();
    ISB_ASSERT(roi.w > 0 && roi.h > 0);
    ISB_ASSERT(type_ == ISB_BLENDER_NO || type_ == ISB_BLENDER_FEATHER);
    cudaStream_t st = current_stream();
    roi_ = roi;
    const size_t n = (size_t)roi.w * roi.h;
    ISB_CUDA(cudaMemsetAsync(dst_.ensure(n * 6), 0, n * 6, st));
    ISB_CUDA(cudaMemsetAsync(dmask_.ensure(n), 0, n, st));
    if (type_ == ISB_BLENDER_FEATHER) ISB_CUDA(cudaMemsetAsync(dweight_.ensure(n * sizeof(float)), 0, n * sizeof(float), st));
    prepared_ = true;
}

void SimpleBlender::feed(const int16_t* img, size_t ipitch, const uint8_t* mask, size_t mpitch, int w, int h, int tlx, int tly)
{
    require_device();
    if (!prepared_) throw Error(ISB_ERR_ASSERT, "Assertion failed: prepare() must be called before feed()");
    if (!img || !mask) throw Error(ISB_ERR_NULL_PTR, "img/mask are null");
    ISB_ASSERT(w > 0 && h > 0 && ipitch >= (size_t)w * 6 && mpitch >= (size_t)w);
    const int dx = tlx - roi_.x, dy = tly - roi_.y;
    ISB_ASSERT(dx >= 0 && dy >= 0 && dx + w <= roi_.w && dy + h <= roi_.h);  // the image must lie inside dst_roi_
    if (type_ == ISB_BLENDER_FEATHER) ISB_ASSERT(sharpness_ >= 1.6e-5f);
    cudaStream_t st = current_stream();
    const int16_t* di = img;
    size_t dip = ipitch;
    if (mem_kind(img) != MemKind::Device) {
        dip = (size_t)w * 6;
        copy2d(img_.ensure(dip * h), dip, img, ipitch, dip, h, st);
        di = img_.as<int16_t>();
    }
    const uint8_t* dm = mask;
    size_t dmp = mpitch;
    if (mem_kind(mask) != MemKind::Device) {
        dmp = w;
        copy2d(mask_.ensure(dmp * h), dmp, mask, mpitch, w, h, st);
        dm = mask_.as<uint8_t>();
    }
    const dim3 grid((w + 31) / 32, (h + 7) / 8);
    if (type_ == ISB_BLENDER_NO) {
        feed_no_kernel<<<grid, 256, 0, st>>>(di, (long long)dip, dm, (long long)dmp, w, h, dst_.as<int16_t>(), dmask_.as<uint8_t>(),
                                             roi_.w, dx, dy);
    } else {
        float* wm = static_cast<float*>(wmap_.ensure((size_t)w * h * sizeof(float)));
        weight_map_device(dm, (long long)dmp, w, h, sharpness_, wm, w, dist_, st);
        feed_feather_kernel<<<grid, 256, 0, st>>>(di, (long long)dip, wm, w, h, dst_.as<int16_t>(), dweight_.as<float>(), roi_.w, dx,
                                                  dy);
    }
    count_launch();
    ISB_CUDA(cudaGetLastError());
    ISB_CUDA(cudaStreamSynchronize(st));  // feed keeps no reference to img/mask after it returns
}

void SimpleBlender::blend(int16_t* dst, size_t dpitch, uint8_t* dmask, size_t mpitch)
{
    require_device();
    if (!prepared_) throw Error(ISB_ERR_ASSERT, "Assertion failed: prepare() must be called before blend()");
    cudaStream_t st = current_stream();
    const int w = roi_.w, h = roi_.h;
    const bool d16 = dst && mem_kind(dst) == MemKind::Device, dmk = dmask && mem_kind(dmask) == MemKind::Device;
    int16_t* o16 = nullptr;
    uint8_t* om = nullptr;
    long long op = 0, omp = 0;
    if (dst) {
        ISB_ASSERT(dpitch >= (size_t)w * 6);
        o16 = d16 ? dst : static_cast<int16_t*>(img_.ensure((size_t)w * 6 * h));
        op = d16 ? (long long)dpitch : (long long)w * 6;
    }
    if (dmask) {
        ISB_ASSERT(mpitch >= (size_t)w);
        om = dmk ? dmask : static_cast<uint8_t*>(mask_.ensure((size_t)w * h));
        omp = dmk ? (long long)mpitch : w;
    }
    const dim3 grid((w + 31) / 32, (h + 7) / 8);
    blend_simple_kernel<<<grid, 256, 0, st>>>(dst_.as<int16_t>(), dmask_.as<uint8_t>(), dweight_.as<float>(),
                                              type_ == ISB_BLENDER_FEATHER ? 1 : 0, w, h, o16, op, om, omp);
    count_launch();
    ISB_CUDA(cudaGetLastError());
    if (dst && !d16) copy2d(dst, dpitch, o16, (size_t)op, (size_t)w * 6, h, st);
    if (dmask && !dmk) copy2d(dmask, mpitch, om, (size_t)omp, w, h, st);
    ISB_CUDA(cudaStreamSynchronize(st));
    prepared_ = false;  // single use per prepare(), like the reference blenders
}

// ---- Timelapser ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) timelapse_copy_kernel(const int16_t* __restrict__ img, long long ipitch, int w, int h,
                                                             int16_t* __restrict__ dst, int dw, int dh, int dx, int dy)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const int X = dx + x, Y = dy + y;  // test_point(): dst_roi_.contains(tl + (x, y))
    if ((unsigned)X >= (unsigned)dw || (unsigned)Y >= (unsigned)dh) return;
    const int16_t* s = reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(img) + y * ipitch) + 3 * x;
    int16_t* d = dst + ((long long)Y * dw + X) * 3;
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
}

void Timelapser::initialize(const int* corners, const int* sizes, int n)
{
    ISB_ASSERT(n > 0);
    ISB_ASSERT(type_ == ISB_TIMELAPSER_AS_IS || type_ == ISB_TIMELAPSER_CROP);
    if (type_ == ISB_TIMELAPSER_AS_IS) roi_ = result_roi(corners, sizes, n);
    else {  // TimelapserCrop::initialize: Rect(Point(max tl), Point(min br)) - cv::Rect_(pt1, pt2) orders the corners itself
        int tlx = INT_MIN, tly = INT_MIN, brx = INT_MAX, bry = INT_MAX;
        for (int i = 0; i < n; ++i) {
            tlx = std::max(tlx, corners[2 * i]); tly = std::max(tly, corners[2 * i + 1]);
            brx = std::min(brx, corners[2 * i] + sizes[2 * i]); bry = std::min(bry, corners[2 * i + 1] + sizes[2 * i + 1]);
        }
        roi_ = Rect{std::min(tlx, brx), std::min(tly, bry), std::max(tlx, brx) - std::min(tlx, brx), std::max(tly, bry) - std::min(tly, bry)};
    }
    ready_ = true;
}

void Timelapser::process(const int16_t* img, size_t ipitch, int w, int h, int tlx, int tly)
{
    require_device();
    if (!ready_) throw Error(ISB_ERR_ASSERT, "Assertion failed: initialize() must be called before process()");
    if (!img) throw Error(ISB_ERR_NULL_PTR, "img is null");
    ISB_ASSERT(w > 0 && h > 0 && ipitch >= (size_t)w * 6);
    cudaStream_t st = current_stream();
    const size_t n = (size_t)roi_.w * roi_.h * 6;
    ISB_CUDA(cudaMemsetAsync(dst_.ensure(std::max<size_t>(n, 1)), 0, std::max<size_t>(n, 1), st));  // dst_.setTo(0)
    if (roi_.w > 0 && roi_.h > 0) {
        const int16_t* di = img;
        size_t dip = ipitch;
        if (mem_kind(img) != MemKind::Device) {
            dip = (size_t)w * 6;
            copy2d(img_.ensure(dip * h), dip, img, ipitch, dip, h, st);
            di = img_.as<int16_t>();
        }
        timelapse_copy_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), 256, 0, st>>>(di, (long long)dip, w, h, dst_.as<int16_t>(), roi_.w,
                                                                               roi_.h, tlx - roi_.x, tly - roi_.y);
        count_launch();
        ISB_CUDA(cudaGetLastError());
    }
    ISB_CUDA(cudaStreamSynchronize(st));
}

void Timelapser::get_dst(int16_t* dst, size_t dpitch)
{
    require_device();
    if (!ready_) throw Error(ISB_ERR_ASSERT, "Assertion failed: initialize() must be called before getDst()");
    if (!dst) throw Error(ISB_ERR_NULL_PTR, "dst is null");
    if (roi_.w <= 0 || roi_.h <= 0) return;
    ISB_ASSERT(dpitch >= (size_t)roi_.w * 6);
    cudaStream_t st = current_stream();
    copy2d(dst, dpitch, dst_.ensure((size_t)roi_.w * roi_.h * 6), (size_t)roi_.w * 6, (size_t)roi_.w * 6, roi_.h, st);
    ISB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace isb
