// kernels.cu - kernels of the per-object API (warp, buildMaps, gain, seam mask, feed packing, rotate, resize) and the
// generic (any size / any storage) pyramid kernels; the fused composer's fast path lives in kernels_fast.cu.
//
// Arithmetic contract (each item pinned against OpenCV 4.13 by the oracle):
//  * every float op of the inverse map / gain / weight pyramid rounds to binary32 on its own
//    (__fmul_rn/__fadd_rn/__fdiv_rn; the file is also built with -fmad=false);
//  * 8-bit bilinear sampling is the 1/32-px fixed-point remap (A.3); masks use nearest/half-even;
//  * 16S pyramids are pure integer (A.5); weight pyramids reproduce OpenCV's SIMD/scalar op order;
//  * accumulate / normalise truncate toward zero and wrap to int16 exactly like the C++ casts (A.6).
// None of this is a dense contraction, so no tensor cores: the kernels are HBM/LSU bound.
#include <climits>
#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "device_math.cuh"
#include "glibc_math.cuh"

namespace isb {

static long long g_launches = 0;
long long launch_count(bool reset)
{
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
#define ISB_COUNT_LAUNCH() (++g_launches)
void count_launch() { ++g_launches; }

static thread_local cudaError_t t_launch_error = cudaSuccess;
void note_launch_error(cudaError_t e)
{
    if (e != cudaSuccess && t_launch_error == cudaSuccess) t_launch_error = e;
}
cudaError_t take_launch_error()
{
    const cudaError_t e = t_launch_error;
    t_launch_error = cudaSuccess;
    return e;
}

// Measurement switches (all default to the production setting) are read from the environment ONCE; tools that flip them
// inside one process (tools/ab_env.py) call isb_reload_env() afterwards.
static EnvSwitches read_env_switches()
{
    auto off = [](const char* name) {
        const char* e = getenv(name);
        return e && e[0] == '0';
    };
    EnvSwitches s;
    s.pdl = !off("ISB_PDL");
    s.blend_tma = !off("ISB_BLEND_TMA");
    s.blend_pipe = !off("ISB_BLEND_PIPE");
    s.staged_stores = getenv("ISB_STAGED_STORES") != nullptr;
    const char* lv = getenv("ISB_PYRDOWN_TMA_LEVELS");
    s.pyrdown_tma_levels = (lv && lv[0] >= '0' && lv[0] <= '9') ? atoi(lv) : 64;
    return s;
}
static EnvSwitches g_env = read_env_switches();
const EnvSwitches& env_switches() { return g_env; }
void reload_env_switches() { g_env = read_env_switches(); }
bool pdl_enabled() { return g_env.pdl; }

// ------------------------------------------------------------------------------------------------
// classic API kernels
// ------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) warp_generic_kernel(ImageDev I, int interp, int border, uint8_t* __restrict__ dst,
                                                           long long dpitch)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= I.roi_w || y >= I.roi_h) return;
    const XY m = inverse_map(I.kr, I.col[x], I.row[y]);
    int v[CH];
    if (interp == 1) {
        if (border == 2) sample_linear<CH, true>(I, m, v);
        else sample_linear<CH, false>(I, m, v);
    } else {
        int ix, iy;
        const bool in = nearest_inside(I, m, ix, iy);
        if (border == 2) {
            ix = reflect(ix, I.sw);
            iy = reflect(iy, I.sh);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = (in || border == 2) ? I.src[(long long)iy * I.spitch + ix * CH + c] : 0;
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) dst[(long long)y * dpitch + x * CH + c] = (uint8_t)v[c];
}

// RotationWarperBase::warpBackward: for every pixel (x, y) of the ORIGINAL image the forward map (u, v) - transcendentals per
// pixel, evaluated with glibc's float algorithms (glibc_math.cuh) - then cv::remap of the warped image at (u - tl.x, v - tl.y)
struct BackwardDev {
    float r_kinv[9];
    float scale;
    int spherical;
    int tlx, tly;
};
template <int CH>
__global__ void __launch_bounds__(256) warp_backward_kernel(ImageDev I, BackwardDev B, int interp, int border, int dw, int dh,
                                                            uint8_t* __restrict__ dst, long long dpitch)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const float fx = (float)x, fy = (float)y;
    const float* r = B.r_kinv;
    const float x_ = __fadd_rn(__fadd_rn(__fmul_rn(r[0], fx), __fmul_rn(r[1], fy)), r[2]);
    const float y_ = __fadd_rn(__fadd_rn(__fmul_rn(r[3], fx), __fmul_rn(r[4], fy)), r[5]);
    const float z_ = __fadd_rn(__fadd_rn(__fmul_rn(r[6], fx), __fmul_rn(r[7], fy)), r[8]);
    const float u = __fmul_rn(B.scale, gm::atan2f_(x_, z_));
    float v;
    if (B.spherical) {
        const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(x_, x_), __fmul_rn(y_, y_)), __fmul_rn(z_, z_));
        const float w = __fdiv_rn(y_, __fsqrt_rn(n2));
        v = __fmul_rn(B.scale, __fsub_rn(3.14159274101257324f, gm::acosf_(w == w ? w : 0.f)));
    } else {
        v = __fdiv_rn(__fmul_rn(B.scale, y_), __fsqrt_rn(__fadd_rn(__fmul_rn(x_, x_), __fmul_rn(z_, z_))));
    }
    const XY m{__fsub_rn(u, (float)B.tlx), __fsub_rn(v, (float)B.tly)};
    int val[CH];
    if (interp == 1) {
        if (border == 2) sample_linear<CH, true>(I, m, val);
        else sample_linear<CH, false>(I, m, val);
    } else {
        int ix, iy;
        const bool in = nearest_inside(I, m, ix, iy);
        if (border == 2) {
            ix = reflect(ix, I.sw);
            iy = reflect(iy, I.sh);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) val[c] = (in || border == 2) ? I.src[(long long)iy * I.spitch + ix * CH + c] : 0;
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) dst[(long long)y * dpitch + x * CH + c] = (uint8_t)val[c];
}

void launch_warp_backward(const ImageDev& warped, int ch, const float* r_kinv, float scale, int spherical, int tlx, int tly, int interp,
                          int border, int dw, int dh, uint8_t* dst, long long dpitch, cudaStream_t st)
{
    BackwardDev B{};
    for (int i = 0; i < 9; ++i) B.r_kinv[i] = r_kinv[i];
    B.scale = scale;
    B.spherical = spherical;
    B.tlx = tlx;
    B.tly = tly;
    dim3 grid((dw + 31) / 32, (dh + 7) / 8);
    if (ch == 1) warp_backward_kernel<1><<<grid, 256, 0, st>>>(warped, B, interp, border, dw, dh, dst, dpitch);
    else warp_backward_kernel<3><<<grid, 256, 0, st>>>(warped, B, interp, border, dw, dh, dst, dpitch);
    ISB_COUNT_LAUNCH();
}

void launch_warp_generic(const ImageDev& img, int ch, int interp, int border, uint8_t* dst, long long dpitch,
                         cudaStream_t st)
{
    dim3 grid((img.roi_w + 31) / 32, (img.roi_h + 7) / 8);
    if (ch == 3) warp_generic_kernel<3><<<grid, 256, 0, st>>>(img, interp, border, dst, dpitch);
    else warp_generic_kernel<1><<<grid, 256, 0, st>>>(img, interp, border, dst, dpitch);
    ISB_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) build_maps_kernel(ImageDev I, float* __restrict__ xmap, float* __restrict__ ymap,
                                                         long long pitch_bytes)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= I.roi_w || y >= I.roi_h) return;
    const XY m = inverse_map(I.kr, I.col[x], I.row[y]);
    reinterpret_cast<float*>(reinterpret_cast<char*>(xmap) + y * pitch_bytes)[x] = m.x;
    reinterpret_cast<float*>(reinterpret_cast<char*>(ymap) + y * pitch_bytes)[x] = m.y;
}

void launch_build_maps(const ImageDev& img, float* xmap, float* ymap, long long pitch_bytes, cudaStream_t st)
{
    dim3 grid((img.roi_w + 31) / 32, (img.roi_h + 7) / 8);
    build_maps_kernel<<<grid, 256, 0, st>>>(img, xmap, ymap, pitch_bytes);
    ISB_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) gain_apply_kernel(uint8_t* __restrict__ img, int w, int h, long long pitch,
                                                         ImageDev I)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const float g = gain_at(I, I.gx[x], I.gy[y]);
    uint8_t* p = img + y * pitch + x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) p[c] = (uint8_t)sat_u8(cv_round(__fmul_rn((float)p[c], g)));
}

void launch_gain_apply(uint8_t* img, int w, int h, long long pitch, const float* gain, int gw, int gh,
                       const LinCoefDev* gx, const LinCoefDev* gy, cudaStream_t st)
{
    ImageDev I{};
    I.gain = gain; I.gw = gw; I.gh = gh; I.gx = gx; I.gy = gy;
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    gain_apply_kernel<<<grid, 256, 0, st>>>(img, w, h, pitch, I);
    ISB_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) dilate3x3_kernel(const uint8_t* __restrict__ src, int w, int h, long long spitch,
                                                        uint8_t* __restrict__ dst)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    int m = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if ((unsigned)yy < (unsigned)h && (unsigned)xx < (unsigned)w) m = max(m, (int)src[yy * spitch + xx]);
        }
    dst[y * w + x] = (uint8_t)m;
}

void launch_dilate3x3(const uint8_t* src, int w, int h, long long spitch, uint8_t* dst, cudaStream_t st)
{
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    dilate3x3_kernel<<<grid, 256, 0, st>>>(src, w, h, spitch, dst);
    ISB_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) seam_and_kernel(const uint8_t* __restrict__ dil, int mw, int mh,
                                                       const uint32_t* __restrict__ mx, const uint32_t* __restrict__ my,
                                                       uint8_t* __restrict__ mask, int w, int h, long long pitch)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    mask[y * pitch + x] &= (uint8_t)seam_at(dil, mw, mh, mx[x], my[y]);
}

void launch_seam_and(const uint8_t* dil, int mw, int mh, const uint32_t* mx, const uint32_t* my, uint8_t* mask, int w,
                     int h, long long pitch, cudaStream_t st)
{
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    seam_and_kernel<<<grid, 256, 0, st>>>(dil, mw, mh, mx, my, mask, w, h, pitch);
    ISB_COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256) pack_tile_kernel(const TileDev* __restrict__ tp, const int16_t* __restrict__ img,
                                                        long long ipitch, const uint8_t* __restrict__ mask,
                                                        long long mpitch)
{
    const TileDev& T = *tp;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= T.w || y >= T.h) return;
    const int rx0 = x - T.left, ry0 = y - T.top;
    const bool in = (unsigned)rx0 < (unsigned)T.roi_w && (unsigned)ry0 < (unsigned)T.roi_h;
    const int rx = reflect(rx0, T.roi_w), ry = reflect(ry0, T.roi_h);
    const int16_t* p = reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(img) + ry * ipitch) + rx * 3;
    int16_t* g = T.G[0] + (long long)y * T.gpitch[0] + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c * T.gplane[0]] = p[c];
    const float inv255 = (float)(1. / 255.);
    T.W[0][(long long)y * T.wpitch[0] + x] = in ? __fmul_rn((float)mask[ry * mpitch + rx], inv255) : 0.f;
}

void launch_pack_tile(const TileDev* tile_dev, const TileDev& tile_host, const int16_t* img, long long ipitch,
                      const uint8_t* mask, long long mpitch, cudaStream_t st)
{
    dim3 grid((tile_host.w + 31) / 32, (tile_host.h + 7) / 8);
    pack_tile_kernel<<<grid, 256, 0, st>>>(tile_dev, img, ipitch, mask, mpitch);
    ISB_COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// ingest pre-steps: cv::rotate(90 CW | 180) and cv::resize(INTER_LINEAR_EXACT), 8U x {1,3} channels
// ------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) rotate_kernel(const uint8_t* __restrict__ src, int w, int h, long long spitch, int code,
                                                     uint8_t* __restrict__ dst, long long dpitch)
{
    // one thread per DESTINATION pixel (coalesced stores); dst is (h x w) for 90 CW, (w x h) for 180
    const int dw = code == 0 ? h : w, dh = code == 0 ? w : h;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const int sx = code == 0 ? y : w - 1 - x, sy = code == 0 ? h - 1 - x : h - 1 - y;
    const uint8_t* p = src + sy * spitch + sx * CH;
    uint8_t* q = dst + y * dpitch + x * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) q[c] = p[c];
}

void launch_rotate(const uint8_t* src, int w, int h, int ch, long long spitch, int code, uint8_t* dst, long long dpitch,
                   cudaStream_t st)
{
    const int dw = code == 0 ? h : w, dh = code == 0 ? w : h;
    dim3 grid((dw + 31) / 32, (dh + 7) / 8);
    if (ch == 3) rotate_kernel<3><<<grid, 256, 0, st>>>(src, w, h, spitch, code, dst, dpitch);
    else rotate_kernel<1><<<grid, 256, 0, st>>>(src, w, h, spitch, code, dst, dpitch);
    ISB_COUNT_LAUNCH();
}

template <int CH>
__global__ void __launch_bounds__(256) resize_exact_kernel(const uint8_t* __restrict__ src, int sw, int sh, long long spitch,
                                                           const uint32_t* __restrict__ tx, const uint32_t* __restrict__ ty,
                                                           uint8_t* __restrict__ dst, int dw, int dh, long long dpitch)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const uint32_t cx = tx[x], cy = ty[y];
    const int c0 = cx >> 16, ax = cx & 0xffff, r0 = cy >> 16, ay = cy & 0xffff;
    const int c1 = min(c0 + 1, sw - 1), r1 = min(r0 + 1, sh - 1);
    const uint8_t* p0 = src + r0 * spitch;
    const uint8_t* p1 = src + r1 * spitch;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int h0 = p0[c0 * CH + c] * (256 - ax) + p0[c1 * CH + c] * ax;
        const int h1 = p1[c0 * CH + c] * (256 - ax) + p1[c1 * CH + c] * ax;
        dst[y * dpitch + x * CH + c] = (uint8_t)((h0 * (256 - ay) + h1 * ay + 32768) >> 16);
    }
}

void launch_resize_exact(const uint8_t* src, int sw, int sh, int ch, long long spitch, const uint32_t* tx, const uint32_t* ty,
                         uint8_t* dst, int dw, int dh, long long dpitch, cudaStream_t st)
{
    dim3 grid((dw + 31) / 32, (dh + 7) / 8);
    if (ch == 3) resize_exact_kernel<3><<<grid, 256, 0, st>>>(src, sw, sh, spitch, tx, ty, dst, dw, dh, dpitch);
    else resize_exact_kernel<1><<<grid, 256, 0, st>>>(src, sw, sh, spitch, tx, ty, dst, dw, dh, dpitch);
    ISB_COUNT_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// plan-time: count valid warped pixels (M of the byte model)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_valid_kernel(const ImageDev* __restrict__ imgs,
                                                          unsigned long long* __restrict__ counts)
{
    const ImageDev& I = imgs[blockIdx.z];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    int cnt = 0;
    if (x < I.roi_w) {
        const F2 c = I.col[x];
        const int y0 = blockIdx.y * 64 + (threadIdx.x >> 5);
        for (int k = 0; k < 8; ++k) {
            const int y = y0 + 8 * k;
            if (y >= I.roi_h) break;
            int ix, iy;
            cnt += nearest_inside(I, inverse_map(I.kr, c, I.row[y]), ix, iy) ? 1 : 0;
        }
    }
    // warp-shuffle reduction, one atomic per warp
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(counts + blockIdx.z, (unsigned long long)cnt);
}

void launch_count_valid(const ImageDev* imgs_dev, int n_img, const int* roi_w_host, const int* roi_h_host,
                        unsigned long long* counts_dev, cudaStream_t st)
{
    int mw = 0, mh = 0;
    for (int i = 0; i < n_img; ++i) {
        mw = max(mw, roi_w_host[i]);
        mh = max(mh, roi_h_host[i]);
    }
    if (mw == 0 || mh == 0) return;
    for (int z0 = 0; z0 < n_img; z0 += 65535) {
        dim3 grid((mw + 31) / 32, (mh + 63) / 64, min(65535, n_img - z0));
        count_valid_kernel<<<grid, 256, 0, st>>>(imgs_dev + z0, counts_dev + z0);
        ISB_COUNT_LAUNCH();
    }
}

// ------------------------------------------------------------------------------------------------
// kernel 2 (generic): pyrDown of both pyramids, one level, any size / storage (odd coarse levels, tiny tiles)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyrdown_tiles_kernel(const WorkItem* __restrict__ work,
                                                            const TileDev* __restrict__ tiles, int l)
{
    pdl_prologue();
    const WorkItem wi = work[blockIdx.x];
    const TileDev& T = tiles[wi.tile];
    const int wl = T.w >> l, hl = T.h >> l, ow = wl >> 1, oh = hl >> 1;
    const int ox = wi.bx * kDownBlockW + (threadIdx.x & 31);
    const int oy = wi.by * kDownBlockH + (threadIdx.x >> 5);
    if (ox >= ow || oy >= oh) return;
    int xs[5], ys[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        xs[i] = reflect101(2 * ox - 2 + i, wl);
        ys[i] = reflect101(2 * oy - 2 + i, hl);
    }
    // 16S planes: integer 5x5 [1 4 6 4 1], (sum + 128) >> 8
    const int gpo = T.gpitch[l + 1];
    int outg[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        int acc = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int h = tile_g(T, l, p, xs[0], ys[j]) + tile_g(T, l, p, xs[4], ys[j]) +
                          4 * (tile_g(T, l, p, xs[1], ys[j]) + tile_g(T, l, p, xs[3], ys[j])) + 6 * tile_g(T, l, p, xs[2], ys[j]);
            acc += (j == 0 || j == 4) ? h : ((j == 2) ? 6 * h : 4 * h);
        }
        outg[p] = (acc + 128) >> 8;
    }
    if (T.packed) T.P[l + 1][oy * T.ppitch[l + 1] + ox] = (uint32_t)outg[0] | ((uint32_t)outg[1] << 8) | ((uint32_t)outg[2] << 16);
    else {
#pragma unroll
        for (int p = 0; p < 3; ++p) T.G[l + 1][p * T.gplane[l + 1] + (long long)oy * gpo + ox] = (int16_t)outg[p];
    }
    // weight plane
    int width0 = (wl - 3) / 2 + 1;
    width0 = min(width0, ow);
    const int simd_h_end = width0 >= 1 ? 1 + 4 * ((width0 - 1) / 4) : 0;
    const bool simd_h = ox >= 1 && ox < simd_h_end;
    const bool simd_v = ox < 4 * (ow / 4);
    float h[5];
#pragma unroll
    for (int j = 0; j < 5; ++j)
        h[j] = wdown_h(tile_w(T, l, xs[0], ys[j]), tile_w(T, l, xs[1], ys[j]), tile_w(T, l, xs[2], ys[j]),
                       tile_w(T, l, xs[3], ys[j]), tile_w(T, l, xs[4], ys[j]), simd_h);
    T.W[l + 1][(long long)oy * T.wpitch[l + 1] + ox] = wdown_v(h[0], h[1], h[2], h[3], h[4], simd_v);
}

void launch_pyrdown_tiles(const WorkItem* work, int n_work, const TileDev* tiles, int level, cudaStream_t st)
{
    if (n_work <= 0) return;
    launch_chained(pyrdown_tiles_kernel, dim3(n_work), dim3(256), 0, st, work, tiles, level);
}

// ------------------------------------------------------------------------------------------------
// kernel 3 (generic): per destination pixel of one level (the coarsest level; classic path fall-back)
//   lap  = sum over covering tiles (feed order) trunc16( sat16(G_l - pyrUp(G_{l+1})) * W_l )     [wraps like short +=]
//   wsum = sum W_l                                                                                [float, feed order]
//   v    = trunc16( lap / (wsum + 1e-5) )  ; v = sat16( pyrUp(C_{l+1}) + v )                      [collapse]
//   level 0: mask = wsum > 1e-5 ; out = mask ? v : 0 ; out8 = sat_u8(out)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blend_level_kernel(DstDev D, const TileDev* __restrict__ tiles, int l, OutDev O)
{
    pdl_prologue();
    const int pw = D.pw >> l, ph = D.ph >> l;
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    // only level 0 is restricted to the rows this process owns; coarser levels cover the whole (sub-)panorama
    const int y = (l == 0 ? D.row0 : 0) + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= pw || y >= (l == 0 ? min(ph, D.row1) : ph)) return;
    const int sh = D.nb - l;
    const int cell = (y >> sh) * D.cells_x + (x >> sh);
    int acc0 = 0, acc1 = 0, acc2 = 0;
    float wsum = 0.f;
    const int e1 = D.cell_start[cell + 1];
    for (int e = D.cell_start[cell]; e < e1; ++e) {
        const TileDev& T = tiles[D.cell_tiles[e]];
        const int lx = x - (T.x0 >> l), ly = y - (T.y0 >> l);
        const float w = tile_w(T, l, lx, ly);
        if (w != 0.f) {
            int v0 = tile_g(T, l, 0, lx, ly), v1 = tile_g(T, l, 1, lx, ly), v2 = tile_g(T, l, 2, lx, ly);
            if (l < D.nb) {
                v0 = sat_s16(v0 - pyrup_tile_at(T, l + 1, 0, lx, ly));
                v1 = sat_s16(v1 - pyrup_tile_at(T, l + 1, 1, lx, ly));
                v2 = sat_s16(v2 - pyrup_tile_at(T, l + 1, 2, lx, ly));
            }
            acc0 += trunc_s16(__fmul_rn((float)v0, w));
            acc1 += trunc_s16(__fmul_rn((float)v1, w));
            acc2 += trunc_s16(__fmul_rn((float)v2, w));
            wsum = __fadd_rn(wsum, w);
        }
    }
    const float den = __fadd_rn(wsum, 1e-5f);
    int r0 = trunc_s16(__fdiv_rn((float)(short)acc0, den));
    int r1 = trunc_s16(__fdiv_rn((float)(short)acc1, den));
    int r2 = trunc_s16(__fdiv_rn((float)(short)acc2, den));
    if (l < D.nb) {
        int up[3];
        pyrup_c_at(D.C[l + 1], D.cpitch[l + 1], D.pw >> (l + 1), D.ph >> (l + 1), x, y, up);
        r0 = sat_s16(up[0] + r0);
        r1 = sat_s16(up[1] + r1);
        r2 = sat_s16(up[2] + r2);
    }
    if (l > 0) {
        D.C[l][y * D.cpitch[l] + x] = c_pack(r0, r1, r2);
        return;
    }
    if (x >= D.fw || y >= D.fh) return;
    const bool on = wsum > 1e-5f;
    if (!on) r0 = r1 = r2 = 0;
    if (O.out8) {
        uint8_t* p = O.out8 + y * O.pitch8 + x * 3;
        p[0] = (uint8_t)sat_u8(r0); p[1] = (uint8_t)sat_u8(r1); p[2] = (uint8_t)sat_u8(r2);
    }
    if (O.mask) O.mask[y * O.mpitch + x] = on ? 255 : 0;
    if (O.out16) {
        int16_t* p = reinterpret_cast<int16_t*>(reinterpret_cast<char*>(O.out16) + y * O.pitch16) + x * 3;
        p[0] = (int16_t)r0; p[1] = (int16_t)r1; p[2] = (int16_t)r2;
    }
}

// ---- depth conversions of the loop: convertTo(CV_16S) (image_stitching.cpp:1164) and the saturate to 8U of imwrite (:1228) ----
__global__ void __launch_bounds__(256) convert_8u16s_kernel(const uint8_t* __restrict__ src, long long spitch, int16_t* __restrict__ dst,
                                                            long long dpitch_bytes, int row_elems, int h)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= row_elems || y >= h) return;
    reinterpret_cast<int16_t*>(reinterpret_cast<char*>(dst) + (long long)y * dpitch_bytes)[x] = src[(long long)y * spitch + x];
}
__global__ void __launch_bounds__(256) convert_16s8u_kernel(const int16_t* __restrict__ src, long long spitch_bytes, uint8_t* __restrict__ dst,
                                                            long long dpitch, int row_elems, int h)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= row_elems || y >= h) return;
    dst[(long long)y * dpitch + x] = (uint8_t)sat_u8(reinterpret_cast<const int16_t*>(reinterpret_cast<const char*>(src) + (long long)y * spitch_bytes)[x]);
}
void launch_convert_8u16s(const uint8_t* src, long long spitch, int16_t* dst, long long dpitch_bytes, int row_elems, int h, cudaStream_t st)
{
    if (row_elems <= 0 || h <= 0) return;
    convert_8u16s_kernel<<<dim3((row_elems + 255) / 256, h), 256, 0, st>>>(src, spitch, dst, dpitch_bytes, row_elems, h);
    ISB_COUNT_LAUNCH();
}
void launch_convert_16s8u(const int16_t* src, long long spitch_bytes, uint8_t* dst, long long dpitch, int row_elems, int h, cudaStream_t st)
{
    if (row_elems <= 0 || h <= 0) return;
    convert_16s8u_kernel<<<dim3((row_elems + 255) / 256, h), 256, 0, st>>>(src, spitch_bytes, dst, dpitch, row_elems, h);
    ISB_COUNT_LAUNCH();
}

void launch_blend_level(const DstDev& dst, const TileDev* tiles, int level, const OutDev& out, cudaStream_t st)
{
    const int pw = dst.pw >> level;
    const int y0 = level == 0 ? dst.row0 : 0, y1 = level == 0 ? min(dst.ph, dst.row1) : (dst.ph >> level);
    if (y1 <= y0 || pw <= 0) return;
    dim3 grid((pw + 31) / 32, (y1 - y0 + 7) / 8);
    launch_chained(blend_level_kernel, grid, dim3(256), 0, st, dst, tiles, level, out);
}

}  // namespace isb
