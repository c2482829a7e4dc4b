// crop.cu - the reference's crop() (image_stitching/cropper.cpp:116-209, helper checkInteriorExterior :6-104) on the device:
// largest-interior-rectangle heuristic on the composited mask.
//
// crop() only consumes three things of cv::findContours / cv::drawContours:
//  (a) which EXTERNAL contour has the most points (cropper.cpp:141-148; first maximum in findContours' order),
//  (b) the sorted x and the sorted y values of that contour's points, with multiplicity (:157-160),
//  (c) the filled contour (:153), probed along the four sides of a shrinking rectangle (:6-104, :171-205).
// None of them needs the sequential border-following trace:
//  * an external contour exists per 8-connected component that touches the OUTSIDE background - the 4-connected background
//    region connected to the image frame (findContours works on a zero-padded copy);
//  * CHAIN_APPROX_NONE emits a pixel once per passage of the border walk, and the walk passes a pixel once per maximal
//    circular run of background pixels in its 8-ring that contains a 4-neighbour and belongs to the outside region
//    (once for an isolated pixel);
//  * findContours lists external contours in reverse raster order of their first pixel, so on equal point counts the
//    component whose first pixel comes LAST in raster order wins;
//  * the filled contour is everything a 4-connected flood from the frame through pixels NOT of that component cannot reach.
// (Each rule is pinned against the reference's own cropper.cpp, compiled into oracle/_ref with cv2 answering the two contour
// calls: tests/test_crop.py.)  So the device does three union-find labellings (8-connected foreground; 4-connected background
// with a frame node; 4-connected "not the chosen component" with a frame node), one pass that counts passages per pixel,
// two histograms, row / column prefix sums of the filled mask's zeros, and then the reference's shrink loop itself as a
// single-warp kernel that jumps over iterations which leave the rectangle unchanged.
#include <algorithm>
#include <climits>
#include <cstring>
#include <vector>

#include "engine.hpp"

namespace isb {

namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;  // pixel outside the labelled set
// node 0 is the frame; pixel i is node i + 1

enum { SET_FG = 0, SET_BG = 1, SET_NOT_COMPONENT = 2 };

struct LabelSrc {
    const uint8_t* mask;      // 8UC1, pitch `pitch`
    long long pitch;
    const uint32_t* fg_parent;  // SET_NOT_COMPONENT: labels of the foreground pass
    uint32_t chosen;            // node of the chosen component's root
};

__device__ __forceinline__ uint32_t root2(const uint32_t* __restrict__ parent, uint32_t node)
{   // after flatten_kernel: a pixel points at its run start, the run start at the root
    return parent[parent[node]];
}

template <int SET>
__device__ __forceinline__ bool in_set(const LabelSrc& S, int W, int x, int y)
{
    const bool fg = S.mask[(long long)y * S.pitch + x] != 0;
    if (SET == SET_FG) return fg;
    if (SET == SET_BG) return !fg;
    if (!fg) return true;
    return root2(S.fg_parent, (uint32_t)((long long)y * W + x) + 1u) != S.chosen;
}

// one CTA per row: parent[pixel] = node of the first pixel of its horizontal run (kNone outside the set)
template <int SET>
__global__ void __launch_bounds__(256) init_runs_kernel(LabelSrc S, int W, int H, uint32_t* __restrict__ parent)
{
    const int y = blockIdx.x;
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = -1;  // x of the last pixel NOT in the set seen so far
    if (y == 0 && threadIdx.x == 0) parent[0] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int x0 = 0; x0 < W; x0 += 256) {
        const int x = x0 + (int)threadIdx.x;
        const bool in = x < W && in_set<SET>(S, W, x, y);
        int last = (x < W && !in) ? x : -1;  // running maximum over x' <= x of the positions outside the set
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, last, d);
            if (lane >= d) last = max(last, o);
        }
        if (lane == 31) s_warp[warp] = last;
        __syncthreads();
        int before = s_carry;
        for (int k = 0; k < warp; ++k) before = max(before, s_warp[k]);
        last = max(last, before);
        if (x < W) parent[(long long)y * W + x + 1] = in ? (uint32_t)((long long)y * W + last + 1) + 1u : kNone;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = last;
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t find_root(const uint32_t* parent, uint32_t a)
{
    uint32_t p = parent[a];
    while (p != a) {
        a = p;
        p = parent[a];
    }
    return a;
}

__device__ void unite(uint32_t* parent, uint32_t a, uint32_t b)
{   // the smaller node becomes the root: the root of a component is its raster-first pixel, the frame (0) wins everything
    for (;;) {
        a = find_root(parent, a);
        b = find_root(parent, b);
        if (a == b) return;
        if (a > b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicMin(parent + b, a);
        if (old == b) return;
        b = old;
    }
}

// vertical (and, for 8-connectivity, diagonal) links between the runs of adjacent rows; frame links.  A link is skipped when
// the pixel to the left makes the same one (its run is this pixel's run).
template <int SET>
__global__ void __launch_bounds__(256) link_kernel(int W, int H, uint32_t* __restrict__ parent)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const long long i = (long long)y * W + x;
    const uint32_t me = (uint32_t)i + 1u;
    if (parent[me] == kNone) return;
    const bool w_in = x > 0 && parent[me - 1] != kNone;
    if (SET != SET_FG) {
        if (x == 0 || x == W - 1 || ((y == 0 || y == H - 1) && !w_in)) unite(parent, me, 0u);
    }
    if (y == 0) return;
    const uint32_t up = me - (uint32_t)W;
    const bool n_in = parent[up] != kNone;
    const bool nw_in = x > 0 && parent[up - 1] != kNone;
    if (n_in) {
        if (!(w_in && nw_in)) unite(parent, me, up);
    } else if (SET == SET_FG) {
        if (nw_in && !w_in) unite(parent, me, up - 1);
        if (x + 1 < W && parent[up + 1] != kNone) unite(parent, me, up + 1);
    }
}

// run starts point at their root; every other pixel keeps pointing at its run start (root2 reads through both)
__global__ void __launch_bounds__(256) flatten_kernel(int W, int H, uint32_t* __restrict__ parent)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint32_t me = (uint32_t)((long long)y * W + x) + 1u;
    if (parent[me] == kNone) return;
    if (x > 0 && parent[me - 1] != kNone) return;  // not a run start
    parent[me] = find_root(parent, me);
}

// passages of the external border walk per foreground pixel -> visits (0..4); summed per component into cnt[root]
__global__ void __launch_bounds__(256) visits_kernel(const uint8_t* __restrict__ mask, long long pitch, int W, int H,
                                                     const uint32_t* __restrict__ fgp, const uint32_t* __restrict__ bgp,
                                                     uint8_t* __restrict__ visits, uint32_t* __restrict__ cnt)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const long long i = (long long)y * W + x;
    if (mask[(long long)y * pitch + x] == 0) {
        visits[i] = 0;
        return;
    }
    // ring in circular order; odd positions are the 4-neighbours
    const int dx[8] = {-1, 0, 1, 1, 1, 0, -1, -1}, dy[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    bool bg[8], out[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int xx = x + dx[k], yy = y + dy[k];
        if (xx < 0 || yy < 0 || xx >= W || yy >= H) {
            bg[k] = out[k] = true;  // the zero frame findContours adds
            continue;
        }
        bg[k] = mask[(long long)yy * pitch + xx] == 0;
        out[k] = false;
        if (bg[k] && (k & 1)) out[k] = root2(bgp, (uint32_t)((long long)yy * W + xx) + 1u) == 0u;
    }
    int v = 0;
    bool all = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) all = all && bg[k];
    if (all) v = out[1] ? 1 : 0;
    else {
#pragma unroll
        for (int j = 1; j < 8; j += 2)
            if (bg[j] && out[j] && !(bg[(j + 7) & 7] && bg[(j + 6) & 7])) ++v;
    }
    visits[i] = (uint8_t)v;
    if (v) atomicAdd(cnt + root2(fgp, (uint32_t)i + 1u), (uint32_t)v);
}

// the external contour with the most points; ties: the component whose first pixel is last in raster order
__global__ void __launch_bounds__(256) pick_kernel(long long n_nodes, const uint32_t* __restrict__ cnt, unsigned long long* __restrict__ best)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    unsigned long long key = 0;
    if (i < n_nodes && cnt[i]) key = ((unsigned long long)cnt[i] << 32) | (unsigned long long)i;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, d));
    if ((threadIdx.x & 31) == 0 && key) atomicMax(best, key);
}

__global__ void __launch_bounds__(256) hist_kernel(int W, int H, const uint32_t* __restrict__ fgp, uint32_t chosen,
                                                   const uint8_t* __restrict__ visits, uint32_t* __restrict__ hx, uint32_t* __restrict__ hy)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    unsigned v = 0;
    if (x < W) {
        const long long i = (long long)y * W + x;
        v = visits[i];
        if (v && root2(fgp, (uint32_t)i + 1u) != chosen) v = 0;
        if (v) atomicAdd(hx + x, v);
    }
    // all threads of a CTA share the row: one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(hy + y, v);
}

// zeros of the filled contour: rowz[y * (W + 1) + x] = number of zeros in row y, columns [0, x); one CTA per row
__global__ void __launch_bounds__(256) row_prefix_kernel(int W, int H, const uint32_t* __restrict__ ncp, uint32_t* __restrict__ rowz,
                                                         uint8_t* __restrict__ zero)
{
    const int y = blockIdx.x;
    __shared__ unsigned s_warp[8];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* __restrict__ o = rowz + (long long)y * (W + 1);
    for (int x0 = 0; x0 < W; x0 += 256) {
        const int x = x0 + (int)threadIdx.x;
        unsigned z = 0;
        if (x < W) {
            const uint32_t me = (uint32_t)((long long)y * W + x) + 1u;
            // outside the chosen component's filled contour <=> the flood from the frame through the other pixels reaches it
            z = (ncp[me] != kNone && root2(ncp, me) == 0u) ? 1u : 0u;
            zero[(long long)y * W + x] = (uint8_t)z;
        }
        unsigned s = z;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) s_warp[warp] = s;
        __syncthreads();
        unsigned before = s_carry;
        for (int k = 0; k < warp; ++k) before += s_warp[k];
        if (x < W) o[x] = before + s - z;  // exclusive
        __syncthreads();
        if (threadIdx.x == 255) s_carry = before + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) o[W] = s_carry;
}

// colz[x * (H + 1) + y] = number of zeros in column x, rows [0, y): a thread walks down one column (coalesced reads)
__global__ void __launch_bounds__(256) col_prefix_kernel(int W, int H, const uint8_t* __restrict__ zero, uint32_t* __restrict__ colz)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    uint32_t* __restrict__ o = colz + (long long)x * (H + 1);
    unsigned s = 0;
    for (int y = 0; y < H; ++y) {
        o[y] = s;
        s += zero[(long long)y * W + x];
    }
    o[H] = s;
}

// The shrink loop of crop() (cropper.cpp:171-205) on one thread, iterations that leave the rectangle unchanged taken in
// one jump.  cx / cy: inclusive prefix sums of the histograms (cx[x] = number of contour points with x' <= x).
struct CropLoopOut { int rect[4]; int iterations; int jumps; int points; };

__device__ __forceinline__ unsigned row_zeros(const uint32_t* rowz, int W, int H, int y, int x0, int n)
{   // zeros among the n pixels at (y, x0 ..) read with Mat::at's pointer arithmetic; outside the buffer counts as zero
    if (n <= 0) return 0;
    if (y < 0 || y >= H) return (unsigned)n;
    const uint32_t* r = rowz + (long long)y * (W + 1);
    return r[x0 + n] - r[x0];
}
__device__ __forceinline__ unsigned col_zeros(const uint32_t* colz, int W, int H, int x, int y0, int n)
{
    if (n <= 0) return 0;
    if (x < 0) {  // Mat::at(y, -1) of a view starting in column 0: the last pixel of the row above
        unsigned z = 0;
        int ya = y0 - 1, m = n;
        if (ya < 0) {  // before the buffer
            z = 1;
            ya = 0;
            --m;
        }
        return z + (m > 0 ? col_zeros(colz, W, H, W - 1, ya, m) : 0u);
    }
    const uint32_t* c = colz + (long long)x * (H + 1);
    return c[y0 + n] - c[y0];
}

__global__ void crop_loop_kernel(int W, int H, const uint32_t* __restrict__ cx, const uint32_t* __restrict__ cy,
                                 const uint32_t* __restrict__ rowz, const uint32_t* __restrict__ colz, CropLoopOut* __restrict__ out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long P = cx[W - 1];  // number of contour points (== cy[H - 1])
    CropLoopOut o{};
    o.points = (int)min(P, (long long)INT_MAX);
    long long a = 0, b = P - 1, c = 0, d = P - 1;
    // values at the four indices: sx[a] = smallest x with cx[x] > a, ...
    int xa = 0, xb = W - 1, yc = 0, yd = H - 1;
    if (P > 0) {
        while (cx[xa] <= a) ++xa;
        while (xb > 0 && cx[xb - 1] > b) --xb;
        while (cy[yc] <= c) ++yc;
        while (yd > 0 && cy[yd - 1] > d) --yd;
    }
    while (a < b && c < d) {
        const int rx = xa, ry = yc, rw = xb - xa, rh = yd - yc;
        o.rect[0] = rx; o.rect[1] = ry; o.rect[2] = rw; o.rect[3] = rh;
        const unsigned top = row_zeros(rowz, W, H, ry, rx, rw), bottom = row_zeros(rowz, W, H, ry + rh - 1, rx, rw);
        const unsigned left = col_zeros(colz, W, H, rx, ry, rh), right = col_zeros(colz, W, H, rx + rw - 1, ry, rh);
        ++o.jumps;
        if (!(top | bottom | left | right)) break;
        int oc_t = 0, oc_b = 0, oc_l = 0, oc_r = 0;
        if (top > bottom) {
            if (top > left && top > right) oc_t = 1;
        } else if (bottom > left) {
            if (bottom > right) oc_b = 1;
        }
        if (left >= right) {
            if (left >= bottom && left >= top) oc_l = 1;
        } else if (right >= top) {
            if (right >= bottom) oc_r = 1;
        }
        // iterations until one of the moving sides reaches the next distinct coordinate (the rectangle - and with it the four
        // counts and the flags - stay the same until then); at least one flag is set whenever a count is non-zero
        long long k = LLONG_MAX;
        if (oc_l) k = min(k, (long long)cx[xa] - a);
        if (oc_r) k = min(k, b - (xb > 0 ? (long long)cx[xb - 1] : 0ll) + 1);
        if (oc_t) k = min(k, (long long)cy[yc] - c);
        if (oc_b) k = min(k, d - (yd > 0 ? (long long)cy[yd - 1] : 0ll) + 1);
        if (k == LLONG_MAX) break;  // cannot happen (see above); never spin
        // the loop condition is evaluated before every iteration: if it fails on the way the last rectangle stands
        long long kk = k;
        if (oc_l || oc_r) kk = min(kk, (b - a + (oc_l + oc_r) - 1) / (oc_l + oc_r));  // iterations until a >= b
        if (oc_t || oc_b) kk = min(kk, (d - c + (oc_t + oc_b) - 1) / (oc_t + oc_b));
        o.iterations += (int)min(kk, (long long)INT_MAX);
        a += kk * oc_l; b -= kk * oc_r; c += kk * oc_t; d -= kk * oc_b;
        if (!(a < b && c < d)) break;
        if (oc_l) while (cx[xa] <= a) ++xa;
        if (oc_r) while (xb > 0 && cx[xb - 1] > b) --xb;
        if (oc_t) while (cy[yc] <= c) ++yc;
        if (oc_b) while (yd > 0 && cy[yd - 1] > d) --yd;
    }
    *out = o;
}

// mask = cvtColor(convertTo(source, CV_8U), COLOR_RGB2GRAY) > 0  (cropper.cpp:118-124): 8-bit fixed-point grey, 14-bit shift
template <typename T>
__global__ void __launch_bounds__(256) gray_mask_kernel(const T* __restrict__ img, long long pitch_bytes, int W, int H, uint8_t* __restrict__ mask)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const T* p = reinterpret_cast<const T*>(reinterpret_cast<const char*>(img) + (long long)y * pitch_bytes) + 3 * x;
    const int c0 = min(max((int)p[0], 0), 255), c1 = min(max((int)p[1], 0), 255), c2 = min(max((int)p[2], 0), 255);
    mask[(long long)y * W + x] = ((c0 * 4899 + c1 * 9617 + c2 * 1868 + (1 << 13)) >> 14) > 0 ? 255 : 0;
}

template <int SET>
void label(const LabelSrc& S, int W, int H, uint32_t* parent, cudaStream_t st)
{
    const dim3 grid((W + 255) / 256, H);
    init_runs_kernel<SET><<<H, 256, 0, st>>>(S, W, H, parent);
    link_kernel<SET><<<grid, 256, 0, st>>>(W, H, parent);
    flatten_kernel<<<grid, 256, 0, st>>>(W, H, parent);
    count_launch(); count_launch(); count_launch();
}

}  // namespace

void crop_rect(const uint8_t* mask, int W, int H, size_t pitch, int rect_xywh[4], int* n_points)
{
    require_device();
    if (!mask || !rect_xywh) throw Error(ISB_ERR_NULL_PTR, "mask/rect are null");
    ISB_ASSERT(W > 0 && H > 0 && pitch >= (size_t)W);
    ISB_ASSERT((long long)W * H < 0xFFFFFFF0ll);  // 32-bit node ids
    cudaStream_t st = current_stream();
    const long long N = (long long)W * H;
    DevBuf mbuf, fgp, bgp, ncp, vis, cnt, hist, rowz, colz, zero, small;
    const uint8_t* m = mask;
    long long mp = (long long)pitch;
    if (mem_kind(mask) != MemKind::Device) {
        void* p = mbuf.ensure((size_t)N);
        copy2d(p, W, mask, pitch, W, H, st);
        m = static_cast<const uint8_t*>(p);
        mp = W;
    }
    const dim3 grid((W + 255) / 256, H);
    uint32_t* fg = static_cast<uint32_t*>(fgp.ensure((size_t)(N + 1) * 4));
    uint32_t* bg = static_cast<uint32_t*>(bgp.ensure((size_t)(N + 1) * 4));
    label<SET_FG>(LabelSrc{m, mp, nullptr, 0}, W, H, fg, st);
    label<SET_BG>(LabelSrc{m, mp, nullptr, 0}, W, H, bg, st);
    uint8_t* v = static_cast<uint8_t*>(vis.ensure((size_t)N));
    uint32_t* c = static_cast<uint32_t*>(cnt.ensure((size_t)(N + 1) * 4));
    ISB_CUDA(cudaMemsetAsync(c, 0, (size_t)(N + 1) * 4, st));
    visits_kernel<<<grid, 256, 0, st>>>(m, mp, W, H, fg, bg, v, c);
    count_launch();
    char* sm = static_cast<char*>(small.ensure(256));
    unsigned long long* best = reinterpret_cast<unsigned long long*>(sm);
    CropLoopOut* lo = reinterpret_cast<CropLoopOut*>(sm + 64);
    ISB_CUDA(cudaMemsetAsync(sm, 0, 256, st));
    pick_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(N + 1, c, best);
    count_launch();
    unsigned long long key = 0;
    ISB_CUDA(cudaMemcpyAsync(&key, best, sizeof(key), cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    if (key == 0) throw Error(ISB_ERR_OUT_OF_RANGE, "crop: the mask holds no contour (the reference's contours.at(0) throws, cropper.cpp:151)");
    const uint32_t chosen = (uint32_t)(key & 0xFFFFFFFFull);
    if (n_points) *n_points = (int)(key >> 32);
    // the background labels are no longer needed: their block takes the third labelling
    uint32_t* nc = bg;
    label<SET_NOT_COMPONENT>(LabelSrc{m, mp, fg, chosen}, W, H, nc, st);
    uint32_t* h = static_cast<uint32_t*>(hist.ensure((size_t)(W + H) * 4));
    ISB_CUDA(cudaMemsetAsync(h, 0, (size_t)(W + H) * 4, st));
    hist_kernel<<<grid, 256, 0, st>>>(W, H, fg, chosen, v, h, h + W);
    count_launch();
    uint32_t* rz = static_cast<uint32_t*>(rowz.ensure((size_t)(W + 1) * H * 4));
    uint32_t* cz = static_cast<uint32_t*>(colz.ensure((size_t)(H + 1) * W * 4));
    uint8_t* z = static_cast<uint8_t*>(zero.ensure((size_t)N));
    row_prefix_kernel<<<H, 256, 0, st>>>(W, H, nc, rz, z);
    col_prefix_kernel<<<(W + 255) / 256, 256, 0, st>>>(W, H, z, cz);
    count_launch(); count_launch();
    // inclusive prefix sums of the two histograms: O(W + H) on the host
    std::vector<uint32_t> hh((size_t)W + H);
    ISB_CUDA(cudaMemcpyAsync(hh.data(), h, hh.size() * 4, cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    for (int x = 1; x < W; ++x) hh[x] += hh[x - 1];
    for (int y = 1; y < H; ++y) hh[(size_t)W + y] += hh[(size_t)W + y - 1];
    ISB_CUDA(cudaMemcpyAsync(h, hh.data(), hh.size() * 4, cudaMemcpyHostToDevice, st));
    crop_loop_kernel<<<1, 32, 0, st>>>(W, H, h, h + W, rz, cz, lo);
    count_launch();
    CropLoopOut out{};
    ISB_CUDA(cudaMemcpyAsync(&out, lo, sizeof(out), cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    ISB_CUDA(cudaGetLastError());
    std::memcpy(rect_xywh, out.rect, sizeof(out.rect));
}

void crop_rect_image(const void* img, int W, int H, size_t pitch, int is_16s, int rect_xywh[4], int* n_points)
{
    require_device();
    if (!img || !rect_xywh) throw Error(ISB_ERR_NULL_PTR, "image/rect are null");
    const size_t es = is_16s ? 6 : 3;
    ISB_ASSERT(W > 0 && H > 0 && pitch >= (size_t)W * es);
    cudaStream_t st = current_stream();
    DevBuf ibuf, mbuf;
    const void* d = img;
    size_t dp = pitch;
    if (mem_kind(img) != MemKind::Device) {
        dp = (size_t)W * es;
        void* p = ibuf.ensure(dp * H);
        copy2d(p, dp, img, pitch, dp, H, st);
        d = p;
    }
    uint8_t* m = static_cast<uint8_t*>(mbuf.ensure((size_t)W * H));
    const dim3 grid((W + 255) / 256, H);
    if (is_16s) gray_mask_kernel<int16_t><<<grid, 256, 0, st>>>(static_cast<const int16_t*>(d), (long long)dp, W, H, m);
    else gray_mask_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(d), (long long)dp, W, H, m);
    count_launch();
    crop_rect(m, W, H, (size_t)W, rect_xywh, n_points);
}

}  // namespace isb
