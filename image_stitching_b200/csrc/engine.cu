// engine.cu - host side of the device pipeline (memory, planning, launches).
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <cuda.h>

#include "kernels.cuh"

namespace isb {

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// ------------------------------------------------------------------------------------------------
// plumbing
// ------------------------------------------------------------------------------------------------
void cuda_check(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return;
    cudaGetLastError();  // clear the sticky-free error state
    throw Error(e == cudaErrorMemoryAllocation ? ISB_ERR_NO_MEM : ISB_ERR_GPU_API,
                std::string(what) + ": " + cudaGetErrorString(e));
}

static thread_local cudaStream_t t_stream = nullptr;
cudaStream_t current_stream() { return t_stream; }
void set_current_stream(cudaStream_t s) { t_stream = s; }

void require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        throw Error(ISB_ERR_GPU_API,
                    "no CUDA device available: image_stitching_b200 has no CPU fallback (the compositing path runs "
                    "only as sm_100a kernels)");
    }
}

MemKind mem_kind(const void* p)
{
    cudaPointerAttributes a{};
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return MemKind::Host;
    }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return MemKind::Device;
    if (a.type == cudaMemoryTypeHost) return MemKind::HostPinned;
    return MemKind::Host;
}

DevBuf::~DevBuf()
{
    if (p_) cudaFree(p_);
}
void DevBuf::release()
{
    if (p_) cudaFree(p_);
    p_ = nullptr;
    cap_ = 0;
}
void* DevBuf::ensure(size_t bytes)
{
    if (bytes <= cap_ && p_) return p_;
    if (p_) {
        ISB_CUDA(cudaFree(p_));  // synchronises with in-flight work that may still read the old block
        p_ = nullptr;
        cap_ = 0;
    }
    size_t want = std::max<size_t>(bytes, 256);
    ISB_CUDA(cudaMalloc(&p_, want));
    cap_ = want;
    return p_;
}

// copy a (rows x row_bytes) block between any two memory kinds on `st`
void copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows, cudaStream_t st)
{
    if (rows == 0 || row_bytes == 0) return;
    if (dpitch == row_bytes && spitch == row_bytes)
        ISB_CUDA(cudaMemcpyAsync(dst, src, row_bytes * rows, cudaMemcpyDefault, st));
    else
        ISB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyDefault, st));
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------
// PyramidEngine
// ------------------------------------------------------------------------------------------------
PyramidEngine::~PyramidEngine()
{
    for (DevBuf* b : extra_) delete b;
}

void PyramidEngine::reset(const BlendGeometry& g, int sub_y0, int sub_h, int own_y0, int own_y1, bool packed)
{
    g_ = g;
    sub_y0_ = sub_y0;
    sub_h_ = sub_h;
    own_y0_ = own_y0;
    own_y1_ = own_y1;
    packed_ = packed;
    tiles_.clear();
    committed_ = 0;
    arena_.begin();
    // classic feed() path: the per-tile allocations of the previous prepare/feed/blend cycle (cudaFree synchronises with
    // whatever may still read them)
    for (DevBuf* b : extra_) delete b;
    extra_.clear();
    img_occ_.clear();
    warp_work_.clear();
    pad_work_.clear();
    down_work_.assign(g.nb, {});
    dst_.cell_start = nullptr;
    ISB_ASSERT(g.nb < kMaxLevels);
}

int PyramidEngine::add_tile(int img_index, int w, int h, int tlx, int tly)
{
    int tl[2], br[2];
    g_.tile_rect(w, h, tlx, tly, tl, br);
    return add_rect(img_index, tl[0] - g_.roi.x, tl[1] - g_.roi.y, br[0] - tl[0], br[1] - tl[1], tlx, tly, w, h);
}

void PyramidEngine::set_image_occupancy(int img, std::vector<uint8_t> grid, int cw, int ch, int gx0, int gy0)
{
    if (img < 0) return;
    if ((int)img_occ_.size() <= img) img_occ_.resize(img + 1);
    OccGrid& G = img_occ_[img];
    G.g = std::move(grid);
    G.cw = cw; G.ch = ch; G.x0 = gx0; G.y0 = gy0;
}

int PyramidEngine::add_rect(int img_index, int X0, int Y0, int W, int H, int tlx, int tly, int roi_w, int roi_h,
                            const uint32_t* need_grid, int grid_X0, int grid_Y0, int grid_cw)
{
    const int cy0 = std::max(Y0, sub_y0_), cy1 = std::min(Y0 + H, sub_y0_ + sub_h_);
    if (cy1 <= cy0 || W <= 0) return -1;
    TileDev t{};
    t.x0 = X0;
    t.y0 = cy0 - sub_y0_;
    t.w = W;
    t.h = cy1 - cy0;
    t.img = img_index;
    t.left = (tlx - g_.roi.x) - X0;
    t.top = (tly - g_.roi.y) - cy0;
    t.roi_w = roi_w;
    t.roi_h = roi_h;
    t.packed = packed_ ? 1 : 0;
    if (need_grid && g_.nb >= 2) {  // cell (0, 0) of this tile inside the image's occupancy grid
        t.need = need_grid + (size_t)((cy0 - grid_Y0) >> g_.nb) * grid_cw + ((X0 - grid_X0) >> g_.nb);
        t.need_cw = grid_cw;
    }
    tiles_.push_back(t);
    return (int)tiles_.size() - 1;
}

void PyramidEngine::commit_tiles(cudaStream_t st)
{
    const int first = committed_, end = (int)tiles_.size();
    if (first == end) return;
    const int nb = g_.nb;
    // storage layout of the new tiles
    Arena local;
    Arena& ar = (first == 0) ? arena_ : local;
    ar.begin();
    const size_t nl = (size_t)nb + 1;
    std::vector<size_t> goff((size_t)(end - first) * nl), woff(goff.size());
    for (int t = first; t < end; ++t) {
        TileDev& T = tiles_[t];
        for (int l = 0; l <= nb; ++l) {
            const int wl = T.w >> l, hl = T.h >> l;
            const size_t k = (size_t)(t - first) * nl + l;
            T.wpitch[l] = round_up(wl, 4);
            if (T.packed) {
                T.ppitch[l] = round_up(wl, 4);
                goff[k] = ar.take((size_t)T.ppitch[l] * hl * sizeof(uint32_t));
                woff[k] = l == 0 ? 0 : ar.take((size_t)T.wpitch[l] * hl * sizeof(float));  // level-0 weight = mask byte
            } else {
                T.gpitch[l] = round_up(wl, 8);
                T.gplane[l] = (long long)T.gpitch[l] * hl;
                goff[k] = ar.take((size_t)T.gplane[l] * 3 * sizeof(int16_t));
                woff[k] = ar.take((size_t)T.wpitch[l] * hl * sizeof(float));
            }
        }
    }
    char* base;
    if (first == 0) base = arena_.commit();
    else {
        DevBuf* b = new DevBuf();
        extra_.push_back(b);
        base = static_cast<char*>(b->ensure(local.used()));
    }
    for (int t = first; t < end; ++t) {
        TileDev& T = tiles_[t];
        for (int l = 0; l <= nb; ++l) {
            const size_t k = (size_t)(t - first) * nl + l;
            if (T.packed) T.P[l] = reinterpret_cast<uint32_t*>(base + goff[k]);
            else T.G[l] = reinterpret_cast<int16_t*>(base + goff[k]);
            T.W[l] = (T.packed && l == 0) ? nullptr : reinterpret_cast<float*>(base + woff[k]);
        }
    }
    // work lists for the new tiles
    last_fast_ = packed_ && first == 0;
    for (const TileDev& T : tiles_) last_fast_ = last_fast_ && (T.w >> nb) >= 2;
    warp_work_.clear();
    pad_work_.clear();
    for (auto& v : down_work_) v.clear();
    {   // REFLECT padding (tile pixels outside the warped ROI): mirrored by a second pass when there is enough of it to pay for
        // the extra launch (large pyramids: the 4-cell dependency margin grows with 2^nb), else computed by kernel 1 itself
        double tile_px = 0, roi_px = 0;
        bool fused = end > first;
        for (int t = first; t < end; ++t) {
            const TileDev& T = tiles_[t];
            fused = fused && T.img >= 0;
            tile_px += (double)T.w * T.h;
            const int ix = std::max(0, std::min(T.w, T.left + T.roi_w) - std::max(0, T.left));
            const int iy = std::max(0, std::min(T.h, T.top + T.roi_h) - std::max(0, T.top));
            roi_px += (double)ix * iy;
        }
        pad_fraction_ = tile_px > 0 ? 1.0 - roi_px / tile_px : 0.0;
        mirror_pad_ = fused && pad_fraction_ > kMirrorPadThreshold;
        // per tile: mirrored only if every padding column / row reflects onto a column / row the tile holds
        auto refl = [](int p, int n) {  // cv::BORDER_REFLECT index (device_math.cuh: reflect)
            if (n == 1) return 0;
            if (p < 0) p = -p - 1;
            p %= 2 * n;
            return p < n ? p : 2 * n - 1 - p;
        };
        for (int t = first; t < end; ++t) {
            TileDev& T = tiles_[t];
            bool ok = mirror_pad_;
            for (int x = 0; ok && x < T.w; ++x) {
                const int rx0 = x - T.left;
                if (rx0 >= 0 && rx0 < T.roi_w) { x = std::max(x, std::min(T.w, T.left + T.roi_w) - 1); continue; }
                const int sx = refl(rx0, T.roi_w) + T.left;
                ok = sx >= 0 && sx < T.w;
            }
            for (int y = 0; ok && y < T.h; ++y) {
                const int ry0 = y - T.top;
                if (ry0 >= 0 && ry0 < T.roi_h) { y = std::max(y, std::min(T.h, T.top + T.roi_h) - 1); continue; }
                const int sy = refl(ry0, T.roi_h) + T.top;
                ok = sy >= 0 && sy < T.h;
            }
            T.mirror = ok ? 1 : 0;
        }
        if (getenv("ISB_DEBUG_PLAN"))
            fprintf(stderr, "[isb plan] tiles %d..%d: %.1f MP, %.1f %% outside the warped ROIs -> %s\n", first, end, tile_px / 1e6,
                    100.0 * pad_fraction_, mirror_pad_ ? "mirror_pad_kernel" : "computed by kernel 1");
    }
    for (int t = first; t < end; ++t) {
        const TileDev& T = tiles_[t];
        for (int by = 0; by < (T.h + kWarpBlockH - 1) / kWarpBlockH; ++by)
            for (int bx = 0; bx < (T.w + kWarpBlockW - 1) / kWarpBlockW; ++bx) {
                // fused path: blocks without a pixel of the warped ROI are pure REFLECT padding (mirror_pad_kernel fills them),
                // blocks that are not entirely inside the ROI hold some
                const int x0 = bx * kWarpBlockW - T.left, x1 = std::min((bx + 1) * kWarpBlockW, T.w) - T.left;
                const int y0 = by * kWarpBlockH - T.top, y1 = std::min((by + 1) * kWarpBlockH, T.h) - T.top;
                const bool touches = x1 > 0 && y1 > 0 && x0 < T.roi_w && y0 < T.roi_h;
                const bool inside = x0 >= 0 && y0 >= 0 && x1 <= T.roi_w && y1 <= T.roi_h;
                if (!T.mirror || touches) warp_work_.push_back(WorkItem{t, bx, by, 0});
                if (T.mirror && !inside) pad_work_.push_back(WorkItem{t, bx, by, 0});
            }
        for (int l = 0; l < nb; ++l) {
            const int ow = T.w >> (l + 1), oh = T.h >> (l + 1);
            // levels with an even output width run the register-rolling kernel (64 x 128 outputs per CTA)
            const int bw = fast_down(l) ? kFastDownCols : kDownBlockW;
            const int bh = fast_down(l) ? fast_rows(l) * kFastDownWarps : kDownBlockH;
            for (int by = 0; by < (oh + bh - 1) / bh; ++by)
                for (int bx = 0; bx < (ow + bw - 1) / bw; ++bx) down_work_[l].push_back(WorkItem{t, bx, by, 0});
        }
    }
    // pyrDown of the fused path goes through TMA-staged tiles at every level l <= nb - 2 (even output widths) at which every tile
    // can hold a full box: level 0 stages the packed pixels (their mask byte is the weight), levels >= 1 the pixels and the f32
    // weight plane.  Maps live at [(2 l + kind) * n_tiles + tile], kind 0 = pixels, 1 = weights.
    use_tma_ = false;
    tma_levels_ = 0;
    if (packed_ && first == 0 && nb >= 2 && encode_tiled_fn()) {
        const size_t nt = tiles_.size();
        std::vector<CUtensorMap> maps(2 * (size_t)nb * nt);
        std::memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
        auto encode = [&](CUtensorMap* m, void* base, int w, int h, size_t stride_bytes) {
            const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
            const cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
            const cuuint32_t box[2] = {(cuuint32_t)kTmaBoxW, (cuuint32_t)kTmaBoxH};
            const cuuint32_t estr[2] = {1, 1};
            return encode_tiled_fn()(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        };
        for (int l = 0; l + 2 <= nb && l < env_switches().pyrdown_tma_levels; ++l) {  // consecutive levels from 0: the first one that does not qualify ends the list
            bool ok = true;
            for (size_t t = 0; ok && t < nt; ++t) {
                const TileDev& T = tiles_[t];
                const int wl = T.w >> l, hl = T.h >> l;
                ok = wl >= kTmaBoxW && hl >= kTmaBoxH && encode(&maps[(2 * (size_t)l) * nt + t], T.P[l], wl, hl, (size_t)T.ppitch[l] * sizeof(uint32_t));
                if (ok && l >= 1) ok = encode(&maps[(2 * (size_t)l + 1) * nt + t], T.W[l], wl, hl, (size_t)T.wpitch[l] * sizeof(float));
            }
            if (!ok) break;
            tma_levels_ = l + 1;
            down_work_[l].clear();
            for (int t = 0; t < (int)nt; ++t) {
                const int ow = tiles_[t].w >> (l + 1), oh = tiles_[t].h >> (l + 1);
                for (int by = 0; by < (oh + kTmaOutH - 1) / kTmaOutH; ++by)
                    for (int bx = 0; bx < (ow + kTmaOutW - 1) / kTmaOutW; ++bx) down_work_[l].push_back(WorkItem{t, bx, by, 0});
            }
        }
        if (tma_levels_ > 0) {
            void* md = tmaps_dev_.ensure(maps.size() * sizeof(CUtensorMap));
            ISB_CUDA(cudaMemcpyAsync(md, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
            ISB_CUDA(cudaStreamSynchronize(st));  // `maps` is a local
            use_tma_ = true;
        }
    }
    // uploads (pageable sources: the runtime stages them before returning, so the vectors may be reused)
    TileDev* td = static_cast<TileDev*>(tiles_dev_.ensure(std::max<size_t>(tiles_.size(), 16) * 2 * sizeof(TileDev)));
    ISB_CUDA(cudaMemcpyAsync(td, tiles_.data(), tiles_.size() * sizeof(TileDev), cudaMemcpyHostToDevice, st));
    if (!warp_work_.empty()) {
        void* p = warp_work_dev_.ensure(warp_work_.size() * sizeof(WorkItem));
        ISB_CUDA(cudaMemcpyAsync(p, warp_work_.data(), warp_work_.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, st));
    }
    if (!pad_work_.empty()) {
        void* p = pad_work_dev_.ensure(pad_work_.size() * sizeof(WorkItem));
        ISB_CUDA(cudaMemcpyAsync(p, pad_work_.data(), pad_work_.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, st));
    }
    size_t total = 0;
    down_off_.assign(nb + 1, 0);
    for (int l = 0; l < nb; ++l) {
        down_off_[l] = total;
        total += down_work_[l].size();
    }
    down_off_[nb] = total;
    if (total) {
        WorkItem* p = static_cast<WorkItem*>(down_work_dev_.ensure(total * sizeof(WorkItem)));
        for (int l = 0; l < nb; ++l)
            if (!down_work_[l].empty())
                ISB_CUDA(cudaMemcpyAsync(p + down_off_[l], down_work_[l].data(), down_work_[l].size() * sizeof(WorkItem),
                                         cudaMemcpyHostToDevice, st));
    }
    committed_ = end;
    dst_.cell_start = nullptr;  // CSR must be rebuilt
}

void PyramidEngine::build_pyramids(int first, int end, cudaStream_t st)
{
    (void)first;
    (void)end;  // the work lists always describe the last committed batch
    const WorkItem* base = down_work_dev_.as<WorkItem>();
    for (int l = 0; l < g_.nb; ++l) {
        const int n = (int)(down_off_[l + 1] - down_off_[l]);
        if (l < tma_levels_) {
            const CUtensorMap* m = tmaps_dev_.as<CUtensorMap>() + (2 * (size_t)l) * tiles_.size();
            launch_pyrdown_tma(base + down_off_[l], n, tiles_dev(), m, m + tiles_.size(), l, st);
        }
        else if (fast_down(l)) launch_pyrdown_fast(base + down_off_[l], n, tiles_dev(), l, packed_, fast_rows(l), l + 1 == g_.nb, st);
        else launch_pyrdown_tiles(base + down_off_[l], n, tiles_dev(), l, st);
    }
}

void PyramidEngine::blend(const OutDev& out, cudaStream_t st)
{
    const int nb = g_.nb;
    if (!dst_.cell_start) {
        // destination pyramid (levels 1..nb) + CSR of covering tiles per macro cell
        dst_ = DstDev{};
        dst_.nb = nb;
        dst_.packed0 = packed_ ? 1 : 0;
        dst_.pw = g_.roi.w;
        dst_.ph = sub_h_;
        dst_.fw = g_.roi_final.w;
        dst_.fh = g_.roi_final.h - sub_y0_;
        dst_.row0 = own_y0_ - sub_y0_;
        dst_.row1 = own_y1_ - sub_y0_;
        dst_.cells_x = dst_.pw >> nb;
        dst_.cells_y = dst_.ph >> nb;
        const int ncell = dst_.cells_x * dst_.cells_y;
        std::vector<int> start(ncell + 1, 0);
        for (const TileDev& T : tiles_)
            for (int cy = T.y0 >> nb; cy < (T.y0 + T.h) >> nb; ++cy)
                for (int cx = T.x0 >> nb; cx < (T.x0 + T.w) >> nb; ++cx) ++start[cy * dst_.cells_x + cx + 1];
        dst_.max_cell_tiles = 0;
        for (int i = 0; i < ncell; ++i) dst_.max_cell_tiles = std::max(dst_.max_cell_tiles, start[i + 1]);
        for (int i = 0; i < ncell; ++i) start[i + 1] += start[i];
        std::vector<int> fill(start.begin(), start.end() - 1), list(std::max(start[ncell], 1));
        for (int t = 0; t < (int)tiles_.size(); ++t) {  // ascending t == feed order
            const TileDev& T = tiles_[t];
            for (int cy = T.y0 >> nb; cy < (T.y0 + T.h) >> nb; ++cy)
                for (int cx = T.x0 >> nb; cx < (T.x0 + T.w) >> nb; ++cx) list[fill[cy * dst_.cells_x + cx]++] = t;
        }
        // can tile t carry weight in panorama cell (cx, cy) [sub-panorama cell coordinates] within `dil` cells?
        auto may_weigh = [&](int t, int cx, int cy, int dil) {
            const TileDev& T = tiles_[t];
            if (T.img < 0 || T.img >= (int)img_occ_.size() || img_occ_[T.img].g.empty()) return true;
            const OccGrid& G = img_occ_[T.img];
            const int gx = ((cx << nb) - G.x0) >> nb, gy = ((cy << nb) + sub_y0_ - G.y0) >> nb;
            for (int y = std::max(0, gy - dil); y <= std::min(G.ch - 1, gy + dil); ++y)
                for (int x = std::max(0, gx - dil); x <= std::min(G.cw - 1, gx + dil); ++x)
                    if (G.g[(size_t)y * G.cw + x]) return true;
            return false;
        };
        size_t cbytes = 0;
        std::vector<size_t> coff(nb + 1, 0);
        for (int l = 1; l <= nb; ++l) {
            dst_.cpitch[l] = round_up(dst_.pw >> l, 2);
            coff[l] = cbytes;
            cbytes += ((size_t)dst_.cpitch[l] * (dst_.ph >> l) * sizeof(uint2) + 255) & ~size_t(255);
        }
        char* cb = static_cast<char*>(dst_buf_.ensure(std::max<size_t>(cbytes, 256)));
        for (int l = 1; l <= nb; ++l) dst_.C[l] = reinterpret_cast<uint2*>(cb + coff[l]);
        int* cd = static_cast<int*>(cells_dev_.ensure((start.size() + list.size()) * sizeof(int)));
        ISB_CUDA(cudaMemcpyAsync(cd, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        ISB_CUDA(cudaMemcpyAsync(cd + start.size(), list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        dst_.cell_start = cd;
        dst_.cell_tiles = cd + start.size();
        dst_.cdesc = nullptr;
        dst_.n_entries = start[ncell];
        if (packed_ && nb > 0 && dst_.n_entries > 0) {
            // compact per-(entry, level) records for kernel 3 (see CellTile)
            std::vector<CellTile> desc((size_t)nb * dst_.n_entries);
            for (int l = 0; l < nb; ++l)
                for (int e = 0; e < dst_.n_entries; ++e) {
                    const TileDev& T = tiles_[list[e]];
                    CellTile& c = desc[(size_t)l * dst_.n_entries + e];
                    c.p0 = T.P[l]; c.p1 = T.P[l + 1]; c.w0 = T.W[l];
                    c.pitch0 = T.ppitch[l]; c.pitch1 = T.ppitch[l + 1];
                    ISB_ASSERT(T.wpitch[l] == T.ppitch[l]);
                    c.ox = T.x0 >> l; c.oy = T.y0 >> l;
                    c.wc = T.w >> (l + 1); c.hc = T.h >> (l + 1);
                    c.tile = list[e];
                }
            CellTile* dd = static_cast<CellTile*>(cdesc_dev_.ensure(desc.size() * sizeof(CellTile)));
            ISB_CUDA(cudaMemcpyAsync(dd, desc.data(), desc.size() * sizeof(CellTile), cudaMemcpyHostToDevice, st));
            ISB_CUDA(cudaStreamSynchronize(st));  // `desc` is a local
            dst_.cdesc = dd;
            // level-0 lists without the tiles whose weights are zero in the cell for sure (no valid pixel of the image there)
            dst_.cell_start0 = nullptr;
            dst_.cdesc0 = nullptr;
            {
                std::vector<int> start0(ncell + 1, 0);
                std::vector<CellTile> desc0;
                for (int c = 0; c < ncell; ++c) {
                    start0[c] = (int)desc0.size();
                    for (int e = start[c]; e < start[c + 1]; ++e)
                        if (may_weigh(list[e], c % dst_.cells_x, c / dst_.cells_x, 0)) desc0.push_back(desc[e]);  // level-0 record
                }
                start0[ncell] = (int)desc0.size();
                if (getenv("ISB_DEBUG_PLAN")) {
                    long long covered = 0;
                    for (int c = 0; c < ncell; ++c) covered += start0[c + 1] > start0[c];
                    fprintf(stderr, "[isb plan] blend lists: %d cells (%lld covered), %d entries, %zu after the occupancy filter (%.2f per covered cell)\n",
                            ncell, covered, dst_.n_entries, desc0.size(), covered ? (double)desc0.size() / covered : 0.0);
                }
                int* s0 = static_cast<int*>(cells0_dev_.ensure(start0.size() * sizeof(int)));
                CellTile* d0 = static_cast<CellTile*>(cdesc0_dev_.ensure(std::max<size_t>(desc0.size(), 1) * sizeof(CellTile)));
                ISB_CUDA(cudaMemcpyAsync(s0, start0.data(), start0.size() * sizeof(int), cudaMemcpyHostToDevice, st));
                if (!desc0.empty()) ISB_CUDA(cudaMemcpyAsync(d0, desc0.data(), desc0.size() * sizeof(CellTile), cudaMemcpyHostToDevice, st));
                ISB_CUDA(cudaStreamSynchronize(st));
                dst_.cell_start0 = s0;
                dst_.cdesc0 = d0;
            }
            // tensor maps of the TMA-fed blend kernels (see DstDev) and the per-block tile lists of the pipelined coarse levels
            dst_.tmap_tiles = dst_.tmap_c = nullptr;
            dst_.tmap_level0 = 0;
            dst_.n_tiles = (int)tiles_.size();
            for (int l = 0; l < kMaxLevels; ++l) {
                dst_.blk_start[l] = nullptr;
                dst_.blk_desc[l] = nullptr;
                dst_.blk_nbx[l] = 0;
            }
            if (encode_tiled_fn() && nb >= 5) {
                const size_t nt = tiles_.size(), per_kind = (size_t)(nb + 1) * nt;
                std::vector<CUtensorMap> maps(3 * per_kind + (nb + 1));
                std::memset(maps.data(), 0, maps.size() * sizeof(CUtensorMap));
                const cuuint32_t estr[2] = {1, 1};
                auto encode = [&](CUtensorMap* m, void* base, int w, int h, size_t stride_bytes, int bw, int bh) {
                    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
                    const cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
                    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
                    return w >= bw && h >= bh &&
                           encode_tiled_fn()(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                };
                // a level is usable by a kernel only if EVERY tile (and the collapsed plane) has a valid map of the kinds it reads
                std::vector<char> okq(nb + 1, 1), okp(nb + 1, 1), okw(nb + 1, 1), okc(nb + 1, 1);
                for (int l = 0; l <= nb; ++l) {
                    for (size_t t = 0; t < nt; ++t) {
                        const TileDev& T = tiles_[t];
                        const int wl = T.w >> l, hl = T.h >> l;
                        okq[l] = okq[l] && encode(&maps[(size_t)l * nt + t], T.P[l], wl, hl, (size_t)T.ppitch[l] * 4, 32, 32);
                        okp[l] = okp[l] && l >= 1 && encode(&maps[per_kind + (size_t)l * nt + t], T.P[l], wl, hl, (size_t)T.ppitch[l] * 4, 24, 18);
                        okw[l] = okw[l] && l >= 1 && encode(&maps[2 * per_kind + (size_t)l * nt + t], T.W[l], wl, hl, (size_t)T.wpitch[l] * 4, 32, 32);
                    }
                    okc[l] = l >= 1 && encode(&maps[3 * per_kind + l], dst_.C[l], 2 * (dst_.pw >> l), dst_.ph >> l, (size_t)dst_.cpitch[l] * sizeof(uint2), 40, 18);
                }
                // the cell kernels read the 24 x 18 maps of level l + 1 and C[l + 1] for every level l with cells of >= 32 px
                bool ok = true;
                for (int l = 0; l + 5 <= nb; ++l) ok = ok && okp[l + 1] && okc[l + 1];
                if (ok) {
                    CUtensorMap* md = static_cast<CUtensorMap*>(tmaps_blend_dev_.ensure(maps.size() * sizeof(CUtensorMap)));
                    ISB_CUDA(cudaMemcpyAsync(md, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st));
                    ISB_CUDA(cudaStreamSynchronize(st));  // `maps` is a local
                    dst_.tmap_tiles = md;
                    dst_.tmap_c = md + 3 * per_kind;
                    dst_.tmap_level0 = okq[0] ? 1 : 0;
                    // Pipelined kernel at levels 1 .. nb - 3 (cells of >= 8 px keep every box origin on a 16-byte boundary): per
                    // 32 x 32 block of the level the sorted union of the tile lists of the macro cells under it
                    std::vector<int> bstart;
                    std::vector<CellTile> bdesc;
                    std::vector<size_t> off_s(nb + 1, 0), off_d(nb + 1, 0);
                    std::vector<char> use(nb + 1, 0);
                    for (int l = 1; l + 3 <= nb; ++l) {
                        if (!(okq[l] && okw[l] && okp[l + 1] && okc[l + 1])) continue;
                        const int pwl = dst_.pw >> l, phl = dst_.ph >> l, sh = nb - l;
                        const int nbx = (pwl + 31) / 32, nby = (phl + 31) / 32;
                        if (nbx * nby < 296) continue;  // too few blocks to feed persistent CTAs: the quad kernel serves the level
                        use[l] = 1;
                        dst_.blk_nbx[l] = nbx;
                        off_s[l] = bstart.size();
                        off_d[l] = bdesc.size();
                        std::vector<int> uni;
                        int count = 0;
                        for (int by = 0; by < nby; ++by)
                            for (int bx = 0; bx < nbx; ++bx) {
                                bstart.push_back(count);
                                uni.clear();
                                const int cx0 = (bx * 32) >> sh, cx1 = std::min(dst_.cells_x - 1, (bx * 32 + 31) >> sh);
                                const int cy0 = (by * 32) >> sh, cy1 = std::min(dst_.cells_y - 1, (by * 32 + 31) >> sh);
                                for (int cy = cy0; cy <= cy1; ++cy)
                                    for (int cx = cx0; cx <= cx1; ++cx) {
                                        const int c = cy * dst_.cells_x + cx;
                                        uni.insert(uni.end(), list.begin() + start[c], list.begin() + start[c + 1]);
                                    }
                                std::sort(uni.begin(), uni.end());
                                uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
                                for (int t : uni) {
                                    bool any = false;  // weights of level l spread by less than one cell beyond the valid pixels
                                    for (int cy = cy0; cy <= cy1 && !any; ++cy)
                                        for (int cx = cx0; cx <= cx1 && !any; ++cx) any = may_weigh(t, cx, cy, 1);
                                    if (!any) continue;
                                    const TileDev& T = tiles_[t];
                                    CellTile c{};
                                    c.p0 = T.P[l]; c.p1 = T.P[l + 1]; c.w0 = T.W[l];
                                    c.pitch0 = T.ppitch[l]; c.pitch1 = T.ppitch[l + 1];
                                    c.ox = T.x0 >> l; c.oy = T.y0 >> l;
                                    c.wc = T.w >> (l + 1); c.hc = T.h >> (l + 1);
                                    c.tile = t;
                                    bdesc.push_back(c);
                                }
                                count = (int)(bdesc.size() - off_d[l]);
                            }
                        bstart.push_back(count);
                    }
                    if (!bstart.empty()) {
                        int* sd = static_cast<int*>(blk_start_dev_.ensure(bstart.size() * sizeof(int)));
                        CellTile* dd2 = static_cast<CellTile*>(blk_desc_dev_.ensure(std::max<size_t>(bdesc.size(), 1) * sizeof(CellTile)));
                        ISB_CUDA(cudaMemcpyAsync(sd, bstart.data(), bstart.size() * sizeof(int), cudaMemcpyHostToDevice, st));
                        if (!bdesc.empty())
                            ISB_CUDA(cudaMemcpyAsync(dd2, bdesc.data(), bdesc.size() * sizeof(CellTile), cudaMemcpyHostToDevice, st));
                        ISB_CUDA(cudaStreamSynchronize(st));
                        for (int l = 1; l + 3 <= nb; ++l)
                            if (use[l]) {
                                dst_.blk_start[l] = sd + off_s[l];
                                dst_.blk_desc[l] = dd2 + off_d[l];
                            }
                    }
                }
            }
        }
    }
    for (int l = nb; l >= 0; --l) {
        if (l < nb) launch_blend_quad(dst_, tiles_dev(), l, out, st);
        else launch_blend_level(dst_, tiles_dev(), l, out, st);
    }
}

// ------------------------------------------------------------------------------------------------
// Warper
// ------------------------------------------------------------------------------------------------
static void check_KR(const float* K, const float* R)
{
    if (!K || !R) throw Error(ISB_ERR_NULL_PTR, "K and R must be 3x3 CV_32F matrices (got a null pointer)");
}

Rect Warper::warp_roi(int sw, int sh, const float* K, const float* R)
{
    check_KR(K, R);
    ISB_ASSERT(sw > 0 && sh > 0);
    Projector p;
    p.set(kind_, scale_, K, R);
    return p.warp_roi(sw, sh);
}

void Warper::warp_point(const float* pt, const float* K, const float* R, float* out, bool backward)
{
    check_KR(K, R);
    Projector p;
    p.set(kind_, scale_, K, R);
    if (backward) p.backward(pt[0], pt[1], out[0], out[1]);
    else p.forward(pt[0], pt[1], out[0], out[1]);
}

void Warper::prepare_image(int sw, int sh, const float* K, const float* R, ImageDev& I, Rect& roi, cudaStream_t st)
{
    check_KR(K, R);
    ISB_ASSERT(sw > 0 && sh > 0);
    Projector p;
    p.set(kind_, scale_, K, R);
    roi = p.warp_roi(sw, sh);
    std::vector<Float2> col, row;
    build_trig_tables(p, roi, col, row);
    F2* t = static_cast<F2*>(tab_.ensure((col.size() + row.size()) * sizeof(F2)));
    ISB_CUDA(cudaMemcpyAsync(t, col.data(), col.size() * sizeof(F2), cudaMemcpyHostToDevice, st));
    ISB_CUDA(cudaMemcpyAsync(t + col.size(), row.data(), row.size() * sizeof(F2), cudaMemcpyHostToDevice, st));
    I = ImageDev{};
    I.sw = sw;
    I.sh = sh;
    I.roi_w = roi.w;
    I.roi_h = roi.h;
    std::memcpy(I.kr, p.k_rinv, sizeof(I.kr));
    I.col = t;
    I.row = t + col.size();
}

Rect Warper::build_maps(int sw, int sh, const float* K, const float* R, float* xmap, float* ymap, size_t pitch)
{
    require_device();
    cudaStream_t st = current_stream();
    ImageDev I;
    Rect roi;
    prepare_image(sw, sh, K, R, I, roi, st);
    if (!xmap || !ymap) throw Error(ISB_ERR_NULL_PTR, "xmap/ymap are null");
    const bool dev = mem_kind(xmap) == MemKind::Device;
    const size_t row_bytes = (size_t)roi.w * sizeof(float);
    float *dx = xmap, *dy = ymap;
    size_t dp = pitch;
    if (!dev) {
        dp = round_up((int)row_bytes, 256);
        dx = static_cast<float*>(xm_.ensure(dp * roi.h));
        dy = static_cast<float*>(ym_.ensure(dp * roi.h));
    }
    launch_build_maps(I, dx, dy, (long long)dp, st);
    if (!dev) {
        copy2d(xmap, pitch, dx, dp, row_bytes, roi.h, st);
        copy2d(ymap, pitch, dy, dp, row_bytes, roi.h, st);
    }
    ISB_CUDA(cudaStreamSynchronize(st));
    return roi;
}

void Warper::warp(const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* K, const float* R, int interp,
                  int border, uint8_t* dst, size_t dpitch, int* corner)
{
    require_device();
    if (!src || !dst) throw Error(ISB_ERR_NULL_PTR, "src/dst are null");
    ISB_ASSERT(ch == 1 || ch == 3);
    ISB_ASSERT(interp == ISB_INTER_NEAREST || interp == ISB_INTER_LINEAR);
    ISB_ASSERT(border == ISB_BORDER_CONSTANT || border == ISB_BORDER_REFLECT);
    ISB_ASSERT(spitch >= (size_t)sw * ch);
    cudaStream_t st = current_stream();
    ImageDev I;
    Rect roi;
    prepare_image(sw, sh, K, R, I, roi, st);
    ISB_ASSERT(dpitch >= (size_t)roi.w * ch);
    if (mem_kind(src) == MemKind::Device) {
        I.src = src;
        I.spitch = (long long)spitch;
    } else {
        const size_t sp = (size_t)sw * ch;
        uint8_t* d = static_cast<uint8_t*>(src_.ensure(sp * sh));
        copy2d(d, sp, src, spitch, sp, sh, st);
        I.src = d;
        I.spitch = (long long)sp;
    }
    const bool ddev = mem_kind(dst) == MemKind::Device;
    uint8_t* dd = dst;
    size_t dp = dpitch;
    if (!ddev) {
        dp = (size_t)roi.w * ch;
        dd = static_cast<uint8_t*>(dst_.ensure(dp * roi.h));
    }
    launch_warp_generic(I, ch, interp, border, dd, (long long)dp, st);
    if (!ddev) copy2d(dst, dpitch, dd, dp, (size_t)roi.w * ch, roi.h, st);
    ISB_CUDA(cudaStreamSynchronize(st));
    if (corner) {
        corner[0] = roi.x;
        corner[1] = roi.y;
    }
}

void Warper::warp_backward(const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* K, const float* R, int interp,
                           int border, int dw, int dh, uint8_t* dst, size_t dpitch)
{
    require_device();
    check_KR(K, R);
    if (!src || !dst) throw Error(ISB_ERR_NULL_PTR, "src/dst are null");
    ISB_ASSERT(ch == 1 || ch == 3);
    ISB_ASSERT(interp == ISB_INTER_NEAREST || interp == ISB_INTER_LINEAR);
    ISB_ASSERT(border == ISB_BORDER_CONSTANT || border == ISB_BORDER_REFLECT);
    ISB_ASSERT(dw > 0 && dh > 0 && sw > 0 && sh > 0 && spitch >= (size_t)sw * ch && dpitch >= (size_t)dw * ch);
    Projector p;
    p.set(kind_, scale_, K, R);
    const Rect roi = p.warp_roi(dw, dh);
    // CV_Assert(src_br.x - src_tl.x + 1 == size.width && src_br.y - src_tl.y + 1 == size.height)
    ISB_ASSERT(roi.w == sw && roi.h == sh);
    cudaStream_t st = current_stream();
    ImageDev I{};
    I.sw = sw;
    I.sh = sh;
    if (mem_kind(src) == MemKind::Device) {
        I.src = src;
        I.spitch = (long long)spitch;
    } else {
        const size_t sp = (size_t)sw * ch;
        uint8_t* d = static_cast<uint8_t*>(src_.ensure(sp * sh));
        copy2d(d, sp, src, spitch, sp, sh, st);
        I.src = d;
        I.spitch = (long long)sp;
    }
    const bool ddev = mem_kind(dst) == MemKind::Device;
    uint8_t* dd = dst;
    size_t dp = dpitch;
    if (!ddev) {
        dp = (size_t)dw * ch;
        dd = static_cast<uint8_t*>(dst_.ensure(dp * dh));
    }
    launch_warp_backward(I, ch, p.r_kinv, scale_, kind_ == ISB_WARP_SPHERICAL ? 1 : 0, roi.x, roi.y, interp, border, dw, dh, dd, (long long)dp, st);
    if (!ddev) copy2d(dst, dpitch, dd, dp, (size_t)dw * ch, dh, st);
    ISB_CUDA(cudaGetLastError());
    ISB_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------------------------------------
// Compensator / seam mask
// ------------------------------------------------------------------------------------------------
void Compensator::set_gains(int n, const float* const* g, const int* gw, const int* gh)
{
    gains_.assign(n, {});
    gw_.assign(gw, gw + n);
    gh_.assign(gh, gh + n);
    for (int i = 0; i < n; ++i) {
        ISB_ASSERT(g[i] != nullptr && gw[i] > 0 && gh[i] > 0);
        gains_[i].assign(g[i], g[i] + (size_t)gw[i] * gh[i]);
    }
}

static void upload_gain_tables(const float* gain, int gw, int gh, int w, int h, DevBuf& aux, const float*& g_dev,
                               const LinCoefDev*& gx_dev, const LinCoefDev*& gy_dev, cudaStream_t st)
{
    std::vector<LinCoef> gx, gy;
    build_linear_f32_table(gw, w, true, gx);
    build_linear_f32_table(gh, h, false, gy);
    const size_t gbytes = ((size_t)gw * gh * sizeof(float) + 15) & ~size_t(15);
    char* base = static_cast<char*>(aux.ensure(gbytes + (gx.size() + gy.size()) * sizeof(LinCoefDev)));
    ISB_CUDA(cudaMemcpyAsync(base, gain, (size_t)gw * gh * sizeof(float), cudaMemcpyHostToDevice, st));
    ISB_CUDA(cudaMemcpyAsync(base + gbytes, gx.data(), gx.size() * sizeof(LinCoefDev), cudaMemcpyHostToDevice, st));
    ISB_CUDA(cudaMemcpyAsync(base + gbytes + gx.size() * sizeof(LinCoefDev), gy.data(), gy.size() * sizeof(LinCoefDev),
                             cudaMemcpyHostToDevice, st));
    g_dev = reinterpret_cast<const float*>(base);
    gx_dev = reinterpret_cast<const LinCoefDev*>(base + gbytes);
    gy_dev = gx_dev + gx.size();
}

void Compensator::apply(int index, uint8_t* image, int w, int h, size_t pitch)
{
    require_device();
    if (!image) throw Error(ISB_ERR_NULL_PTR, "image is null");
    if (index < 0 || index >= count()) throw Error(ISB_ERR_OUT_OF_RANGE, "compensator index out of range");
    ISB_ASSERT(w > 0 && h > 0 && pitch >= (size_t)w * 3);
    cudaStream_t st = current_stream();
    const float* g;
    const LinCoefDev *gx, *gy;
    upload_gain_tables(gains_[index].data(), gw_[index], gh_[index], w, h, aux_, g, gx, gy, st);
    const bool dev = mem_kind(image) == MemKind::Device;
    uint8_t* d = image;
    size_t dp = pitch;
    if (!dev) {
        dp = (size_t)w * 3;
        d = static_cast<uint8_t*>(img_.ensure(dp * h));
        copy2d(d, dp, image, pitch, dp, h, st);
    }
    launch_gain_apply(d, w, h, (long long)dp, g, gw_[index], gh_[index], gx, gy, st);
    if (!dev) copy2d(image, pitch, d, dp, dp, h, st);
    ISB_CUDA(cudaStreamSynchronize(st));
}

void seam_mask_apply(const uint8_t* seam, int mw, int mh, size_t spitch, uint8_t* mask, int w, int h, size_t pitch)
{
    require_device();
    if (!seam || !mask) throw Error(ISB_ERR_NULL_PTR, "seam/mask are null");
    ISB_ASSERT(mw > 0 && mh > 0 && w > 0 && h > 0 && spitch >= (size_t)mw && pitch >= (size_t)w);
    cudaStream_t st = current_stream();
    std::vector<uint32_t> mx, my;
    build_linear_exact_table(mw, w, mx);
    build_linear_exact_table(mh, h, my);
    DevBuf buf;
    const size_t raw = ((size_t)mw * mh + 255) & ~size_t(255), mb = ((size_t)w * h + 255) & ~size_t(255);
    char* base = static_cast<char*>(buf.ensure(2 * raw + mb + (mx.size() + my.size()) * sizeof(uint32_t)));
    uint8_t* sraw = reinterpret_cast<uint8_t*>(base);
    uint8_t* sdil = sraw + raw;
    uint8_t* mdev = sdil + raw;
    uint32_t* tx = reinterpret_cast<uint32_t*>(mdev + mb);
    uint32_t* ty = tx + mx.size();
    copy2d(sraw, mw, seam, spitch, mw, mh, st);
    ISB_CUDA(cudaMemcpyAsync(tx, mx.data(), mx.size() * 4, cudaMemcpyHostToDevice, st));
    ISB_CUDA(cudaMemcpyAsync(ty, my.data(), my.size() * 4, cudaMemcpyHostToDevice, st));
    const bool dev = mem_kind(mask) == MemKind::Device;
    uint8_t* m = mask;
    size_t mp = pitch;
    if (!dev) {
        m = mdev;
        mp = w;
        copy2d(m, mp, mask, pitch, w, h, st);
    }
    launch_dilate3x3(sraw, mw, mh, mw, sdil, st);
    launch_seam_and(sdil, mw, mh, tx, ty, m, w, h, (long long)mp, st);
    if (!dev) copy2d(mask, pitch, m, mp, w, h, st);
    ISB_CUDA(cudaStreamSynchronize(st));
}

void rotate_image(const uint8_t* src, int w, int h, int ch, size_t spitch, int code, uint8_t* dst, size_t dpitch)
{
    require_device();
    if (!src || !dst) throw Error(ISB_ERR_NULL_PTR, "src/dst are null");
    ISB_ASSERT(w > 0 && h > 0 && (ch == 1 || ch == 3) && (code == 0 || code == 1));
    const int dw = code == 0 ? h : w, dh = code == 0 ? w : h;
    ISB_ASSERT(spitch >= (size_t)w * ch && dpitch >= (size_t)dw * ch);
    cudaStream_t st = current_stream();
    DevBuf sb, db;
    const uint8_t* s = src;
    size_t sp = spitch;
    if (mem_kind(src) != MemKind::Device) {
        sp = (size_t)w * ch;
        void* p = sb.ensure(sp * h);
        copy2d(p, sp, src, spitch, sp, h, st);
        s = static_cast<const uint8_t*>(p);
    }
    const bool ddev = mem_kind(dst) == MemKind::Device;
    uint8_t* d = dst;
    size_t dp = dpitch;
    if (!ddev) {
        dp = (size_t)dw * ch;
        d = static_cast<uint8_t*>(db.ensure(dp * dh));
    }
    launch_rotate(s, w, h, ch, (long long)sp, code, d, (long long)dp, st);
    if (!ddev) copy2d(dst, dpitch, d, dp, (size_t)dw * ch, dh, st);
    ISB_CUDA(cudaStreamSynchronize(st));
}

void resize_linear_exact(const uint8_t* src, int sw, int sh, int ch, size_t spitch, uint8_t* dst, int dw, int dh, size_t dpitch,
                         double fx, double fy)
{
    require_device();
    if (!src || !dst) throw Error(ISB_ERR_NULL_PTR, "src/dst are null");
    ISB_ASSERT(sw > 0 && sh > 0 && dw > 0 && dh > 0 && (ch == 1 || ch == 3));
    ISB_ASSERT(spitch >= (size_t)sw * ch && dpitch >= (size_t)dw * ch);
    cudaStream_t st = current_stream();
    std::vector<uint32_t> tx, ty;
    build_linear_exact_table(sw, dw, tx, fx);
    build_linear_exact_table(sh, dh, ty, fy);
    DevBuf sb, db, tb;
    uint32_t* t = static_cast<uint32_t*>(tb.ensure((tx.size() + ty.size()) * sizeof(uint32_t)));
    ISB_CUDA(cudaMemcpyAsync(t, tx.data(), tx.size() * 4, cudaMemcpyHostToDevice, st));
    ISB_CUDA(cudaMemcpyAsync(t + tx.size(), ty.data(), ty.size() * 4, cudaMemcpyHostToDevice, st));
    const uint8_t* s = src;
    size_t sp = spitch;
    if (mem_kind(src) != MemKind::Device) {
        sp = (size_t)sw * ch;
        void* p = sb.ensure(sp * sh);
        copy2d(p, sp, src, spitch, sp, sh, st);
        s = static_cast<const uint8_t*>(p);
    }
    const bool ddev = mem_kind(dst) == MemKind::Device;
    uint8_t* d = dst;
    size_t dp = dpitch;
    if (!ddev) {
        dp = (size_t)dw * ch;
        d = static_cast<uint8_t*>(db.ensure(dp * dh));
    }
    launch_resize_exact(s, sw, sh, ch, (long long)sp, t, t + tx.size(), d, dw, dh, (long long)dp, st);
    if (!ddev) copy2d(dst, dpitch, d, dp, (size_t)dw * ch, dh, st);
    ISB_CUDA(cudaStreamSynchronize(st));
}

// ------------------------------------------------------------------------------------------------
// Blender (classic prepare / feed / blend surface)
// ------------------------------------------------------------------------------------------------
void Blender::prepare(const Rect& roi)
{
    ISB_ASSERT(roi.w > 0 && roi.h > 0);
    ISB_ASSERT(requested_ >= 0);
    BlendGeometry g;
    g.prepare(roi, requested_);
    eng_.reset(g, 0, g.roi.h, 0, g.roi.h, /*packed=*/false);
    prepared_ = true;
}

void Blender::feed(const int16_t* img, size_t ipitch, const uint8_t* mask, size_t mpitch, int w, int h, int tlx, int tly)
{
    require_device();
    if (!prepared_) throw Error(ISB_ERR_ASSERT, "Assertion failed: prepare() must be called before feed()");
    if (!img || !mask) throw Error(ISB_ERR_NULL_PTR, "img/mask are null");
    ISB_ASSERT(w > 0 && h > 0 && ipitch >= (size_t)w * 6 && mpitch >= (size_t)w);
    cudaStream_t st = current_stream();
    const int t = eng_.add_tile(-1, w, h, tlx, tly);
    if (t < 0) return;
    eng_.commit_tiles(st);
    const int16_t* di = img;
    size_t dip = ipitch;
    if (mem_kind(img) != MemKind::Device) {
        dip = (size_t)w * 6;
        void* p = img_.ensure(dip * h);
        copy2d(p, dip, img, ipitch, dip, h, st);
        di = static_cast<const int16_t*>(p);
    }
    const uint8_t* dm = mask;
    size_t dmp = mpitch;
    if (mem_kind(mask) != MemKind::Device) {
        dmp = w;
        void* p = mask_.ensure(dmp * h);
        copy2d(p, dmp, mask, mpitch, w, h, st);
        dm = static_cast<const uint8_t*>(p);
    }
    launch_pack_tile(eng_.tiles_dev() + t, eng_.tiles()[t], di, (long long)dip, dm, (long long)dmp, st);
    eng_.build_pyramids(t, t + 1, st);
    ISB_CUDA(take_launch_error());
    ISB_CUDA(cudaStreamSynchronize(st));  // feed keeps no reference to img/mask after it returns
}

void Blender::blend(int16_t* dst, size_t dpitch, uint8_t* dmask, size_t mpitch)
{
    require_device();
    if (!prepared_) throw Error(ISB_ERR_ASSERT, "Assertion failed: prepare() must be called before blend()");
    cudaStream_t st = current_stream();
    const Rect rf = eng_.geom().roi_final;
    OutDev o{};
    const bool d16 = dst && mem_kind(dst) == MemKind::Device, dmk = dmask && mem_kind(dmask) == MemKind::Device;
    if (dst) {
        ISB_ASSERT(dpitch >= (size_t)rf.w * 6);
        o.out16 = d16 ? dst : static_cast<int16_t*>(out16_.ensure((size_t)rf.w * 6 * rf.h));
        o.pitch16 = d16 ? (long long)dpitch : (long long)rf.w * 6;
    }
    if (dmask) {
        ISB_ASSERT(mpitch >= (size_t)rf.w);
        o.mask = dmk ? dmask : static_cast<uint8_t*>(outm_.ensure((size_t)rf.w * rf.h));
        o.mpitch = dmk ? (long long)mpitch : rf.w;
    }
    eng_.blend(o, st);
    ISB_CUDA(take_launch_error());
    if (dst && !d16) copy2d(dst, dpitch, o.out16, (size_t)o.pitch16, (size_t)rf.w * 6, rf.h, st);
    if (dmask && !dmk) copy2d(dmask, mpitch, o.mask, (size_t)o.mpitch, rf.w, rf.h, st);
    ISB_CUDA(cudaStreamSynchronize(st));
    prepared_ = false;  // single use per prepare(), as MultiBandBlender::blend releases its pyramids
}

// ------------------------------------------------------------------------------------------------
// Composer (fused loop)
// ------------------------------------------------------------------------------------------------
enum { ST_H2D = 0, ST_WARP, ST_PYRDOWN, ST_BLEND, ST_D2H, ST_COUNT };
const char* Composer::stage_name(int i)
{
    static const char* names[ST_COUNT] = {"h2d", "warp", "pyrdown", "blend", "d2h"};
    return (i >= 0 && i < ST_COUNT) ? names[i] : nullptr;
}

Composer::~Composer()
{
    if (ev_init_)
        for (auto& e : ev_) cudaEventDestroy(e);
    if (copy_stream_) {
        cudaStreamSynchronize(copy_stream_);
        for (int k = 0; k < 2; ++k) {
            cudaEventDestroy(ev_done_[k]);
            cudaEventDestroy(ev_copied_[k]);
        }
        cudaStreamDestroy(copy_stream_);
    }
}

bool Composer::same_plan(const isb_camera* cams, const int* sizes_wh, int n) const
{
    if (!planned_ || (int)cams_.size() != n) return false;
    return std::memcmp(cams_.data(), cams, sizeof(isb_camera) * n) == 0 &&
           std::memcmp(src_sizes_.data(), sizes_wh, sizeof(int) * 2 * n) == 0;
}

void Composer::plan(const isb_camera* cams, const int* sizes_wh, int n, int* corners, int* sizes, int* dst_roi)
{
    if (!cams || !sizes_wh) throw Error(ISB_ERR_NULL_PTR, "cams/sizes are null");
    ISB_ASSERT(n > 0);
    ISB_ASSERT(cfg_.warp_kind == ISB_WARP_SPHERICAL || cfg_.warp_kind == ISB_WARP_CYLINDRICAL);
    ISB_ASSERT(cfg_.strip_count >= 1 && cfg_.strip_index >= 0 && cfg_.strip_index < cfg_.strip_count);
    ISB_ASSERT(cfg_.num_bands >= 0);
    if (!(cfg_.cache_plan && same_plan(cams, sizes_wh, n))) {
        require_device();
        cudaStream_t st = current_stream();
        planned_ = false;
        cams_.assign(cams, cams + n);
        src_sizes_.assign(sizes_wh, sizes_wh + 2 * n);
        img_.assign(n, {});
        // per-image projector + ROI (border forward maps: O(perimeter) atan2f/acosf) - spread over host threads
        auto work = [&](int i) {
            ImagePlan& P = img_[i];
            float K[9];
            isb_camera_K(&cams[i], K);
            // with ingest pre-steps the caller passes decoded sizes; the path works on sz = cvRound(rotated size * compose_scale)
            ingest_size(sizes_wh[2 * i], sizes_wh[2 * i + 1], P.src_w, P.src_h);
            P.proj.set(cfg_.warp_kind, cfg_.warped_image_scale, K, cams[i].R);
            P.roi = P.proj.warp_roi(P.src_w, P.src_h);
        };
        // cv::remap itself requires source sizes below SHRT_MAX; the fused sampler relies on it too
        for (int i = 0; i < n; ++i)
            ISB_ASSERT(sizes_wh[2 * i] > 0 && sizes_wh[2 * i + 1] > 0 && sizes_wh[2 * i] < 32767 && sizes_wh[2 * i + 1] < 32767);
        const int nthr = std::max(1, std::min<int>(n, std::min(16u, std::thread::hardware_concurrency())));
        if (nthr <= 1) {
            for (int i = 0; i < n; ++i) work(i);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nthr; ++t)
                pool.emplace_back([&, t] { for (int i = t; i < n; i += nthr) work(i); });
            for (auto& th : pool) th.join();
        }
        std::vector<int> cs(2 * n), ss(2 * n);
        for (int i = 0; i < n; ++i) {
            cs[2 * i] = img_[i].roi.x; cs[2 * i + 1] = img_[i].roi.y;
            ss[2 * i] = img_[i].roi.w; ss[2 * i + 1] = img_[i].roi.h;
        }
        dst_roi_ = result_roi(cs.data(), ss.data(), n);
        // which blender (image_stitching.cpp:1173-1193)
        eff_blend_type_ = ISB_BLENDER_MULTI_BAND;
        eff_num_bands_ = cfg_.num_bands;
        if (cfg_.use_blend_rule) {
            ISB_ASSERT(cfg_.blend_type == ISB_BLENDER_NO || cfg_.blend_type == ISB_BLENDER_FEATHER || cfg_.blend_type == ISB_BLENDER_MULTI_BAND);
            const float blend_width = std::sqrt(static_cast<float>((long long)dst_roi_.w * dst_roi_.h)) * cfg_.blend_strength / 100.f;
            eff_blend_type_ = cfg_.blend_type;
            if (blend_width < 1.f) eff_blend_type_ = ISB_BLENDER_NO;
            else if (eff_blend_type_ == ISB_BLENDER_MULTI_BAND)
                eff_num_bands_ = static_cast<int>(std::ceil(std::log(blend_width) / std::log(2.)) - 1.);
            else if (eff_blend_type_ == ISB_BLENDER_FEATHER) eff_sharpness_ = 1.f / blend_width;
        }
        ISB_ASSERT(eff_num_bands_ >= 0);
        if (eff_blend_type_ != ISB_BLENDER_MULTI_BAND) {
            // Blender::NO / FeatherBlender: no pyramids to plan; run() drives the per-call kernels over the planned ROIs
            ISB_ASSERT(cfg_.strip_count == 1);
            planned_ = true;
            for (int i = 0; i < n; ++i) {
                if (corners) { corners[2 * i] = img_[i].roi.x; corners[2 * i + 1] = img_[i].roi.y; }
                if (sizes) { sizes[2 * i] = img_[i].roi.w; sizes[2 * i + 1] = img_[i].roi.h; }
            }
            if (dst_roi) { dst_roi[0] = dst_roi_.x; dst_roi[1] = dst_roi_.y; dst_roi[2] = dst_roi_.w; dst_roi[3] = dst_roi_.h; }
            return;
        }
        BlendGeometry g;
        g.prepare(dst_roi_, eff_num_bands_);
        // strip: owned rows on the 2^nb grid + halo rows computed redundantly.  With m = 2^nb and strip cuts on the m grid,
        // output rows [y0, y1) read collapsed rows [y0 / 2^l - 2, y1 / 2^l + 1] of level l (fine rows [a, b] read coarse rows
        // [a/2 - 1, b/2 + 1]); a level-l row q of an image pyramid reads level-0 rows [2^l q - 2 (2^l - 1), 2^l q + 2 (2^l - 1)],
        // and the Laplacian at row q additionally level l + 1 rows q/2 - 1 .. q/2 + 1.  The extremes are reached at the top
        // level: rows [y0 - 4 m + 2, y1 + 3 m - 2).  So the halo is FOUR cells above and THREE cells below the strip; data
        // beyond it - and the artificial border rules at the cuts - cannot reach an owned row.
        // (the cuts themselves are chosen below, once the work per cell row is known)
        eng_.reset(g, 0, g.roi.h, 0, g.roi.h, /*packed=*/true);  // geometry only; re-issued with the strip's rows before tiles are added

        // separable trig tables (every image) + slots for the per-run coefficient tables
        tables_.begin();
        for (int i = 0; i < n; ++i) {
            ImagePlan& P = img_[i];
            P.col_off = tables_.take(P.roi.w * sizeof(F2));
            P.row_off = tables_.take(P.roi.h * sizeof(F2));
            P.gx_off = tables_.take(P.roi.w * sizeof(LinCoefDev));
            P.gy_off = tables_.take(P.roi.h * sizeof(LinCoefDev));
            P.mx_off = tables_.take(P.roi.w * sizeof(uint32_t));
            P.my_off = tables_.take(P.roi.h * sizeof(uint32_t));
            P.gain_w = P.gain_h = P.seam_w = P.seam_h = -1;
        }
        char* tb = tables_.commit();
        auto tabs = [&](int i) { build_trig_tables(img_[i].proj, img_[i].roi, img_[i].col, img_[i].row); };
        if (nthr <= 1) {
            for (int i = 0; i < n; ++i) tabs(i);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nthr; ++t)
                pool.emplace_back([&, t] { for (int i = t; i < n; i += nthr) tabs(i); });
            for (auto& th : pool) th.join();
        }
        for (int i = 0; i < n; ++i) {
            ImagePlan& P = img_[i];
            ISB_CUDA(cudaMemcpyAsync(tb + P.col_off, P.col.data(), P.col.size() * sizeof(F2), cudaMemcpyHostToDevice, st));
            ISB_CUDA(cudaMemcpyAsync(tb + P.row_off, P.row.data(), P.row.size() * sizeof(F2), cudaMemcpyHostToDevice, st));
        }

        // Support culling.  The device marks, per macro cell (2^nb x 2^nb px) of every image's feed() tile, whether
        // any warped pixel is valid.  Weight pyramids are exactly zero farther than 2*2^nb px from the support and the
        // Laplacian there never reaches the output, so pyramids are only built on rectangles that cover the occupied
        // cells dilated by 4 cells (>= the 4*2^nb px dependency radius): a wrap-around image whose ROI
        // spans the whole panorama shrinks to its two real ends.  Rect edges that coincide with the tile's own edges
        // keep OpenCV's border rules; the other cuts are too far from any non-zero weight to matter.
        struct FullTile { int X0, Y0, W, H; };
        std::vector<FullTile> full(n);
        std::vector<OccTile>& occ_tiles = occ_tiles_;
        occ_tiles.assign(n, OccTile{});
        last_seam_blk_.clear();
        std::vector<ImageDev> idev(n);
        size_t occ_bytes = 0;
        int max_w = 0, max_h = 0;
        for (int i = 0; i < n; ++i) {
            const ImagePlan& P = img_[i];
            int tl[2], br[2];
            g.tile_rect(P.roi.w, P.roi.h, P.roi.x, P.roi.y, tl, br);
            full[i] = FullTile{tl[0] - g.roi.x, tl[1] - g.roi.y, br[0] - tl[0], br[1] - tl[1]};
            occ_tiles[i] = OccTile{i, P.roi.x - tl[0], P.roi.y - tl[1], full[i].W, full[i].H, (long long)occ_bytes};
            occ_bytes += (size_t)(full[i].W >> g.nb) * (full[i].H >> g.nb);
            max_w = std::max(max_w, full[i].W);
            max_h = std::max(max_h, full[i].H);
            ImageDev& I = idev[i];
            I = ImageDev{};
            I.sw = P.src_w; I.sh = P.src_h; I.roi_w = P.roi.w; I.roi_h = P.roi.h;
            std::memcpy(I.kr, P.proj.k_rinv, sizeof(I.kr));
            I.col = reinterpret_cast<const F2*>(tb + P.col_off);
            I.row = reinterpret_cast<const F2*>(tb + P.row_off);
        }
        std::vector<uint8_t> occ(occ_bytes, 0);
        {
            ImageDev* idp = static_cast<ImageDev*>(imgs_dev_.ensure(n * sizeof(ImageDev)));
            ISB_CUDA(cudaMemcpyAsync(idp, idev.data(), n * sizeof(ImageDev), cudaMemcpyHostToDevice, st));
            last_idev_.clear();  // imgs_dev_ now holds the plan-time subset of the descriptors
            OccTile* otp = static_cast<OccTile*>(occ_tiles_dev_.ensure(n * sizeof(OccTile)));
            ISB_CUDA(cudaMemcpyAsync(otp, occ_tiles.data(), n * sizeof(OccTile), cudaMemcpyHostToDevice, st));
            // the valid occupancy stays on the device: every run ANDs it with the seam masks' support (launch_seam_prep)
            uint8_t* od = static_cast<uint8_t*>(occ_valid_dev_.ensure(std::max<size_t>(occ_bytes, 1)));
            ISB_CUDA(cudaMemsetAsync(need_dev_.ensure(std::max<size_t>(occ_bytes, 1) * sizeof(uint32_t)), 0,
                                     std::max<size_t>(occ_bytes, 1) * sizeof(uint32_t), st));
            need_gen_ = 0;
            ISB_CUDA(cudaMemsetAsync(od, 0, std::max<size_t>(occ_bytes, 1), st));
            for (int z0 = 0; z0 < n; z0 += 32768)
                launch_occupancy(otp + z0, std::min(32768, n - z0), max_w, max_h, idp, g.nb, od, st);
            ISB_CUDA(cudaMemcpyAsync(occ.data(), od, occ_bytes, cudaMemcpyDeviceToHost, st));
            ISB_CUDA(cudaStreamSynchronize(st));
        }
        tiles_of_image_.assign(n, {});
        const int kDilate = 4;
        struct RectSpec { int img, X0, Y0, W, H; };  // level-0 pixels, padded-panorama coordinates, on the 2^nb grid
        std::vector<RectSpec> rects;
        for (int i = 0; i < n; ++i) {
            const int cw = full[i].W >> g.nb, chh = full[i].H >> g.nb;
            const uint8_t* o = occ.data() + occ_tiles[i].occ_off;
            // dilated column / row extents
            std::vector<int> col_lo(cw, INT_MAX), col_hi(cw, INT_MIN);
            for (int cy = 0; cy < chh; ++cy)
                for (int cx = 0; cx < cw; ++cx)
                    if (o[(size_t)cy * cw + cx])
                        for (int dx = std::max(0, cx - kDilate); dx <= std::min(cw - 1, cx + kDilate); ++dx) {
                            col_lo[dx] = std::min(col_lo[dx], std::max(0, cy - kDilate));
                            col_hi[dx] = std::max(col_hi[dx], std::min(chh - 1, cy + kDilate));
                        }
            for (int cx = 0; cx < cw;) {
                if (col_hi[cx] < col_lo[cx]) { ++cx; continue; }
                int x1 = cx, lo = col_lo[cx], hi = col_hi[cx];
                while (x1 + 1 < cw && col_hi[x1 + 1] >= col_lo[x1 + 1]) {
                    ++x1;
                    lo = std::min(lo, col_lo[x1]);
                    hi = std::max(hi, col_hi[x1]);
                }
                rects.push_back(RectSpec{i, full[i].X0 + (cx << g.nb), full[i].Y0 + (lo << g.nb), (x1 - cx + 1) << g.nb, (hi - lo + 1) << g.nb});
                cx = x1 + 1;
            }
        }
        // Strip cuts, balanced by WORK.  Per cell row: W[r] = cells of all pyramid rectangles in that row (kernels 1 and 2 and
        // the coarse blend levels scale with it), and the level-0 blend scales with the panorama width.  A strip [a, b) of cell
        // rows warps rows [a - 4, b + 3) and blends rows [a, b); the cuts minimise the largest strip cost (greedy partition
        // under a bisected bound).  Every rank evaluates the same integers, so every rank finds the same cuts.
        int y0 = 0, y1 = g.roi.h;
        const int m = 1 << g.nb;
        if (cfg_.strip_count > 1) {
            const int cells_y = g.roi.h >> g.nb, cells_x = g.roi.w >> g.nb, N = cfg_.strip_count;
            strip_rows(g.roi.h, g.nb, cfg_.strip_index, N, y0, y1);  // arithmetic cuts: the fall-back for very short panoramas
            if (cells_y >= 2 * N + 7) {
                std::vector<double> PW(cells_y + 1, 0.0);
                {
                    std::vector<double> Wr(cells_y, 0.0);
                    for (const RectSpec& r : rects)
                        for (int cy = r.Y0 >> g.nb; cy < (r.Y0 + r.H) >> g.nb; ++cy) Wr[cy] += (double)(r.W >> g.nb);
                    for (int r = 0; r < cells_y; ++r) PW[r + 1] = PW[r] + Wr[r];
                }
                const double alpha = 0.85, beta = 1.0;
                auto cost = [&](int a, int b) {
                    const int ha = a == 0 ? 0 : std::max(0, a - 4), hb = b == cells_y ? cells_y : std::min(cells_y, b + 3);
                    return alpha * (PW[hb] - PW[ha]) + beta * (double)cells_x * (b - a);
                };
                auto partition = [&](double T, std::vector<int>* cuts) {  // greedy: strips as long as the bound allows
                    int a = 0, k = 0;
                    if (cuts) cuts->assign(1, 0);
                    while (a < cells_y) {
                        int b = a + 1;
                        if (cost(a, b) > T) return N + 1;  // a single row exceeds the bound
                        while (b < cells_y && cost(a, b + 1) <= T) ++b;
                        // leave at least one row for each of the remaining strips
                        b = std::min(b, cells_y - (N - 1 - k));
                        b = std::max(b, a + 1);
                        if (cuts) cuts->push_back(b);
                        a = b;
                        ++k;
                        if (k > N) return k;
                    }
                    return k;
                };
                double lo_t = 0.0, hi_t = cost(0, cells_y);
                for (int it = 0; it < 60; ++it) {
                    const double mid = 0.5 * (lo_t + hi_t);
                    if (partition(mid, nullptr) <= N) hi_t = mid;
                    else lo_t = mid;
                }
                std::vector<int> cuts;
                if (partition(hi_t, &cuts) <= N) {
                    while ((int)cuts.size() < N + 1) {  // fewer strips than ranks: split the tallest strip
                        int best = 0;
                        for (int k = 1; k + 1 < (int)cuts.size(); ++k)
                            if (cuts[k + 1] - cuts[k] > cuts[best + 1] - cuts[best]) best = k;
                        if (cuts[best + 1] - cuts[best] < 2) break;
                        cuts.insert(cuts.begin() + best + 1, (cuts[best] + cuts[best + 1]) / 2);
                    }
                    if ((int)cuts.size() == N + 1) {
                        y0 = cuts[cfg_.strip_index] << g.nb;
                        y1 = cuts[cfg_.strip_index + 1] << g.nb;
                    }
                }
            }
        }
        const int halo_top = cfg_.strip_count > 1 ? 4 * m : 0, halo_bot = cfg_.strip_count > 1 ? 3 * m : 0;
        const int sy0 = std::max(0, y0 - halo_top), sy1 = std::min(g.roi.h, (y1 + m - 1) / m * m + halo_bot);
        eng_.reset(g, sy0, std::max(sy1 - sy0, 0), y0, y1, /*packed=*/true);
        for (int i = 0; i < n; ++i) {
            const int cw = full[i].W >> g.nb, chh = full[i].H >> g.nb;
            const uint8_t* o = occ.data() + occ_tiles[i].occ_off;
            eng_.set_image_occupancy(i, std::vector<uint8_t>(o, o + (size_t)cw * chh), cw, chh, full[i].X0, full[i].Y0);
        }
        for (const RectSpec& r : rects) {
            const int i = r.img;
            const int t = eng_.add_rect(i, r.X0, r.Y0, r.W, r.H, img_[i].roi.x, img_[i].roi.y, img_[i].roi.w, img_[i].roi.h,
                                        need_dev_.as<uint32_t>() + occ_tiles[i].occ_off, full[i].X0, full[i].Y0, full[i].W >> g.nb);
            if (t >= 0) tiles_of_image_[i].push_back(t);
        }
        eng_.commit_tiles(st);
        // strip-sharded runs: the band of source rows this strip can read from every image (host sources are uploaded
        // band-wise).  imgs_dev_ holds the plan-time descriptors (geometry only), which is all the kernel needs.
        src_band_.assign(2 * (size_t)n, 0);
        for (int i = 0; i < n; ++i) src_band_[2 * i + 1] = img_[i].src_h - 1;
        if (cfg_.strip_count > 1 && !eng_.warp_work().empty()) {
            std::vector<int> init(2 * (size_t)n);
            for (int i = 0; i < n; ++i) { init[2 * i] = INT_MAX; init[2 * i + 1] = INT_MIN; }
            int* bd = static_cast<int*>(band_dev_.ensure(init.size() * sizeof(int)));
            ISB_CUDA(cudaMemcpyAsync(bd, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            launch_src_band(eng_.warp_work_dev(), (int)eng_.warp_work().size(), eng_.tiles_dev(), imgs_dev_.as<ImageDev>(), bd, st);
            ISB_CUDA(cudaMemcpyAsync(init.data(), bd, init.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
            ISB_CUDA(cudaStreamSynchronize(st));
            for (int i = 0; i < n; ++i) {
                if (init[2 * i] > init[2 * i + 1]) continue;  // image without tiles in this strip
                // whole multiples of 8 rows keep the (virtual) base address of row 0 8-byte aligned
                src_band_[2 * i] = std::max(0, init[2 * i]) & ~7;
                src_band_[2 * i + 1] = std::min(img_[i].src_h - 1, init[2 * i + 1]);
            }
        }
        valid_counts_.clear();
        planned_ = true;
    }
    for (int i = 0; i < n; ++i) {
        if (corners) { corners[2 * i] = img_[i].roi.x; corners[2 * i + 1] = img_[i].roi.y; }
        if (sizes) { sizes[2 * i] = img_[i].roi.w; sizes[2 * i + 1] = img_[i].roi.h; }
    }
    if (dst_roi) { dst_roi[0] = dst_roi_.x; dst_roi[1] = dst_roi_.y; dst_roi[2] = dst_roi_.w; dst_roi[3] = dst_roi_.h; }
}

void Composer::run(const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seams, int n, isb_pano* out)
{
    require_device();
    if (!planned_) throw Error(ISB_ERR_ASSERT, "Assertion failed: isb_composer_plan() must precede isb_composer_run()");
    if (!imgs || !out) throw Error(ISB_ERR_NULL_PTR, "imgs/out are null");
    ISB_ASSERT(n == (int)img_.size());
    std::vector<isb_image> ingested;
    if (ingest_active()) {
        ingest(imgs, n, ingested, current_stream());
        imgs = ingested.data();
    }
    if (eff_blend_type_ != ISB_BLENDER_MULTI_BAND) {
        run_simple(imgs, gains, seams, n, out);
        return;
    }
    cudaStream_t st = current_stream();
    if (!ev_init_) {
        for (auto& e : ev_) ISB_CUDA(cudaEventCreate(&e));
        ev_init_ = true;
    }
    const Rect rf = eng_.geom().roi_final;
    const int oy0 = std::min(eng_.own_y0(), rf.h), oy1 = std::min(eng_.own_y1(), rf.h);

    // ---- stage 0: inputs to the device -------------------------------------------------------
    ISB_CUDA(cudaEventRecord(ev_[0], st));
    dyn_.begin();
    h2d_bytes_ = 0;
    std::vector<size_t> src_off(n, 0), gain_off(n, 0), sraw_off(n, 0), sdil_off(n, 0);
    std::vector<char> src_dev(n, 0), gain_dev(n, 0), seam_dev(n, 0);  // one pointer-attribute query per buffer and run
    for (int i = 0; i < n; ++i) {
        if (tiles_of_image_[i].empty()) continue;
        const isb_image& im = imgs[i];
        if (!im.data) throw Error(ISB_ERR_NULL_PTR, "image data is null");
        ISB_ASSERT(im.width == img_[i].src_w && im.height == img_[i].src_h && im.pitch >= (size_t)im.width * 3);
        src_dev[i] = mem_kind(im.data) == MemKind::Device;
        if (!src_dev[i])  // rows [band lo, band hi] only, + one row: room for the sampler's 16-byte
            src_off[i] = dyn_.take((size_t)im.width * 3 * (src_band_[2 * i + 1] - src_band_[2 * i] + 2) + 32);  // windows and its speculative loads
        if (gains && gains[i].data) {
            ISB_ASSERT(gains[i].width > 0 && gains[i].height > 0);
            gain_dev[i] = mem_kind(gains[i].data) == MemKind::Device;
            if (!gain_dev[i])
                gain_off[i] = dyn_.take((size_t)gains[i].width * gains[i].height * sizeof(float));
        }
        if (seams && seams[i].data) {
            ISB_ASSERT(seams[i].width > 0 && seams[i].height > 0 && seams[i].pitch >= (size_t)seams[i].width);
            seam_dev[i] = mem_kind(seams[i].data) == MemKind::Device;
            if (!seam_dev[i]) sraw_off[i] = dyn_.take((size_t)seams[i].width * seams[i].height);
            sdil_off[i] = dyn_.take((size_t)seams[i].width * seams[i].height);
        }
    }
    char* db = dyn_.commit();
    char* tb = tables_.base();
    std::vector<ImageDev> idev(n);
    std::vector<LinCoef> gx, gy;
    std::vector<uint32_t> mx, my;
    for (int i = 0; i < n; ++i) {
        ImageDev& I = idev[i];
        I = ImageDev{};
        if (tiles_of_image_[i].empty()) continue;
        ImagePlan& P = img_[i];
        const isb_image& im = imgs[i];
        I.sw = P.src_w; I.sh = P.src_h; I.roi_w = P.roi.w; I.roi_h = P.roi.h;
        std::memcpy(I.kr, P.proj.k_rinv, sizeof(I.kr));
        I.col = reinterpret_cast<const F2*>(tb + P.col_off);
        I.row = reinterpret_cast<const F2*>(tb + P.row_off);
        {   // |x|, |y|, |z| of the inverse map are at most 3 * max|kr| * max(1, |y_|): the shared-reciprocal division of
            // the fused warp is exact while they stay below 2^40 (and z above 2^-40)
            if (P.row_bmax < 0.f) {
                float m = 0.f;
                for (const Float2& r : P.row) m = std::fabs(r.b) <= m ? m : std::fabs(r.b);  // NaN propagates
                P.row_bmax = m;
            }
            float kmax = 0.f;
            bool finite = std::isfinite(P.row_bmax);
            for (float k : I.kr) {
                finite = finite && std::isfinite(k);
                kmax = std::max(kmax, std::fabs(k));
            }
            const double bound = 3.0 * (double)kmax * std::max(1.0, (double)P.row_bmax);
            I.zlo = (finite && bound < 0x1p40) ? 0x1p-40f : INFINITY;
        }
        if (src_dev[i]) {
            I.src = im.data;
            I.spitch = (long long)im.pitch;
        } else {
            const size_t rb = (size_t)im.width * 3;
            const int lo = src_band_[2 * i], hi = src_band_[2 * i + 1];
            copy2d(db + src_off[i], rb, im.data + (size_t)lo * im.pitch, im.pitch, rb, (size_t)(hi - lo + 1), st);
            // the address row 0 would have: rows outside [lo, hi] are never read (src_band_kernel)
            I.src = reinterpret_cast<const uint8_t*>(db + src_off[i]) - (size_t)lo * rb;
            I.spitch = (long long)rb;
            I.band_lo = lo;
            h2d_bytes_ += rb * (size_t)(hi - lo + 1);
        }
        {   // the vectorised sampler reads aligned 16-byte windows: needs an 8-B aligned base and 32-bit offsets
            const unsigned long long extent = (unsigned long long)I.spitch * (I.sh - 1) + (unsigned long long)I.sw * 3;
            I.sbytes = ((reinterpret_cast<uintptr_t>(I.src) & 7) == 0 && extent < 0xFFFFFF00ull) ? (unsigned)extent : 0u;
            // window of tap row y0 + 1 at x0 = sw - 2 ends at (y0 + 1) * pitch + 3 * (sw - 2) + 16 at most
            const long long room = (long long)I.sbytes - 16 - 3ll * (I.sw - 2);
            I.fast_h = (I.sbytes && I.sw >= 2 && room >= I.spitch) ? (int)std::min<long long>(room / I.spitch, I.sh - 1) : 0;
            // staged host sources end in a spare row: every tap row of the band may use the vectorised sampler
            if (I.sbytes && I.sw >= 2 && !src_dev[i]) I.fast_h = I.sh - 1;
        }
        if (gains && gains[i].data) {
            const isb_gainmap& g = gains[i];
            const bool gdev = gain_dev[i];  // device-resident gain maps are used in place
            if (!gdev) ISB_CUDA(cudaMemcpyAsync(db + gain_off[i], g.data, (size_t)g.width * g.height * sizeof(float), cudaMemcpyDefault, st));
            if (P.gain_w != g.width || P.gain_h != g.height) {  // coefficient tables depend on the sizes only
                build_linear_f32_table(g.width, P.roi.w, true, gx);
                build_linear_f32_table(g.height, P.roi.h, false, gy);
                ISB_CUDA(cudaMemcpyAsync(tb + P.gx_off, gx.data(), gx.size() * sizeof(LinCoefDev), cudaMemcpyHostToDevice, st));
                ISB_CUDA(cudaMemcpyAsync(tb + P.gy_off, gy.data(), gy.size() * sizeof(LinCoefDev), cudaMemcpyHostToDevice, st));
                P.gain_w = g.width;
                P.gain_h = g.height;
            }
            I.gain = gdev ? g.data : reinterpret_cast<const float*>(db + gain_off[i]);
            I.gw = g.width; I.gh = g.height;
            I.gx = reinterpret_cast<const LinCoefDev*>(tb + P.gx_off);
            I.gy = reinterpret_cast<const LinCoefDev*>(tb + P.gy_off);
        }
        if (seams && seams[i].data) {
            const isb_mask& m = seams[i];
            if (seam_dev[i]) {
                I.seam_raw = m.data;
                I.seam_raw_pitch = (int)m.pitch;
            } else {
                copy2d(db + sraw_off[i], m.width, m.data, m.pitch, m.width, m.height, st);
                I.seam_raw = reinterpret_cast<const uint8_t*>(db + sraw_off[i]);
                I.seam_raw_pitch = m.width;
            }
            if (P.seam_w != m.width || P.seam_h != m.height) {
                build_linear_exact_table(m.width, P.roi.w, mx);
                build_linear_exact_table(m.height, P.roi.h, my);
                ISB_CUDA(cudaMemcpyAsync(tb + P.mx_off, mx.data(), mx.size() * 4, cudaMemcpyHostToDevice, st));
                ISB_CUDA(cudaMemcpyAsync(tb + P.my_off, my.data(), my.size() * 4, cudaMemcpyHostToDevice, st));
                P.seam_w = m.width;
                P.seam_h = m.height;
            }
            I.seam = reinterpret_cast<const uint8_t*>(db + sdil_off[i]);
            I.mw = m.width; I.mh = m.height;
            I.mx = reinterpret_cast<const uint32_t*>(tb + P.mx_off);
            I.my = reinterpret_cast<const uint32_t*>(tb + P.my_off);
        }
    }
    ImageDev* idp = static_cast<ImageDev*>(imgs_dev_.ensure(n * sizeof(ImageDev)));
    // the descriptors hold pointers and sizes only: device-resident callers repeat them run after run, and the copy that is
    // already on the device (stream-ordered behind the previous run) is then the one to use
    if (last_idev_.size() != idev.size() || std::memcmp(last_idev_.data(), idev.data(), n * sizeof(ImageDev)) != 0) {
        ISB_CUDA(cudaMemcpyAsync(idp, idev.data(), n * sizeof(ImageDev), cudaMemcpyHostToDevice, st));
        last_idev_ = idev;
    }

    // ---- stage 1: seam dilate + fused warp (kernel 1) ---------------------------------------
    ISB_CUDA(cudaEventRecord(ev_[1], st));
    const bool cull = seams && eng_.geom().nb >= 2;
    if (cull) ++need_gen_;
    {   // dense block lists of the seam preparation launch (see seam_prep_kernel); uploaded when they change
        std::vector<int> blk(2 * (size_t)n + 2, 0);
        for (int i = 0; i < n; ++i)
            blk[i + 1] = blk[i] + (idev[i].seam ? ((idev[i].mw + 127) / 128) * ((idev[i].mh + 7) / 8) : 0);
        blk[n + 1] = blk[n];
        for (int i = 0; i < n; ++i) {
            const int cw = occ_tiles_[i].w >> eng_.geom().nb, ch = occ_tiles_[i].h >> eng_.geom().nb;
            blk[n + 2 + i] = blk[n + 1 + i] + (cull ? ((cw + 31) / 32) * ((ch + 7) / 8) : 0);
        }
        int* bd = static_cast<int*>(seam_blk_dev_.ensure(blk.size() * sizeof(int)));
        if (blk != last_seam_blk_) {
            ISB_CUDA(cudaMemcpyAsync(bd, blk.data(), blk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            last_seam_blk_ = blk;
        }
        launch_seam_prep(idp, n, bd, blk.back(), occ_tiles_dev_.as<OccTile>(), eng_.geom().nb, occ_valid_dev_.as<uint8_t>(),
                         need_dev_.as<uint32_t>(), need_gen_, st);
    }
    bool banded = false;
    for (const ImageDev& I : idev) banded = banded || I.band_lo != 0;
    launch_warp_tiles_packed(eng_.warp_work_dev(), (int)eng_.warp_work().size(), eng_.tiles_dev(), idp, eng_.geom().nb, need_gen_, banded,
                             eng_.mirror_pad(), st);
    if (eng_.mirror_pad()) launch_mirror_pad(eng_.pad_work_dev(), (int)eng_.pad_work().size(), eng_.tiles_dev(), st);
    // ---- stage 2: pyramids (kernel 2) ---------------------------------------------------------
    ISB_CUDA(cudaEventRecord(ev_[2], st));
    eng_.build_pyramids(0, (int)eng_.tiles().size(), st);
    // ---- stage 3: accumulate + normalise + collapse (kernel 3) --------------------------------
    ISB_CUDA(cudaEventRecord(ev_[3], st));
    OutDev o{};
    const bool d8 = out->data && mem_kind(out->data) == MemKind::Device;
    const bool dm = out->mask && mem_kind(out->mask) == MemKind::Device;
    const bool d16 = out->data16 && mem_kind(out->data16) == MemKind::Device;
    const int sub0 = eng_.sub_y0();
    // device-side output addressed in sub-panorama rows: row y of the kernel == panorama row y + sub0
    if (out->data) {
        ISB_ASSERT(out->pitch >= (size_t)rf.w * 3);
        if (d8) { o.out8 = out->data + (long long)sub0 * out->pitch; o.pitch8 = (long long)out->pitch; }
        else {
            o.pitch8 = (long long)rf.w * 3;
            o.out8 = static_cast<uint8_t*>(out8_.ensure((size_t)o.pitch8 * std::max(oy1 - oy0, 1))) - (long long)(oy0 - sub0) * o.pitch8;
        }
    }
    if (out->mask) {
        ISB_ASSERT(out->mask_pitch >= (size_t)rf.w);
        if (dm) { o.mask = out->mask + (long long)sub0 * out->mask_pitch; o.mpitch = (long long)out->mask_pitch; }
        else {
            o.mpitch = rf.w;
            o.mask = static_cast<uint8_t*>(outm_.ensure((size_t)o.mpitch * std::max(oy1 - oy0, 1))) - (long long)(oy0 - sub0) * o.mpitch;
        }
    }
    if (out->data16) {
        ISB_ASSERT(out->pitch16 >= (size_t)rf.w * 6);
        if (d16) { o.out16 = reinterpret_cast<int16_t*>(reinterpret_cast<char*>(out->data16) + (long long)sub0 * out->pitch16); o.pitch16 = (long long)out->pitch16; }
        else {
            o.pitch16 = (long long)rf.w * 6;
            o.out16 = reinterpret_cast<int16_t*>(static_cast<char*>(out16_.ensure((size_t)o.pitch16 * std::max(oy1 - oy0, 1))) - (long long)(oy0 - sub0) * o.pitch16);
        }
    }
    // strip-sharded runs usually write into rank 0's panorama over NVLink (peer-mapped pointers cannot be told apart from
    // local ones reliably): their stores go out as staged 16-byte vectors; a single GPU is served better by direct stores
    o.peer = d8 && cfg_.strip_count > 1 && cfg_.gather_mode != ISB_GATHER_LOCAL;
    // ISB_GATHER_COPY_ENGINE: the strip is composed into a LOCAL double-buffered staging block (direct stores) and pushed into the
    // caller's panorama - rank 0's, peer-mapped - by the copy engine on a second stream, so the transfer overlaps the next
    // run's kernels instead of stalling this run's last kernel on NVLink back-pressure.
    const bool gcopy = cfg_.gather_mode == ISB_GATHER_COPY_ENGINE && cfg_.strip_count > 1 && d8 && dm && !out->data16 && oy1 > oy0;
    int slot = 0;
    if (gcopy) {
        if (!copy_stream_) {
            ISB_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
            for (int k = 0; k < 2; ++k) {
                ISB_CUDA(cudaEventCreateWithFlags(&ev_done_[k], cudaEventDisableTiming));
                ISB_CUDA(cudaEventCreateWithFlags(&ev_copied_[k], cudaEventDisableTiming));
            }
        }
        slot = (int)(run_count_ & 1);
        if (run_count_ >= 2) ISB_CUDA(cudaStreamWaitEvent(st, ev_copied_[slot], 0));  // the slot's previous strip has left
        const size_t rows = (size_t)(oy1 - oy0);
        for (int k = 0; k < 2; ++k) {  // both staging blocks at once: an allocation stalls everything in flight on the device
            strip8_[k].ensure(out->pitch * rows);
            stripm_[k].ensure(out->mask_pitch * rows);
        }
        uint8_t* s8 = strip8_[slot].as<uint8_t>();
        uint8_t* sm = stripm_[slot].as<uint8_t>();
        o.out8 = s8 - (long long)(oy0 - sub0) * o.pitch8;
        o.mask = sm - (long long)(oy0 - sub0) * o.mpitch;
        o.peer = 0;
    }
    eng_.blend(o, st);
    ISB_CUDA(take_launch_error());  // a kernel of the chain that could not be launched must not pass as a finished run
    if (gcopy) {
        const size_t rows = (size_t)(oy1 - oy0);
        ISB_CUDA(cudaEventRecord(ev_done_[slot], st));
        ISB_CUDA(cudaStreamWaitEvent(copy_stream_, ev_done_[slot], 0));
        copy2d(out->data + (size_t)oy0 * out->pitch, out->pitch, strip8_[slot].as<uint8_t>(), out->pitch, (size_t)rf.w * 3, rows, copy_stream_);
        copy2d(out->mask + (size_t)oy0 * out->mask_pitch, out->mask_pitch, stripm_[slot].as<uint8_t>(), out->mask_pitch, (size_t)rf.w, rows, copy_stream_);
        ISB_CUDA(cudaEventRecord(ev_copied_[slot], copy_stream_));
        copies_pending_ = true;
    }
    ++run_count_;
    // ---- stage 4: results to the host ---------------------------------------------------------
    ISB_CUDA(cudaEventRecord(ev_[4], st));
    const int rows = std::max(oy1 - oy0, 0);
    if (out->data && !d8) copy2d(out->data + (size_t)oy0 * out->pitch, out->pitch, out8_.as<uint8_t>(), (size_t)o.pitch8, (size_t)rf.w * 3, rows, st);
    if (out->mask && !dm) copy2d(out->mask + (size_t)oy0 * out->mask_pitch, out->mask_pitch, outm_.as<uint8_t>(), (size_t)o.mpitch, rf.w, rows, st);
    if (out->data16 && !d16)
        copy2d(reinterpret_cast<char*>(out->data16) + (size_t)oy0 * out->pitch16, out->pitch16, out16_.as<char>(), (size_t)o.pitch16, (size_t)rf.w * 6, rows, st);
    ISB_CUDA(cudaEventRecord(ev_[5], st));
    out->roi_xywh[0] = rf.x; out->roi_xywh[1] = rf.y; out->roi_xywh[2] = rf.w; out->roi_xywh[3] = rf.h;
    out->strip_y0 = oy0;
    out->strip_y1 = oy1;
    // host-visible results (or host-owned inputs) => the call is synchronous; all-device calls stay stream-ordered
    bool any_host = (out->data && !d8) || (out->mask && !dm) || (out->data16 && !d16);
    for (int i = 0; i < n && !any_host; ++i)
        if (!tiles_of_image_[i].empty() && !src_dev[i] && mem_kind(imgs[i].data) == MemKind::Host) any_host = true;
    if (any_host && !cfg_.async_mode) ISB_CUDA(cudaStreamSynchronize(st));
}

// ---- ingest pre-steps of the loop (image_stitching.cpp:1093-1103, 1143-1146) inside the composer -------------------------
bool Composer::ingest_active() const
{
    return cfg_.ingest_rotate != 0 || (cfg_.compose_scale > 0 && std::abs(cfg_.compose_scale - 1) > 1e-1);
}

void Composer::ingest_size(int w, int h, int& ow, int& oh) const
{
    ISB_ASSERT(cfg_.ingest_rotate >= 0 && cfg_.ingest_rotate <= 2);
    int rw = w, rh = h;
    if (cfg_.ingest_rotate == 1 + ISB_ROTATE_90_CLOCKWISE) { rw = h; rh = w; }
    ow = rw;
    oh = rh;
    if (cfg_.compose_scale > 0 && std::abs(cfg_.compose_scale - 1) > 1e-1) {  // :1130-1133 / cv::resize(Size(), fx, fy): cvRound
        ow = (int)std::nearbyint(rw * cfg_.compose_scale);
        oh = (int)std::nearbyint(rh * cfg_.compose_scale);
    }
}

void Composer::ingest(const isb_image* imgs, int n, std::vector<isb_image>& out, cudaStream_t st)
{
    const bool scale = cfg_.compose_scale > 0 && std::abs(cfg_.compose_scale - 1) > 1e-1;
    out.assign(imgs, imgs + n);
    ing_out_.begin();
    ing_tab_.begin();
    std::vector<size_t> off(n), toff(n);
    for (int i = 0; i < n; ++i) {
        off[i] = ing_out_.take((size_t)img_[i].src_w * 3 * img_[i].src_h + 32);
        toff[i] = ing_tab_.take((size_t)(img_[i].src_w + img_[i].src_h) * sizeof(uint32_t));
    }
    char* ob = ing_out_.commit();
    char* tb = ing_tab_.commit();
    std::vector<uint32_t> tx, ty;
    for (int i = 0; i < n; ++i) {
        const isb_image& im = imgs[i];
        if (!im.data) throw Error(ISB_ERR_NULL_PTR, "image data is null");
        ISB_ASSERT(im.width > 0 && im.height > 0 && im.pitch >= (size_t)im.width * 3);
        int ew, eh;
        ingest_size(im.width, im.height, ew, eh);
        ISB_ASSERT(ew == img_[i].src_w && eh == img_[i].src_h);  // the decoded size the plan was made for
        const uint8_t* cur = im.data;
        size_t cp = im.pitch;
        int cw = im.width, chh = im.height;
        if (mem_kind(cur) != MemKind::Device) {
            const size_t rb = (size_t)cw * 3;
            uint8_t* up = static_cast<uint8_t*>(ing_up_.ensure(rb * chh));
            copy2d(up, rb, cur, cp, rb, chh, st);
            cur = up;
            cp = rb;
        }
        uint8_t* dst = reinterpret_cast<uint8_t*>(ob + off[i]);
        if (cfg_.ingest_rotate) {
            const int code = cfg_.ingest_rotate - 1;
            const int rw = code == ISB_ROTATE_90_CLOCKWISE ? chh : cw, rh = code == ISB_ROTATE_90_CLOCKWISE ? cw : chh;
            uint8_t* r = scale ? static_cast<uint8_t*>(ing_rot_.ensure((size_t)rw * 3 * rh)) : dst;
            launch_rotate(cur, cw, chh, 3, (long long)cp, code, r, (long long)rw * 3, st);
            cur = r;
            cp = (size_t)rw * 3;
            cw = rw;
            chh = rh;
        }
        if (scale) {
            build_linear_exact_table(cw, ew, tx, cfg_.compose_scale);
            build_linear_exact_table(chh, eh, ty, cfg_.compose_scale);
            uint32_t* t = reinterpret_cast<uint32_t*>(tb + toff[i]);
            ISB_CUDA(cudaMemcpyAsync(t, tx.data(), tx.size() * 4, cudaMemcpyHostToDevice, st));
            ISB_CUDA(cudaMemcpyAsync(t + tx.size(), ty.data(), ty.size() * 4, cudaMemcpyHostToDevice, st));
            launch_resize_exact(cur, cw, chh, 3, (long long)cp, t, t + tx.size(), dst, ew, eh, (long long)ew * 3, st);
        }
        out[i].data = dst;
        out[i].width = ew;
        out[i].height = eh;
        out[i].pitch = (size_t)ew * 3;
    }
}

// Blender::NO / FeatherBlender (image_stitching.cpp:1086-1229 with blend_type no / feather): the loop call by call, every
// intermediate on the device - warp of the image and of the all-255 mask, compensator apply, convertTo(16S), seam mask
// dilate + INTER_LINEAR_EXACT + AND, feed, blend, saturate to 8U.
void Composer::run_simple(const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seams, int n, isb_pano* out)
{
    cudaStream_t st = current_stream();
    SimpleBlender blender(eff_blend_type_, eff_sharpness_);
    blender.prepare(dst_roi_);
    Warper warper(cfg_.warp_kind, cfg_.warped_image_scale);
    DevBuf warped, ones, maskw, img16;
    for (int i = 0; i < n; ++i) {
        const isb_image& im = imgs[i];
        if (!im.data) throw Error(ISB_ERR_NULL_PTR, "image data is null");
        ISB_ASSERT(im.width == img_[i].src_w && im.height == img_[i].src_h && im.pitch >= (size_t)im.width * 3);
        const Rect roi = img_[i].roi;
        float K[9];
        isb_camera_K(&cams_[i], K);
        uint8_t* w8 = static_cast<uint8_t*>(warped.ensure((size_t)roi.w * 3 * roi.h));
        uint8_t* m1 = static_cast<uint8_t*>(ones.ensure((size_t)im.width * im.height));
        uint8_t* mw = static_cast<uint8_t*>(maskw.ensure((size_t)roi.w * roi.h));
        int16_t* i16 = static_cast<int16_t*>(img16.ensure((size_t)roi.w * 6 * roi.h));
        ISB_CUDA(cudaMemsetAsync(m1, 255, (size_t)im.width * im.height, st));
        warper.warp(im.data, im.width, im.height, 3, im.pitch, K, cams_[i].R, ISB_INTER_LINEAR, ISB_BORDER_REFLECT, w8, (size_t)roi.w * 3, nullptr);
        warper.warp(m1, im.width, im.height, 1, (size_t)im.width, K, cams_[i].R, ISB_INTER_NEAREST, ISB_BORDER_CONSTANT, mw, (size_t)roi.w, nullptr);
        if (gains && gains[i].data) {
            const isb_gainmap& g = gains[i];
            ISB_ASSERT(g.width > 0 && g.height > 0);
            std::vector<float> gh((size_t)g.width * g.height);
            ISB_CUDA(cudaMemcpyAsync(gh.data(), g.data, gh.size() * sizeof(float), cudaMemcpyDefault, st));
            ISB_CUDA(cudaStreamSynchronize(st));
            Compensator comp(64, 64);
            const float* gp = gh.data();
            comp.set_gains(1, &gp, &g.width, &g.height);
            comp.apply(0, w8, roi.w, roi.h, (size_t)roi.w * 3);
        }
        if (seams && seams[i].data) {
            const isb_mask& m = seams[i];
            ISB_ASSERT(m.width > 0 && m.height > 0 && m.pitch >= (size_t)m.width);
            seam_mask_apply(m.data, m.width, m.height, m.pitch, mw, roi.w, roi.h, (size_t)roi.w);
        }
        launch_convert_8u16s(w8, (long long)roi.w * 3, i16, (long long)roi.w * 6, roi.w * 3, roi.h, st);
        blender.feed(i16, (size_t)roi.w * 6, mw, (size_t)roi.w, roi.w, roi.h, roi.x, roi.y);
    }
    const int W = dst_roi_.w, H = dst_roi_.h;
    int16_t* r16 = static_cast<int16_t*>(out16_.ensure((size_t)W * 6 * H));
    const bool dm = out->mask && mem_kind(out->mask) == MemKind::Device;
    uint8_t* rm = dm ? out->mask : static_cast<uint8_t*>(outm_.ensure((size_t)W * H));
    const size_t rmp = dm ? out->mask_pitch : (size_t)W;
    if (out->mask) ISB_ASSERT(out->mask_pitch >= (size_t)W);
    blender.blend(r16, (size_t)W * 6, rm, rmp);
    if (out->mask && !dm) copy2d(out->mask, out->mask_pitch, rm, rmp, W, H, st);
    if (out->data) {
        ISB_ASSERT(out->pitch >= (size_t)W * 3);
        const bool d8 = mem_kind(out->data) == MemKind::Device;
        uint8_t* r8 = d8 ? out->data : static_cast<uint8_t*>(out8_.ensure((size_t)W * 3 * H));
        const size_t p8 = d8 ? out->pitch : (size_t)W * 3;
        launch_convert_16s8u(r16, (long long)W * 6, r8, (long long)p8, W * 3, H, st);
        if (!d8) copy2d(out->data, out->pitch, r8, p8, (size_t)W * 3, H, st);
    }
    if (out->data16) {
        ISB_ASSERT(out->pitch16 >= (size_t)W * 6);
        copy2d(out->data16, out->pitch16, r16, (size_t)W * 6, (size_t)W * 6, H, st);
    }
    ISB_CUDA(cudaGetLastError());
    ISB_CUDA(cudaStreamSynchronize(st));
    out->roi_xywh[0] = dst_roi_.x; out->roi_xywh[1] = dst_roi_.y; out->roi_xywh[2] = W; out->roi_xywh[3] = H;
    out->strip_y0 = 0;
    out->strip_y1 = H;
}

void Composer::join()
{
    if (!copies_pending_) return;
    cudaStream_t st = current_stream();
    for (int k = 0; k < 2; ++k)
        if (run_count_ > (unsigned long long)k) ISB_CUDA(cudaStreamWaitEvent(st, ev_copied_[k], 0));
}

void Composer::sync()
{
    if (ev_init_) ISB_CUDA(cudaEventSynchronize(ev_[5]));
    if (copies_pending_) {
        ISB_CUDA(cudaStreamSynchronize(copy_stream_));
        copies_pending_ = false;
    }
}

int Composer::timings(float* ms, int cap)
{
    if (!ev_init_) return 0;
    ISB_CUDA(cudaEventSynchronize(ev_[5]));
    for (int i = 0; i < ST_COUNT; ++i) {
        float v = 0.f;
        ISB_CUDA(cudaEventElapsedTime(&v, ev_[i], ev_[i + 1]));
        last_ms_[i] = v;
        if (ms && i < cap) ms[i] = v;
    }
    return ST_COUNT;
}

// ------------------------------------------------------------------------------------------------
// ComposerPool
// ------------------------------------------------------------------------------------------------
ComposerPool::ComposerPool(const isb_config& cfg)
{
    const int depth = std::max(1, std::min(cfg.pipeline_depth, 8));
    isb_config c = cfg;
    if (depth > 1) c.async_mode = 1;  // a slot never blocks the host: that is what sync() is for
    for (int k = 0; k < depth; ++k) comps_.push_back(new Composer(c));
    busy_.assign(depth, 0);
}

ComposerPool::~ComposerPool()
{
    for (size_t k = 0; k < streams_.size(); ++k) {
        cudaStreamSynchronize(streams_[k]);
        cudaEventDestroy(fork_[k]);
        cudaEventDestroy(done_[k]);
    }
    for (Composer* c : comps_) delete c;
    for (cudaStream_t s : streams_) cudaStreamDestroy(s);
}

void ComposerPool::plan(const isb_camera* cams, const int* sizes_wh, int n, int* corners, int* sizes, int* dst_roi)
{
    for (Composer* c : comps_) c->plan(cams, sizes_wh, n, corners, sizes, dst_roi);
}

void ComposerPool::run(const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seams, int n, isb_pano* out)
{
    if (comps_.size() == 1) {
        comps_[0]->run(imgs, gains, seams, n, out);
        return;
    }
    if (streams_.empty()) {
        require_device();
        for (size_t k = 0; k < comps_.size(); ++k) {
            cudaStream_t s;
            cudaEvent_t a, b;
            ISB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            ISB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
            ISB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
            streams_.push_back(s);
            fork_.push_back(a);
            done_.push_back(b);
        }
    }
    const int k = (int)(next_++ % comps_.size());
    cudaStream_t caller = current_stream();
    // the slot starts behind everything the caller has enqueued so far (inputs produced on its stream, the plan's uploads)
    ISB_CUDA(cudaEventRecord(fork_[k], caller));
    ISB_CUDA(cudaStreamWaitEvent(streams_[k], fork_[k], 0));
    set_current_stream(streams_[k]);
    try {
        comps_[k]->run(imgs, gains, seams, n, out);
    } catch (...) {
        set_current_stream(caller);
        throw;
    }
    set_current_stream(caller);
    ISB_CUDA(cudaEventRecord(done_[k], streams_[k]));
    busy_[k] = 1;
    last_ = k;
}

void ComposerPool::join()
{
    cudaStream_t caller = current_stream();
    for (size_t k = 0; k < comps_.size(); ++k) {
        comps_[k]->join();  // copy-engine gather of the slot's runs (enqueues waits on the caller's stream)
        if (!streams_.empty() && busy_[k]) ISB_CUDA(cudaStreamWaitEvent(caller, done_[k], 0));
    }
}

void ComposerPool::sync()
{
    for (size_t k = 0; k < comps_.size(); ++k) {
        if (!streams_.empty() && busy_[k]) {
            ISB_CUDA(cudaEventSynchronize(done_[k]));
            busy_[k] = 0;
        }
        comps_[k]->sync();
    }
}

void Composer::byte_model(double* S, double* M, double* Ap, double* B)
{
    if (!planned_) throw Error(ISB_ERR_ASSERT, "Assertion failed: plan() first");
    require_device();
    const int n = (int)img_.size();
    cudaStream_t st = current_stream();
    if (valid_counts_.empty()) {
        // M = number of non-zero pixels of the nearest-warped masks, counted by the device inverse map
        std::vector<ImageDev> idev(n);
        std::vector<int> rw(n), rh(n);
        DevBuf tabs;
        Arena a;
        a.begin();
        std::vector<size_t> co(n), ro(n);
        std::vector<std::vector<Float2>> cols(n), rows(n);
        for (int i = 0; i < n; ++i) {
            build_trig_tables(img_[i].proj, img_[i].roi, cols[i], rows[i]);
            co[i] = a.take(cols[i].size() * sizeof(F2));
            ro[i] = a.take(rows[i].size() * sizeof(F2));
        }
        char* base = a.commit();
        for (int i = 0; i < n; ++i) {
            ISB_CUDA(cudaMemcpyAsync(base + co[i], cols[i].data(), cols[i].size() * sizeof(F2), cudaMemcpyHostToDevice, st));
            ISB_CUDA(cudaMemcpyAsync(base + ro[i], rows[i].data(), rows[i].size() * sizeof(F2), cudaMemcpyHostToDevice, st));
            ImageDev& I = idev[i];
            I = ImageDev{};
            I.sw = img_[i].src_w; I.sh = img_[i].src_h; I.roi_w = rw[i] = img_[i].roi.w; I.roi_h = rh[i] = img_[i].roi.h;
            std::memcpy(I.kr, img_[i].proj.k_rinv, sizeof(I.kr));
            I.col = reinterpret_cast<const F2*>(base + co[i]);
            I.row = reinterpret_cast<const F2*>(base + ro[i]);
        }
        DevBuf idb;
        ImageDev* idp = static_cast<ImageDev*>(idb.ensure(n * sizeof(ImageDev)));
        ISB_CUDA(cudaMemcpyAsync(idp, idev.data(), n * sizeof(ImageDev), cudaMemcpyHostToDevice, st));
        unsigned long long* cd = static_cast<unsigned long long*>(counts_dev_.ensure(n * sizeof(unsigned long long)));
        ISB_CUDA(cudaMemsetAsync(cd, 0, n * sizeof(unsigned long long), st));
        launch_count_valid(idp, n, rw.data(), rh.data(), cd, st);
        valid_counts_.resize(n);
        ISB_CUDA(cudaMemcpyAsync(valid_counts_.data(), cd, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        ISB_CUDA(cudaStreamSynchronize(st));
    }
    double s = 0, m = 0;
    for (int i = 0; i < n; ++i) {
        s += (double)img_[i].src_w * img_[i].src_h;
        m += (double)valid_counts_[i];
    }
    const double ap = (double)dst_roi_.w * dst_roi_.h;
    double P = 0;
    for (int l = 0; l <= eng_.geom().nb; ++l) P += std::pow(4.0, -l);
    if (S) *S = s;
    if (M) *M = m;
    if (Ap) *Ap = ap;
    if (B) *B = 3 * s + 40 * P * m + 10 * P * ap + 4 * ap;
}

}  // namespace isb
