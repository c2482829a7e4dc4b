// engine.hpp - host-side objects behind the C ABI: device buffers, the pyramid engine shared by the
// classic MultiBandBlender surface and the fused composer, and the warper / compensator mirrors.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/image_stitching.h"
#include "device_types.hpp"
#include "geometry.hpp"
#include "kernels.cuh"

namespace isb {

// Errors travel as exceptions inside the library and are converted to codes at the C boundary.
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define ISB_ASSERT(cond)                                                                                  \
    do {                                                                                                  \
        if (!(cond)) throw ::isb::Error(ISB_ERR_ASSERT, std::string("Assertion failed: ") + #cond);      \
    } while (0)
void cuda_check(cudaError_t e, const char* what);
#define ISB_CUDA(call) ::isb::cuda_check((call), #call)

cudaStream_t current_stream();
void set_current_stream(cudaStream_t s);
void require_device();  // throws ISB_ERR_GPU_API when no CUDA device is usable (no CPU fallback exists)

enum class MemKind { Host, HostPinned, Device };
MemKind mem_kind(const void* p);
// copy a (rows x row_bytes) block between any two memory kinds on `st`
void copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows, cudaStream_t st);

// growable device allocation (never shrinks; reused across calls)
class DevBuf {
public:
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf();
    void* ensure(size_t bytes);
    void release();  // frees the block (synchronises like cudaFree)
    template <typename T> T* as() const { return static_cast<T*>(p_); }
    size_t capacity() const { return cap_; }
private:
    void* p_ = nullptr;
    size_t cap_ = 0;
};

// bump allocator over one device block; sized by a dry run, then replayed
class Arena {
public:
    void begin() { off_ = 0; }
    size_t take(size_t bytes) { size_t o = off_; off_ = (off_ + bytes + 255) & ~size_t(255); return o; }
    size_t used() const { return off_; }
    char* commit() { return static_cast<char*>(buf_.ensure(off_ ? off_ : 256)); }
    char* base() const { return buf_.as<char>(); }
private:
    DevBuf buf_;
    size_t off_ = 0;
};

// per-image inputs of the fused warp that do not depend on pixel data (built at plan time)
struct ImagePlan {
    Projector proj;
    Rect roi;            // warpRoi
    int src_w = 0, src_h = 0;
    std::vector<Float2> col, row;
    // offsets into the composer's table arena
    size_t col_off = 0, row_off = 0, gx_off = 0, gy_off = 0, mx_off = 0, my_off = 0;
    int gain_w = -1, gain_h = -1, seam_w = -1, seam_h = -1;  // sizes the cached coefficient tables were built for
    float row_bmax = -1.f;  // max |row[j].b| (bounds the inverse-map operands), computed on first use
};

// The pyramid engine: owns the per-tile pyramids and the destination pyramid, runs kernels 2 and 3.
class PyramidEngine {
public:
    // Start a blend over the sub-panorama rows [sub_y0, sub_y0 + sub_h) of the padded ROI; rows
    // [own_y0, own_y1) (padded coords) are the ones written to the output.
    void reset(const BlendGeometry& g, int sub_y0, int sub_h, int own_y0, int own_y1, bool packed);
    // Add the tile feed() would use for an image at `tl` of size (w x h); clipped to the sub-panorama.
    // Returns the tile index or -1 when the tile does not touch the sub-panorama.
    int add_tile(int img_index, int w, int h, int tlx, int tly);
    // Same for an explicit rectangle (padded-ROI coordinates, on the 2^nb grid) of the image whose warped ROI
    // has its top-left at (tlx, tly) in panorama coordinates.
    int add_rect(int img_index, int X0, int Y0, int W, int H, int tlx, int tly, int roi_w, int roi_h,
                 const uint32_t* need_grid = nullptr, int grid_X0 = 0, int grid_Y0 = 0, int grid_cw = 0);
    // pyrDown l -> l+1 runs the register-rolling kernel: always when the output width is even (l + 1 < nb); the last level
    // of packed tiles (odd widths possible) when every tile is at least two macro cells wide (last_fast_, commit_tiles)
    bool fast_down(int l) const { return l + 1 < g_.nb || (last_fast_ && l >= 1); }
    // output rows per warp of the register-rolling pyrDown: long runs amortise the 3-row halo, short runs give the small
    // levels more CTAs (they are latency-bound)
    static int fast_rows(int l) { return l == 0 ? kFastDownRows : l == 1 ? kFastDownRowsLevel1 : kFastDownRowsSmall; }
    // Allocate pyramid storage for tiles [first, end) (device pointers filled in), upload descriptors.
    void commit_tiles(cudaStream_t st);
    // kernel 2 for tiles [first, end): all levels
    void build_pyramids(int first, int end, cudaStream_t st);
    // kernel 3 for all levels; writes the owned rows
    void blend(const OutDev& out, cudaStream_t st);

    const BlendGeometry& geom() const { return g_; }
    std::vector<TileDev>& tiles() { return tiles_; }
    const TileDev* tiles_dev() const { return tiles_dev_.as<TileDev>(); }
    int sub_y0() const { return sub_y0_; }
    int own_y0() const { return own_y0_; }
    int own_y1() const { return own_y1_; }
    const std::vector<WorkItem>& warp_work() const { return warp_work_; }
    const WorkItem* warp_work_dev() const { return warp_work_dev_.as<WorkItem>(); }
    bool mirror_pad() const { return mirror_pad_; }
    double pad_fraction() const { return pad_fraction_; }
    const std::vector<WorkItem>& pad_work() const { return pad_work_; }
    const WorkItem* pad_work_dev() const { return pad_work_dev_.as<WorkItem>(); }
    size_t pyramid_bytes() const { return arena_.used(); }
    // Plan-time occupancy (which macro cells of an image's full feed() tile hold a valid pixel): lets the blend skip, per cell
    // and level, tiles whose weights are provably zero there (level 0: unoccupied cells; levels 1 .. nb-1: farther than one cell
    // from an occupied one, because W_l spreads by 2 (2^l - 1) < 2^nb pixels).  grid: cw x ch cells whose cell (0, 0) sits at
    // padded-panorama pixel (gx0, gy0).  Tiles without a registered grid are taken as occupied everywhere.
    void set_image_occupancy(int img, std::vector<uint8_t> grid, int cw, int ch, int gx0, int gy0);
    bool uses_tma() const { return use_tma_; }

private:
    BlendGeometry g_;
    int sub_y0_ = 0, sub_h_ = 0, own_y0_ = 0, own_y1_ = 0;
    std::vector<TileDev> tiles_;
    int committed_ = 0;           // tiles [0, committed_) have storage
    Arena arena_;                 // fused path: one block for all tiles
    std::vector<DevBuf*> extra_;  // classic path: one allocation per late-added tile
    DevBuf tiles_dev_, warp_work_dev_, down_work_dev_, cells_dev_, cdesc_dev_, dst_buf_, tmaps_dev_, tmaps_blend_dev_, blk_start_dev_, blk_desc_dev_;
    bool last_fast_ = false;
    bool use_tma_ = false;  // pyrDown staged by TMA at levels [0, tma_levels_) (packed tiles large enough for a full box there)
    int tma_levels_ = 0;
    std::vector<WorkItem> warp_work_, pad_work_;  // blocks kernel 1 computes / blocks that hold REFLECT padding (kernel 1b)
    DevBuf pad_work_dev_;
    bool mirror_pad_ = false;
    double pad_fraction_ = 0.0;
    static constexpr double kMirrorPadThreshold = 0.12;  // share of the tile area outside the warped ROIs
    std::vector<std::vector<WorkItem>> down_work_;  // per level, for tiles of the last commit
    std::vector<size_t> down_off_;
    DstDev dst_{};
    bool packed_ = false;
    struct OccGrid { std::vector<uint8_t> g; int cw = 0, ch = 0, x0 = 0, y0 = 0; };
    std::vector<OccGrid> img_occ_;
    DevBuf cells0_dev_, cdesc0_dev_;
public:
    ~PyramidEngine();
};

class Warper {
public:
    Warper(int kind, float scale) : kind_(kind), scale_(scale) {}
    int kind() const { return kind_; }
    float scale() const { return scale_; }
    void set_scale(float s) { scale_ = s; }
    Rect warp_roi(int sw, int sh, const float* K, const float* R);
    void warp_point(const float* pt, const float* K, const float* R, float* out, bool backward);
    Rect build_maps(int sw, int sh, const float* K, const float* R, float* xmap, float* ymap, size_t pitch);
    void warp(const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* K, const float* R, int interp,
              int border, uint8_t* dst, size_t dpitch, int* corner);
    // warpBackward(src = warped image of size warpRoi(dst_size), ..., dst_size, dst)
    void warp_backward(const uint8_t* src, int sw, int sh, int ch, size_t spitch, const float* K, const float* R, int interp,
                       int border, int dw, int dh, uint8_t* dst, size_t dpitch);
private:
    void prepare_image(int sw, int sh, const float* K, const float* R, ImageDev& I, Rect& roi, cudaStream_t st);
    int kind_;
    float scale_;
    DevBuf src_, dst_, tab_, xm_, ym_;
};

class Compensator {
public:
    Compensator(int bw, int bh) : bw_(bw), bh_(bh) {}
    void set_gains(int n, const float* const* g, const int* gw, const int* gh);
    int count() const { return (int)gains_.size(); }
    const std::vector<float>& gain(int i, int& w, int& h) const { w = gw_[i]; h = gh_[i]; return gains_[i]; }
    void apply(int index, uint8_t* image, int w, int h, size_t pitch);
private:
    int bw_, bh_;
    std::vector<std::vector<float>> gains_;
    std::vector<int> gw_, gh_;
    DevBuf img_, aux_;
};

// ingest pre-steps (image_stitching.cpp:1093-1103, 1143-1146)
void rotate_image(const uint8_t* src, int w, int h, int ch, size_t spitch, int code, uint8_t* dst, size_t dpitch);
void resize_linear_exact(const uint8_t* src, int sw, int sh, int ch, size_t spitch, uint8_t* dst, int dw, int dh, size_t dpitch,
                         double fx, double fy);

// crop() of the reference (cropper.cpp:116-209) on the device: crop.cu
void crop_rect(const uint8_t* mask, int W, int H, size_t pitch, int rect_xywh[4], int* n_points);
void crop_rect_image(const void* img, int W, int H, size_t pitch, int is_16s, int rect_xywh[4], int* n_points);

// cv::imwrite("result.jpg", result) of the reference (image_stitching.cpp:1228) on the device: jpeg.cu
void jpeg_encode(const void* image, int W, int H, size_t pitch, int is_16s, int quality, uint8_t* out, size_t capacity, size_t* out_size);
void jpeg_release_workspace();  // frees the calling thread's cached work buffers of jpeg_encode

void seam_mask_apply(const uint8_t* seam, int mw, int mh, size_t spitch, uint8_t* mask, int w, int h, size_t pitch);

class Blender {
public:
    explicit Blender(int nb) : requested_(nb) {}
    void set_num_bands(int nb) { requested_ = nb; }
    int num_bands() const { return requested_; }
    int actual_bands() const { return eng_.geom().nb; }
    bool prepared() const { return prepared_; }
    void prepare(const Rect& roi);
    const BlendGeometry& geom() const { return eng_.geom(); }
    void feed(const int16_t* img, size_t ipitch, const uint8_t* mask, size_t mpitch, int w, int h, int tlx, int tly);
    void blend(int16_t* dst, size_t dpitch, uint8_t* dmask, size_t mpitch);
private:
    int requested_;
    bool prepared_ = false;
    PyramidEngine eng_;
    DevBuf img_, mask_, out16_, outm_;
};

// cv::detail::Blender (type NO) and cv::detail::FeatherBlender (image_stitching.cpp:1175-1191): destination-resident
// 16SC3 image, 8U mask and (feather) f32 weight sum; kernels in simple_blend.cu.
class SimpleBlender {
public:
    SimpleBlender(int type, float sharpness) : type_(type), sharpness_(sharpness) {}
    int type() const { return type_; }
    float sharpness() const { return sharpness_; }
    void set_sharpness(float s) { sharpness_ = s; }
    const Rect& roi() const { return roi_; }
    void prepare(const Rect& roi);
    void feed(const int16_t* img, size_t ipitch, const uint8_t* mask, size_t mpitch, int w, int h, int tlx, int tly);
    void blend(int16_t* dst, size_t dpitch, uint8_t* dmask, size_t mpitch);
    // cv::detail::createWeightMap(mask, sharpness, weight) on its own
    static void weight_map(const uint8_t* mask, size_t mpitch, int w, int h, float sharpness, float* weight, size_t wpitch);
private:
    int type_;
    float sharpness_;
    bool prepared_ = false;
    Rect roi_{};
    DevBuf dst_, dmask_, dweight_, img_, mask_, wmap_, dist_;
};

// cv::detail::Timelapser / TimelapserCrop (image_stitching.cpp:1194-1215): every process() call clears the canvas and copies
// the part of one warped image that lies inside dst_roi_ (AS_IS: resultRoi of all images; CROP: cv's Rect(max tl, min br)).
class Timelapser {
public:
    explicit Timelapser(int type) : type_(type) {}
    void initialize(const int* corners_xy, const int* sizes_wh, int n);
    const Rect& roi() const { return roi_; }
    void process(const int16_t* img, size_t ipitch, int w, int h, int tlx, int tly);
    void get_dst(int16_t* dst, size_t dpitch);
private:
    int type_;
    bool ready_ = false;
    Rect roi_{};
    DevBuf dst_, img_;
};

class Composer {
public:
    explicit Composer(const isb_config& cfg) : cfg_(cfg) {}
    void plan(const isb_camera* cams, const int* sizes_wh, int n, int* corners, int* sizes, int* dst_roi);
    void run(const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seams, int n, isb_pano* out);
    void sync();
    void join();
    int timings(float* ms, int cap);
    void byte_model(double* S, double* M, double* Ap, double* B);
    static const char* stage_name(int i);
    // the blender the reference's rule selects for the planned panorama (image_stitching.cpp:1173-1193)
    int effective_blend_type() const { return eff_blend_type_; }
    int effective_num_bands() const { return eff_num_bands_; }
    float effective_sharpness() const { return eff_sharpness_; }
private:
    // ingest pre-steps (rotate + compose-scale resize) on the device: returns device-resident stand-ins of `imgs`
    bool ingest_active() const;
    void ingest_size(int w, int h, int& ow, int& oh) const;
    void ingest(const isb_image* imgs, int n, std::vector<isb_image>& out, cudaStream_t st);
    Arena ing_out_, ing_tab_;
    DevBuf ing_up_, ing_rot_;
    void run_simple(const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seams, int n, isb_pano* out);
    int eff_blend_type_ = ISB_BLENDER_MULTI_BAND, eff_num_bands_ = 0;
    float eff_sharpness_ = 0.02f;
    bool same_plan(const isb_camera* cams, const int* sizes_wh, int n) const;
    isb_config cfg_;
    bool planned_ = false;
    std::vector<isb_camera> cams_;
    std::vector<int> src_sizes_;
    std::vector<ImagePlan> img_;
    std::vector<std::vector<int>> tiles_of_image_;
    PyramidEngine eng_;
    Rect dst_roi_;
    Arena tables_;              // trig tables (static per plan)
    Arena dyn_;                 // per-run uploads: sources, gains, seam masks, their coefficient tables
    DevBuf imgs_dev_, counts_dev_, out8_, outm_, out16_;
    // seam-aware culling: plan-time valid occupancy per macro cell (device), per-run map of needed cells
    DevBuf occ_valid_dev_, need_dev_, occ_tiles_dev_, seam_blk_dev_;
    std::vector<OccTile> occ_tiles_;      // one per image: its full feed() tile on the 2^nb grid
    std::vector<int> last_seam_blk_;      // block prefix sums of the seam preparation launch as last uploaded
    uint32_t need_gen_ = 0;  // run counter stamped into the need map
    std::vector<unsigned long long> valid_counts_;
    cudaEvent_t ev_[8] = {};
    bool ev_init_ = false;
    // descriptors of the last run as uploaded to imgs_dev_ (a run with identical descriptors skips the upload)
    std::vector<ImageDev> last_idev_;
    float last_ms_[8] = {};
    // strip-sharded runs
    std::vector<int> src_band_;   // per image [lo, hi]: source rows this strip can read (plan time, src_band_kernel)
    DevBuf band_dev_;
    size_t h2d_bytes_ = 0;        // source bytes the last run uploaded
    // ISB_GATHER_COPY_ENGINE: local double-buffered strip + copy-engine push into the caller's (peer) panorama
    DevBuf strip8_[2], stripm_[2];
    cudaStream_t copy_stream_ = nullptr;
    cudaEvent_t ev_done_[2] = {}, ev_copied_[2] = {};
    unsigned long long run_count_ = 0;
    bool copies_pending_ = false;
public:
    ~Composer();
    size_t last_h2d_bytes() const { return h2d_bytes_; }
    // rows [y0, y1) of the final panorama this (strip-sharded) composer produces; valid after plan()
    void planned_rows(int& y0, int& y1) const
    {
        const int fh = eng_.geom().roi_final.h;
        y0 = std::min(eng_.own_y0(), fh);
        y1 = std::min(eng_.own_y1(), fh);
    }
    const std::vector<int>& src_band() const { return src_band_; }
};

// isb_config::pipeline_depth > 1: that many Composers (own per-image pyramids, tables and stream each) served round-robin, so
// that consecutive isb_composer_run() calls overlap on the device: the latency-bound coarse-level kernels of step k leave most
// SMs idle, and the issue-bound kernels of step k + 1 fill them.  A run is then asynchronous with respect to the caller's
// stream: its output is valid behind isb_composer_join() (stream-ordered) or after isb_composer_sync() (host), and runs in
// flight must write to different output buffers.  Depth 1 is a plain Composer on the caller's stream.
class ComposerPool {
public:
    explicit ComposerPool(const isb_config& cfg);
    ~ComposerPool();
    ComposerPool(const ComposerPool&) = delete;
    ComposerPool& operator=(const ComposerPool&) = delete;
    void plan(const isb_camera* cams, const int* sizes_wh, int n, int* corners, int* sizes, int* dst_roi);
    void run(const isb_image* imgs, const isb_gainmap* gains, const isb_mask* seams, int n, isb_pano* out);
    void sync();
    void join();
    Composer& first() { return *comps_[0]; }
    Composer& last() { return *comps_[last_]; }
    int depth() const { return (int)comps_.size(); }
private:
    std::vector<Composer*> comps_;
    std::vector<cudaStream_t> streams_;
    std::vector<cudaEvent_t> fork_, done_;
    std::vector<char> busy_;
    unsigned long long next_ = 0;
    int last_ = 0;
};

}  // namespace isb
