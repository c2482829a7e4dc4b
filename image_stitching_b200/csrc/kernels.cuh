// kernels.cuh - launch wrappers of the sm_100a kernels (definitions in kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "device_types.hpp"

namespace isb {

long long launch_count(bool reset);
void count_launch();

// ---- programmatic dependent launch (PDL) of the composer's kernel chain ------------------------------------------
// The fused step is a chain of 14+ kernels of which nine are tiny (coarse pyramid levels, seam maps).  Every kernel of
// the chain starts with pdl_prologue(): it first releases its own dependents (so the next kernel's CTAs may become
// resident while this one drains its last wave - they hold no resources a CTA of this grid still waits for, because the
// trigger only fires once ALL CTAs of this grid have executed it) and then waits until the grid it depends on has
// completed and flushed its memory.  Launched without the attribute (classic per-call API) both instructions are no-ops.
bool pdl_enabled();  // ISB_PDL=0 turns the launch attribute off (A/B measurements)
// run-time measurement switches, read from the environment once per process (reload_env_switches() re-reads them)
struct EnvSwitches {
    bool pdl;            // ISB_PDL=0: launch the chain without programmatic stream serialization
    bool blend_tma;      // ISB_BLEND_TMA=0: LDG cell kernels instead of the TMA-staged ones
    bool blend_pipe;     // ISB_BLEND_PIPE=0: no pipelined (persistent) blend kernels
    bool staged_stores;  // ISB_STAGED_STORES: force the staged 16-byte store path of the level-0 blend on local memory
    int pyrdown_tma_levels;  // ISB_PYRDOWN_TMA_LEVELS=n: TMA-staged pyrDown at the first n levels only (0: none; default: all that qualify)
};
const EnvSwitches& env_switches();
void reload_env_switches();
// first error a chained launch returned since the last call (cudaSuccess if none); the engine turns it into ISB_ERR_GPU_API
void note_launch_error(cudaError_t e);
cudaError_t take_launch_error();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_prologue()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// kernel<<<grid, block, smem, st>>>(args...) with programmatic stream serialization allowed
template <typename... KArgs, typename... Args>
inline void launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    note_launch_error(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    count_launch();
}
#endif

// ---- classic per-call API ---------------------------------------------------------------------
// RotationWarper::warp: dst(rect_h x rect_w, `ch` channels) from src via the separable-table inverse map.
void launch_warp_generic(const ImageDev& img, int ch, int interp, int border, uint8_t* dst, long long dpitch,
                         cudaStream_t st);
// RotationWarper::warpBackward: dst (dw x dh, the original image size) from the warped image `warped` (src / sw / sh / spitch
// of the descriptor) through the forward map r_kinv = R * K^-1; (tlx, tly) = top-left of the warped ROI
void launch_warp_backward(const ImageDev& warped, int ch, const float* r_kinv, float scale, int spherical, int tlx, int tly, int interp,
                          int border, int dw, int dh, uint8_t* dst, long long dpitch, cudaStream_t st);
// RotationWarper::buildMaps (maps are only ever stored for this diagnostic/parity entry point)
void launch_build_maps(const ImageDev& img, float* xmap, float* ymap, long long pitch_bytes, cudaStream_t st);
// BlocksGainCompensator::apply on 8UC3 in place
void launch_gain_apply(uint8_t* img, int w, int h, long long pitch, const float* gain, int gw, int gh,
                       const LinCoefDev* gx, const LinCoefDev* gy, cudaStream_t st);
// dilate 3x3 (8UC1): src pitch given, dst tight
void launch_dilate3x3(const uint8_t* src, int w, int h, long long spitch, uint8_t* dst, cudaStream_t st);
// mask &= resize_linear_exact(dilated)
void launch_seam_and(const uint8_t* dil, int mw, int mh, const uint32_t* mx, const uint32_t* my, uint8_t* mask, int w,
                     int h, long long pitch, cudaStream_t st);
// feed(): copyMakeBorder(REFLECT) of a 16SC3 image / CONSTANT of mask*(1/255) into a tile's level 0
void launch_pack_tile(const TileDev* tile_dev, const TileDev& tile_host, const int16_t* img, long long ipitch,
                      const uint8_t* mask, long long mpitch, cudaStream_t st);

// element-wise depth conversions (row_elems = width * channels): convertTo(CV_16S) and saturate_cast<uchar>
void launch_convert_8u16s(const uint8_t* src, long long spitch, int16_t* dst, long long dpitch_bytes, int row_elems, int h, cudaStream_t st);
void launch_convert_16s8u(const int16_t* src, long long spitch_bytes, uint8_t* dst, long long dpitch, int row_elems, int h, cudaStream_t st);

// ingest pre-steps of the compositing loop (image_stitching.cpp:1093-1103, 1143-1146)
// cv::rotate: code 0 = ROTATE_90_CLOCKWISE (dst is h x w), 1 = ROTATE_180
void launch_rotate(const uint8_t* src, int w, int h, int ch, long long spitch, int code, uint8_t* dst, long long dpitch,
                   cudaStream_t st);
// cv::resize(INTER_LINEAR_EXACT) on 8U, 1 or 3 channels; tx/ty = per-destination (ofs << 16) | alpha tables
void launch_resize_exact(const uint8_t* src, int sw, int sh, int ch, long long spitch, const uint32_t* tx, const uint32_t* ty,
                         uint8_t* dst, int dw, int dh, long long dpitch, cudaStream_t st);

// ---- batched pyramid pipeline -----------------------------------------------------------------
// number of valid (mask != 0) warped pixels of every image -> counts[img] (unsigned long long)
void launch_count_valid(const ImageDev* imgs_dev, int n_img, const int* roi_w_host, const int* roi_h_host,
                        unsigned long long* counts_dev, cudaStream_t st);
// G[l+1], W[l+1] = pyrDown(G[l], W[l]) for every tile block in `work`
void launch_pyrdown_tiles(const WorkItem* work, int n_work, const TileDev* tiles, int level, cudaStream_t st);
// accumulate (image order) + normalise + collapse one level of the destination; level 0 writes the output
void launch_blend_level(const DstDev& dst, const TileDev* tiles, int level, const OutDev& out, cudaStream_t st);

// ---- fast path of the fused composer (kernels_fast.cu): byte-packed level 0 ------------------------
// occupancy of the valid mask per macro cell (2^nb x 2^nb px, tile-aligned) of every listed tile -> occ (bytes)
struct OccTile { int img; int left, top; int w, h; long long occ_off; };
void launch_occupancy(const OccTile* tiles_dev, int n_tiles, int max_w, int max_h, const ImageDev* imgs, int nb,
                      uint8_t* occ, cudaStream_t st);
// fused warp -> packed level 0
// banded_sources: some ImageDev::band_lo is non-zero (see ImageDev); mirrored_padding: `work` omits the blocks outside the warped
// ROIs and launch_mirror_pad follows (the kernel then skips every pixel outside the ROI)
void launch_warp_tiles_packed(const WorkItem* work, int n_work, const TileDev* tiles, const ImageDev* imgs, int nb, uint32_t gen,
                              bool banded_sources, bool mirrored_padding, cudaStream_t st);
// kernel 1b: mirrors the tile pixels outside the warped ROI (REFLECT padding of feed()) from the pixels kernel 1 stored; `work`
// lists the kWarpBlockW x kWarpBlockH blocks that hold padding
void launch_mirror_pad(const WorkItem* work, int n_work, const TileDev* tiles, cudaStream_t st);
// per-run seam preparation in one launch: (1) cv::dilate(3x3) of every image's seam mask, imgs[i].seam_raw -> imgs[i].seam;
// (2) seam-aware culling: every macro cell that holds a valid pixel (plan-time occupancy) and in which the upsampled
// dilated seam mask can be non-zero stamps `gen` into need[] for all cells within 4 cells of it.  One OccTile per image
// (its full feed() tile on the 2^nb grid).  blk_dev = 2 (n_img + 1) block prefix sums: dilate blocks of 128 x 8 mask pixels
// per image, then - continuing the count - culling blocks of 32 x 8 cells per image (none when culling is off);
// n_blocks = the last entry.
void launch_seam_prep(const ImageDev* imgs_dev, int n_img, const int* blk_dev, int n_blocks, const OccTile* tiles_dev, int nb,
                      const uint8_t* occ_valid, uint32_t* need, uint32_t gen, cudaStream_t st);
// register-rolling separable pyrDown, 2 outputs per thread (packed or planar storage)
// odd_width: the level's output width may be odd (last level of packed tiles at least two macro cells wide)
void launch_pyrdown_fast(const WorkItem* work, int n_work, const TileDev* tiles, int level, bool packed, int rows_per_warp,
                         bool odd_width, cudaStream_t st);
// TMA-staged pyrDown of packed tiles, level -> level + 1: pmaps / wmaps = one CUtensorMap per tile over the level's packed pixels /
// f32 weights (device memory; wmaps is not read at level 0, whose weights are the mask bytes of the pixels)
void launch_pyrdown_tma(const WorkItem* work, int n_work, const TileDev* tiles, const void* pmaps, const void* wmaps, int level, cudaStream_t st);
constexpr int kTmaOutW = 64, kTmaOutH = 32;                          // outputs per CTA
constexpr int kTmaBoxW = 2 * kTmaOutW + 8, kTmaBoxH = 2 * kTmaOutH + 3;  // 136 x 67 input box (x origin 2*ox0 - 4)
// 2x2-quad accumulate + normalise + collapse for level < nb
void launch_blend_quad(const DstDev& dst, const TileDev* tiles, int level, const OutDev& out, cudaStream_t st);
// rows [y0, y1) of `level` a run produces (level 0: the owned rows; coarser levels: what those rows reach through pyrUp)
void blend_level_rows(const DstDev& dst, int level, int& y0, int& y1);
// plan time: per-image [min, max] of the source rows the fused warp can read over the given blocks -> band[2 img], band[2 img + 1]
// (initialise to INT_MAX / INT_MIN)
void launch_src_band(const WorkItem* work, int n_work, const TileDev* tiles, const ImageDev* imgs, int* band, cudaStream_t st);

constexpr int kFastDownCols = 64;   // output columns per warp (2 per lane)
#ifndef ISB_DOWN_ROWS_L1
#define ISB_DOWN_ROWS_L1 4
#endif
#ifndef ISB_DOWN_ROWS_SMALL
#define ISB_DOWN_ROWS_SMALL 4
#endif
constexpr int kFastDownRows = 16;   // output rows per warp (level 0 -> 1)
constexpr int kFastDownRowsLevel1 = ISB_DOWN_ROWS_L1;    // level 1 -> 2
constexpr int kFastDownRowsSmall = ISB_DOWN_ROWS_SMALL;  // coarser levels, where parallelism matters more than halo reuse
constexpr int kFastDownWarps = 8;   // warps per CTA, stacked in y  -> CTA block = 64 x 128 outputs

// geometry of the CTA blocks the planner must use when it builds work lists
#ifndef ISB_WARP_BLOCK_H
#define ISB_WARP_BLOCK_H 128
#endif
constexpr int kWarpBlockW = 64, kWarpBlockH = ISB_WARP_BLOCK_H;
constexpr int kDownBlockW = 32, kDownBlockH = 8;  // in OUTPUT (level l+1) pixels

}  // namespace isb
