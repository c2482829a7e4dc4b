// device_types.hpp - plain structs shared by the host planner and the CUDA kernels.
#pragma once
#include <cstdint>

namespace isb {

constexpr int kMaxLevels = 16;  // pyramid levels 0..nb, nb <= 15

struct F2 { float a, b; };
struct LinCoefDev { int ofs; float frac; };

// One source image + everything the fused warp needs to evaluate, at any ROI pixel, the value the
// reference's  warp -> mask warp -> compensator.apply -> convertTo(16S) -> seam & mask  chain produces
// (image_stitching.cpp:1154-1171), without stored xmap/ymap.
struct ImageDev {
    const uint8_t* src;   // 8UC3 interleaved (or 8UC1 for mask warps)
    long long spitch;     // bytes
    unsigned sbytes;      // spitch * sh when the vectorised sampler may be used (8-B aligned base, < 4 GB), else 0
    int fast_h;           // rows y0 in [0, fast_h) whose two 16-byte tap windows stay inside the buffer for every x0 <= sw - 2
    int band_lo;          // first source row present behind `src` (strip-sharded runs upload a row band of host sources; `src`
                          // is then the address row 0 WOULD have, and only rows the plan-time band covers are ever read)
    int sw, sh;           // source size
    int roi_w, roi_h;     // warped size (warpRoi)
    float kr[9];          // k_rinv = K * R^T
    float zlo;            // the fused warp divides by z with a shared reciprocal when z > zlo (2^-40, or +inf when the
                          // operands are not provably below 2^40: then every pixel takes the IEEE division path)
    const F2* col;        // [roi_w] (sin u', cos u')
    const F2* row;        // [roi_h] (sin(pi - v'), cos(pi - v')) | (1, v')
    const float* gain;    // gain grid gh x gw (nullptr: no gain)
    int gw, gh;
    const LinCoefDev* gx; // [roi_w]
    const LinCoefDev* gy; // [roi_h]
    const uint8_t* seam;  // dilated low-res seam mask mh x mw, tight pitch (nullptr: all 255)
    const uint8_t* seam_raw;  // the caller's seam mask (input of the batched 3x3 dilate), pitch seam_raw_pitch
    int seam_raw_pitch;
    int mw, mh;
    const uint32_t* mx;   // [roi_w] (ofs << 16) | alpha
    const uint32_t* my;   // [roi_h]
};

// One pyramid-building unit: a rectangle of the padded panorama (aligned to the 2^nb grid) on which one
// image's Gaussian (16S x3 planar) and weight (f32) pyramids live.  Level l has size (w >> l, h >> l).
struct TileDev {
    int x0, y0, w, h;        // level-0 rect, relative to the padded dst_roi_ origin
    int img;                 // index into ImageDev[] (fused path) or -1 (classic feed)
    int left, top;           // ROI top-left relative to the tile origin (copyMakeBorder's left/top)
    int roi_w, roi_h;        // warped image size (tile px outside are REFLECT padding, weight 0)
    // fused path with seam masks: need[cy * need_cw + cx] == generation of the current run iff macro cell (cx, cy) of this
    // tile lies within 4 cells of a cell in which the blend weight (valid & seam) can be non-zero; other cells never reach
    // the output and the warp kernel zero-fills them instead of computing them.  Stamped on the device every run.
    const uint32_t* need;
    int need_cw;
    // REFLECT padding (tile pixels outside the warped ROI) of this tile is mirrored from computed pixels by mirror_pad_kernel:
    // set when every padding pixel's mirror image lies inside the tile (a tile cut by a strip or by support culling may lack
    // the rows / columns its padding reflects onto; kernel 1 then computes the padding itself)
    int mirror;
    // Fused path (`packed` != 0): the warped image is 8-bit, so every Gaussian level stays in [0,255] and is stored
    // byte-packed, one uint32 per pixel = b | g<<8 | r<<16 (| m<<24 at level 0, m = the 8-bit blend mask, whose
    // weight is m * (1/255) exactly as feed() forms it).  Level 0 therefore costs 4 B/px instead of 10 B/px and one
    // 32-bit load fetches all channels.  Weights of levels >= 1 are f32 planes (W).  The classic feed() path
    // (arbitrary 16S input) keeps planar 16S (G) + f32 (W) at every level.
    uint32_t* P[kMaxLevels];
    int ppitch[kMaxLevels];  // elements
    int packed;
    int16_t* G[kMaxLevels];  // plane p at G[l] + p * gplane[l]
    float* W[kMaxLevels];
    int gpitch[kMaxLevels];  // elements
    int wpitch[kMaxLevels];
    long long gplane[kMaxLevels];
};

// A CTA-sized piece of work: block (bx, by) of tile `tile` at the kernel's level.
struct WorkItem { int tile, bx, by, pad; };

// Compact per-(tile, level) view used by kernel 3 on packed tiles: what a quad of level l needs from one covering tile.
// Built once per plan for every CSR entry and level (DstDev::cdesc), so that the kernel reads ONE 48-byte record after
// the cell list instead of chasing cell_tiles -> TileDev -> its per-level arrays.
struct alignas(16) CellTile {
    const uint32_t* p0;   // packed level l
    const uint32_t* p1;   // packed level l + 1
    const float* w0;      // f32 weights of level l (levels >= 1; level 0 carries them in the mask byte), pitch0 as well
    int pitch0, pitch1;   // elements (a packed tile's weight plane has the pitch of its pixel plane)
    int ox, oy;           // tile origin at level l
    int wc, hc;           // size of level l + 1
    int tile, pad[3];     // tile index (selects the tensor map of its level l + 1 plane in the TMA-staged cell kernel)
};
static_assert(sizeof(CellTile) == 64, "kernel 3 moves a record as four 16-byte words");

// Destination state of one blend (MultiBandBlender::prepare)
struct DstDev {
    int nb;
    int pw, ph;                // padded level-0 size
    int fw, fh;                // final (unpadded) size
    uint2* C[kMaxLevels];      // collapsed Laplacian levels 1..nb: 16S x4 interleaved (b, g, r, 0) = 8 B per pixel
    int cpitch[kMaxLevels];    // elements (pixels)
    int cells_x, cells_y;      // macro cells of 2^nb x 2^nb level-0 px
    const int* cell_start;     // CSR: tiles covering each macro cell, ascending feed order
    const int* cell_tiles;
    const CellTile* cdesc;     // packed tiles: record of CSR entry e at level l (< nb) = cdesc[l * n_entries + e]; else nullptr
    int n_entries;
    int row0, row1;            // level-0 rows [row0,row1) this process owns (strip)
    int packed0;               // all tiles carry the byte-packed level 0 (fused composer)
    int max_cell_tiles;        // longest tile list of any macro cell
    // TMA-fed blend kernels: tensor maps (CUtensorMap, device memory), three kinds of (nb + 1) * n_tiles maps each, indexed
    // [kind][level][tile]: kind 0 = packed planes with 32 x 32 boxes (a block's own pixels), kind 1 = packed planes with 24 x 18
    // boxes (the coarser level's neighbourhood), kind 2 = f32 weight planes with 32 x 32 boxes (levels >= 1); tmap_c[l] = the
    // collapsed level C[l] viewed as uint32 pairs, 40 x 18-word boxes.  Levels whose planes are smaller than a box have no
    // valid map and are never addressed.  nullptr: the driver has no tensor-map encoder, the LDG kernels are used.
    const void* tmap_tiles;
    const void* tmap_c;
    int n_tiles;
    int tmap_level0;           // the level-0 maps of kind 0 are valid: the pipelined level-0 kernel applies
    // pipelined kernel at levels l >= 1: covering tiles per 32 x 32 block of the level (sorted union over the macro cells under
    // the block, feed order), CSR start + one CellTile record per entry; nullptr for levels served by the other kernels
    // level 0 of the pipelined kernel: the cell lists without the tiles that cannot carry weight in the cell (plan-time occupancy)
    const int* cell_start0;
    const CellTile* cdesc0;
    const int* blk_start[kMaxLevels];
    const CellTile* blk_desc[kMaxLevels];
    int blk_nbx[kMaxLevels];
};

struct OutDev {
    uint8_t* out8; long long pitch8;     // 8UC3 interleaved, may be null
    uint8_t* mask; long long mpitch;     // 8UC1, may be null
    int16_t* out16; long long pitch16;   // 16SC3 interleaved (bytes pitch), may be null
    int fast8;                           // set by the launcher: the packed 8-bit store path applies
    int odd;                             // set by the launcher: some pointer or pitch of out8 / mask is odd (rows may start at odd addresses)
    int staged;                          // set by the launcher: strip-sharded output, staged vector stores (any alignment)
    int peer;                            // set by the composer: strip-sharded run, the output may live on another GPU
};

}  // namespace isb
