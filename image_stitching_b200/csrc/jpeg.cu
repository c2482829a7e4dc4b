// Output side of the compositing loop: the bytes of cv::imwrite("result.jpg", result) (image_stitching.cpp:1228) on the device.
//
// OpenCV's JPEG writer drives libjpeg(-turbo) with its defaults: quality 95, YCbCr 4:2:0, baseline sequential DCT (JDCT_ISLOW),
// the Huffman tables of ITU T.81 Annex K, no restart markers, JFIF 1.01.  Every step is integer arithmetic, so the stream is
// reproduced byte for byte (oracle/jpeg_oracle.py restates it on the CPU; tests/test_jpeg.py pins both against cv2.imencode):
//   1 jpeg_planes_kernel   BGR (8U, or 16S saturated like imwrite's convertTo) -> Y and h2v2 Cb / Cr planes, padded to whole MCUs
//                          with libjpeg's edge rules (jccolor.c, jcsample.c, jcprepct.c)
//   2 jpeg_dct_kernel      one thread per 8 x 8 block: jfdctint.c forward DCT, jcdctmgr.c quantisation, zig-zag order; the dummy
//                          blocks of the last MCU column / row copy the DC of their neighbour (jccoefct.c)
//   3 jpeg_bits_kernel     entropy-coded length of every block (jchuff.c encode_one_block) -> exclusive scan = bit offsets
//   4 jpeg_emit_kernel     every block writes its code words at its offset (whole words stored, the shared first / last word OR-ed)
//   5 jpeg_stuff_kernel    0xFF -> 0xFF 0x00 byte stuffing: count per word, scan, scatter behind the header
// The entropy coder is sequential in libjpeg only through the DC prediction (previous block of the component) and the bit
// position; the former is a neighbour lookup, the latter a prefix sum.
#include <cstring>
#include <vector>

#include "engine.hpp"
#include "kernels.cuh"

namespace isb {
namespace {

// ---- tables (ITU T.81 Annex K; jcparam.c) ------------------------------------------------------------------------------------
const uint8_t kStdLumaQ[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                               18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kStdChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1,
    0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a,
    0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3,
    0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1,
    0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca,
    0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kZigzagHost[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JpegTables {
    uint16_t q8[2][64];      // 8 * quantisation table (natural order): the divisor of the scaled DCT output
    uint32_t dc[2][12];      // Huffman code | length << 16, by category
    uint32_t ac[2][256];     // by (run << 4 | size)
};

// jpeg_set_quality(quality, force_baseline = TRUE) -> jpeg_add_quant_table
void quant_table(const uint8_t* std_tbl, int quality, uint8_t out[64])
{
    quality = std::min(std::max(quality, 1), 100);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; ++i) {
        long t = ((long)std_tbl[i] * scale + 50L) / 100L;
        out[i] = (uint8_t)std::min(std::max(t, 1L), 255L);
    }
}

// jchuff.c jpeg_make_c_derived_tbl
void derive(const uint8_t bits[16], const uint8_t* vals, uint32_t* table)
{
    unsigned code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i) table[vals[k++]] = code++ | ((uint32_t)len << 16);
        code <<= 1;
    }
}

// ---- 1: colour conversion + subsampling ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_bgr(const T* __restrict__ src, long long pitch, int x, int y, int& b, int& g, int& r)
{
    const T* p = reinterpret_cast<const T*>(reinterpret_cast<const char*>(src) + (long long)y * pitch) + 3 * x;
    b = p[0]; g = p[1]; r = p[2];
    if (sizeof(T) == 2) {  // imwrite converts a 16S image with saturate_cast<uchar>
        b = min(max(b, 0), 255); g = min(max(g, 0), 255); r = min(max(r, 0), 255);
    }
}
__device__ __forceinline__ int ycc_y(int b, int g, int r) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
__device__ __forceinline__ int ycc_cb(int b, int g, int r) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
__device__ __forceinline__ int ycc_cr(int b, int g, int r) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

// One thread per chroma sample of the padded grid (= one 2 x 2 quad of the padded luma plane).
// Edge rules: luma replicates the last column and row; chroma replicates the full-resolution pixels along a row and, down the
// image, only up to an even row count - beyond that the last DOWNSAMPLED row repeats (jcprepct.c pads the downsampled output).
template <typename T>
__global__ void __launch_bounds__(256) jpeg_planes_kernel(const T* __restrict__ src, long long pitch, int w, int h, int pw, int ph,
                                                          uint8_t* __restrict__ Y, uint8_t* __restrict__ Cb, uint8_t* __restrict__ Cr)
{
    const int cx = blockIdx.x * 32 + (threadIdx.x & 31), cy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (cx >= pw / 2 || cy >= ph / 2) return;
    const int xa = min(2 * cx, w - 1), xb = min(2 * cx + 1, w - 1);
    const int ya = min(2 * cy, h - 1), yb = min(2 * cy + 1, h - 1);
    int b[4], g[4], r[4];
    load_bgr(src, pitch, xa, ya, b[0], g[0], r[0]);
    load_bgr(src, pitch, xb, ya, b[1], g[1], r[1]);
    load_bgr(src, pitch, xa, yb, b[2], g[2], r[2]);
    load_bgr(src, pitch, xb, yb, b[3], g[3], r[3]);
    uint8_t* y0 = Y + (size_t)(2 * cy) * pw + 2 * cx;
    *reinterpret_cast<uint16_t*>(y0) = (uint16_t)(ycc_y(b[0], g[0], r[0]) | (ycc_y(b[1], g[1], r[1]) << 8));
    *reinterpret_cast<uint16_t*>(y0 + pw) = (uint16_t)(ycc_y(b[2], g[2], r[2]) | (ycc_y(b[3], g[3], r[3]) << 8));
    const int ch = (h + 1) / 2;
    if (cy >= ch) {  // below the image: the chroma row of the last image rows
        const int cyl = ch - 1, y2a = min(2 * cyl, h - 1), y2b = min(2 * cyl + 1, h - 1);
        load_bgr(src, pitch, xa, y2a, b[0], g[0], r[0]);
        load_bgr(src, pitch, xb, y2a, b[1], g[1], r[1]);
        load_bgr(src, pitch, xa, y2b, b[2], g[2], r[2]);
        load_bgr(src, pitch, xb, y2b, b[3], g[3], r[3]);
    }
    const int bias = 1 + (cx & 1);
    int sb = bias, sr = bias;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sb += ycc_cb(b[k], g[k], r[k]);
        sr += ycc_cr(b[k], g[k], r[k]);
    }
    Cb[(size_t)cy * (pw / 2) + cx] = (uint8_t)(sb >> 2);
    Cr[(size_t)cy * (pw / 2) + cx] = (uint8_t)(sr >> 2);
}

// ---- 2: forward DCT + quantisation --------------------------------------------------------------------------------------------
__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c: one 8-point pass in place on d[0], d[s], ..., d[7 s]
template <bool FIRST>
__device__ __forceinline__ void fdct8(int* d, int s)
{
    constexpr int C = 13, P = 2, N = FIRST ? C - P : C + P;
    const int tmp0 = d[0] + d[7 * s], tmp7 = d[0] - d[7 * s];
    const int tmp1 = d[s] + d[6 * s], tmp6 = d[s] - d[6 * s];
    const int tmp2 = d[2 * s] + d[5 * s], tmp5 = d[2 * s] - d[5 * s];
    const int tmp3 = d[3 * s] + d[4 * s], tmp4 = d[3 * s] - d[4 * s];
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    if (FIRST) {
        d[0] = (tmp10 + tmp11) << P;
        d[4 * s] = (tmp10 - tmp11) << P;
    } else {
        d[0] = descale(tmp10 + tmp11, P);
        d[4 * s] = descale(tmp10 - tmp11, P);
    }
    int z1 = (tmp12 + tmp13) * 4433;
    d[2 * s] = descale(z1 + tmp13 * 6270, N);
    d[6 * s] = descale(z1 + tmp12 * (-15137), N);
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    const int z5 = (z3 + z4) * 9633;
    const int t4 = tmp4 * 2446, t5 = tmp5 * 16819, t6 = tmp6 * 25172, t7 = tmp7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    d[7 * s] = descale(t4 + z1 + z3, N);
    d[5 * s] = descale(t5 + z2 + z4, N);
    d[3 * s] = descale(t6 + z2 + z3, N);
    d[s] = descale(t7 + z1 + z4, N);
}

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct JpegGeom {
    int w, h, pw, ph;        // image and MCU-padded size
    int mcu_x, mcu_y;
    int yb_w, yb_h;          // luma blocks that hold image pixels
    long long n_blocks;      // 6 per MCU, scan order: Y00 Y01 Y10 Y11 Cb Cr
};

// One thread per block in scan order; coefficients leave in zig-zag order (int16 x 64 per block).
__global__ void __launch_bounds__(128) jpeg_dct_kernel(JpegGeom G, const uint8_t* __restrict__ Y, const uint8_t* __restrict__ Cb,
                                                       const uint8_t* __restrict__ Cr, const JpegTables* __restrict__ tb, int16_t* __restrict__ coef)
{
    __shared__ int sblk[128][65];  // the transform runs in shared memory: one padded row per thread, no bank conflicts
    const long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
    if (gid >= G.n_blocks) return;
    const int k = (int)(gid % 6);
    const long long mcu = gid / 6;
    const int mx = (int)(mcu % G.mcu_x), my = (int)(mcu / G.mcu_x);
    const uint8_t* plane;
    int pitch, r, c;
    bool dummy = false;
    if (k < 4) {
        plane = Y; pitch = G.pw;
        r = 2 * my + (k >> 1); c = 2 * mx + (k & 1);
        if (r >= G.yb_h) {          // a row of dummy blocks below the image: the DC of the MCU's block (0, 1) - itself a copy of
            dummy = true;           // block (0, 0) when it lies beyond the last column
            r = 2 * my; c = min(2 * mx + 1, G.yb_w - 1);
        } else if (c >= G.yb_w) {   // dummy block right of the image: the DC of its left neighbour
            dummy = true;
            c = G.yb_w - 1;
        }
    } else {
        plane = k == 4 ? Cb : Cr; pitch = G.pw / 2;
        r = my; c = mx;
    }
    int* d = sblk[threadIdx.x];
    const uint8_t* p = plane + (size_t)(8 * r) * pitch + 8 * c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint2 v = *reinterpret_cast<const uint2*>(p + (size_t)j * pitch);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d[8 * j + i] = (int)((v.x >> (8 * i)) & 0xFF) - 128;
            d[8 * j + 4 + i] = (int)((v.y >> (8 * i)) & 0xFF) - 128;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) fdct8<true>(d + 8 * j, 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) fdct8<false>(d + i, 8);
    const uint16_t* __restrict__ q8 = tb->q8[k < 4 ? 0 : 1];
    uint32_t* out = reinterpret_cast<uint32_t*>(coef + gid * 64);
#pragma unroll 4
    for (int z = 0; z < 64; z += 2) {
        int v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = c_zigzag[z + e];
            const int x = d[n], q = q8[n];
            const int a = (abs(x) + (q >> 1)) / q;
            v[e] = (dummy && z + e > 0) ? 0 : (x < 0 ? -a : a);
        }
        out[z >> 1] = ((uint32_t)v[0] & 0xFFFFu) | ((uint32_t)v[1] << 16);
    }
}

// ---- 3, 4: entropy coding ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int bit_length(int v) { return 32 - __clz(v); }  // v >= 0

// DC predictor: the previous block of the same component in scan order
__device__ __forceinline__ int pred_dc(const int16_t* __restrict__ coef, long long gid)
{
    const int k = (int)(gid % 6);
    if (k >= 1 && k <= 3) return coef[(gid - 1) * 64];
    if (gid < 6) return 0;
    return coef[(gid - (k == 0 ? 3 : 6)) * 64];
}

// jchuff.c encode_one_block; Sink::put(code, length)
template <typename Sink>
__device__ __forceinline__ void encode_block(const int16_t* __restrict__ blk, int last_dc, const uint32_t* __restrict__ dc_tab,
                                             const uint32_t* __restrict__ ac_tab, Sink& sink)
{
    const uint4* b4 = reinterpret_cast<const uint4*>(blk);
    int run = 0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const uint4 w = __ldg(b4 + q);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int v = (int)(short)(ww[e >> 1] >> (16 * (e & 1)));
            if (q == 0 && e == 0) {
                int t = v - last_dc, t2 = t;
                if (t < 0) { t = -t; --t2; }
                const int nb = bit_length(t);
                const uint32_t c = dc_tab[nb];
                sink.put(((c & 0xFFFFu) << nb) | ((uint32_t)t2 & ((1u << nb) - 1u)), (int)(c >> 16) + nb);
                continue;
            }
            if (v == 0) { ++run; continue; }
            while (run > 15) {
                const uint32_t c = ac_tab[0xF0];
                sink.put(c & 0xFFFFu, (int)(c >> 16));
                run -= 16;
            }
            int v2 = v;
            if (v < 0) { v = -v; --v2; }
            const int nb = bit_length(v);
            const uint32_t c = ac_tab[(run << 4) + nb];
            sink.put(((c & 0xFFFFu) << nb) | ((uint32_t)v2 & ((1u << nb) - 1u)), (int)(c >> 16) + nb);
            run = 0;
        }
    }
    if (run > 0) {
        const uint32_t c = ac_tab[0];
        sink.put(c & 0xFFFFu, (int)(c >> 16));
    }
}

struct CountSink {
    uint32_t bits = 0;
    __device__ __forceinline__ void put(uint32_t, int len) { bits += (uint32_t)len; }
};

__global__ void __launch_bounds__(256) jpeg_bits_kernel(long long n, const int16_t* __restrict__ coef, const JpegTables* __restrict__ tb,
                                                        uint32_t* __restrict__ bits)
{
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    if (gid >= n) return;
    const int t = gid % 6 < 4 ? 0 : 1;
    CountSink s;
    encode_block(coef + gid * 64, pred_dc(coef, gid), tb->dc[t], tb->ac[t], s);
    bits[gid] = s.bits;
}

// big-endian bit stream in 32-bit words: bit 0 of the stream is the top bit of word 0
struct WordSink {
    uint32_t* words;
    unsigned long long acc = 0;  // the low `nb` bits are pending; the first `lead` of them belong to the previous block (zeros here)
    int nb;
    long long word;
    bool first = true;
    __device__ WordSink(uint32_t* w, unsigned long long pos) : words(w), nb((int)(pos & 31)), word((long long)(pos >> 5)) {}
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        acc = (acc << len) | code;
        nb += len;
        if (nb >= 32) {
            const uint32_t v = (uint32_t)(acc >> (nb - 32));
            if (first) atomicOr(words + word, v);  // shared with the previous block
            else words[word] = v;                  // all 32 bits are this block's
            first = false;
            ++word;
            nb -= 32;
            acc &= (1ull << nb) - 1ull;
        }
    }
    __device__ __forceinline__ void flush()
    {
        if (nb > 0) atomicOr(words + word, (uint32_t)(acc << (32 - nb)));  // shared with the next block (or the tail padding)
    }
};

__global__ void __launch_bounds__(256) jpeg_emit_kernel(long long n, const int16_t* __restrict__ coef, const JpegTables* __restrict__ tb,
                                                        const unsigned long long* __restrict__ offs, const uint32_t* __restrict__ bits,
                                                        uint32_t* __restrict__ words)
{
    const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
    if (gid >= n) return;
    const int t = gid % 6 < 4 ? 0 : 1;
    WordSink s(words, offs[gid]);
    encode_block(coef + gid * 64, pred_dc(coef, gid), tb->dc[t], tb->ac[t], s);
    if (gid == n - 1) {  // jchuff.c flush_bits: the last byte is filled up with one bits
        const unsigned long long total = offs[gid] + bits[gid];
        const int pad = (int)((8 - (total & 7)) & 7);
        if (pad) s.put((1u << pad) - 1u, pad);
    }
    s.flush();
}

// ---- exclusive scan of 32-bit counts into 64-bit offsets -------------------------------------------------------------------------
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(256) scan_partials_kernel(const uint32_t* __restrict__ in, long long n, unsigned long long* __restrict__ part)
{
    __shared__ unsigned long long s[256];
    const long long base = (long long)blockIdx.x * kScanBlock;
    unsigned long long v = 0;
    for (int i = threadIdx.x; i < kScanBlock; i += 256)
        if (base + i < n) v += in[base + i];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}
// one CTA: exclusive scan of the partial sums in place; part[np] = total
__global__ void __launch_bounds__(1024) scan_top_kernel(unsigned long long* part, long long np)
{
    __shared__ unsigned long long s[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < np; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < np ? part[i] : 0ull;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const unsigned long long a = (int)threadIdx.x >= d ? s[threadIdx.x - d] : 0ull;
            __syncthreads();
            s[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < np) part[i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[np] = carry;
}
__global__ void __launch_bounds__(256) scan_apply_kernel(const uint32_t* __restrict__ in, long long n, const unsigned long long* __restrict__ part,
                                                         unsigned long long* __restrict__ out)
{
    __shared__ unsigned long long s[256];
    const long long base = (long long)blockIdx.x * kScanBlock;
    // thread t owns elements [4 t, 4 t + 4) of the block
    uint32_t v[4];
    unsigned long long sum = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long long i = base + 4 * threadIdx.x + e;
        v[e] = i < n ? in[i] : 0u;
        sum += v[e];
    }
    s[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        const unsigned long long a = (int)threadIdx.x >= d ? s[threadIdx.x - d] : 0ull;
        __syncthreads();
        s[threadIdx.x] += a;
        __syncthreads();
    }
    unsigned long long run = part[blockIdx.x] + s[threadIdx.x] - sum;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const long long i = base + 4 * threadIdx.x + e;
        if (i < n) out[i] = run;
        run += v[e];
    }
}

void exclusive_scan(const uint32_t* in, long long n, unsigned long long* out, unsigned long long* part, cudaStream_t st)
{
    const long long np = (n + kScanBlock - 1) / kScanBlock;
    scan_partials_kernel<<<(unsigned)np, 256, 0, st>>>(in, n, part);
    scan_top_kernel<<<1, 1024, 0, st>>>(part, np);
    scan_apply_kernel<<<(unsigned)np, 256, 0, st>>>(in, n, part, out);
    count_launch(); count_launch(); count_launch();
}

// ---- 5: byte stuffing -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) jpeg_ff_count_kernel(const uint32_t* __restrict__ words, long long n_words, long long n_bytes,
                                                            uint32_t* __restrict__ cnt)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_words) return;
    const uint32_t w = words[i];
    uint32_t c = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (4 * i + b < n_bytes && ((w >> (24 - 8 * b)) & 0xFFu) == 0xFFu) ++c;
    cnt[i] = c;
}
__global__ void __launch_bounds__(256) jpeg_stuff_kernel(const uint32_t* __restrict__ words, long long n_words, long long n_bytes,
                                                         const unsigned long long* __restrict__ ff_before, uint8_t* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_words) return;
    const uint32_t w = words[i];
    uint8_t* o = out + 4 * i + ff_before[i];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (4 * i + b >= n_bytes) break;
        const uint8_t v = (uint8_t)(w >> (24 - 8 * b));
        *o++ = v;
        if (v == 0xFF) *o++ = 0;
    }
}

void put_segment(std::vector<uint8_t>& o, uint8_t marker, const std::vector<uint8_t>& payload)
{
    o.push_back(0xFF);
    o.push_back(marker);
    const size_t len = payload.size() + 2;
    o.push_back((uint8_t)(len >> 8));
    o.push_back((uint8_t)len);
    o.insert(o.end(), payload.begin(), payload.end());
}

// jcmarker.c: SOI, APP0 (JFIF 1.01, density 1:1), DQT x 2, SOF0, DHT x 4, SOS
std::vector<uint8_t> jpeg_header(int w, int h, const uint8_t ql[64], const uint8_t qc[64])
{
    std::vector<uint8_t> o = {0xFF, 0xD8};
    put_segment(o, 0xE0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int t = 0; t < 2; ++t) {
        std::vector<uint8_t> p = {(uint8_t)t};
        for (int z = 0; z < 64; ++z) p.push_back((t ? qc : ql)[kZigzagHost[z]]);
        put_segment(o, 0xDB, p);
    }
    put_segment(o, 0xC0, {8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
    struct H { uint8_t id; const uint8_t* bits; const uint8_t* vals; int n; };
    const H hs[4] = {{0x00, kDcLumaBits, kDcVals, 12}, {0x10, kAcLumaBits, kAcLumaVals, 162}, {0x01, kDcChromaBits, kDcVals, 12},
                     {0x11, kAcChromaBits, kAcChromaVals, 162}};
    for (const H& t : hs) {
        std::vector<uint8_t> p = {t.id};
        p.insert(p.end(), t.bits, t.bits + 16);
        p.insert(p.end(), t.vals, t.vals + t.n);
        put_segment(o, 0xC4, p);
    }
    put_segment(o, 0xDA, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
    return o;
}

struct JpegWorkspace {
    DevBuf b[11];
    int tables_quality = -1;  // quality whose tables b[2] holds (uploaded once: an upload per call would queue behind whatever
                              // large host -> device copy another stream has in flight on the copy engine)
    void* pinned = nullptr;  // host staging of the finished stream: a pageable destination is filled from here
    size_t pinned_cap = 0;
    uint8_t* host(size_t bytes)
    {
        if (bytes > pinned_cap) {
            if (pinned) cudaFreeHost(pinned);
            pinned = nullptr;
            pinned_cap = 0;
            ISB_CUDA(cudaHostAlloc(&pinned, bytes + bytes / 4, cudaHostAllocDefault));
            pinned_cap = bytes + bytes / 4;
        }
        return static_cast<uint8_t*>(pinned);
    }
    void release()
    {
        for (DevBuf& d : b) d.release();
        tables_quality = -1;
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr;
        pinned_cap = 0;
    }
    ~JpegWorkspace() { if (pinned) cudaFreeHost(pinned); }
};
JpegWorkspace& workspace()
{
    static thread_local JpegWorkspace ws;
    static thread_local int ws_device = -1;
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != ws_device) {  // the blocks belong to the device they were allocated on
        if (ws_device >= 0) {
            cudaSetDevice(ws_device);
            ws.release();
            cudaSetDevice(dev);
        }
        ws_device = dev;
    }
    return ws;
}

}  // namespace

void jpeg_release_workspace()
{
    workspace().release();
}

void jpeg_encode(const void* image, int W, int H, size_t pitch, int is_16s, int quality, uint8_t* out, size_t capacity, size_t* out_size)
{
    require_device();
    if (!image || !out_size) throw Error(ISB_ERR_NULL_PTR, "image/out_size are null");
    const size_t es = is_16s ? 6 : 3;
    ISB_ASSERT(W > 0 && H > 0 && pitch >= (size_t)W * es);
    if (W > 65500 || H > 65500) throw Error(ISB_ERR_OUT_OF_RANGE, "jpeg: image larger than libjpeg's JPEG_MAX_DIMENSION (65500)");
    cudaStream_t st = current_stream();
    JpegGeom G{};
    G.w = W; G.h = H;
    G.mcu_x = (W + 15) / 16; G.mcu_y = (H + 15) / 16;
    G.pw = 16 * G.mcu_x; G.ph = 16 * G.mcu_y;
    G.yb_w = (W + 7) / 8; G.yb_h = (H + 7) / 8;
    G.n_blocks = 6ll * G.mcu_x * G.mcu_y;
    uint8_t ql[64], qc[64];
    quant_table(kStdLumaQ, quality, ql);
    quant_table(kStdChromaQ, quality, qc);
    JpegTables tb{};
    for (int i = 0; i < 64; ++i) {
        tb.q8[0][i] = (uint16_t)(ql[i] << 3);
        tb.q8[1][i] = (uint16_t)(qc[i] << 3);
    }
    derive(kDcLumaBits, kDcVals, tb.dc[0]);
    derive(kDcChromaBits, kDcVals, tb.dc[1]);
    derive(kAcLumaBits, kAcLumaVals, tb.ac[0]);
    derive(kAcChromaBits, kAcChromaVals, tb.ac[1]);
    const std::vector<uint8_t> hdr = jpeg_header(W, H, ql, qc);

    // grow-only work buffers, kept per calling thread between calls (an encode of a 60 MP panorama needs ~ 0.4 GB of them, and
    // allocating and freeing that much costs more than the kernels); isb_jpeg_release_workspace() gives them back
    JpegWorkspace& ws = workspace();
    DevBuf &ibuf = ws.b[0], &planes = ws.b[1], &tbuf = ws.b[2], &cbuf = ws.b[3], &bbuf = ws.b[4], &obuf = ws.b[5], &pbuf = ws.b[6], &wbuf = ws.b[7],
           &fbuf = ws.b[8], &fobuf = ws.b[9], &dout = ws.b[10];
    const void* d = image;
    size_t dp = pitch;
    if (mem_kind(image) != MemKind::Device) {
        dp = (size_t)W * es;
        void* p = ibuf.ensure(dp * H);
        copy2d(p, dp, image, pitch, dp, H, st);
        d = p;
    }
    const size_t ysz = (size_t)G.pw * G.ph, csz = ysz / 4;
    uint8_t* Yp = static_cast<uint8_t*>(planes.ensure(ysz + 2 * csz));
    uint8_t* Cbp = Yp + ysz;
    uint8_t* Crp = Cbp + csz;
    JpegTables* td = static_cast<JpegTables*>(tbuf.ensure(sizeof(JpegTables)));
    const int qkey = std::min(std::max(quality, 1), 100);
    if (ws.tables_quality != qkey) {
        ISB_CUDA(cudaMemcpyAsync(td, &tb, sizeof(tb), cudaMemcpyHostToDevice, st));
        ISB_CUDA(cudaStreamSynchronize(st));  // `tb` is a local
        ws.tables_quality = qkey;
    }
    {
        const dim3 grid((G.pw / 2 + 31) / 32, (G.ph / 2 + 7) / 8);
        if (is_16s) jpeg_planes_kernel<int16_t><<<grid, 256, 0, st>>>(static_cast<const int16_t*>(d), (long long)dp, W, H, G.pw, G.ph, Yp, Cbp, Crp);
        else jpeg_planes_kernel<uint8_t><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(d), (long long)dp, W, H, G.pw, G.ph, Yp, Cbp, Crp);
        count_launch();
    }
    const long long n = G.n_blocks;
    int16_t* coef = static_cast<int16_t*>(cbuf.ensure((size_t)n * 64 * sizeof(int16_t)));
    jpeg_dct_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(G, Yp, Cbp, Crp, td, coef);
    count_launch();
    uint32_t* bits = static_cast<uint32_t*>(bbuf.ensure((size_t)n * 4));
    jpeg_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, coef, td, bits);
    count_launch();
    unsigned long long* offs = static_cast<unsigned long long*>(obuf.ensure((size_t)n * 8));
    const long long np = (n + kScanBlock - 1) / kScanBlock;
    unsigned long long* part = static_cast<unsigned long long*>(pbuf.ensure((size_t)(np + 1) * 8));
    exclusive_scan(bits, n, offs, part, st);
    unsigned long long total_bits = 0;
    ISB_CUDA(cudaMemcpyAsync(&total_bits, part + np, 8, cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    const long long n_bytes = (long long)((total_bits + 7) / 8), n_words = (n_bytes + 3) / 4;
    // (the data-dependent buffers are requested with a quarter of headroom: a video's next frame is rarely larger than that)
    const size_t room = (size_t)n_words + (size_t)n_words / 4 + 16;
    uint32_t* words = static_cast<uint32_t*>(wbuf.ensure(room * 4));
    ISB_CUDA(cudaMemsetAsync(words, 0, (size_t)(n_words + 1) * 4, st));
    jpeg_emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, coef, td, offs, bits, words);
    count_launch();
    // byte stuffing: count, scan, scatter
    uint32_t* ffc = static_cast<uint32_t*>(fbuf.ensure(room * 4));
    unsigned long long* ffo = static_cast<unsigned long long*>(fobuf.ensure(room * 8));
    const long long npw = (n_words + kScanBlock - 1) / kScanBlock;
    unsigned long long* partw = static_cast<unsigned long long*>(pbuf.ensure((size_t)(std::max(np, npw) + npw / 4 + 2) * 8));
    jpeg_ff_count_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(words, n_words, n_bytes, ffc);
    count_launch();
    exclusive_scan(ffc, n_words, ffo, partw, st);
    unsigned long long total_ff = 0;
    ISB_CUDA(cudaMemcpyAsync(&total_ff, partw + npw, 8, cudaMemcpyDeviceToHost, st));
    ISB_CUDA(cudaStreamSynchronize(st));
    const size_t body = (size_t)n_bytes + (size_t)total_ff, total = hdr.size() + body + 2;
    *out_size = total;
    if (!out || capacity < total) throw Error(ISB_ERR_OUT_OF_RANGE, "jpeg: the output buffer is smaller than the stream (out_size holds the size needed)");
    const bool odev = mem_kind(out) == MemKind::Device;
    const uint8_t eoi[2] = {0xFF, 0xD9};
    if (odev) {
        ISB_CUDA(cudaMemcpyAsync(out, hdr.data(), hdr.size(), cudaMemcpyHostToDevice, st));
        jpeg_stuff_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(words, n_words, n_bytes, ffo, out + hdr.size());
        count_launch();
        ISB_CUDA(cudaMemcpyAsync(out + hdr.size() + body, eoi, 2, cudaMemcpyHostToDevice, st));
        ISB_CUDA(cudaStreamSynchronize(st));  // hdr / eoi are locals
    } else {
        // host destination: only the entropy-coded body crosses the bus; the header and the EOI marker are written by the host
        // (no host -> device copy at all: it would wait on the copy engine behind other streams' uploads)
        uint8_t* dbody = static_cast<uint8_t*>(dout.ensure(body + body / 4));
        jpeg_stuff_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(words, n_words, n_bytes, ffo, dbody);
        count_launch();
        uint8_t* hp = mem_kind(out) == MemKind::HostPinned ? out + hdr.size() : ws.host(body);
        ISB_CUDA(cudaMemcpyAsync(hp, dbody, body, cudaMemcpyDeviceToHost, st));
        std::memcpy(out, hdr.data(), hdr.size());
        ISB_CUDA(cudaStreamSynchronize(st));
        if (hp != out + hdr.size()) std::memcpy(out + hdr.size(), hp, body);
        std::memcpy(out + hdr.size() + body, eoi, 2);
    }
    ISB_CUDA(cudaGetLastError());
}

}  // namespace isb
