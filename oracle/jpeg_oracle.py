"""TEST INFRASTRUCTURE - CPU restatement of what `cv::imwrite("result.jpg", result)` produces (image_stitching.cpp:1228).

OpenCV's JPEG writer (modules/imgcodecs/src/grfmt_jpeg.cpp, not vendored in /root/reference; OpenCV 4.13 bundles libjpeg-turbo
3.1) drives libjpeg with its defaults: quality 95 (`jpeg_set_quality(95, TRUE)`), YCbCr 4:2:0 (h2v2 chroma), baseline sequential
DCT (`JDCT_ISLOW`), the standard Huffman tables of ITU T.81 Annex K (no optimisation), no restart markers, a JFIF 1.01 header with
density 1:1.  Everything in that pipeline is integer arithmetic, so the byte stream is reproducible exactly:

  colour      jccolor.c rgb_ycc_convert   16-bit fixed point, ONE_HALF rounding, CBCR_OFFSET + ONE_HALF - 1 for Cb / Cr
  edges       jcprepct.c / jcsample.c     last column replicated at full resolution up to the padded width; rows replicated at full
                                          resolution up to an even count, then the last downsampled row up to the padded height
  subsample   jcsample.c h2v2_downsample  (a + b + c + d + bias) >> 2, bias 1, 2, 1, 2 ... along a row
  DCT         jfdctint.c jpeg_fdct_islow  13-bit constants, two passes, output scaled by 8
  quantise    jcdctmgr.c                  sign * ((|c| + q8 / 2) / q8), q8 = 8 * table entry
  dummy       jccoefct.c compress_data    blocks beyond a component's block grid inside the last MCU column / row: AC = 0,
                                          DC = the previous block's quantised DC
  entropy     jchuff.c encode_one_block   DC difference categories, (run, size) AC symbols, ZRL, EOB, 0xFF byte stuffing,
                                          final byte padded with one bits

Pinned by tests/test_jpeg.py against cv2.imencode (the dependency itself) byte for byte.  Only tests may import this module.
"""
import numpy as np

STD_LUMA_Q = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87,
                       80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92,
                       95, 98, 112, 100, 103, 99], np.int64)
STD_CHROMA_Q = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99,
                         99, 99, 99] + [99] * 32, np.int64)
ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42,
                   49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])

DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d]
AC_LUMA_VALS = [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
    0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
    0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
    0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
    0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
    0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
    0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
    0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
    0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
    0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
    0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
    0xfa]


def quant_tables(quality=95):
    """jpeg_set_quality(quality, force_baseline = TRUE): (luma, chroma) in natural order."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - quality * 2
    out = []
    for std in (STD_LUMA_Q, STD_CHROMA_Q):
        t = (std * scale + 50) // 100
        out.append(np.clip(t, 1, 255))
    return out


def huff_codes(bits, vals):
    """jchuff.c jpeg_make_c_derived_tbl: symbol -> (code, length)."""
    sizes = []
    for l, n in enumerate(bits, start=1):
        sizes += [l] * n
    codes, code, si = {}, 0, sizes[0] if sizes else 0
    k = 0
    while k < len(sizes):
        while k < len(sizes) and sizes[k] == si:
            codes[vals[k]] = (code, si)
            code += 1
            k += 1
        code <<= 1
        si += 1
    return codes


def rgb_to_ycc(bgr):
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    half, off = 1 << 15, 128 << 16
    y = (19595 * r + 38470 * g + 7471 * b + half) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + off + half - 1) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + off + half - 1) >> 16
    return y, cb, cr


def pad_edge(a, h, w):
    return np.pad(a, ((0, h - a.shape[0]), (0, w - a.shape[1])), mode="edge")


def h2v2(a):
    s = a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]
    bias = np.where(np.arange(s.shape[1]) % 2 == 0, 1, 2)
    return (s + bias[None, :]) >> 2


def descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _dct_1d(d, first):
    C, P = 13, 2
    tmp0, tmp7 = d[0] + d[7], d[0] - d[7]
    tmp1, tmp6 = d[1] + d[6], d[1] - d[6]
    tmp2, tmp5 = d[2] + d[5], d[2] - d[5]
    tmp3, tmp4 = d[3] + d[4], d[3] - d[4]
    tmp10, tmp13 = tmp0 + tmp3, tmp0 - tmp3
    tmp11, tmp12 = tmp1 + tmp2, tmp1 - tmp2
    o = [None] * 8
    if first:
        o[0] = (tmp10 + tmp11) << P
        o[4] = (tmp10 - tmp11) << P
        n = C - P
    else:
        o[0] = descale(tmp10 + tmp11, P)
        o[4] = descale(tmp10 - tmp11, P)
        n = C + P
    z1 = (tmp12 + tmp13) * 4433
    o[2] = descale(z1 + tmp13 * 6270, n)
    o[6] = descale(z1 + tmp12 * (-15137), n)
    z1, z2, z3, z4 = tmp4 + tmp7, tmp5 + tmp6, tmp4 + tmp6, tmp5 + tmp7
    z5 = (z3 + z4) * 9633
    tmp4, tmp5, tmp6, tmp7 = tmp4 * 2446, tmp5 * 16819, tmp6 * 25172, tmp7 * 12299
    z1, z2, z3, z4 = z1 * (-7373), z2 * (-20995), z3 * (-16069), z4 * (-3196)
    z3, z4 = z3 + z5, z4 + z5
    o[7] = descale(tmp4 + z1 + z3, n)
    o[5] = descale(tmp5 + z2 + z4, n)
    o[3] = descale(tmp6 + z2 + z3, n)
    o[1] = descale(tmp7 + z1 + z4, n)
    return o


def fdct_islow(blocks):
    """blocks: (..., 8, 8) int64 samples already level-shifted by -128; returns coefficients scaled by 8."""
    rows = _dct_1d([blocks[..., :, k] for k in range(8)], True)        # pass 1: along each row
    t = np.stack(rows, axis=-1)
    cols = _dct_1d([t[..., k, :] for k in range(8)], False)            # pass 2: down each column
    return np.stack(cols, axis=-2)


def quantize(coef, q):
    q8 = (q.reshape(8, 8) << 3).astype(np.int64)
    a = np.abs(coef)
    v = (a + (q8 >> 1)) // q8
    return np.where(coef < 0, -v, v)


def component_blocks(plane, q):
    """plane: padded to a multiple of 8 both ways.  Returns (rows_in_blocks, cols_in_blocks, 64) quantised coefficients."""
    h, w = plane.shape
    b = (plane.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3) - 128).astype(np.int64)
    return quantize(fdct_islow(b), q).reshape(h // 8, w // 8, 64)


class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, code, size):
        if size == 0:
            return
        self.acc = (self.acc << size) | (code & ((1 << size) - 1))
        self.n += size
        while self.n >= 8:
            byte = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(byte)
            if byte == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        self.put(0x7F, 7)
        self.acc = 0
        self.n = 0


def encode_block(bw, blk, last_dc, dc_tab, ac_tab):
    t = int(blk[0]) - last_dc
    t2 = t
    if t < 0:
        t = -t
        t2 -= 1
    nbits = t.bit_length()
    bw.put(*dc_tab[nbits])
    bw.put(t2, nbits)
    r = 0
    for k in range(1, 64):
        v = int(blk[ZIGZAG[k]])
        if v == 0:
            r += 1
            continue
        while r > 15:
            bw.put(*ac_tab[0xF0])
            r -= 16
        v2 = v
        if v < 0:
            v = -v
            v2 -= 1
        nb = v.bit_length()
        bw.put(*ac_tab[(r << 4) + nb])
        bw.put(v2, nb)
        r = 0
    if r > 0:
        bw.put(*ac_tab[0])
    return int(blk[0])


def header(w, h, ql, qc):
    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload
    out = bytes([0xFF, 0xD8])
    out += seg(0xE0, b"JFIF\0" + bytes([1, 1, 0, 0, 1, 0, 1, 0, 0]))
    out += seg(0xDB, bytes([0]) + bytes(int(ql[z]) for z in ZIGZAG))
    out += seg(0xDB, bytes([1]) + bytes(int(qc[z]) for z in ZIGZAG))
    out += seg(0xC0, bytes([8]) + h.to_bytes(2, "big") + w.to_bytes(2, "big") + bytes([3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1]))
    for cls_id, bits, vals in ((0x00, DC_LUMA_BITS, DC_VALS), (0x10, AC_LUMA_BITS, AC_LUMA_VALS), (0x01, DC_CHROMA_BITS, DC_VALS),
                               (0x11, AC_CHROMA_BITS, AC_CHROMA_VALS)):
        out += seg(0xC4, bytes([cls_id]) + bytes(bits) + bytes(vals))
    out += seg(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return out


def encode(bgr, quality=95):
    """bgr: HxWx3 uint8 (OpenCV channel order).  Returns the bytes cv2.imencode('.jpg', bgr) produces."""
    h, w = bgr.shape[:2]
    ql, qc = quant_tables(quality)
    y, cb, cr = rgb_to_ycc(bgr)
    mcu_x, mcu_y = (w + 15) // 16, (h + 15) // 16
    # real block grids: luma ceil(w / 8) x ceil(h / 8); chroma ceil(ceil(w / 2) / 8) x ceil(ceil(h / 2) / 8)
    yb_w, yb_h = (w + 7) // 8, (h + 7) // 8
    cw, ch = (w + 1) // 2, (h + 1) // 2
    cb_w, cb_h = (cw + 7) // 8, (ch + 7) // 8
    Y = component_blocks(pad_edge(y, yb_h * 8, yb_w * 8), ql)
    # chroma: along a row the FULL-resolution pixels are replicated up to twice the padded chroma width and then averaged
    # (jcsample.c expand_right_edge); down the image the full-resolution rows are replicated only up to an even count
    # (jcprepct.c, one row group), and it is the last DOWNSAMPLED row that fills the rest of the block row
    def chroma(p):
        return pad_edge(h2v2(pad_edge(p, 2 * ch, cb_w * 16)), cb_h * 8, cb_w * 8)
    Cb = component_blocks(chroma(cb), qc)
    Cr = component_blocks(chroma(cr), qc)
    dcl, acl = huff_codes(DC_LUMA_BITS, DC_VALS), huff_codes(AC_LUMA_BITS, AC_LUMA_VALS)
    dcc, acc = huff_codes(DC_CHROMA_BITS, DC_VALS), huff_codes(AC_CHROMA_BITS, AC_CHROMA_VALS)
    bw = BitWriter()
    last = [0, 0, 0]
    zero = np.zeros(64, np.int64)
    for my in range(mcu_y):
        for mx in range(mcu_x):
            prev = None  # the previous block of this MCU (dummy blocks copy its DC)
            for by in range(2):
                for bx in range(2):
                    r, c = 2 * my + by, 2 * mx + bx
                    if r < yb_h and c < yb_w:
                        blk = Y[r, c]
                    else:
                        blk = zero.copy()
                        blk[0] = prev[0]
                    prev = blk
                    last[0] = encode_block(bw, blk, last[0], dcl, acl)
            for ci, comp in ((1, Cb), (2, Cr)):
                blk = comp[my, mx] if (my < cb_h and mx < cb_w) else None
                assert blk is not None  # the chroma grid always covers the MCU grid (1 x 1 blocks per MCU)
                last[ci] = encode_block(bw, blk, last[ci], dcc, acc)
    bw.flush()
    return header(w, h, ql, qc) + bytes(bw.out) + bytes([0xFF, 0xD9])
