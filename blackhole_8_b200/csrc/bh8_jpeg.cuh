// bh8_jpeg.cuh -- baseline JPEG encoder of the frame sink, own kernels.
//
// What cv::VideoWriter(fourcc 'MJPG')::write(frame) does per frame on the host (blackhole_solution_test.cc:71-72,
// 334) happens here on the device-resident BGR8 frame: colour conversion BGR -> YCbCr (JFIF), 4:2:0 chroma
// subsampling, 8x8 forward DCT, quantisation with the IJG tables scaled by the quality factor, Huffman coding
// with the standard tables of ITU-T T.81 Annex K.  Only the bitstream leaves the GPU.
//
// Parallel entropy coding.  Huffman coding is sequential in its bit position; JPEG's restart markers are the
// standard's own way to cut that chain: with a restart interval of `ri` MCUs (DRI segment) every interval
// starts byte-aligned with the DC predictors reset, so intervals are coded independently -- one warp per
// interval writes its bytes (0xFF stuffing included) into its own slot, a scan of the slot lengths gives
// every interval its place in the file, and a gather kernel packs them with the RSTm markers in between.
// Every baseline decoder (libjpeg / OpenCV / ffmpeg) reads such files.  Cost: 2 bytes of marker per interval
// and absolute instead of differential DC values at every interval start (+6 % file size at ri = 2, quality 95).
//
//   jpeg_transform_kernel   one 64-thread CTA per 16x16 MCU: pixels -> 6 blocks of 64 quantised coefficients,
//                           zigzag order, int16
//   jpeg_entropy_kernel     one warp per restart interval, lanes share a block's coefficients: -> bytes in the
//                           interval's slot
//                           interval's slot; per-CTA byte totals
//   jpeg_gather_kernel      one warp per interval: the CTA sums the totals of the CTAs before it (its place in
//                           the stream), then slot -> stream, RSTm / EOI marker
//
// The per-MCU arithmetic and the bit writer are __host__ __device__: tests/host_harness runs the very same
// code on the CPU and OpenCV decodes the result (tests/test_jpeg_host.py); the shipped library only calls
// them from the kernels.
#ifndef BH8_JPEG_CUH_
#define BH8_JPEG_CUH_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BH8J_HD __host__ __device__ __forceinline__
#else
#define BH8J_HD inline
#endif

namespace bh8jpeg {

constexpr int kSlotBytesPerMcu = 3072;  // room for the worst case of 384 coefficients at 27 bits each, stuffed

// natural (row-major) index -> position in the zigzag sequence (T.81 Figure A.6)
static const uint8_t kZigzagOfNatural[64] = {
    0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

// IJG / T.81 Annex K.1 quantisation tables (natural order), the ones libjpeg and OpenCV scale by the quality
static const uint8_t kBaseQuant[2][64] = {
    {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,  69,  56,
     14, 17, 22, 29, 51,  87,  80,  62,  18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

// T.81 Annex K.3 Huffman tables: BITS (codes per length 1..16) and HUFFVAL
static const uint8_t kDcBits[2][16] = {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}};
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcBits[2][16] = {{0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d},
                                       {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}};
static const uint8_t kAcVals[2][162] = {
    {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81,
     0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18,
     0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
     0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75,
     0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
     0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
     0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5,
     0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa},
    {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08,
     0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25,
     0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47,
     0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74,
     0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97,
     0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
     0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4,
     0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa}};

// Everything the kernels look up, built once per sink on the host (make_tables) and copied to the device.
struct Tables {
  float qrecip[2][64];      // 1 / quantiser, natural order; [0] luminance, [1] chrominance
  uint8_t zigzag[64];       // natural index -> zigzag position
  uint32_t dc_code[2][12];  // Huffman code | length << 16, by magnitude category
  uint32_t ac_code[2][256]; // ... by (run << 4 | category)
};

// Geometry of a frame as the encoder sees it.
struct Geometry {
  int32_t width, height;
  int32_t mcus_x, mcus_y, n_mcus;  // 16x16 MCUs (4:2:0), the last row / column padded by edge replication
  int32_t ri, n_intervals;         // restart interval in MCUs and their number
  int32_t header_bytes;            // SOI .. SOS segment, already at the start of the output buffer
};

inline Geometry make_geometry(int width, int height, int ri, int header_bytes) {
  Geometry g;
  g.width = width;
  g.height = height;
  g.mcus_x = (width + 15) / 16;
  g.mcus_y = (height + 15) / 16;
  g.n_mcus = g.mcus_x * g.mcus_y;
  g.ri = ri;
  g.n_intervals = (g.n_mcus + ri - 1) / ri;
  g.header_bytes = header_bytes;
  return g;
}

inline void quant_table(int quality, int which, uint8_t out_natural[64]) {  // jpeg_quality_scaling + jpeg_add_quant_table
  const int q = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
  const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
  for (int i = 0; i < 64; ++i) {
    int v = (kBaseQuant[which][i] * scale + 50) / 100;
    out_natural[i] = (uint8_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
  }
}

inline void huffman_codes(const uint8_t bits[16], const uint8_t* vals, int n_vals, uint32_t* code_by_symbol) {
  uint32_t code = 0;
  int k = 0;
  for (int len = 1; len <= 16; ++len) {
    for (int i = 0; i < bits[len - 1] && k < n_vals; ++i, ++k) code_by_symbol[vals[k]] = code++ | ((uint32_t)len << 16);
    code <<= 1;
  }
}

inline void make_tables(int quality, Tables* t) {
  memset(t, 0, sizeof *t);
  for (int w = 0; w < 2; ++w) {
    uint8_t q[64];
    quant_table(quality, w, q);
    for (int i = 0; i < 64; ++i) t->qrecip[w][i] = 1.0f / (float)q[i];
    huffman_codes(kDcBits[w], kDcVals, 12, t->dc_code[w]);
    huffman_codes(kAcBits[w], kAcVals[w], 162, t->ac_code[w]);
  }
  memcpy(t->zigzag, kZigzagOfNatural, 64);
}

// SOI, APP0 (JFIF), DQT x2, SOF0, DHT x4, DRI, SOS: everything before the entropy-coded data.
inline int make_header(int width, int height, int quality, int ri, uint8_t* out /* >= 700 bytes */) {
  uint8_t* p = out;
  const auto put16 = [&](int v) {
    *p++ = (uint8_t)(v >> 8);
    *p++ = (uint8_t)v;
  };
  *p++ = 0xFF; *p++ = 0xD8;
  *p++ = 0xFF; *p++ = 0xE0; put16(16);
  memcpy(p, "JFIF\0", 5); p += 5;
  *p++ = 1; *p++ = 1; *p++ = 0; put16(1); put16(1); *p++ = 0; *p++ = 0;
  for (int w = 0; w < 2; ++w) {
    uint8_t q[64], zz[64];
    quant_table(quality, w, q);
    for (int i = 0; i < 64; ++i) zz[kZigzagOfNatural[i]] = q[i];
    *p++ = 0xFF; *p++ = 0xDB; put16(67); *p++ = (uint8_t)w;
    memcpy(p, zz, 64); p += 64;
  }
  *p++ = 0xFF; *p++ = 0xC0; put16(17); *p++ = 8; put16(height); put16(width); *p++ = 3;
  *p++ = 1; *p++ = 0x22; *p++ = 0;  // Y: 2x2, table 0
  *p++ = 2; *p++ = 0x11; *p++ = 1;  // Cb
  *p++ = 3; *p++ = 0x11; *p++ = 1;  // Cr
  for (int w = 0; w < 2; ++w) {
    *p++ = 0xFF; *p++ = 0xC4; put16(2 + 1 + 16 + 12); *p++ = (uint8_t)w;
    memcpy(p, kDcBits[w], 16); p += 16;
    memcpy(p, kDcVals, 12); p += 12;
    *p++ = 0xFF; *p++ = 0xC4; put16(2 + 1 + 16 + 162); *p++ = (uint8_t)(0x10 | w);
    memcpy(p, kAcBits[w], 16); p += 16;
    memcpy(p, kAcVals[w], 162); p += 162;
  }
  *p++ = 0xFF; *p++ = 0xDD; put16(4); put16(ri);
  *p++ = 0xFF; *p++ = 0xDA; put16(12); *p++ = 3;
  *p++ = 1; *p++ = 0x00; *p++ = 2; *p++ = 0x11; *p++ = 3; *p++ = 0x11;
  *p++ = 0; *p++ = 63; *p++ = 0;
  return (int)(p - out);
}

// ---- per-MCU transform ----------------------------------------------------------------------------------

// JFIF colour conversion of one BGR pixel to 8-bit Y, Cb, Cr in libjpeg's fixed-point arithmetic
// (jccolor.c rgb_ycc_convert: FIX(x) = x * 2^16 rounded), so that the encoder starts from the very samples
// OpenCV's encoder starts from; the level shift (-128) happens where the blocks are formed.
BH8J_HD void ycc(uint8_t b, uint8_t g, uint8_t r, float* y, float* cb, float* cr) {
  const int ib = b, ig = g, ir = r;
  *y = (float)((19595 * ir + 38470 * ig + 7471 * ib + 32768) >> 16);
  *cb = (float)((-11059 * ir - 21709 * ig + 32768 * ib + (128 << 16) + 32767) >> 16);
  *cr = (float)((32768 * ir - 27439 * ig - 5329 * ib + (128 << 16) + 32767) >> 16);
}

// libjpeg's h2v2_downsample (jcsample.c): the 2x2 box average with the bias alternating 1, 2 along a row,
// level-shifted for the DCT.  The inputs are whole numbers 0..255.
BH8J_HD float chroma_2x2(float a, float b, float c, float d, int out_col) {
  const int sum = (int)a + (int)b + (int)c + (int)d + ((out_col & 1) ? 2 : 1);
  return (float)(sum >> 2) - 128.0f;
}

// Quantisation: coefficient / quantiser rounded half AWAY from zero, as libjpeg's forward_DCT does -- with a
// hair of bias so that a coefficient the float DCT leaves an ulp below a tie (flat areas produce exact ties)
// rounds as the exact value would.
BH8J_HD int quantise(float coef, float qrecip) {
  const float v = fmaf(coef, qrecip, coef < 0.0f ? -0.5001f : 0.5001f);
  return (int)v;  // truncation toward zero
}

// One row or column of the 8-point DCT: out[u] = sum_x in[x * stride] D[u][x] with the orthonormal DCT-II matrix
// D[u][x] = 0.5 C(u) cos((2x+1) u pi / 16), C(0) = 1/sqrt(2), as float literals (instruction immediates on the
// device; the same bits on the host, so both sides round alike).
BH8J_HD void dct8(const float* in, int stride, float* out) {
  constexpr float D[8][8] = {
    {0.353553385f, 0.353553385f, 0.353553385f, 0.353553385f, 0.353553385f, 0.353553385f, 0.353553385f, 0.353553385f},
    {0.490392625f, 0.415734798f, 0.277785122f, 0.0975451618f, -0.0975451618f, -0.277785122f, -0.415734798f, -0.490392625f},
    {0.461939752f, 0.191341713f, -0.191341713f, -0.461939752f, -0.461939752f, -0.191341713f, 0.191341713f, 0.461939752f},
    {0.415734798f, -0.0975451618f, -0.490392625f, -0.277785122f, 0.277785122f, 0.490392625f, 0.0975451618f, -0.415734798f},
    {0.353553385f, -0.353553385f, -0.353553385f, 0.353553385f, 0.353553385f, -0.353553385f, -0.353553385f, 0.353553385f},
    {0.277785122f, -0.490392625f, 0.0975451618f, 0.415734798f, -0.415734798f, -0.0975451618f, 0.490392625f, -0.277785122f},
    {0.191341713f, -0.461939752f, 0.461939752f, -0.191341713f, -0.191341713f, 0.461939752f, -0.461939752f, 0.191341713f},
    {0.0975451618f, -0.277785122f, 0.415734798f, -0.490392625f, 0.490392625f, -0.415734798f, 0.277785122f, -0.0975451618f}};
  float v[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) v[x] = in[x * stride];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float s = 0.0f;
#pragma unroll
    for (int x = 0; x < 8; ++x) s = fmaf(v[x], D[u][x], s);
    out[u] = s;
  }
}

// ---- entropy coding ------------------------------------------------------------------------------------------

// Bits into bytes with JPEG's 0xFF 0x00 stuffing.  `room` is checked by the caller per block.
struct BitWriter {
  uint8_t* out;
  uint32_t pos;
  uint64_t acc;  // `n` valid bits, right-aligned
  int n;
  BH8J_HD void put(uint32_t code, int len) {
    acc = (acc << len) | (code & ((1u << len) - 1u));
    n += len;
    while (n >= 8) {
      const uint8_t byte = (uint8_t)(acc >> (n - 8));
      out[pos++] = byte;
      if (byte == 0xFF) out[pos++] = 0x00;
      n -= 8;
    }
  }
  BH8J_HD void flush() {  // pad the last byte with 1-bits (T.81 F.1.2.3)
    if (n > 0) put(0x7Fu, 8 - n);
  }
};

BH8J_HD int magnitude_category(int v) {  // number of bits of |v|
  const unsigned a = (unsigned)(v < 0 ? -v : v);
#if defined(__CUDA_ARCH__)
  return 32 - __clz(a);
#else
  int c = 0;
  for (unsigned x = a; x; x >>= 1) ++c;
  return c;
#endif
}

// One 8x8 block of quantised coefficients in zigzag order (T.81 F.1.2); returns the new DC predictor.
BH8J_HD int encode_block(const Tables& t, int which, const int16_t* zz, int pred, BitWriter& w) {
  const int dc = zz[0], diff = dc - pred;
  int cat = magnitude_category(diff);
  const uint32_t dcode = t.dc_code[which][cat];
  w.put(dcode & 0xFFFFu, (int)(dcode >> 16));
  if (cat) w.put((uint32_t)(diff < 0 ? diff - 1 : diff), cat);
  int run = 0;
  for (int k = 1; k < 64; ++k) {
    const int v = zz[k];
    if (v == 0) {
      ++run;
      continue;
    }
    while (run > 15) {
      const uint32_t zrl = t.ac_code[which][0xF0];
      w.put(zrl & 0xFFFFu, (int)(zrl >> 16));
      run -= 16;
    }
    cat = magnitude_category(v);
    const uint32_t acode = t.ac_code[which][(run << 4) | cat];
    w.put(acode & 0xFFFFu, (int)(acode >> 16));
    w.put((uint32_t)(v < 0 ? v - 1 : v), cat);
    run = 0;
  }
  if (run > 0) {
    const uint32_t eob = t.ac_code[which][0x00];
    w.put(eob & 0xFFFFu, (int)(eob >> 16));
  }
  return dc;
}

// One restart interval: MCUs first .. first + count - 1, coefficients at coef + mcu * 384 (Y0 Y1 Y2 Y3 Cb Cr).
// Returns the bytes written into `slot` (0xFF stuffing included, padded to a byte), or -1 if the slot is full.
BH8J_HD int encode_interval(const Tables& t, const int16_t* coef, int first, int count, uint8_t* slot, int slot_bytes) {
  BitWriter w{slot, 0u, 0ull, 0};
  int pred[3] = {0, 0, 0};
  for (int m = first; m < first + count; ++m) {
    for (int b = 0; b < 6; ++b) {
      if ((int)w.pos + 512 > slot_bytes) return -1;  // cannot happen for 8-bit baseline data (see kSlotBytesPerMcu)
      const int comp = b < 4 ? 0 : b - 3;
      pred[comp] = encode_block(t, comp ? 1 : 0, coef + ((size_t)m * 6 + b) * 64, pred[comp], w);
    }
  }
  w.flush();
  return (int)w.pos;
}

#if defined(__CUDACC__)
// ---- kernels ----------------------------------------------------------------------------------------------------

// One 16x16 MCU per 64-thread CTA: BGR8 pixels (edge replication beyond the frame) -> YCbCr, chroma 2x2 box
// average, level shift, row DCTs, column DCTs, quantisation, zigzag order.
__global__ void __launch_bounds__(64) jpeg_transform_kernel(const uint8_t* __restrict__ bgr, const Tables* __restrict__ tp,
                                                            Geometry g, int16_t* __restrict__ coef, uint32_t* status) {
  if (blockIdx.x == 0 && threadIdx.x == 0) status[1] = 0u;  // the frame's overflow flag (set by the entropy kernel)
  __shared__ float sy[256], scb[256], scr[256];  // full-resolution planes of the MCU
  __shared__ float blk[6][64];                   // the six blocks, then their row transforms (in place per row)
  __shared__ __align__(16) uint32_t raw[192];    // the MCU's 16 rows of 48 bytes; later its 384 coefficients (int16)
  const Tables& t = *tp;
  const int mcu = blockIdx.x, tid = threadIdx.x;
  const int x0 = (mcu % g.mcus_x) * 16, y0 = (mcu / g.mcus_x) * 16;
  const bool whole = x0 + 16 <= g.width && y0 + 16 <= g.height && (g.width & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(bgr) & 3u) == 0;
  if (whole) {  // 12 aligned words per row, 3 words per thread: coalesced
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int wi = tid + 64 * k, row = wi / 12, col = wi - row * 12;
      raw[wi] = reinterpret_cast<const uint32_t*>(bgr + ((size_t)(y0 + row) * g.width + x0) * 3)[col];
    }
    __syncthreads();
    const uint8_t* rb = reinterpret_cast<const uint8_t*>(raw);
    for (int p = tid; p < 256; p += 64) ycc(rb[3 * p], rb[3 * p + 1], rb[3 * p + 2], &sy[p], &scb[p], &scr[p]);
  } else {  // MCUs on the right / bottom edge: pixels beyond the frame repeat the last column / row
    for (int p = tid; p < 256; p += 64) {
      const int row = p >> 4, col = p & 15;
      const int gx = min(x0 + col, g.width - 1), gy = min(y0 + row, g.height - 1);
      const uint8_t* px = bgr + ((size_t)gy * g.width + gx) * 3;
      ycc(px[0], px[1], px[2], &sy[p], &scb[p], &scr[p]);
    }
  }
  __syncthreads();
  {  // blocks: Y0 Y1 / Y2 Y3 are the quadrants of the luminance plane; Cb, Cr the 2x2 averages
    const int r = tid >> 3, c = tid & 7;
#pragma unroll
    for (int b = 0; b < 4; ++b) blk[b][tid] = sy[((b >> 1) * 8 + r) * 16 + (b & 1) * 8 + c] - 128.0f;
    const int q = (2 * r) * 16 + 2 * c;
    blk[4][tid] = chroma_2x2(scb[q], scb[q + 1], scb[q + 16], scb[q + 17], c);
    blk[5][tid] = chroma_2x2(scr[q], scr[q + 1], scr[q + 16], scr[q + 17], c);
  }
  __syncthreads();
  float tmp[8];
  const int b = tid >> 3, k = tid & 7;  // threads 0..47: block b, row / column k
  if (tid < 48) dct8(&blk[b][k * 8], 1, tmp);  // row k: horizontal frequencies
  __syncthreads();
  if (tid < 48) {
#pragma unroll
    for (int u = 0; u < 8; ++u) blk[b][k * 8 + u] = tmp[u];
  }
  __syncthreads();
  int16_t* staged = reinterpret_cast<int16_t*>(raw);  // 6 x 64 coefficients in zigzag order
  if (tid < 48) {
    dct8(&blk[b][k], 8, tmp);  // column k (horizontal frequency k): vertical frequencies
    const int which = b < 4 ? 0 : 1;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int n = v * 8 + k;
      staged[b * 64 + t.zigzag[n]] = (int16_t)quantise(tmp[v], t.qrecip[which][n]);
    }
  }
  __syncthreads();
  if (tid < 48) reinterpret_cast<uint4*>(coef + (size_t)mcu * 384)[tid] = reinterpret_cast<const uint4*>(raw)[tid];
}

// Entropy coding, one WARP per restart interval (the bytes equal encode_interval()'s, which the host harness
// runs sequentially).  Per 8x8 block lane l holds the zigzag coefficients l and l + 32.  Two ballots give the
// 64-bit map of non-zero coefficients, from which every lane reads the zero run in front of its coefficients
// (the distance to the next lower set bit); it forms the complete code of each -- ZRL codes for runs over 15,
// the (run, size) Huffman code, the value bits, and the end-of-block code behind the block's last non-zero
// coefficient -- at most 63 bits.  One packed warp scan of the code lengths places them; the bits go into a
// per-warp bit buffer in shared memory with atomicOr (a code spans at most three 32-bit words).  When the
// interval is complete the buffer is padded with 1-bits to a byte and copied to the interval's slot with
// the 0xFF 0x00 stuffing: 32 bytes per step, a ballot tells each lane how many stuffed bytes precede its own.
constexpr int kEntropyWarps = 8;
constexpr int kBitWordsPerMcu = 6 * 64 * 27 / 32 + 8;  // 384 coefficients at 27 bits, in 32-bit words

__device__ __forceinline__ void put_bits(uint32_t* words, uint32_t off, uint64_t code, int len) {
  // `code` (len <= 63 bits, right-aligned) at bit offset `off` of a big-endian bit stream held in 32-bit words
  const uint32_t w = off >> 5, r = off & 31u;
  const int room = 64 - (int)r;  // bits of the 64-bit window that starts at word w, from the stream position on
  if (len <= room) {
    const uint64_t v = code << (room - len);
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    if (hi) atomicOr(&words[w], hi);
    if (lo) atomicOr(&words[w + 1], lo);
  } else {
    const int spill = len - room;  // 1 .. 30 bits go into the third word
    const uint64_t v = code >> spill;
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t tail = (uint32_t)(code & ((1ull << spill) - 1ull)) << (32 - spill);
    if (hi) atomicOr(&words[w], hi);
    if (lo) atomicOr(&words[w + 1], lo);
    if (tail) atomicOr(&words[w + 2], tail);
  }
}

__global__ void __launch_bounds__(kEntropyWarps * 32) jpeg_entropy_kernel(const Tables* __restrict__ tp, Geometry g,
                                                                         const int16_t* __restrict__ coef,
                                                                         uint8_t* __restrict__ slots, int32_t* lens,
                                                                         uint32_t* cta_bytes, uint32_t* status) {
  extern __shared__ uint32_t sh_bits[];  // kEntropyWarps x (ri * kBitWordsPerMcu) words
  __shared__ uint32_t sh_dc[2][12], sh_ac[2][256];  // the Huffman tables: looked up with data-dependent indices
  for (int k = threadIdx.x; k < 2 * 256; k += blockDim.x) (&sh_ac[0][0])[k] = (&tp->ac_code[0][0])[k];
  if (threadIdx.x < 24) (&sh_dc[0][0])[threadIdx.x] = (&tp->dc_code[0][0])[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kEntropyWarps + warp;
  if (i < g.n_intervals) {
  const int n_words = g.ri * kBitWordsPerMcu;
  uint32_t* words = sh_bits + (size_t)warp * n_words;
  for (int k = lane; k < n_words; k += 32) words[k] = 0u;
  __syncwarp();
  const int first = i * g.ri, count = min(g.ri, g.n_mcus - first);
  uint32_t total = 0;  // bits so far (warp-uniform)
  int pred0 = 0, pred1 = 0, pred2 = 0;  // DC predictors (lane 0's are the ones used)
  // All 384 coefficients of the NEXT MCU are fetched (twelve independent loads per lane) while this one is
  // coded from registers: one trip to L2 per MCU is exposed at most, instead of one per block.
  const int16_t* zz0 = coef + (size_t)first * 384;
  int16_t nx0[6], nx1[6];
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    nx0[b] = zz0[b * 64 + lane];
    nx1[b] = zz0[b * 64 + lane + 32];
  }
  for (int m = first; m < first + count; ++m) {
    int16_t cur0[6], cur1[6];
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      cur0[b] = nx0[b];
      cur1[b] = nx1[b];
    }
    if (m + 1 < first + count) {
      const int16_t* zn = coef + (size_t)(m + 1) * 384;
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        nx0[b] = zn[b * 64 + lane];
        nx1[b] = zn[b * 64 + lane + 32];
      }
    }
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const int which = b < 4 ? 0 : 1;
      const int c0 = cur0[b], c1 = cur1[b];
      const uint32_t nz_lo = __ballot_sync(0xffffffffu, c0 != 0) | 1u;  // bit 0: the DC position bounds the first run
      const uint32_t nz_hi = __ballot_sync(0xffffffffu, c1 != 0);
      const uint64_t nz = ((uint64_t)nz_hi << 32) | nz_lo;
      const int last = 63 - __clzll((long long)nz);  // position of the last non-zero coefficient (0: only DC)
      const uint32_t eob = sh_ac[which][0x00];
      uint64_t code[2] = {0ull, 0ull};
      int len[2] = {0, 0};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && nz_hi == 0u) break;  // (warp-uniform) coefficients 32..63 are all zero in most blocks
        const int k = lane + 32 * h, v = h ? c1 : c0;
        if (k == 0) {  // DC difference (T.81 F.1.2.1)
          const int pred = b < 4 ? pred0 : (b == 4 ? pred1 : pred2);
          const int diff = v - pred;
          const int cat = magnitude_category(diff);
          const uint32_t dcode = sh_dc[which][cat];
          code[h] = ((uint64_t)(dcode & 0xFFFFu) << cat) | ((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << cat) - 1u));
          len[h] = (int)(dcode >> 16) + cat;
          if (b < 4) pred0 = v; else if (b == 4) pred1 = v; else pred2 = v;
        } else if (v != 0) {  // AC coefficient: zero run, then (run, size) code and value bits (F.1.2.2)
          const uint64_t below = nz & ((1ull << k) - 1ull);
          int run = k - (63 - __clzll((long long)below)) - 1;
          const uint32_t zrl = sh_ac[which][0xF0];
          while (run > 15) {
            code[h] = (code[h] << (zrl >> 16)) | (zrl & 0xFFFFu);
            len[h] += (int)(zrl >> 16);
            run -= 16;
          }
          const int cat = magnitude_category(v);
          const uint32_t acode = sh_ac[which][(run << 4) | cat];
          code[h] = (code[h] << ((acode >> 16) + cat)) | ((uint64_t)(acode & 0xFFFFu) << cat) |
                    ((uint32_t)(v < 0 ? v - 1 : v) & ((1u << cat) - 1u));
          len[h] += (int)(acode >> 16) + cat;
        }
        if (k == last && last < 63) {  // end of block behind the last non-zero coefficient (or the DC)
          code[h] = (code[h] << (eob >> 16)) | (eob & 0xFFFFu);
          len[h] += (int)(eob >> 16);
        }
      }
      // exclusive scan of the lengths in coefficient order: first halves (k = lane), then second halves
      uint32_t packed = (uint32_t)len[0] | ((uint32_t)len[1] << 16), incl = packed;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      const uint32_t sums = __shfl_sync(0xffffffffu, incl, 31), excl = incl - packed;
      const uint32_t off0 = total + (excl & 0xFFFFu), off1 = total + (sums & 0xFFFFu) + (excl >> 16);
      if (len[0]) put_bits(words, off0, code[0], len[0]);
      if (len[1]) put_bits(words, off1, code[1], len[1]);
      total += (sums & 0xFFFFu) + (sums >> 16);
    }
  }
  __syncwarp();
  if (lane == 0 && (total & 7u)) put_bits(words, total, (1ull << (8 - (total & 7u))) - 1ull, 8 - (int)(total & 7u));
  __syncwarp();
  const uint32_t n_bytes = (total + 7u) >> 3;
  uint8_t* slot = slots + (size_t)i * g.ri * kSlotBytesPerMcu;
  uint32_t out = 0;
  for (uint32_t base = 0; base < n_bytes; base += 32) {
    const uint32_t j = base + lane;
    const bool have = j < n_bytes;
    const uint32_t byte = have ? (words[j >> 2] >> (24 - 8 * (j & 3u))) & 0xFFu : 0u;
    const uint32_t ff = __ballot_sync(0xffffffffu, have && byte == 0xFFu);
    const uint32_t pos = out + lane + __popc(ff & ((1u << lane) - 1u));
    if (have) {
      slot[pos] = (uint8_t)byte;
      if (byte == 0xFFu) slot[pos + 1] = 0;
    }
    out += min(32u, n_bytes - base) + __popc(ff);
  }
  if (lane == 0) lens[i] = (int32_t)out;
  }
  // What this CTA's intervals add to the stream (their bytes + one 2-byte marker each): the gather kernel
  // sums these per-CTA totals instead of scanning 4 000 interval lengths.
  __shared__ uint32_t sh_cta_bytes;
  if (threadIdx.x == 0) sh_cta_bytes = 0u;
  __syncthreads();
  if (lane == 0 && i < g.n_intervals) {
    const int32_t n = lens[i];
    if (n < 0) atomicOr(&status[1], 1u);  // the interval overflowed its slot: the host fails the frame
    atomicAdd(&sh_cta_bytes, (uint32_t)max(n, 0) + 2u);
  }
  __syncthreads();
  if (threadIdx.x == 0) cta_bytes[blockIdx.x] = sh_cta_bytes;
}

// One warp per interval, launched with the entropy kernel's grid: the CTA sums the byte totals of the CTAs
// before it (its place in the stream), its warps take their intervals in order, each copies its slot and
// appends the RSTm / EOI marker.  status[0] = total bytes of the JPEG (written by the last CTA).
__global__ void __launch_bounds__(kEntropyWarps * 32) jpeg_gather_kernel(Geometry g, const uint8_t* __restrict__ slots,
                                                                        const int32_t* __restrict__ lens,
                                                                        const uint32_t* __restrict__ cta_bytes,
                                                                        uint8_t* __restrict__ out, uint32_t* status,
                                                                        size_t cap) {
  __shared__ uint32_t sh_warp_sum[kEntropyWarps], sh_base, sh_off[kEntropyWarps + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t sum = 0;
  for (int k = threadIdx.x; k < (int)blockIdx.x; k += blockDim.x) sum += cta_bytes[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) sh_warp_sum[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t at = (uint32_t)g.header_bytes;
    for (int w = 0; w < kEntropyWarps; ++w) at += sh_warp_sum[w];
    sh_base = at;
    for (int w = 0; w < kEntropyWarps; ++w) {
      sh_off[w] = at;
      const int k = blockIdx.x * kEntropyWarps + w;
      if (k < g.n_intervals) at += (uint32_t)max(lens[k], 0) + 2u;
    }
    sh_off[kEntropyWarps] = at;
    if (blockIdx.x == gridDim.x - 1) status[0] = at;
  }
  __syncthreads();
  const int i = blockIdx.x * kEntropyWarps + warp;
  if (i >= g.n_intervals) return;
  const uint32_t offset = sh_off[warp];
  const int n = max(lens[i], 0);
  if ((size_t)offset + n + 2 > cap) return;  // the host sees total > cap in the status word and fails the frame
  const uint8_t* src = slots + (size_t)i * g.ri * kSlotBytesPerMcu;
  uint8_t* dst = out + offset;
  for (int k = lane; k < n; k += 32) dst[k] = src[k];
  if (lane == 0) {
    dst[n] = 0xFF;
    dst[n + 1] = (i == g.n_intervals - 1) ? 0xD9 : (uint8_t)(0xD0 + (i & 7));  // EOI / RSTm
  }
}
#endif  // __CUDACC__

}  // namespace bh8jpeg

#endif  // BH8_JPEG_CUH_
