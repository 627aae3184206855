// TEST-ONLY harness: the sink's JPEG encoder (blackhole_8_b200/csrc/bh8_jpeg.cuh) on the CPU -- the same
// __host__ __device__ functions the kernels call (colour conversion, 8-point DCT, bit writer, block and
// interval coder), driven by plain loops that mirror the kernels' thread mapping.  tests/test_jpeg_host.py
// lets OpenCV decode the result; tests/test_video_sink.py (GPU) requires the device's bitstream to be equal.
// Never part of libbh8.so.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "bh8_jpeg.cuh"

extern "C" long bh8_jpeg_host_encode(const uint8_t* bgr, int width, int height, int quality, int ri, uint8_t* out, long cap) {
  using namespace bh8jpeg;
  Tables t;
  make_tables(quality, &t);
  std::vector<uint8_t> header(1024);
  header.resize(make_header(width, height, quality, ri, header.data()));
  const Geometry g = make_geometry(width, height, ri, (int)header.size());
  std::vector<int16_t> coef((size_t)g.n_mcus * 384);
  for (int mcu = 0; mcu < g.n_mcus; ++mcu) {  // jpeg_transform_kernel
    float sy[256], scb[256], scr[256], blk[6][64];
    const int x0 = (mcu % g.mcus_x) * 16, y0 = (mcu / g.mcus_x) * 16;
    for (int p = 0; p < 256; ++p) {
      const int row = p >> 4, col = p & 15;
      const int gx = std::min(x0 + col, width - 1), gy = std::min(y0 + row, height - 1);
      const uint8_t* px = bgr + ((size_t)gy * width + gx) * 3;
      ycc(px[0], px[1], px[2], &sy[p], &scb[p], &scr[p]);
    }
    for (int tid = 0; tid < 64; ++tid) {
      const int r = tid >> 3, c = tid & 7;
      for (int b = 0; b < 4; ++b) blk[b][tid] = sy[((b >> 1) * 8 + r) * 16 + (b & 1) * 8 + c] - 128.0f;
      const int q = (2 * r) * 16 + 2 * c;
      blk[4][tid] = chroma_2x2(scb[q], scb[q + 1], scb[q + 16], scb[q + 17], c);
      blk[5][tid] = chroma_2x2(scr[q], scr[q + 1], scr[q + 16], scr[q + 17], c);
    }
    float rows[6][64];
    for (int b = 0; b < 6; ++b)
      for (int k = 0; k < 8; ++k) dct8(&blk[b][k * 8], 1, &rows[b][k * 8]);
    for (int b = 0; b < 6; ++b)
      for (int k = 0; k < 8; ++k) {
        float tmp[8];
        dct8(&rows[b][k], 8, tmp);
        const int which = b < 4 ? 0 : 1;
        int16_t* o = coef.data() + ((size_t)mcu * 6 + b) * 64;
        for (int v = 0; v < 8; ++v) {
          const int n = v * 8 + k;
          o[t.zigzag[n]] = (int16_t)quantise(tmp[v], t.qrecip[which][n]);
        }
      }
  }
  std::vector<uint8_t> stream(header);
  std::vector<uint8_t> slot((size_t)ri * kSlotBytesPerMcu);
  for (int i = 0; i < g.n_intervals; ++i) {  // jpeg_entropy_kernel + jpeg_gather_kernel
    const int first = i * ri, count = std::min(ri, g.n_mcus - first);
    const int n = encode_interval(t, coef.data(), first, count, slot.data(), (int)slot.size());
    if (n < 0) return -1;
    stream.insert(stream.end(), slot.begin(), slot.begin() + n);
    stream.push_back(0xFF);
    stream.push_back(i == g.n_intervals - 1 ? 0xD9 : (uint8_t)(0xD0 + (i & 7)));
  }
  if ((long)stream.size() > cap) return -2;
  std::memcpy(out, stream.data(), stream.size());
  return (long)stream.size();
}
