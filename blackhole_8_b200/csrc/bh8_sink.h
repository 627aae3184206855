// bh8_sink.h -- frame sink: Motion-JPEG AVI writer fed by GPU-side JPEG encoding.
//
// Replaces cv::VideoWriter(path, fourcc('M','J','P','G'), fps, size, true) and out_capture.write(frame)
// of blackhole_solution_test.cc:71-72,334.  At thousands of rendered frames per second the raw frame
// read-back (8.3 MB per 1080p frame) and a host-side JPEG encoder are the bottleneck; here the frame
// never leaves the GPU uncompressed: nvJPEG encodes the device-resident BGR8 frame (the reference's
// CV_8UC3 layout) and only the bitstream crosses PCIe and goes into the AVI file.
//
// Included at the end of bh8_lib.cu (one translation unit: it uses bh8_ctx, Device, launch_frame).
// The container part needs no GPU and is what the CPU tests exercise (bh8_sink_append_jpeg).
#ifndef BH8_SINK_H_
#define BH8_SINK_H_

#include <nvjpeg.h>

#include "bh8_hud.h"
#include "bh8_jpeg.cuh"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

// AVI 1.0 with one MJPG video stream and an idx1 index: RIFF('AVI ' LIST('hdrl' avih LIST('strl'
// strh strf)) LIST('movi' 00dc...) idx1).  Sizes and the frame count are patched in at close.
class AviMjpgWriter {
 public:
  bool open(const char* path, int width, int height, double fps, std::string* err) {
    fp_ = std::fopen(path, "wb");
    if (!fp_) {
      *err = std::string("cannot open ") + path;
      return false;
    }
    width_ = width;
    height_ = height;
    const uint32_t rate = (uint32_t)(fps * 1000.0 + 0.5), scale = 1000;
    fourcc("RIFF"); riff_size_at_ = tell(); u32(0); fourcc("AVI ");
    fourcc("LIST"); const long hdrl_size_at = tell(); u32(0); const long hdrl_start = tell(); fourcc("hdrl");
    fourcc("avih"); u32(56);
    u32((uint32_t)(1e6 / fps + 0.5));      // dwMicroSecPerFrame
    u32(0);                                // dwMaxBytesPerSec
    u32(0);                                // dwPaddingGranularity
    u32(0x10);                             // dwFlags: AVIF_HASINDEX
    frames_at_[0] = tell(); u32(0);        // dwTotalFrames
    u32(0);                                // dwInitialFrames
    u32(1);                                // dwStreams
    u32((uint32_t)width * height * 3);     // dwSuggestedBufferSize
    u32(width); u32(height);
    u32(0); u32(0); u32(0); u32(0);
    fourcc("LIST"); const long strl_size_at = tell(); u32(0); const long strl_start = tell(); fourcc("strl");
    fourcc("strh"); u32(56);
    fourcc("vids"); fourcc("MJPG");
    u32(0);                                // dwFlags
    u32(0);                                // wPriority, wLanguage
    u32(0);                                // dwInitialFrames
    u32(scale); u32(rate);                 // dwScale, dwRate: rate / scale = fps
    u32(0);                                // dwStart
    frames_at_[1] = tell(); u32(0);        // dwLength
    u32((uint32_t)width * height * 3);     // dwSuggestedBufferSize
    u32(0xFFFFFFFFu);                      // dwQuality
    u32(0);                                // dwSampleSize
    u16(0); u16(0); u16((uint16_t)width); u16((uint16_t)height);  // rcFrame
    fourcc("strf"); u32(40);
    u32(40); u32(width); u32(height); u16(1); u16(24); fourcc("MJPG");
    u32((uint32_t)width * height * 3); u32(0); u32(0); u32(0); u32(0);
    patch(strl_size_at, (uint32_t)(tell() - strl_start));
    patch(hdrl_size_at, (uint32_t)(tell() - hdrl_start));
    fourcc("LIST"); movi_size_at_ = tell(); u32(0); movi_start_ = tell(); fourcc("movi");
    return check(err);
  }

  bool append(const uint8_t* jpeg, size_t bytes, std::string* err) {
    if (!fp_) {
      *err = "sink has no file";
      return false;
    }
    if ((uint64_t)tell() + bytes + 16ull * (index_.size() + 2) + 64 > 0x7FF00000ull) {
      *err = "AVI 1.0 size limit (2 GiB) reached: close this sink and open another file";
      return false;
    }
    const long at = tell();
    fourcc("00dc"); u32((uint32_t)bytes);
    std::fwrite(jpeg, 1, bytes, fp_);
    if (bytes & 1) std::fputc(0, fp_);
    index_.push_back({(uint32_t)(at - movi_start_), (uint32_t)bytes});
    return check(err);
  }

  bool close(uint64_t* file_bytes, std::string* err) {
    if (!fp_) return true;
    patch(movi_size_at_, (uint32_t)(tell() - movi_start_));
    fourcc("idx1"); u32((uint32_t)(16 * index_.size()));
    for (const Entry& e : index_) {
      fourcc("00dc"); u32(0x10);  // AVIIF_KEYFRAME
      u32(e.offset); u32(e.size);
    }
    const long end = tell();
    patch(riff_size_at_, (uint32_t)(end - 8));
    patch(frames_at_[0], (uint32_t)index_.size());
    patch(frames_at_[1], (uint32_t)index_.size());
    const bool ok = check(err);
    std::fclose(fp_);
    fp_ = nullptr;
    if (file_bytes) *file_bytes = (uint64_t)end;
    return ok;
  }

  bool is_open() const { return fp_ != nullptr; }
  uint64_t frames() const { return index_.size(); }
  void abandon() {  // close without finishing (the caller removes the file)
    if (fp_) std::fclose(fp_);
    fp_ = nullptr;
  }
  ~AviMjpgWriter() {
    if (fp_) std::fclose(fp_);
  }

 private:
  struct Entry {
    uint32_t offset, size;
  };
  long tell() { return std::ftell(fp_); }
  void fourcc(const char* c) { std::fwrite(c, 1, 4, fp_); }
  void u32(uint32_t v) {
    const uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
    std::fwrite(b, 1, 4, fp_);
  }
  void u16(uint16_t v) {
    const uint8_t b[2] = {(uint8_t)v, (uint8_t)(v >> 8)};
    std::fwrite(b, 1, 2, fp_);
  }
  void patch(long at, uint32_t v) {
    const long here = tell();
    std::fseek(fp_, at, SEEK_SET);
    u32(v);
    std::fseek(fp_, here, SEEK_SET);
  }
  bool check(std::string* err) {
    if (std::ferror(fp_)) {
      *err = "write error on the AVI file";
      return false;
    }
    return true;
  }
  FILE* fp_ = nullptr;
  int width_ = 0, height_ = 0;
  long riff_size_at_ = 0, movi_size_at_ = 0, movi_start_ = 0, frames_at_[2] = {0, 0};
  std::vector<Entry> index_;
};

// Reader for the files AviMjpgWriter produces (and any AVI 1.0 whose video chunks are '00dc' / '00db'
// inside LIST 'movi'): frame sizes and offsets from a walk over the chunk tree.
class AviMjpgReader {
 public:
  struct Frame {
    long offset;
    uint32_t size;
  };
  bool open(const char* path, std::string* err) {
    fp_ = std::fopen(path, "rb");
    if (!fp_) {
      *err = std::string("cannot open ") + path;
      return false;
    }
    char id[4];
    uint32_t size = 0;
    if (!tag(id) || std::memcmp(id, "RIFF", 4) != 0 || !u32(&size) || !tag(id) || std::memcmp(id, "AVI ", 4) != 0) {
      *err = std::string(path) + " is not a RIFF AVI file";
      return false;
    }
    std::fseek(fp_, 0, SEEK_END);
    const long file_end = std::ftell(fp_);
    if (!walk(12, file_end, 0, err)) return false;
    if (width_ <= 0 || height_ <= 0 || !(fps_ > 0)) {
      *err = std::string(path) + ": no avih / strh header found";
      return false;
    }
    return true;
  }
  bool read(size_t k, std::vector<uint8_t>* out, std::string* err) {
    out->resize(frames_[k].size);
    std::fseek(fp_, frames_[k].offset, SEEK_SET);
    if (std::fread(out->data(), 1, out->size(), fp_) != out->size()) {
      *err = "short read on an AVI part";
      return false;
    }
    return true;
  }
  size_t frames() const { return frames_.size(); }
  int width() const { return width_; }
  int height() const { return height_; }
  double fps() const { return fps_; }
  ~AviMjpgReader() {
    if (fp_) std::fclose(fp_);
  }

 private:
  bool tag(char* id) { return std::fread(id, 1, 4, fp_) == 4; }
  bool u32(uint32_t* v) {
    uint8_t b[4];
    if (std::fread(b, 1, 4, fp_) != 4) return false;
    *v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    return true;
  }
  bool walk(long at, long end, int depth, std::string* err) {
    if (depth > 8) {  // AVI nests LISTs two deep; a crafted file must not recurse without bound
      *err = "LIST chunks nested too deeply in an AVI part";
      return false;
    }
    while (at + 8 <= end) {
      std::fseek(fp_, at, SEEK_SET);
      char id[4];
      uint32_t size = 0;
      if (!tag(id) || !u32(&size)) break;
      const long body = at + 8;
      if (body + (long)size > end) {
        *err = "truncated chunk in an AVI part";
        return false;
      }
      if (std::memcmp(id, "LIST", 4) == 0) {
        char kind[4];
        if (!tag(kind)) break;
        if (!walk(body + 4, body + (long)size, depth + 1, err)) return false;
      } else if (std::memcmp(id, "avih", 4) == 0 && size >= 40) {
        uint32_t v[10];
        for (int i = 0; i < 10; ++i) u32(&v[i]);
        width_ = (int)v[8];
        height_ = (int)v[9];
      } else if (std::memcmp(id, "strh", 4) == 0 && size >= 28) {
        uint32_t v[7];
        for (int i = 0; i < 7; ++i) u32(&v[i]);
        if (v[5]) fps_ = (double)v[6] / (double)v[5];  // dwRate / dwScale
      } else if (id[0] == '0' && id[1] == '0' && id[2] == 'd' && (id[3] == 'c' || id[3] == 'b')) {
        frames_.push_back({body, size});
      }
      at = body + (long)size + (size & 1);
    }
    return true;
  }
  FILE* fp_ = nullptr;
  int width_ = 0, height_ = 0;
  double fps_ = 0;
  std::vector<Frame> frames_;
};

const char* nvjpeg_status_name(nvjpegStatus_t s) {
  switch (s) {
    case NVJPEG_STATUS_SUCCESS: return "success";
    case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
    case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
    case NVJPEG_STATUS_BAD_JPEG: return "bad jpeg";
    case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "jpeg not supported";
    case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
    case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
    case NVJPEG_STATUS_ARCH_MISMATCH: return "arch mismatch";
    case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
    default: return "unknown status";
  }
}

}  // namespace

namespace {

// HUD text on the device: the glyph table of bh8_hud_font.h in device memory, one CTA per line, one thread
// per character (bh8_hud_draw, bh8_hud.h, is the host form of the same blit).
struct HudDevice {
  uint16_t* glyph = nullptr;  // [95][16]
  uint8_t* advance = nullptr; // [95]
};

__global__ void __launch_bounds__(BH8_HUD_MAX_TEXT) bh8_hud_kernel(uint8_t* bgr, int rows, int cols, const bh8_hud_line* lines,
                                                                    const uint16_t* glyph, const uint8_t* advance) {
  const bh8_hud_line& ln = lines[blockIdx.x];
  const int me = threadIdx.x;
  int x = ln.x;
  for (int k = 0; k < me; ++k) {  // pen position of this character; the line ends at its NUL
    const unsigned char ch = (unsigned char)ln.text[k];
    if (!ch) return;
    x += advance[((ch < BH8_HUD_FIRST || ch > BH8_HUD_LAST) ? '?' : ch) - BH8_HUD_FIRST];
  }
  const unsigned char ch = (unsigned char)ln.text[me];
  if (!ch) return;
  const uint16_t* gl = glyph + (((ch < BH8_HUD_FIRST || ch > BH8_HUD_LAST) ? '?' : ch) - BH8_HUD_FIRST) * BH8_HUD_CELL;
  for (int gr = 0; gr < BH8_HUD_CELL; ++gr) {
    const int yy = ln.y + BH8_HUD_TOP + gr;
    const unsigned bits = gl[gr];
    if (!bits || yy < 0 || yy >= rows) continue;
    for (int gx = 0; gx < BH8_HUD_CELL; ++gx) {
      const int xx = x + gx;
      if (!((bits >> gx) & 1u) || xx < 0 || xx >= cols) continue;
      uint8_t* px = bgr + ((size_t)yy * cols + xx) * 3;
      px[0] = ln.b;
      px[1] = ln.g;
      px[2] = ln.r;
    }
  }
}

// The sink's own JPEG encoder (bh8_jpeg.cuh): per frame slot the intermediate buffers and a pinned landing
// zone for the bitstream.
struct JpegSlot {
  int16_t* d_coef = nullptr;
  uint8_t* d_slots = nullptr;
  int32_t* d_lens = nullptr;
  uint32_t* d_offsets = nullptr;  // per entropy-kernel CTA: bytes its intervals add to the stream
  uint8_t* d_out = nullptr;    // [0, 8): status (total bytes, overflow flag); the JPEG starts at byte 16
  uint8_t* h_out = nullptr;    // pinned mirror of d_out
  bh8_hud_line* d_hud = nullptr;
  size_t cap = 0;              // bytes of JPEG the output buffers hold
  size_t copied = 0;           // bytes of JPEG queued for read-back with the frame in flight
  uint64_t hud_version = 0;    // which bh8_sink_hud() text d_hud holds
};
constexpr size_t kJpegOutPrefix = 16;

}  // namespace

struct bh8_sink {
  bh8_ctx* ctx = nullptr;  // nullptr: host-only sink (container fed with ready JPEGs)
  int width = 0, height = 0, quality = 95;
  // own encoder (default) or nvJPEG ($BH8_SINK_NVJPEG=1, kept for A/B measurements)
  bool use_nvjpeg = false;
  bh8jpeg::Geometry geo{};
  bh8jpeg::Tables* d_tables = nullptr;
  std::vector<uint8_t> header;
  JpegSlot jslot[3];       // [0], [1]: bh8_sink_submit's slots; [2]: the synchronous calls
  size_t copy_hint = 0;    // how much of a frame's output buffer is read back blindly (grows with the frames seen)
  HudDevice hud_dev;
  std::vector<bh8_hud_line> hud;  // text drawn into every frame before it is encoded (bh8_sink_hud)
  uint64_t hud_version = 1;
  AviMjpgWriter avi;
  std::string err;
  std::vector<uint8_t> jpeg;  // bitstream of the last frame
  uint64_t frames = 0, jpeg_bytes = 0;
  double encode_ms = 0;  // device time spent in nvJPEG, CUDA events
  // GPU side (device 0 of the context)
  nvjpegHandle_t handle = nullptr;
  nvjpegEncoderState_t state = nullptr;
  nvjpegEncoderParams_t params = nullptr;
  void* d_frame = nullptr;  // BGR8 staging frame for bh8_sink_render
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // bh8_sink_submit: two frames in flight, each with its own stream, staging frame and encoder
  // state -- the kernel of frame k+1 runs while frame k is being encoded and its bitstream read back.
  struct Slot {
    nvjpegEncoderState_t state = nullptr;
    void* d_frame = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool busy = false;
  } slot[2];
  uint64_t submitted = 0;
};

namespace {

int sink_fail(bh8_sink* s, int code, const std::string& msg) {
  if (s) s->err = msg;
  return code;
}

#define BH8_NVJPEG(s, call)                                                                          \
  do {                                                                                               \
    const nvjpegStatus_t st__ = (call);                                                              \
    if (st__ != NVJPEG_STATUS_SUCCESS)                                                               \
      return sink_fail(s, BH8_ECUDA, std::string(#call) + ": nvJPEG " + nvjpeg_status_name(st__));   \
  } while (0)
#define BH8_SINK_CUDA(s, call)                                                                       \
  do {                                                                                               \
    const cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) return sink_fail(s, BH8_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

int jpeg_slot_init(bh8_sink* s, JpegSlot& js) {
  if (js.d_coef) return BH8_OK;
  const bh8jpeg::Geometry& g = s->geo;
  js.cap = static_cast<size_t>(g.n_mcus) * 1024 + 4096;  // raw 4:2:0 data is 384 bytes per MCU
  BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&js.d_coef), static_cast<size_t>(g.n_mcus) * 384 * sizeof(int16_t)));
  BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&js.d_slots),
                              static_cast<size_t>(g.n_intervals) * g.ri * bh8jpeg::kSlotBytesPerMcu));
  BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&js.d_lens), static_cast<size_t>(g.n_intervals) * sizeof(int32_t)));
  BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&js.d_offsets), static_cast<size_t>(g.n_intervals) * sizeof(uint32_t)));
  BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&js.d_out), kJpegOutPrefix + js.cap));
  BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&js.d_hud), BH8_HUD_MAX_LINES * sizeof(bh8_hud_line)));
  BH8_SINK_CUDA(s, cudaHostAlloc(reinterpret_cast<void**>(&js.h_out), kJpegOutPrefix + js.cap, cudaHostAllocDefault));
  BH8_SINK_CUDA(s, cudaMemset(js.d_out, 0, kJpegOutPrefix));
  BH8_SINK_CUDA(s, cudaMemcpy(js.d_out + kJpegOutPrefix, s->header.data(), s->header.size(), cudaMemcpyHostToDevice));
  return BH8_OK;
}

void jpeg_slot_free(JpegSlot& js) {
  cudaFree(js.d_coef);
  cudaFree(js.d_slots);
  cudaFree(js.d_lens);
  cudaFree(js.d_offsets);
  cudaFree(js.d_out);
  cudaFree(js.d_hud);
  if (js.h_out) cudaFreeHost(js.h_out);
  js = JpegSlot();
}

// HUD (if any) and the four encoder kernels for the BGR8 frame at d_bgr, then the read-back of the
// status words and of the first copy_hint bytes of the stream: all queued on `st`, nothing waits.
int jpeg_encode_async(bh8_sink* s, JpegSlot& js, void* d_bgr, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
  const bh8jpeg::Geometry& g = s->geo;
  if (!s->hud.empty()) {
    if (js.hud_version != s->hud_version) {  // the slot's copy of the text is stale (stream order keeps frames in flight intact)
      BH8_SINK_CUDA(s, cudaMemcpyAsync(js.d_hud, s->hud.data(), s->hud.size() * sizeof(bh8_hud_line), cudaMemcpyHostToDevice, st));
      js.hud_version = s->hud_version;
    }
    bh8_hud_kernel<<<static_cast<unsigned>(s->hud.size()), BH8_HUD_MAX_TEXT, 0, st>>>(
        static_cast<uint8_t*>(d_bgr), s->height, s->width, js.d_hud, s->hud_dev.glyph, s->hud_dev.advance);
  }
  BH8_SINK_CUDA(s, cudaEventRecord(ev0, st));
  bh8jpeg::jpeg_transform_kernel<<<g.n_mcus, 64, 0, st>>>(static_cast<const uint8_t*>(d_bgr), s->d_tables, g, js.d_coef,
                                                           reinterpret_cast<uint32_t*>(js.d_out));
  const size_t bit_smem = static_cast<size_t>(bh8jpeg::kEntropyWarps) * g.ri * bh8jpeg::kBitWordsPerMcu * sizeof(uint32_t);
  bh8jpeg::jpeg_entropy_kernel<<<(g.n_intervals + bh8jpeg::kEntropyWarps - 1) / bh8jpeg::kEntropyWarps,
                                 bh8jpeg::kEntropyWarps * 32, bit_smem, st>>>(s->d_tables, g, js.d_coef, js.d_slots, js.d_lens,
                                                                              js.d_offsets, reinterpret_cast<uint32_t*>(js.d_out));
  bh8jpeg::jpeg_gather_kernel<<<(g.n_intervals + bh8jpeg::kEntropyWarps - 1) / bh8jpeg::kEntropyWarps,
                                bh8jpeg::kEntropyWarps * 32, 0, st>>>(g, js.d_slots, js.d_lens, js.d_offsets,
                                                                      js.d_out + kJpegOutPrefix,
                                                                      reinterpret_cast<uint32_t*>(js.d_out), js.cap);
  BH8_SINK_CUDA(s, cudaGetLastError());
  BH8_SINK_CUDA(s, cudaEventRecord(ev1, st));
  s->ctx->launches += s->hud.empty() ? 3 : 4;
  js.copied = s->copy_hint < js.cap ? s->copy_hint : js.cap;
  BH8_SINK_CUDA(s, cudaMemcpyAsync(js.h_out, js.d_out, kJpegOutPrefix + js.copied, cudaMemcpyDeviceToHost, st));
  return BH8_OK;
}

// Wait for the frame queued by jpeg_encode_async and leave its bitstream in s->jpeg.
int jpeg_collect(bh8_sink* s, JpegSlot& js, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
  BH8_SINK_CUDA(s, cudaStreamSynchronize(st));
  uint32_t status[2];
  std::memcpy(status, js.h_out, sizeof status);
  const size_t total = status[0];
  if (status[1] || total > js.cap || total < s->header.size() + 2)
    return sink_fail(s, BH8_ECUDA, "JPEG encoder: a restart interval or the frame overflowed its buffer");
  if (total > js.copied) {  // the blind read-back was too short: fetch the rest, and read more next time
    BH8_SINK_CUDA(s, cudaMemcpyAsync(js.h_out + kJpegOutPrefix + js.copied, js.d_out + kJpegOutPrefix + js.copied,
                                     total - js.copied, cudaMemcpyDeviceToHost, st));
    BH8_SINK_CUDA(s, cudaStreamSynchronize(st));
  }
  const size_t want = total + total / 4 + 65536;
  if (want > s->copy_hint) s->copy_hint = want;
  s->jpeg.assign(js.h_out + kJpegOutPrefix, js.h_out + kJpegOutPrefix + total);
  float ms = 0;
  BH8_SINK_CUDA(s, cudaEventElapsedTime(&ms, ev0, ev1));
  s->encode_ms += ms;
  return BH8_OK;
}

int sink_gpu_init(bh8_sink* s) {
  Device& d = s->ctx->dev[0];
  BH8_SINK_CUDA(s, cudaSetDevice(d.ordinal));
  BH8_SINK_CUDA(s, cudaEventCreate(&s->ev0));
  BH8_SINK_CUDA(s, cudaEventCreate(&s->ev1));
  {  // HUD glyphs
    BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&s->hud_dev.glyph), sizeof bh8_hud_glyph));
    BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&s->hud_dev.advance), sizeof bh8_hud_advance));
    BH8_SINK_CUDA(s, cudaMemcpy(s->hud_dev.glyph, bh8_hud_glyph, sizeof bh8_hud_glyph, cudaMemcpyHostToDevice));
    BH8_SINK_CUDA(s, cudaMemcpy(s->hud_dev.advance, bh8_hud_advance, sizeof bh8_hud_advance, cudaMemcpyHostToDevice));
  }
  if (const char* e = std::getenv("BH8_SINK_NVJPEG")) s->use_nvjpeg = std::atoi(e) != 0;
  if (!s->use_nvjpeg) {
    int ri = 2;  // restart interval in MCUs: see bh8_jpeg.cuh
    if (const char* e = std::getenv("BH8_SINK_RESTART_INTERVAL")) ri = std::atoi(e);
    if (ri < 1 || ri > 16) return sink_fail(s, BH8_EINVAL, "BH8_SINK_RESTART_INTERVAL must be in 1..16 (MCUs)");
    s->header.resize(1024);
    s->header.resize(static_cast<size_t>(bh8jpeg::make_header(s->width, s->height, s->quality, ri, s->header.data())));
    s->geo = bh8jpeg::make_geometry(s->width, s->height, ri, static_cast<int>(s->header.size()));
    bh8jpeg::Tables t;
    bh8jpeg::make_tables(s->quality, &t);
    BH8_SINK_CUDA(s, cudaMalloc(reinterpret_cast<void**>(&s->d_tables), sizeof t));
    BH8_SINK_CUDA(s, cudaMemcpy(s->d_tables, &t, sizeof t, cudaMemcpyHostToDevice));
    s->copy_hint = 1 << 20;
    return BH8_OK;
  }
  BH8_NVJPEG(s, nvjpegCreateSimple(&s->handle));
  BH8_NVJPEG(s, nvjpegEncoderStateCreate(s->handle, &s->state, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsCreate(s->handle, &s->params, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsSetQuality(s->params, s->quality, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsSetSamplingFactors(s->params, NVJPEG_CSS_420, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsSetOptimizedHuffman(s->params, 0, d.stream));
  return BH8_OK;
}

// Encode the BGR8 frame at d_bgr (device 0 of the context, already complete on the context's
// stream or produced earlier on it) and keep the bitstream in s->jpeg.
int sink_encode(bh8_sink* s, const void* d_bgr) {
  Device& d = s->ctx->dev[0];
  BH8_SINK_CUDA(s, cudaSetDevice(d.ordinal));
  if (!s->use_nvjpeg) {
    JpegSlot& js = s->jslot[2];
    int rc = jpeg_slot_init(s, js);
    if (rc == BH8_OK) rc = jpeg_encode_async(s, js, const_cast<void*>(d_bgr), d.stream, s->ev0, s->ev1);
    return rc != BH8_OK ? rc : jpeg_collect(s, js, d.stream, s->ev0, s->ev1);
  }
  nvjpegImage_t img{};
  img.channel[0] = static_cast<unsigned char*>(const_cast<void*>(d_bgr));
  img.pitch[0] = static_cast<size_t>(s->width) * 3;
  BH8_SINK_CUDA(s, cudaEventRecord(s->ev0, d.stream));
  BH8_NVJPEG(s, nvjpegEncodeImage(s->handle, s->state, s->params, &img, NVJPEG_INPUT_BGRI, s->width, s->height,
                                  d.stream));
  BH8_SINK_CUDA(s, cudaEventRecord(s->ev1, d.stream));
  size_t length = 0;
  BH8_NVJPEG(s, nvjpegEncodeRetrieveBitstream(s->handle, s->state, nullptr, &length, d.stream));
  s->jpeg.resize(length);
  BH8_NVJPEG(s, nvjpegEncodeRetrieveBitstream(s->handle, s->state, s->jpeg.data(), &length, d.stream));
  BH8_SINK_CUDA(s, cudaStreamSynchronize(d.stream));
  s->jpeg.resize(length);
  float ms = 0;
  BH8_SINK_CUDA(s, cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  s->encode_ms += ms;
  return BH8_OK;
}

int sink_commit(bh8_sink* s) {
  if (s->avi.is_open() && !s->avi.append(s->jpeg.data(), s->jpeg.size(), &s->err)) return BH8_EINVAL;
  s->frames++;  // only frames that reached the file (or the caller, for a sink without one) are counted
  s->jpeg_bytes += s->jpeg.size();
  return BH8_OK;
}

// Everything sink_gpu_init / the first frames created on the device (bh8_sink_close, and bh8_sink_open's
// failure path).
void sink_gpu_teardown(bh8_sink* s) {
  if (!s->ctx) return;
  cudaSetDevice(s->ctx->dev[0].ordinal);
  cudaStreamSynchronize(s->ctx->dev[0].stream);
  for (bh8_sink::Slot& sl : s->slot) {
    if (sl.stream) cudaStreamSynchronize(sl.stream);
    if (sl.state) nvjpegEncoderStateDestroy(sl.state);
    if (sl.d_frame) cudaFree(sl.d_frame);
    if (sl.ev0) cudaEventDestroy(sl.ev0);
    if (sl.ev1) cudaEventDestroy(sl.ev1);
    if (sl.stream) cudaStreamDestroy(sl.stream);
    sl = bh8_sink::Slot();
  }
  for (JpegSlot& js : s->jslot) jpeg_slot_free(js);
  cudaFree(s->d_tables);
  cudaFree(s->hud_dev.glyph);
  cudaFree(s->hud_dev.advance);
  if (s->params) nvjpegEncoderParamsDestroy(s->params);
  if (s->state) nvjpegEncoderStateDestroy(s->state);
  if (s->handle) nvjpegDestroy(s->handle);
  if (s->d_frame) cudaFree(s->d_frame);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  s->d_tables = nullptr;
  s->hud_dev = HudDevice();
  s->params = nullptr;
  s->state = nullptr;
  s->handle = nullptr;
  s->d_frame = nullptr;
  s->ev0 = s->ev1 = nullptr;
}

// Fetch the bitstream of the frame in flight on this slot and append it to the file.
int sink_drain(bh8_sink* s, bh8_sink::Slot& sl) {
  if (!sl.busy) return BH8_OK;
  BH8_SINK_CUDA(s, cudaSetDevice(s->ctx->dev[0].ordinal));
  if (!s->use_nvjpeg) {
    const int rc = jpeg_collect(s, s->jslot[&sl - s->slot], sl.stream, sl.ev0, sl.ev1);
    sl.busy = false;
    return rc != BH8_OK ? rc : sink_commit(s);
  }
  sl.busy = false;
  size_t length = 0;
  BH8_NVJPEG(s, nvjpegEncodeRetrieveBitstream(s->handle, sl.state, nullptr, &length, sl.stream));
  s->jpeg.resize(length);
  BH8_NVJPEG(s, nvjpegEncodeRetrieveBitstream(s->handle, sl.state, s->jpeg.data(), &length, sl.stream));
  BH8_SINK_CUDA(s, cudaStreamSynchronize(sl.stream));
  s->jpeg.resize(length);
  float ms = 0;
  BH8_SINK_CUDA(s, cudaEventElapsedTime(&ms, sl.ev0, sl.ev1));
  s->encode_ms += ms;
  return sink_commit(s);
}

// Oldest frame first: frames reach the file in the order they were submitted.
int sink_flush(bh8_sink* s) {
  if (!s->ctx) return BH8_OK;
  for (int k = 0; k < 2; ++k) {
    const int rc = sink_drain(s, s->slot[(s->submitted + k) & 1]);
    if (rc != BH8_OK) return rc;
  }
  return BH8_OK;
}

}  // namespace

extern "C" {

int bh8_sink_open(bh8_ctx* ctx, const char* avi_path, int width, int height, double fps, int quality,
                  bh8_sink** out) {
  if (!out) return BH8_EINVAL;
  *out = nullptr;
  if (width < 1 || height < 1 || width > 65535 || height > 65535 || !(fps > 0) || quality < 1 || quality > 100)
    return fail(ctx, BH8_EINVAL, "bh8_sink_open: width/height 1..65535, fps > 0, quality 1..100");
  bh8_sink* s = new bh8_sink;
  s->ctx = ctx;
  s->width = width;
  s->height = height;
  s->quality = quality;
  if (avi_path && !s->avi.open(avi_path, width, height, fps, &s->err)) {
    const int rc = fail(ctx, BH8_EINVAL, "bh8_sink_open: " + s->err);
    delete s;
    return rc;
  }
  if (ctx) {
    const int rc = sink_gpu_init(s);
    if (rc != BH8_OK) {
      fail(ctx, rc, "bh8_sink_open: " + s->err);
      sink_gpu_teardown(s);
      s->avi.abandon();  // no half-written file is left behind
      if (avi_path) std::remove(avi_path);
      delete s;
      return rc;
    }
  }
  *out = s;
  return BH8_OK;
}

int bh8_sink_hud(bh8_sink* s, const bh8_hud_line* lines, int n_lines) {
  if (!s || n_lines < 0 || n_lines > BH8_HUD_MAX_LINES || (n_lines > 0 && !lines)) return BH8_EINVAL;
  if (!s->ctx) return sink_fail(s, BH8_EINVAL, "this sink was opened without a context");
  if (s->use_nvjpeg && n_lines > 0) return sink_fail(s, BH8_EUNSUPPORTED, "the HUD needs the sink's own encoder");
  // frames in flight may still be copying the old text out of s->hud (pageable memory is staged at the call,
  // so they are not, but a drained pipeline keeps this obviously right)
  const int rcf = sink_flush(s);
  if (rcf != BH8_OK) return rcf;
  s->hud.assign(lines, lines + n_lines);
  for (bh8_hud_line& ln : s->hud) ln.text[BH8_HUD_MAX_TEXT - 1] = 0;
  s->hud_version++;
  return BH8_OK;
}

int bh8_hud_draw_device(bh8_ctx* ctx, void* d_bgr_frame, int width, int height, const bh8_hud_line* lines, int n_lines) {
  if (!ctx) return BH8_EINVAL;
  if (!d_bgr_frame || width < 1 || height < 1 || n_lines < 1 || n_lines > BH8_HUD_MAX_LINES || !lines)
    return fail(ctx, BH8_EINVAL, "bad arguments to bh8_hud_draw_device");
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  uint16_t* glyph = nullptr;
  uint8_t* advance = nullptr;
  bh8_hud_line* d_lines = nullptr;
  std::vector<bh8_hud_line> copy(lines, lines + n_lines);
  for (bh8_hud_line& ln : copy) ln.text[BH8_HUD_MAX_TEXT - 1] = 0;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&glyph), sizeof bh8_hud_glyph);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&advance), sizeof bh8_hud_advance);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_lines), copy.size() * sizeof(bh8_hud_line));
  if (e == cudaSuccess) e = cudaMemcpyAsync(glyph, bh8_hud_glyph, sizeof bh8_hud_glyph, cudaMemcpyHostToDevice, d.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(advance, bh8_hud_advance, sizeof bh8_hud_advance, cudaMemcpyHostToDevice, d.stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(d_lines, copy.data(), copy.size() * sizeof(bh8_hud_line), cudaMemcpyHostToDevice, d.stream);
  if (e == cudaSuccess) {
    bh8_hud_kernel<<<static_cast<unsigned>(n_lines), BH8_HUD_MAX_TEXT, 0, d.stream>>>(static_cast<uint8_t*>(d_bgr_frame), height,
                                                                                     width, d_lines, glyph, advance);
    e = cudaGetLastError();
    ctx->launches++;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
  cudaFree(glyph);
  cudaFree(advance);
  cudaFree(d_lines);
  if (e != cudaSuccess) return fail(ctx, BH8_ECUDA, std::string("bh8_hud_draw_device: ") + cudaGetErrorString(e));
  return BH8_OK;
}

int bh8_sink_write_device(bh8_sink* s, const void* d_bgr_frame) {
  if (!s || !d_bgr_frame) return BH8_EINVAL;
  if (!s->ctx) return sink_fail(s, BH8_EINVAL, "this sink was opened without a context: it only takes ready JPEGs");
  const int rcf = sink_flush(s);
  if (rcf != BH8_OK) return rcf;
  const int rc = sink_encode(s, d_bgr_frame);
  return rc != BH8_OK ? rc : sink_commit(s);
}

int bh8_sink_render(bh8_sink* s, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* params) {
  if (!s || !scene || !cam) return BH8_EINVAL;
  if (!s->ctx) return sink_fail(s, BH8_EINVAL, "this sink was opened without a context");
  if (cam->width != s->width || cam->height != s->height)
    return sink_fail(s, BH8_EINVAL, "camera size differs from the sink's frame size");
  Device& d = s->ctx->dev[0];
  const int rcf = sink_flush(s);
  if (rcf != BH8_OK) return rcf;
  if (!s->d_frame) {
    BH8_SINK_CUDA(s, cudaSetDevice(d.ordinal));
    BH8_SINK_CUDA(s, cudaMalloc(&s->d_frame, static_cast<size_t>(s->width) * s->height * 3));
  }
  bh8_params prm{};
  if (params) prm = *params;
  prm.pixel_format = BH8_PIXEL_BGR8;  // cv::Mat CV_8UC3, the layout VideoWriter::write takes
  const int rc = launch_frame(s->ctx, d, scene, cam, &prm, s->d_frame, nullptr, nullptr, nullptr);
  if (rc != BH8_OK) return sink_fail(s, rc, s->ctx->err);
  const int rc2 = sink_encode(s, s->d_frame);  // same stream: ordered after the kernel
  return rc2 != BH8_OK ? rc2 : sink_commit(s);
}

int bh8_sink_submit(bh8_sink* s, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* params) {
  if (!s || !scene || !cam) return BH8_EINVAL;
  if (!s->ctx) return sink_fail(s, BH8_EINVAL, "this sink was opened without a context");
  if (cam->width != s->width || cam->height != s->height)
    return sink_fail(s, BH8_EINVAL, "camera size differs from the sink's frame size");
  Device& d = s->ctx->dev[0];
  bh8_sink::Slot& sl = s->slot[s->submitted & 1];
  const int rcd = sink_drain(s, sl);  // the frame submitted two calls ago
  if (rcd != BH8_OK) return rcd;
  BH8_SINK_CUDA(s, cudaSetDevice(d.ordinal));
  if (!sl.stream) {
    BH8_SINK_CUDA(s, cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    BH8_SINK_CUDA(s, cudaEventCreate(&sl.ev0));
    BH8_SINK_CUDA(s, cudaEventCreate(&sl.ev1));
    if (s->use_nvjpeg) BH8_NVJPEG(s, nvjpegEncoderStateCreate(s->handle, &sl.state, sl.stream));
    BH8_SINK_CUDA(s, cudaMalloc(&sl.d_frame, static_cast<size_t>(s->width) * s->height * 3));
  }
  bh8_params prm{};
  if (params) prm = *params;
  prm.pixel_format = BH8_PIXEL_BGR8;
  const int rc = launch_frame(s->ctx, d, scene, cam, &prm, sl.d_frame, nullptr, nullptr, nullptr, sl.stream);
  if (rc != BH8_OK) return sink_fail(s, rc, s->ctx->err);
  if (!s->use_nvjpeg) {
    JpegSlot& js = s->jslot[s->submitted & 1];
    int rcj = jpeg_slot_init(s, js);
    if (rcj == BH8_OK) rcj = jpeg_encode_async(s, js, sl.d_frame, sl.stream, sl.ev0, sl.ev1);
    if (rcj != BH8_OK) return rcj;
    sl.busy = true;
    s->submitted++;
    return BH8_OK;
  }
  nvjpegImage_t img{};
  img.channel[0] = static_cast<unsigned char*>(sl.d_frame);
  img.pitch[0] = static_cast<size_t>(s->width) * 3;
  BH8_SINK_CUDA(s, cudaEventRecord(sl.ev0, sl.stream));
  BH8_NVJPEG(s, nvjpegEncodeImage(s->handle, sl.state, s->params, &img, NVJPEG_INPUT_BGRI, s->width, s->height,
                                  sl.stream));
  BH8_SINK_CUDA(s, cudaEventRecord(sl.ev1, sl.stream));
  sl.busy = true;
  s->submitted++;
  return BH8_OK;
}

int bh8_sink_flush(bh8_sink* s) {
  if (!s) return BH8_EINVAL;
  return sink_flush(s);
}

int bh8_sink_merge(const char* const* part_paths, int n_parts, const char* out_path, uint64_t* frames,
                   uint64_t* file_bytes) {
  if (frames) *frames = 0;
  if (file_bytes) *file_bytes = 0;
  if (!part_paths || n_parts < 1 || n_parts > 1024 || !out_path)
    return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: bad arguments");
  std::vector<AviMjpgReader> parts(n_parts);
  std::string err;
  uint64_t total = 0;
  for (int i = 0; i < n_parts; ++i) {
    if (!part_paths[i] || !parts[i].open(part_paths[i], &err)) return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: " + err);
    if (parts[i].width() != parts[0].width() || parts[i].height() != parts[0].height() ||
        parts[i].fps() != parts[0].fps())
      return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: the parts differ in frame size or rate");
    total += parts[i].frames();
  }
  // frame k of the job is frame k / n of part k % n (sharding.frames_of: whole frames round-robin)
  for (int i = 0; i < n_parts; ++i)
    if (parts[i].frames() != (total + n_parts - 1 - i) / n_parts)
      return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: frame counts of the parts are not a round-robin split");
  AviMjpgWriter out;
  if (!out.open(out_path, parts[0].width(), parts[0].height(), parts[0].fps(), &err))
    return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: " + err);
  std::vector<uint8_t> buf;
  const auto give_up = [&]() {  // no half-written video is left behind
    out.abandon();
    std::remove(out_path);
    if (file_bytes) *file_bytes = 0;
    return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: " + err);
  };
  for (uint64_t k = 0; k < total; ++k) {
    if (!parts[k % n_parts].read(k / n_parts, &buf, &err) || !out.append(buf.data(), buf.size(), &err)) return give_up();
  }
  if (!out.close(file_bytes, &err)) return give_up();
  if (frames) *frames = total;
  return BH8_OK;
}

int bh8_sink_append_jpeg(bh8_sink* s, const uint8_t* jpeg, size_t bytes) {
  if (!s || !jpeg || bytes < 4) return BH8_EINVAL;
  if (jpeg[0] != 0xFF || jpeg[1] != 0xD8) return sink_fail(s, BH8_EINVAL, "not a JPEG stream (no SOI marker)");
  const int rcf = sink_flush(s);
  if (rcf != BH8_OK) return rcf;
  s->jpeg.assign(jpeg, jpeg + bytes);
  return sink_commit(s);
}

int bh8_sink_last_jpeg(bh8_sink* s, const uint8_t** data, size_t* bytes) {
  if (!s || !data || !bytes) return BH8_EINVAL;
  *data = s->jpeg.data();
  *bytes = s->jpeg.size();
  return BH8_OK;
}

int bh8_sink_stats(const bh8_sink* s, uint64_t* frames, uint64_t* jpeg_bytes, double* encode_ms) {
  if (!s) return BH8_EINVAL;
  if (frames) *frames = s->frames;
  if (jpeg_bytes) *jpeg_bytes = s->jpeg_bytes;
  if (encode_ms) *encode_ms = s->encode_ms;
  return BH8_OK;
}

const char* bh8_sink_last_error(const bh8_sink* s) { return s ? s->err.c_str() : ""; }

int bh8_sink_close(bh8_sink* s, uint64_t* file_bytes) {
  if (!s) return BH8_EINVAL;
  if (file_bytes) *file_bytes = 0;
  int rc = sink_flush(s);  // frames still in flight belong to the file
  if (s->avi.is_open() && !s->avi.close(file_bytes, &s->err)) rc = BH8_EINVAL;
  sink_gpu_teardown(s);
  delete s;
  return rc;
}

}  // extern "C"

#endif  // BH8_SINK_H_
