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

struct bh8_sink {
  bh8_ctx* ctx = nullptr;  // nullptr: host-only sink (container fed with ready JPEGs)
  int width = 0, height = 0, quality = 95;
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

int sink_gpu_init(bh8_sink* s) {
  Device& d = s->ctx->dev[0];
  BH8_SINK_CUDA(s, cudaSetDevice(d.ordinal));
  BH8_NVJPEG(s, nvjpegCreateSimple(&s->handle));
  BH8_NVJPEG(s, nvjpegEncoderStateCreate(s->handle, &s->state, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsCreate(s->handle, &s->params, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsSetQuality(s->params, s->quality, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsSetSamplingFactors(s->params, NVJPEG_CSS_420, d.stream));
  BH8_NVJPEG(s, nvjpegEncoderParamsSetOptimizedHuffman(s->params, 0, d.stream));
  BH8_SINK_CUDA(s, cudaEventCreate(&s->ev0));
  BH8_SINK_CUDA(s, cudaEventCreate(&s->ev1));
  return BH8_OK;
}

// Encode the BGR8 frame at d_bgr (device 0 of the context, already complete on the context's
// stream or produced earlier on it) and keep the bitstream in s->jpeg.
int sink_encode(bh8_sink* s, const void* d_bgr) {
  Device& d = s->ctx->dev[0];
  BH8_SINK_CUDA(s, cudaSetDevice(d.ordinal));
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
  s->frames++;
  s->jpeg_bytes += s->jpeg.size();
  if (s->avi.is_open() && !s->avi.append(s->jpeg.data(), s->jpeg.size(), &s->err)) return BH8_EINVAL;
  return BH8_OK;
}

// Fetch the bitstream of the frame in flight on this slot and append it to the file.
int sink_drain(bh8_sink* s, bh8_sink::Slot& sl) {
  if (!sl.busy) return BH8_OK;
  sl.busy = false;
  BH8_SINK_CUDA(s, cudaSetDevice(s->ctx->dev[0].ordinal));
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
      delete s;
      return rc;
    }
  }
  *out = s;
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
    BH8_NVJPEG(s, nvjpegEncoderStateCreate(s->handle, &sl.state, sl.stream));
    BH8_SINK_CUDA(s, cudaMalloc(&sl.d_frame, static_cast<size_t>(s->width) * s->height * 3));
  }
  bh8_params prm{};
  if (params) prm = *params;
  prm.pixel_format = BH8_PIXEL_BGR8;
  const int rc = launch_frame(s->ctx, d, scene, cam, &prm, sl.d_frame, nullptr, nullptr, nullptr, sl.stream);
  if (rc != BH8_OK) return sink_fail(s, rc, s->ctx->err);
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
  for (uint64_t k = 0; k < total; ++k) {
    if (!parts[k % n_parts].read(k / n_parts, &buf, &err) || !out.append(buf.data(), buf.size(), &err))
      return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: " + err);
  }
  if (!out.close(file_bytes, &err)) return fail(nullptr, BH8_EINVAL, "bh8_sink_merge: " + err);
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
  if (s->ctx) {
    cudaSetDevice(s->ctx->dev[0].ordinal);
    for (bh8_sink::Slot& sl : s->slot) {
      if (sl.stream) cudaStreamSynchronize(sl.stream);
      if (sl.state) nvjpegEncoderStateDestroy(sl.state);
      if (sl.d_frame) cudaFree(sl.d_frame);
      if (sl.ev0) cudaEventDestroy(sl.ev0);
      if (sl.ev1) cudaEventDestroy(sl.ev1);
      if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    if (s->params) nvjpegEncoderParamsDestroy(s->params);
    if (s->state) nvjpegEncoderStateDestroy(s->state);
    if (s->handle) nvjpegDestroy(s->handle);
    if (s->d_frame) cudaFree(s->d_frame);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
  }
  delete s;
  return rc;
}

}  // extern "C"

#endif  // BH8_SINK_H_
