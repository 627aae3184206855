// bh8_lib.cu -- the C ABI of include/bh8.h: contexts, textures, launches, copies.
//
// Host side of the boundary that replaces the inline pixel loop of
// blackhole_solution_test.cc:161-308.  No CPU rendering path exists in this library: every entry
// point that produces pixels launches bh8_render_kernel on a CUDA device or fails.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#define BH8_HOST_BUILD 1
#include "bh8_kernel.cuh"

namespace {

thread_local std::string g_create_error;

struct Device {
  int ordinal = 0;
  cudaStream_t stream = nullptr;       // kernels
  cudaStream_t copy_stream = nullptr;  // D2H of finished frames
  // bh8_submit: one stream per staging slot carries a frame's kernel AND its read-back, so the
  // kernel of frame k+1 (other slot, other stream) starts on the SMs frame k's last wave leaves idle
  // and runs under frame k's copy.
  cudaStream_t slot_stream[2] = {nullptr, nullptr};
  cudaArray_t tex_array[BH8_MAX_TEXTURES] = {};
  cudaTextureObject_t tex_obj[BH8_MAX_TEXTURES] = {};
  unsigned long long* d_stats = nullptr;  // 16 counters (bh8::kStatSlots used)
  // Ticket counters of the persistent-warp render kernel (bh8::render_warps): one pair per stream that
  // launches it (kernels on different streams of one device may overlap), each pair on its own 128-byte line.
  // The kernel leaves them zero.
  static constexpr int kSchedSlots = 16, kSchedStride = 32;
  unsigned* d_sched = nullptr;
  cudaStream_t sched_stream[kSchedSlots] = {};
  int n_sched = 0;
  int sm_count = 0;
  // double-buffered frame staging for bh8_render()
  void* d_pix[2] = {nullptr, nullptr};
  uint8_t* d_cls[2] = {nullptr, nullptr};
  int8_t* d_key[2] = {nullptr, nullptr};
  uint16_t* d_steps[2] = {nullptr, nullptr};
  size_t cap_pixels = 0;  // pixels each staging buffer can hold
  cudaEvent_t ev_kernel_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_copy_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  // bh8_render of ONE frame on one device: the frame is drawn as kMaxBands row bands, band k read back
  // while band k+1 is drawn
  static constexpr int kMaxBands = 8;
  cudaEvent_t ev_band[kMaxBands] = {};
  bool timing_open = false;
  bool slot_busy[2] = {false, false};  // bh8_submit: a frame whose read-back has not been waited for
};

}  // namespace

struct bh8_ctx {
  int n_dev = 0;
  Device dev[BH8_MAX_DEVICES];
  int tex_rows[BH8_MAX_TEXTURES] = {};
  int tex_cols[BH8_MAX_TEXTURES] = {};
  std::string err;
  uint64_t launches = 0;
  uint64_t next_ticket = 0;  // bh8_submit
  bool submit_slot_streams = true;  // $BH8_SUBMIT_ONE_STREAM=1: kernels on one stream, copies on another (A/B knob)
  int render_bands = 5;             // $BH8_RENDER_BANDS: row bands of a single-frame bh8_render (1 = whole frame, then the
                                    // copy); measured best of 1..8 at 1920x1080 (profiles/r02bh_render_bands.txt)
  bool band_two_streams = true;     // $BH8_RENDER_BANDS_ONE_STREAM=1: all bands on one stream (A/B knob)
  int resolve_wait = 0x7fffffff;    // $BH8_RESOLVE_WAIT: tuning knob for the batching window, read once at creation
  size_t extra_smem = 0;            // $BH8_EXTRA_SMEM: unused dynamic shared memory per CTA (occupancy experiments)
  std::vector<void*> owned;  // bh8_frame_alloc results (device 0)
};

namespace {

int fail(bh8_ctx* ctx, int code, const std::string& msg) {
  if (ctx)
    ctx->err = msg;
  else
    g_create_error = msg;
  return code;
}

#define BH8_CUDA(ctx, call)                                                                     \
  do {                                                                                          \
    const cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                                     \
      return fail(ctx, BH8_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
  } while (0)

struct Launch {
  bh8::Bh8Out out;
  dim3 grid;
};

// Grid for the rows this shard renders (see the kernel's tile-origin computation).
int make_grid(const Bh8Frame& f, dim3* grid) {
  const int gx = (f.width + bh8::kTileW - 1) / bh8::kTileW;
  int gy;
  if (f.shard_count > 1) {
    if (f.stripe_rows % bh8::kTileH != 0) return BH8_EINVAL;
    const int n_stripes = (f.height + f.stripe_rows - 1) / f.stripe_rows;
    const int n_local = (n_stripes - f.shard_index + f.shard_count - 1) / f.shard_count;
    gy = n_local * (f.stripe_rows / bh8::kTileH);
  } else {
    gy = (f.height + bh8::kTileH - 1) / bh8::kTileH;
  }
  *grid = dim3(gx, gy > 0 ? gy : 0, 1);
  return BH8_OK;
}

// Launch the render kernel for frame constants that are already built (by bh8_build_frame here, or
// on the device by a script and read back once at its creation).
int launch_built(bh8_ctx* ctx, Device& d, Bh8Frame& f, void* d_pixels, void* d_cls, void* d_key, void* d_steps,
                 cudaStream_t st) {
  dim3 grid;
  if (make_grid(f, &grid) != BH8_OK)
    return fail(ctx, BH8_EINVAL, "stripe_rows must be a multiple of 8 when shard_count > 1");
  if (!d_pixels) return fail(ctx, BH8_EINVAL, "null pixel buffer");
  if (grid.y == 0) return BH8_OK;
  bh8::Bh8Tex tex;
  for (int i = 0; i < BH8_MAX_TEXTURES; ++i) tex.obj[i] = d.tex_obj[i];
  bh8::Bh8Out out;
  out.pixels = static_cast<uint8_t*>(d_pixels);
  out.cls = static_cast<uint8_t*>(d_cls);
  out.key = static_cast<int8_t*>(d_key);
  out.steps = static_cast<uint16_t*>(d_steps);
  out.stats = d.d_stats;
  const bool aligned = reinterpret_cast<uintptr_t>(d_pixels) % 16 == 0;
  out.vec_ok = aligned && (f.pixel_format == BH8_PIXEL_BGR8 ? f.width % 32 == 0 : f.width % 4 == 0);
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  out.sched = nullptr;
  out.tiles_x = static_cast<int32_t>(grid.x);
  out.n_patches = static_cast<int32_t>(grid.x * grid.y * (bh8::kTileW / bh8::kPatchW) * (bh8::kTileH / bh8::kPatchH));
  if (f.tracer == BH8_TRACER_LINEAR) {
    bh8::bh8_linear_kernel<<<grid, bh8::kThreads, 0, st>>>(f, tex, out);
    BH8_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return BH8_OK;
  }
  // One instantiation per number of non-central planes with an FP32 side filter; scenes with more
  // planes than filter slots take the generic instantiation (exact test on every gated step).
  // BH8_FLAG_STATS launches take the instantiation that also counts the warp schedule.
  const int nn = f.n_nc <= bh8::kMaxFilterPlanes ? f.n_nc : -1;
  // Persistent warps: one wave of CTAs, every warp draws 8x4-pixel patches from the stream's ticket counter.
  {
    int slot = -1;
    for (int i = 0; i < d.n_sched; ++i)
      if (d.sched_stream[i] == st) slot = i;
    if (slot < 0) {
      if (d.n_sched == Device::kSchedSlots) return fail(ctx, BH8_EINVAL, "too many streams render on one device");
      slot = d.n_sched++;
      d.sched_stream[slot] = st;
    }
    out.sched = d.d_sched + slot * Device::kSchedStride;
  }
#if defined(BH8_PERSISTENT_WARPS)  // A/B variant: one wave of CTAs, patches drawn from the ticket counter
  const unsigned tiles = grid.x * grid.y, wave = static_cast<unsigned>(d.sm_count) * BH8_MIN_BLOCKS;
  grid = dim3(tiles < wave ? tiles : wave, 1, 1);
#endif
#ifndef BH8_FINE_UPV
#define BH8_FINE_UPV 3  // updates per round of votes at fine steps (A/B knob)
#endif
#define BH8_LAUNCH2(NN_, ST_)                                                              \
  do {                                                                                     \
    if (f.nstep >= bh8::kFineNstep)                                                        \
      bh8::bh8_render_kernel<NN_, ST_, BH8_FINE_UPV><<<grid, bh8::kThreads, ctx->extra_smem, st>>>(f, tex, out);    \
    else                                                                                   \
      bh8::bh8_render_kernel<NN_, ST_, 2><<<grid, bh8::kThreads, ctx->extra_smem, st>>>(f, tex, out);    \
  } while (0)
#define BH8_LAUNCH(NN_)                                                                    \
  do {                                                                                     \
    if (f.flags & BH8_FLAG_STATS)                                                          \
      BH8_LAUNCH2(NN_, true);                                                              \
    else                                                                                   \
      BH8_LAUNCH2(NN_, false);                                                             \
  } while (0)
  switch (nn) {
    case 0: BH8_LAUNCH(0); break;
    case 1: BH8_LAUNCH(1); break;
    case 2: BH8_LAUNCH(2); break;
    case 3: BH8_LAUNCH(3); break;
    case 4: BH8_LAUNCH(4); break;
    default: BH8_LAUNCH(-1); break;
  }
#undef BH8_LAUNCH
#undef BH8_LAUNCH2
  BH8_CUDA(ctx, cudaGetLastError());
  ctx->launches++;
  return BH8_OK;
}

int launch_frame(bh8_ctx* ctx, Device& d, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* prm,
                 void* d_pixels, void* d_cls, void* d_key, void* d_steps, cudaStream_t st = nullptr) {
  if (!st) st = d.stream;
  Bh8Frame f;
  char msg[192];
  const int rc = bh8_build_frame(scene, cam, prm, ctx->tex_rows, ctx->tex_cols, &f, msg);
  if (rc != BH8_OK) return fail(ctx, rc, msg);
  if (ctx->resolve_wait != 0x7fffffff) f.resolve_wait = ctx->resolve_wait;  // tuning knob
  if (prm->flags & BH8_FLAG_NO_BATCHING) f.resolve_wait = -1;  // exact tests run at once
  return launch_built(ctx, d, f, d_pixels, d_cls, d_key, d_steps, st);
}

int ensure_staging(bh8_ctx* ctx, Device& d, size_t pixels) {
  if (d.cap_pixels >= pixels) return BH8_OK;
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  BH8_CUDA(ctx, cudaDeviceSynchronize());
  for (int b = 0; b < 2; ++b) {
    cudaFree(d.d_pix[b]);
    cudaFree(d.d_cls[b]);
    cudaFree(d.d_key[b]);
    cudaFree(d.d_steps[b]);
    d.d_pix[b] = nullptr;
    d.d_cls[b] = nullptr;
    d.d_key[b] = nullptr;
    d.d_steps[b] = nullptr;
  }
  d.cap_pixels = 0;
  for (int b = 0; b < 2; ++b) {
    BH8_CUDA(ctx, cudaMalloc(&d.d_pix[b], pixels * 4));
    BH8_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d.d_cls[b]), pixels));
    BH8_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d.d_key[b]), pixels));
    BH8_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d.d_steps[b]), pixels * 2));
  }
  d.cap_pixels = pixels;
  return BH8_OK;
}

int read_stats(bh8_ctx* ctx, bh8_stats* s) {
  for (int i = 0; i < ctx->n_dev; ++i) {
    Device& d = ctx->dev[i];
    unsigned long long h[16];
    BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
    // launches that add to the counters may sit on any of the device's streams
    BH8_CUDA(ctx, cudaStreamSynchronize(d.copy_stream));
    for (int b = 0; b < 2; ++b) BH8_CUDA(ctx, cudaStreamSynchronize(d.slot_stream[b]));
    BH8_CUDA(ctx, cudaMemcpyAsync(h, d.d_stats, sizeof h, cudaMemcpyDeviceToHost, d.stream));
    BH8_CUDA(ctx, cudaMemsetAsync(d.d_stats, 0, sizeof h, d.stream));
    BH8_CUDA(ctx, cudaStreamSynchronize(d.stream));
    s->rays += h[0];
    s->steps += h[1];
    for (int c = 0; c < 4; ++c) s->class_count[c] += h[2 + c];
    s->tex_oob += h[6];
    s->warps += h[7];
    s->update_slots += h[8];
    s->resolve_passes += h[9];
    s->exact_tests += h[10];
  }
  return BH8_OK;
}

}  // namespace

extern "C" {

int bh8_abi_version(void) { return BH8_ABI_VERSION; }

size_t bh8_pixel_bytes(int pixel_format) { return pixel_format == BH8_PIXEL_BGR8 ? 3 : 4; }

size_t bh8_launch_param_bytes(void) { return sizeof(Bh8Frame) + sizeof(bh8::Bh8Tex) + sizeof(bh8::Bh8Out); }

const char* bh8_last_error(const bh8_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

uint64_t bh8_launch_count(const bh8_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bh8_create(bh8_ctx** out, const int* devices, int n_dev) {
  if (!out || n_dev < 1 || n_dev > BH8_MAX_DEVICES) return fail(nullptr, BH8_EINVAL, "bad arguments to bh8_create");
  *out = nullptr;
  int count = 0;
  const cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1)
    return fail(nullptr, BH8_ENODEVICE,
                std::string("no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
  bh8_ctx* ctx = new (std::nothrow) bh8_ctx();
  if (!ctx) return fail(nullptr, BH8_ENOMEM, "out of host memory");
  ctx->n_dev = n_dev;
  if (const char* one = std::getenv("BH8_SUBMIT_ONE_STREAM")) ctx->submit_slot_streams = std::atoi(one) == 0;
  if (const char* w = std::getenv("BH8_RESOLVE_WAIT")) ctx->resolve_wait = std::atoi(w);
  if (const char* o = std::getenv("BH8_RENDER_BANDS_ONE_STREAM")) ctx->band_two_streams = std::atoi(o) == 0;
  if (const char* nb = std::getenv("BH8_RENDER_BANDS")) {
    const int v = std::atoi(nb);
    ctx->render_bands = v < 1 ? 1 : (v > Device::kMaxBands ? Device::kMaxBands : v);
  }
  if (const char* x = std::getenv("BH8_EXTRA_SMEM")) ctx->extra_smem = static_cast<size_t>(std::atoi(x));
  for (int i = 0; i < n_dev; ++i) {
    Device& d = ctx->dev[i];
    d.ordinal = devices ? devices[i] : i;
    if (d.ordinal < 0 || d.ordinal >= count) {
      delete ctx;
      return fail(nullptr, BH8_EINVAL, "device ordinal out of range");
    }
  }
#define BH8_CREATE_CUDA(call)                                                                  \
  do {                                                                                         \
    const cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                                  \
      const std::string m__ = std::string(#call) + ": " + cudaGetErrorString(e__);             \
      bh8_destroy(ctx);                                                                        \
      return fail(nullptr, BH8_ECUDA, m__);                                                    \
    }                                                                                          \
  } while (0)
  for (int i = 0; i < n_dev; ++i) {
    Device& d = ctx->dev[i];
    BH8_CREATE_CUDA(cudaSetDevice(d.ordinal));
    cudaDeviceProp prop;
    BH8_CREATE_CUDA(cudaGetDeviceProperties(&prop, d.ordinal));
    if (prop.major < 10) {
      bh8_destroy(ctx);
      return fail(nullptr, BH8_ENODEVICE, "device is not sm_100 (Blackwell B200); the kernels are sm_100a only");
    }
    BH8_CREATE_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    BH8_CREATE_CUDA(cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) BH8_CREATE_CUDA(cudaStreamCreateWithFlags(&d.slot_stream[b], cudaStreamNonBlocking));
    BH8_CREATE_CUDA(cudaMalloc(reinterpret_cast<void**>(&d.d_stats), 16 * sizeof(unsigned long long)));
    BH8_CREATE_CUDA(cudaMemset(d.d_stats, 0, 16 * sizeof(unsigned long long)));
    d.sm_count = prop.multiProcessorCount;
    BH8_CREATE_CUDA(cudaMalloc(reinterpret_cast<void**>(&d.d_sched),
                               Device::kSchedSlots * Device::kSchedStride * sizeof(unsigned)));
    BH8_CREATE_CUDA(cudaMemset(d.d_sched, 0, Device::kSchedSlots * Device::kSchedStride * sizeof(unsigned)));
    for (int b = 0; b < 2; ++b) {
      BH8_CREATE_CUDA(cudaEventCreateWithFlags(&d.ev_kernel_done[b], cudaEventDisableTiming));
      BH8_CREATE_CUDA(cudaEventCreateWithFlags(&d.ev_copy_done[b], cudaEventDisableTiming));
    }
    for (int b = 0; b < Device::kMaxBands; ++b)
      BH8_CREATE_CUDA(cudaEventCreateWithFlags(&d.ev_band[b], cudaEventDisableTiming));
    BH8_CREATE_CUDA(cudaEventCreate(&d.ev_t0));
    BH8_CREATE_CUDA(cudaEventCreate(&d.ev_t1));
    // Frames are gathered on device 0: the other devices store into its memory over NVLink.  An ordinal
    // listed twice is a second logical device on the same GPU (own streams, buffers and textures; the
    // gather is then an ordinary store): the sharding logic can be exercised on a one-GPU box.
    if (i > 0 && d.ordinal != ctx->dev[0].ordinal) {
      int can = 0;
      BH8_CREATE_CUDA(cudaDeviceCanAccessPeer(&can, d.ordinal, ctx->dev[0].ordinal));
      if (!can) {
        bh8_destroy(ctx);
        return fail(nullptr, BH8_ECUDA, "no peer access to device 0 of the context");
      }
      const cudaError_t pe = cudaDeviceEnablePeerAccess(ctx->dev[0].ordinal, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) BH8_CREATE_CUDA(pe);
      (void)cudaGetLastError();
    }
  }
#undef BH8_CREATE_CUDA
  *out = ctx;
  return BH8_OK;
}

void bh8_destroy(bh8_ctx* ctx) {
  if (!ctx) return;
  for (int i = 0; i < ctx->n_dev; ++i) {
    Device& d = ctx->dev[i];
    if (cudaSetDevice(d.ordinal) != cudaSuccess) continue;
    cudaDeviceSynchronize();
    for (int t = 0; t < BH8_MAX_TEXTURES; ++t) {
      if (d.tex_obj[t]) cudaDestroyTextureObject(d.tex_obj[t]);
      if (d.tex_array[t]) cudaFreeArray(d.tex_array[t]);
    }
    for (int b = 0; b < 2; ++b) {
      cudaFree(d.d_pix[b]);
      cudaFree(d.d_cls[b]);
      cudaFree(d.d_key[b]);
      cudaFree(d.d_steps[b]);
      if (d.ev_kernel_done[b]) cudaEventDestroy(d.ev_kernel_done[b]);
      if (d.ev_copy_done[b]) cudaEventDestroy(d.ev_copy_done[b]);
    }
    for (int b = 0; b < Device::kMaxBands; ++b)
      if (d.ev_band[b]) cudaEventDestroy(d.ev_band[b]);
    if (d.ev_t0) cudaEventDestroy(d.ev_t0);
    if (d.ev_t1) cudaEventDestroy(d.ev_t1);
    cudaFree(d.d_stats);
    cudaFree(d.d_sched);
    if (i == 0)
      for (void* p : ctx->owned) cudaFree(p);
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    for (int b = 0; b < 2; ++b)
      if (d.slot_stream[b]) cudaStreamDestroy(d.slot_stream[b]);
  }
  delete ctx;
}

int bh8_set_texture(bh8_ctx* ctx, int tex_id, const uint8_t* bgr, int rows, int cols, size_t row_stride_bytes) {
  if (!ctx) return BH8_EINVAL;
  if (tex_id < 0 || tex_id >= BH8_MAX_TEXTURES || !bgr || rows < 1 || cols < 1 || rows > 32768 || cols > 32768)
    return fail(ctx, BH8_EINVAL, "bad texture arguments");
  if (row_stride_bytes == 0) row_stride_bytes = static_cast<size_t>(cols) * 3;
  if (row_stride_bytes < static_cast<size_t>(cols) * 3) return fail(ctx, BH8_EINVAL, "row stride smaller than a row");
  // cv::Mat CV_8UC3 (B,G,R) -> uchar4 texels (B,G,R,255)
  std::vector<uchar4> texels(static_cast<size_t>(rows) * cols);
  for (int r = 0; r < rows; ++r) {
    const uint8_t* src = bgr + static_cast<size_t>(r) * row_stride_bytes;
    uchar4* dst = texels.data() + static_cast<size_t>(r) * cols;
    for (int c = 0; c < cols; ++c) dst[c] = make_uchar4(src[3 * c], src[3 * c + 1], src[3 * c + 2], 255);
  }
  for (int i = 0; i < ctx->n_dev; ++i) {
    Device& d = ctx->dev[i];
    BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
    BH8_CUDA(ctx, cudaDeviceSynchronize());
    if (d.tex_obj[tex_id]) {
      cudaDestroyTextureObject(d.tex_obj[tex_id]);
      d.tex_obj[tex_id] = 0;
    }
    if (d.tex_array[tex_id]) {
      cudaFreeArray(d.tex_array[tex_id]);
      d.tex_array[tex_id] = nullptr;
    }
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc<uchar4>();
    BH8_CUDA(ctx, cudaMallocArray(&d.tex_array[tex_id], &desc, cols, rows));
    BH8_CUDA(ctx, cudaMemcpy2DToArray(d.tex_array[tex_id], 0, 0, texels.data(), static_cast<size_t>(cols) * 4,
                                      static_cast<size_t>(cols) * 4, rows, cudaMemcpyHostToDevice));
    cudaResourceDesc res;
    std::memset(&res, 0, sizeof res);
    res.resType = cudaResourceTypeArray;
    res.res.array.array = d.tex_array[tex_id];
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof td);
    td.addressMode[0] = cudaAddressModeClamp;
    td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;       // nearest texel, like the reference's truncation
    td.readMode = cudaReadModeElementType;     // raw bytes
    td.normalizedCoords = 0;
    BH8_CUDA(ctx, cudaCreateTextureObject(&d.tex_obj[tex_id], &res, &td, nullptr));
  }
  ctx->tex_rows[tex_id] = rows;
  ctx->tex_cols[tex_id] = cols;
  return BH8_OK;
}

int bh8_render_device(bh8_ctx* ctx, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* params,
                      void* d_pixels, void* d_class, void* d_key, void* d_steps) {
  if (!ctx) return BH8_EINVAL;
  return launch_frame(ctx, ctx->dev[0], scene, cam, params, d_pixels, d_class, d_key, d_steps);
}

int bh8_sync(bh8_ctx* ctx) {
  if (!ctx) return BH8_EINVAL;
  for (int i = 0; i < ctx->n_dev; ++i) {
    BH8_CUDA(ctx, cudaSetDevice(ctx->dev[i].ordinal));
    BH8_CUDA(ctx, cudaStreamSynchronize(ctx->dev[i].stream));
    BH8_CUDA(ctx, cudaStreamSynchronize(ctx->dev[i].copy_stream));
    for (int b = 0; b < 2; ++b) BH8_CUDA(ctx, cudaStreamSynchronize(ctx->dev[i].slot_stream[b]));
  }
  return BH8_OK;
}

int bh8_timer_begin(bh8_ctx* ctx) {
  if (!ctx) return BH8_EINVAL;
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  BH8_CUDA(ctx, cudaEventRecord(d.ev_t0, d.stream));
  d.timing_open = true;
  return BH8_OK;
}

int bh8_timer_end_ms(bh8_ctx* ctx, double* ms) {
  if (!ctx || !ms) return BH8_EINVAL;
  Device& d = ctx->dev[0];
  if (!d.timing_open) return fail(ctx, BH8_EINVAL, "bh8_timer_end_ms without bh8_timer_begin");
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  BH8_CUDA(ctx, cudaEventRecord(d.ev_t1, d.stream));
  BH8_CUDA(ctx, cudaEventSynchronize(d.ev_t1));
  float f = 0;
  BH8_CUDA(ctx, cudaEventElapsedTime(&f, d.ev_t0, d.ev_t1));
  *ms = f;
  d.timing_open = false;
  return BH8_OK;
}

int bh8_read_stats(bh8_ctx* ctx, bh8_stats* stats) {
  if (!ctx || !stats) return BH8_EINVAL;
  std::memset(stats, 0, sizeof *stats);
  return read_stats(ctx, stats);
}

namespace {
// After a failure in the middle of a batch: nothing may still be writing into the caller's buffers
// when the call returns.
void drain_all(bh8_ctx* ctx) {
  for (int i = 0; i < ctx->n_dev; ++i) {
    Device& d = ctx->dev[i];
    if (cudaSetDevice(d.ordinal) != cudaSuccess) continue;
    cudaStreamSynchronize(d.stream);
    cudaStreamSynchronize(d.copy_stream);
    for (int b = 0; b < 2; ++b) cudaStreamSynchronize(d.slot_stream[b]);
  }
  (void)cudaGetLastError();
}
int render_batch(bh8_ctx* ctx, const bh8_scene* scenes, const bh8_camera* cams, int n_frames, const bh8_params* params,
                 uint8_t* out_pixels, uint8_t* out_class, int8_t* out_key, uint16_t* out_steps, bh8_stats* stats);
}  // namespace

int bh8_render(bh8_ctx* ctx, const bh8_scene* scenes, const bh8_camera* cams, int n_frames, const bh8_params* params,
               uint8_t* out_pixels, uint8_t* out_class, int8_t* out_key, uint16_t* out_steps, bh8_stats* stats) {
  if (!ctx) return BH8_EINVAL;
  // bh8_render shares the staging buffers with bh8_submit: frames still in flight there finish first
  for (int i = 0; i < ctx->n_dev; ++i)
    for (int b = 0; b < 2; ++b)
      if (ctx->dev[i].slot_busy[b]) {
        BH8_CUDA(ctx, cudaSetDevice(ctx->dev[i].ordinal));
        BH8_CUDA(ctx, cudaEventSynchronize(ctx->dev[i].ev_copy_done[b]));
        ctx->dev[i].slot_busy[b] = false;
      }
  const int rc = render_batch(ctx, scenes, cams, n_frames, params, out_pixels, out_class, out_key, out_steps, stats);
  if (rc != BH8_OK) {
    const std::string keep = ctx->err;
    drain_all(ctx);
    ctx->err = keep;
  }
  return rc;
}

} // extern "C"

namespace {
int render_batch(bh8_ctx* ctx, const bh8_scene* scenes, const bh8_camera* cams, int n_frames, const bh8_params* params,
                 uint8_t* out_pixels, uint8_t* out_class, int8_t* out_key, uint16_t* out_steps, bh8_stats* stats) {
  if (!scenes || !cams || !params || !out_pixels || n_frames < 1) return fail(ctx, BH8_EINVAL, "bad arguments to bh8_render");
  const auto wall0 = std::chrono::steady_clock::now();
  const size_t bpp = bh8_pixel_bytes(params->pixel_format);
  size_t max_pixels = 0;
  for (int fi = 0; fi < n_frames; ++fi) {
    const size_t p = static_cast<size_t>(cams[fi].width) * cams[fi].height;
    if (cams[fi].width != cams[0].width || cams[fi].height != cams[0].height)
      return fail(ctx, BH8_EINVAL, "all frames of one bh8_render call must have the same size");
    if (p > max_pixels) max_pixels = p;
  }
  const size_t npx = max_pixels;
  const bool striped = ctx->n_dev > 1 && n_frames < ctx->n_dev;
  const int n_active = striped ? ctx->n_dev : (n_frames < ctx->n_dev ? n_frames : ctx->n_dev);
  // In striped mode every device stores into device 0's staging buffer (peer memory).
  for (int i = 0; i < (striped ? 1 : n_active); ++i) {
    const int rc = ensure_staging(ctx, ctx->dev[i], npx);
    if (rc != BH8_OK) return rc;
  }
  bh8_params prm = *params;
  if (stats) prm.flags |= BH8_FLAG_STATS;

  for (int i = 0; i < n_active; ++i) {
    BH8_CUDA(ctx, cudaSetDevice(ctx->dev[i].ordinal));
    BH8_CUDA(ctx, cudaEventRecord(ctx->dev[i].ev_t0, ctx->dev[i].stream));
  }

  if (striped) {
    Device& d0 = ctx->dev[0];
    for (int fi = 0; fi < n_frames; ++fi) {
      const int b = fi & 1;
      if (fi >= 2) {  // buffer b is free once its previous D2H finished
        for (int i = 0; i < ctx->n_dev; ++i) {
          BH8_CUDA(ctx, cudaSetDevice(ctx->dev[i].ordinal));
          BH8_CUDA(ctx, cudaStreamWaitEvent(ctx->dev[i].stream, d0.ev_copy_done[b], 0));
        }
      }
      bh8_params p = prm;
      p.stripe_rows = params->stripe_rows > 0 ? params->stripe_rows : 16;
      p.shard_count = ctx->n_dev;
      for (int i = 0; i < ctx->n_dev; ++i) {
        Device& d = ctx->dev[i];
        p.shard_index = i;
        const int rc = launch_frame(ctx, d, &scenes[fi], &cams[fi], &p, d0.d_pix[b], out_class ? d0.d_cls[b] : nullptr,
                                    out_key ? d0.d_key[b] : nullptr, out_steps ? d0.d_steps[b] : nullptr);
        if (rc != BH8_OK) return rc;
        BH8_CUDA(ctx, cudaEventRecord(d.ev_kernel_done[b], d.stream));
        BH8_CUDA(ctx, cudaSetDevice(d0.ordinal));
        BH8_CUDA(ctx, cudaStreamWaitEvent(d0.copy_stream, d.ev_kernel_done[b], 0));
      }
      BH8_CUDA(ctx, cudaSetDevice(d0.ordinal));
      BH8_CUDA(ctx, cudaMemcpyAsync(out_pixels + fi * npx * bpp, d0.d_pix[b], npx * bpp, cudaMemcpyDeviceToHost,
                                    d0.copy_stream));
      if (out_class)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_class + fi * npx, d0.d_cls[b], npx, cudaMemcpyDeviceToHost, d0.copy_stream));
      if (out_key)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_key + fi * npx, d0.d_key[b], npx, cudaMemcpyDeviceToHost, d0.copy_stream));
      if (out_steps)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_steps + fi * npx, d0.d_steps[b], npx * 2, cudaMemcpyDeviceToHost,
                                      d0.copy_stream));
      BH8_CUDA(ctx, cudaEventRecord(d0.ev_copy_done[b], d0.copy_stream));
    }
  } else if (n_frames == 1 && ctx->render_bands > 1 && cams[0].height >= 64 * ctx->render_bands) {
    // ONE frame, one device: nothing of the next frame to draw under this one's read-back, so the frame
    // itself is cut into row bands (the stripe mapping with one stripe per band): all bands are launched
    // back to back, band k leaves for the host while band k+1 is drawn, and the call ends one band's copy
    // after the last kernel instead of one frame's.
    Device& d = ctx->dev[0];
    BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
    const int W = cams[0].width, H = cams[0].height, nb = ctx->render_bands;
    const int band_rows = (((H + nb - 1) / nb) + bh8::kTileH - 1) / bh8::kTileH * bh8::kTileH;
    bh8_params p = prm;
    p.stripe_rows = band_rows;
    p.shard_count = nb;
    p.shard_index = 0;
    Bh8Frame f;  // the frame constants are built once; a band differs in its shard index only
    char msg[192];
    const int rcb = bh8_build_frame(&scenes[0], &cams[0], &p, ctx->tex_rows, ctx->tex_cols, &f, msg);
    if (rcb != BH8_OK) return fail(ctx, rcb, msg);
    if (ctx->resolve_wait != 0x7fffffff) f.resolve_wait = ctx->resolve_wait;
    if (p.flags & BH8_FLAG_NO_BATCHING) f.resolve_wait = -1;
    // Bands alternate between two streams, so that band k+1 starts on the SMs band k's last wave leaves idle
    // (as bh8_submit does with whole frames); the second stream starts after ev_t0 and joins before ev_t1.
    cudaStream_t lane[2] = {d.stream, ctx->band_two_streams ? d.slot_stream[0] : d.stream};
    if (lane[1] != lane[0]) BH8_CUDA(ctx, cudaStreamWaitEvent(lane[1], d.ev_t0, 0));
    for (int k = 0; k < nb; ++k) {
      f.shard_index = k;
      const int rc = launch_built(ctx, d, f, d.d_pix[0], out_class ? d.d_cls[0] : nullptr,
                                  out_key ? d.d_key[0] : nullptr, out_steps ? d.d_steps[0] : nullptr, lane[k & 1]);
      if (rc != BH8_OK) return rc;
      BH8_CUDA(ctx, cudaEventRecord(d.ev_band[k], lane[k & 1]));
    }
    if (lane[1] != lane[0])
      for (int k = 1; k < nb; k += 2) BH8_CUDA(ctx, cudaStreamWaitEvent(d.stream, d.ev_band[k], 0));
    for (int k = 0; k < nb; ++k) {
      const int y0 = k * band_rows, y1 = y0 + band_rows < H ? y0 + band_rows : H;
      if (y0 >= H) break;
      const size_t off = static_cast<size_t>(y0) * W, n = static_cast<size_t>(y1 - y0) * W;
      BH8_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_band[k], 0));
      BH8_CUDA(ctx, cudaMemcpyAsync(out_pixels + off * bpp, static_cast<uint8_t*>(d.d_pix[0]) + off * bpp, n * bpp,
                                    cudaMemcpyDeviceToHost, d.copy_stream));
      if (out_class)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_class + off, d.d_cls[0] + off, n, cudaMemcpyDeviceToHost, d.copy_stream));
      if (out_key)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_key + off, d.d_key[0] + off, n, cudaMemcpyDeviceToHost, d.copy_stream));
      if (out_steps)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_steps + off, d.d_steps[0] + off, n * 2, cudaMemcpyDeviceToHost,
                                      d.copy_stream));
    }
    BH8_CUDA(ctx, cudaEventRecord(d.ev_copy_done[0], d.copy_stream));
  } else {
    // whole frames dealt round-robin; per device: kernel(f) -> D2H(f) overlaps kernel(f+1)
    std::vector<int> issued(ctx->n_dev, 0);
    for (int fi = 0; fi < n_frames; ++fi) {
      Device& d = ctx->dev[fi % n_active];
      const int b = issued[fi % n_active]++ & 1;
      BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
      if (issued[fi % n_active] > 2) BH8_CUDA(ctx, cudaStreamWaitEvent(d.stream, d.ev_copy_done[b], 0));
      bh8_params p = prm;
      p.shard_count = 0;
      p.shard_index = 0;
      const int rc = launch_frame(ctx, d, &scenes[fi], &cams[fi], &p, d.d_pix[b], out_class ? d.d_cls[b] : nullptr,
                                  out_key ? d.d_key[b] : nullptr, out_steps ? d.d_steps[b] : nullptr);
      if (rc != BH8_OK) return rc;
      BH8_CUDA(ctx, cudaEventRecord(d.ev_kernel_done[b], d.stream));
      BH8_CUDA(ctx, cudaStreamWaitEvent(d.copy_stream, d.ev_kernel_done[b], 0));
      BH8_CUDA(ctx, cudaMemcpyAsync(out_pixels + fi * npx * bpp, d.d_pix[b], npx * bpp, cudaMemcpyDeviceToHost,
                                    d.copy_stream));
      if (out_class)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_class + fi * npx, d.d_cls[b], npx, cudaMemcpyDeviceToHost, d.copy_stream));
      if (out_key)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_key + fi * npx, d.d_key[b], npx, cudaMemcpyDeviceToHost, d.copy_stream));
      if (out_steps)
        BH8_CUDA(ctx, cudaMemcpyAsync(out_steps + fi * npx, d.d_steps[b], npx * 2, cudaMemcpyDeviceToHost,
                                      d.copy_stream));
      BH8_CUDA(ctx, cudaEventRecord(d.ev_copy_done[b], d.copy_stream));
    }
  }

  double kernel_ms = 0;
  for (int i = 0; i < n_active; ++i) {
    Device& d = ctx->dev[i];
    BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
    BH8_CUDA(ctx, cudaEventRecord(d.ev_t1, d.stream));
  }
  for (int i = 0; i < n_active; ++i) {
    Device& d = ctx->dev[i];
    BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
    BH8_CUDA(ctx, cudaStreamSynchronize(d.stream));
    BH8_CUDA(ctx, cudaStreamSynchronize(d.copy_stream));
    float ms = 0;
    BH8_CUDA(ctx, cudaEventElapsedTime(&ms, d.ev_t0, d.ev_t1));
    if (ms > kernel_ms) kernel_ms = ms;
  }
  if (stats) {
    std::memset(stats, 0, sizeof *stats);
    const int rc = read_stats(ctx, stats);
    if (rc != BH8_OK) return rc;
    stats->kernel_ms = kernel_ms;
    stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  }
  return BH8_OK;
}
}  // namespace

extern "C" {

int bh8_submit(bh8_ctx* ctx, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* params,
               uint8_t* out_pixels, uint64_t* ticket) {
  if (!ctx) return BH8_EINVAL;
  if (!scene || !cam || !params || !out_pixels || !ticket) return fail(ctx, BH8_EINVAL, "bad arguments to bh8_submit");
  const uint64_t t = ctx->next_ticket;
  Device& d = ctx->dev[t % ctx->n_dev];
  const int b = static_cast<int>((t / ctx->n_dev) & 1);
  const size_t npx = static_cast<size_t>(cam->width) * cam->height;
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  if (d.cap_pixels < npx) {
    const int rc = ensure_staging(ctx, d, npx);
    if (rc != BH8_OK) return rc;
    d.slot_busy[0] = d.slot_busy[1] = false;
  }
  if (d.slot_busy[b]) {  // the frame submitted two rounds ago on this device still owns the buffer
    BH8_CUDA(ctx, cudaEventSynchronize(d.ev_copy_done[b]));
    d.slot_busy[b] = false;
  }
  bh8_params p = *params;
  p.shard_count = 0;
  p.shard_index = 0;
  cudaStream_t st = ctx->submit_slot_streams ? d.slot_stream[b] : d.stream;
  const int rc = launch_frame(ctx, d, scene, cam, &p, d.d_pix[b], nullptr, nullptr, nullptr, st);
  if (rc != BH8_OK) return rc;
  if (!ctx->submit_slot_streams) {
    BH8_CUDA(ctx, cudaEventRecord(d.ev_kernel_done[b], st));
    st = d.copy_stream;
    BH8_CUDA(ctx, cudaStreamWaitEvent(st, d.ev_kernel_done[b], 0));
  }
  BH8_CUDA(ctx, cudaMemcpyAsync(out_pixels, d.d_pix[b], npx * bh8_pixel_bytes(p.pixel_format), cudaMemcpyDeviceToHost, st));
  BH8_CUDA(ctx, cudaEventRecord(d.ev_copy_done[b], st));
  d.slot_busy[b] = true;
  *ticket = t;
  ctx->next_ticket = t + 1;
  return BH8_OK;
}

int bh8_wait(bh8_ctx* ctx, uint64_t ticket) {
  if (!ctx) return BH8_EINVAL;
  if (ticket >= ctx->next_ticket) return fail(ctx, BH8_EINVAL, "bh8_wait: unknown ticket");
  if (ticket + 2ull * ctx->n_dev < ctx->next_ticket) return BH8_OK;  // its buffer was already recycled => done
  Device& d = ctx->dev[ticket % ctx->n_dev];
  const int b = static_cast<int>((ticket / ctx->n_dev) & 1);
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  BH8_CUDA(ctx, cudaEventSynchronize(d.ev_copy_done[b]));
  d.slot_busy[b] = false;
  return BH8_OK;
}

int bh8_frame_alloc(bh8_ctx* ctx, size_t bytes, void** d_ptr) {
  if (!ctx || !d_ptr || bytes == 0) return BH8_EINVAL;
  BH8_CUDA(ctx, cudaSetDevice(ctx->dev[0].ordinal));
  void* p = nullptr;
  BH8_CUDA(ctx, cudaMalloc(&p, bytes));
  ctx->owned.push_back(p);
  *d_ptr = p;
  return BH8_OK;
}

int bh8_frame_free(bh8_ctx* ctx, void* d_ptr) {
  if (!ctx) return BH8_EINVAL;
  for (size_t i = 0; i < ctx->owned.size(); ++i)
    if (ctx->owned[i] == d_ptr) {
      BH8_CUDA(ctx, cudaSetDevice(ctx->dev[0].ordinal));
      BH8_CUDA(ctx, cudaDeviceSynchronize());
      BH8_CUDA(ctx, cudaFree(d_ptr));
      ctx->owned.erase(ctx->owned.begin() + i);
      return BH8_OK;
    }
  return fail(ctx, BH8_EINVAL, "pointer was not allocated by bh8_frame_alloc");
}

int bh8_ipc_export(bh8_ctx* ctx, void* d_ptr, uint8_t handle[64]) {
  if (!ctx || !d_ptr || !handle) return BH8_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  cudaIpcMemHandle_t h;
  BH8_CUDA(ctx, cudaSetDevice(ctx->dev[0].ordinal));
  BH8_CUDA(ctx, cudaIpcGetMemHandle(&h, d_ptr));
  std::memcpy(handle, &h, 64);
  return BH8_OK;
}

int bh8_ipc_import(bh8_ctx* ctx, const uint8_t handle[64], void** d_ptr) {
  if (!ctx || !d_ptr || !handle) return BH8_EINVAL;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  BH8_CUDA(ctx, cudaSetDevice(ctx->dev[0].ordinal));
  BH8_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return BH8_OK;
}

int bh8_ipc_close(bh8_ctx* ctx, void* d_ptr) {
  if (!ctx || !d_ptr) return BH8_EINVAL;
  BH8_CUDA(ctx, cudaSetDevice(ctx->dev[0].ordinal));
  BH8_CUDA(ctx, cudaStreamSynchronize(ctx->dev[0].stream));
  BH8_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
  return BH8_OK;
}

int bh8_memcpy_d2h(bh8_ctx* ctx, void* host, const void* d_ptr, size_t bytes) {
  if (!ctx || !host || !d_ptr) return BH8_EINVAL;
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  BH8_CUDA(ctx, cudaMemcpyAsync(host, d_ptr, bytes, cudaMemcpyDeviceToHost, d.stream));
  BH8_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return BH8_OK;
}

int bh8_memset_d(bh8_ctx* ctx, void* d_ptr, int value, size_t bytes) {
  if (!ctx || !d_ptr) return BH8_EINVAL;
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  BH8_CUDA(ctx, cudaMemsetAsync(d_ptr, value, bytes, d.stream));
  return BH8_OK;
}

int bh8_host_alloc(void** p, size_t bytes) {
  // portable: pinned for every CUDA context of the process (multi-device contexts read back into one buffer)
  return cudaHostAlloc(p, bytes, cudaHostAllocPortable) == cudaSuccess ? BH8_OK : BH8_ENOMEM;
}

int bh8_host_alloc_flags(void** p, size_t bytes, unsigned flags) {
  unsigned f = cudaHostAllocPortable;
  if (flags & BH8_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
  return cudaHostAlloc(p, bytes, f) == cudaSuccess ? BH8_OK : BH8_ENOMEM;
}

int bh8_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return BH8_EINVAL;
  const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e == cudaSuccess || e == cudaErrorHostMemoryAlreadyRegistered) {
    (void)cudaGetLastError();
    return BH8_OK;
  }
  (void)cudaGetLastError();
  return e == cudaErrorMemoryAllocation ? BH8_ENOMEM : BH8_ECUDA;
}

int bh8_host_unregister(void* p) {
  if (!p) return BH8_EINVAL;
  const cudaError_t e = cudaHostUnregister(p);
  (void)cudaGetLastError();
  return e == cudaSuccess ? BH8_OK : BH8_ECUDA;
}

int bh8_measure_d2h(bh8_ctx* ctx, uint8_t* host_a, uint8_t* host_b, size_t bytes, int reps, double* seconds) {
  if (!ctx || !host_a || !host_b || bytes == 0 || reps < 1 || !seconds) return BH8_EINVAL;
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  {
    const int rc = ensure_staging(ctx, d, (bytes + 3) / 4);
    if (rc != BH8_OK) return rc;
  }
  for (int b = 0; b < 2; ++b) BH8_CUDA(ctx, cudaStreamSynchronize(d.slot_stream[b]));
  // the copies bh8_submit queues, without the kernels: staging slot b -> host buffer b on slot stream b,
  // at most one copy per slot in flight
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; ++i) {
    const int b = i & 1;
    if (i >= 2) BH8_CUDA(ctx, cudaEventSynchronize(d.ev_copy_done[b]));
    BH8_CUDA(ctx, cudaMemcpyAsync(b ? host_b : host_a, d.d_pix[b], bytes, cudaMemcpyDeviceToHost, d.slot_stream[b]));
    BH8_CUDA(ctx, cudaEventRecord(d.ev_copy_done[b], d.slot_stream[b]));
  }
  for (int b = 0; b < 2; ++b) BH8_CUDA(ctx, cudaStreamSynchronize(d.slot_stream[b]));
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return BH8_OK;
}

int bh8_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? BH8_OK : BH8_ECUDA; }

int bh8_measure_stepping(bh8_ctx* ctx, int fp32, double* updates_per_s) {
  if (!ctx || !updates_per_s) return BH8_EINVAL;
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  cudaDeviceProp prop;
  BH8_CUDA(ctx, cudaGetDeviceProperties(&prop, d.ordinal));
  double* sink = nullptr;
  BH8_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&sink), sizeof(double)));
  const int blocks = prop.multiProcessorCount * 5 * 8, threads = 256, updates = 390;
  cudaEvent_t e0, e1;
  BH8_CUDA(ctx, cudaEventCreate(&e0));
  BH8_CUDA(ctx, cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    BH8_CUDA(ctx, cudaEventRecord(e0, d.stream));
    // M = 10, camera at r0 = 2040, b = 300: G stays positive over 390 steps of du = 4e-6
    if (fp32)
      bh8::bh8_stepping_probe_kernel<float><<<blocks, threads, 0, d.stream>>>(sink, updates, 20.0, 4.9e-4, 4e-6, 1.1e-5);
    else
      bh8::bh8_stepping_probe_kernel<double><<<blocks, threads, 0, d.stream>>>(sink, updates, 20.0, 4.9e-4, 4e-6, 1.1e-5);
    BH8_CUDA(ctx, cudaGetLastError());
    BH8_CUDA(ctx, cudaEventRecord(e1, d.stream));
    BH8_CUDA(ctx, cudaEventSynchronize(e1));
    ctx->launches++;
    float ms = 0;
    BH8_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    const double rate = static_cast<double>(blocks) * threads * updates / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *updates_per_s = best;
  return BH8_OK;
}

int bh8_measure_fp64_peak(bh8_ctx* ctx, double* flops_per_s, double* seconds_run) {
  if (!ctx || !flops_per_s) return BH8_EINVAL;
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  cudaDeviceProp prop;
  BH8_CUDA(ctx, cudaGetDeviceProperties(&prop, d.ordinal));
  double* sink = nullptr;
  BH8_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&sink), sizeof(double)));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  cudaEvent_t e0, e1;
  BH8_CUDA(ctx, cudaEventCreate(&e0));
  BH8_CUDA(ctx, cudaEventCreate(&e1));
  double best = 0, total_s = 0;
  int iters = 2000;
  for (int rep = 0; rep < 6; ++rep) {
    BH8_CUDA(ctx, cudaEventRecord(e0, d.stream));
    bh8::bh8_dfma_peak_kernel<<<blocks, threads, 0, d.stream>>>(sink, iters, 0.999999, 1e-7);
    BH8_CUDA(ctx, cudaGetLastError());
    BH8_CUDA(ctx, cudaEventRecord(e1, d.stream));
    BH8_CUDA(ctx, cudaEventSynchronize(e1));
    ctx->launches++;
    float ms = 0;
    BH8_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * static_cast<double>(blocks) * threads;
    const double rate = flops / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
    total_s += ms * 1e-3;
    if (ms < 20.0) iters *= 4;  // aim for tens of ms per launch
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *flops_per_s = best;
  if (seconds_run) *seconds_run = total_s;
  return BH8_OK;
}

}  // extern "C"

#include "bh8_sink.h"
#include "bh8_script.h"
