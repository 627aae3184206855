// bh8_script.h -- scripted animation behind bh8_script_* (SURVEY.md 8f-3): host side of bh8_anim.cuh.
//
// Included at the end of bh8_lib.cu (one translation unit: it uses bh8_ctx, Device, make_grid).
#ifndef BH8_SCRIPT_H_
#define BH8_SCRIPT_H_

#include <vector>

#include "bh8_anim.cuh"
#include "bh8_hud.h"

struct bh8_script {
  bh8_ctx* ctx = nullptr;
  int n_frames = 0, n_obj = 0, bh_index = -1;
  bh8_params prm{};
  int width = 0, height = 0;
  Bh8Frame* d_frames = nullptr;
  bh8_camera* d_cams = nullptr;
  bh8_object* d_objs = nullptr;
  // The frame constants the device built, read back ONCE at creation: a launch passes frame k's as the
  // kernel's __grid_constant__ parameter, exactly as a host-built frame travels.  (They used to be copied
  // device-to-device into one module-wide __constant__ symbol before each launch; two contexts
  // rendering scripts on the same GPU could then overwrite each other's constants under a running
  // kernel.  A parameter belongs to its launch.)
  std::vector<Bh8Frame> h_frames;
  // bh8_script_render_range: the captured launch sequence of the last range drawn
  cudaGraphExec_t graph = nullptr;
  int g_first = -1, g_count = 0;
  void* g_base = nullptr;
  size_t g_stride = 0;
};

namespace {

void script_free(bh8_script* s) {
  if (!s) return;
  if (s->ctx && cudaSetDevice(s->ctx->dev[0].ordinal) == cudaSuccess) {
    cudaStreamSynchronize(s->ctx->dev[0].stream);
    if (s->graph) cudaGraphExecDestroy(s->graph);
    cudaFree(s->d_frames);
    cudaFree(s->d_cams);
    cudaFree(s->d_objs);
  }
  delete s;
}

}  // namespace

extern "C" {

size_t bh8_frame_bytes(void) { return sizeof(Bh8Frame); }

int bh8_host_frame_constants(const bh8_ctx* ctx, const bh8_scene* scene, const bh8_camera* cam,
                             const bh8_params* params, void* out, size_t bytes) {
  if (!ctx || !out || bytes != sizeof(Bh8Frame)) return BH8_EINVAL;
  Bh8Frame f;
  char msg[192];
  const int rc = bh8_build_frame(scene, cam, params, ctx->tex_rows, ctx->tex_cols, &f, msg);
  if (rc != BH8_OK) return fail(const_cast<bh8_ctx*>(ctx), rc, msg);
  if (params->flags & BH8_FLAG_NO_BATCHING) f.resolve_wait = -1;
  std::memcpy(out, &f, sizeof f);
  return BH8_OK;
}

int bh8_script_create(bh8_ctx* ctx, const bh8_scene* scene0, const bh8_basis* obj_basis, const bh8_camera* cam0,
                      const bh8_params* params, const bh8_action* actions, int n_actions, int n_frames,
                      bh8_script** out) {
  if (!ctx) return BH8_EINVAL;
  if (!out) return fail(ctx, BH8_EINVAL, "bh8_script_create: null out");
  *out = nullptr;
  if (!scene0 || !scene0->obj || !cam0 || !params || n_frames < 1 || n_frames > (1 << 16) || n_actions < 0 ||
      (n_actions > 0 && !actions))
    return fail(ctx, BH8_EINVAL, "bad arguments to bh8_script_create");
  if (scene0->n_obj < 1 || scene0->n_obj > BH8_MAX_OBJECTS) return fail(ctx, BH8_EINVAL, "n_obj out of range");
  const int n_obj = scene0->n_obj;
  {  // frame 0 must be a valid scene: the common mistakes get their message before anything is launched
    Bh8Frame f;
    char msg[192];
    const int rc = bh8_build_frame(scene0, cam0, params, ctx->tex_rows, ctx->tex_cols, &f, msg);
    if (rc != BH8_OK) return fail(ctx, rc, msg);
  }
  std::vector<Bh8Action> acts(static_cast<size_t>(n_actions) + 1);
  if (const char* why = bh8a_prepare_actions(actions, n_actions, n_obj, acts.data())) return fail(ctx, BH8_EINVAL, why);
  std::vector<Bh8Entity> ent(static_cast<size_t>(n_obj) + 1);
  bh8a_prepare_entities(scene0, obj_basis, cam0, ent.data());

  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  bh8_script* s = new (std::nothrow) bh8_script();
  if (!s) return fail(ctx, BH8_ENOMEM, "out of host memory");
  s->ctx = ctx;
  s->n_frames = n_frames;
  s->n_obj = n_obj;
  s->bh_index = scene0->bh_index;
  s->prm = *params;
  s->width = cam0->width;
  s->height = cam0->height;
  Bh8Entity* d_ent = nullptr;
  Bh8Action* d_act = nullptr;
  bh8_object* d_proto = nullptr;
  int* d_status = nullptr;
  std::vector<int> status(2 * static_cast<size_t>(n_frames));
  s->h_frames.resize(static_cast<size_t>(n_frames));
  cudaError_t e = cudaSuccess;
  const auto step = [&](cudaError_t r) {
    if (e == cudaSuccess) e = r;
    return e == cudaSuccess;
  };
  step(cudaMalloc(reinterpret_cast<void**>(&d_ent), ent.size() * sizeof(Bh8Entity))) &&
      step(cudaMalloc(reinterpret_cast<void**>(&d_act), acts.size() * sizeof(Bh8Action))) &&
      step(cudaMalloc(reinterpret_cast<void**>(&d_proto), n_obj * sizeof(bh8_object))) &&
      step(cudaMalloc(reinterpret_cast<void**>(&d_status), status.size() * sizeof(int))) &&
      step(cudaMalloc(reinterpret_cast<void**>(&s->d_cams), static_cast<size_t>(n_frames) * sizeof(bh8_camera))) &&
      step(cudaMalloc(reinterpret_cast<void**>(&s->d_objs), static_cast<size_t>(n_frames) * n_obj * sizeof(bh8_object))) &&
      step(cudaMalloc(reinterpret_cast<void**>(&s->d_frames), static_cast<size_t>(n_frames) * sizeof(Bh8Frame))) &&
      step(cudaMemcpyAsync(d_ent, ent.data(), ent.size() * sizeof(Bh8Entity), cudaMemcpyHostToDevice, d.stream)) &&
      step(cudaMemcpyAsync(d_act, acts.data(), acts.size() * sizeof(Bh8Action), cudaMemcpyHostToDevice, d.stream)) &&
      step(cudaMemcpyAsync(d_proto, scene0->obj, n_obj * sizeof(bh8_object), cudaMemcpyHostToDevice, d.stream));
  if (e == cudaSuccess) {
    bh8::bh8_animate_kernel<<<1, 32, 0, d.stream>>>(d_ent, d_act, n_actions, n_obj, n_frames, d_proto, *cam0, s->d_cams,
                                                    s->d_objs);
    step(cudaGetLastError());
    bh8::Bh8TexSizes ts;
    for (int i = 0; i < BH8_MAX_TEXTURES; ++i) {
      ts.rows[i] = ctx->tex_rows[i];
      ts.cols[i] = ctx->tex_cols[i];
    }
    const int resolve_wait = ctx->resolve_wait;  // tuning knob, as launch_frame
    if (e == cudaSuccess) {
      bh8::bh8_build_frames_kernel<<<(n_frames + 63) / 64, 64, 0, d.stream>>>(
          s->d_cams, s->d_objs, n_obj, s->bh_index, s->prm, ts, resolve_wait, s->d_frames, d_status, n_frames);
      step(cudaGetLastError());
      ctx->launches += 2;
    }
    step(cudaMemcpyAsync(status.data(), d_status, status.size() * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
    step(cudaMemcpyAsync(s->h_frames.data(), s->d_frames, s->h_frames.size() * sizeof(Bh8Frame),
                         cudaMemcpyDeviceToHost, d.stream));
    step(cudaStreamSynchronize(d.stream));
  }
  cudaFree(d_ent);
  cudaFree(d_act);
  cudaFree(d_proto);
  cudaFree(d_status);
  if (e != cudaSuccess) {
    script_free(s);
    return fail(ctx, BH8_ECUDA, std::string("bh8_script_create: ") + cudaGetErrorString(e));
  }
  for (int k = 0; k < n_frames; ++k) {
    if (status[2 * k] != BH8F_OK) {
      const int reason = status[2 * k];
      script_free(s);
      return fail(ctx, bh8_frame_error_code(reason),
                  "bh8_script_create: frame " + std::to_string(k) + ": " + bh8_frame_error_text(reason));
    }
  }
  *out = s;
  return BH8_OK;
}

int bh8_script_frames(const bh8_script* s) { return s ? s->n_frames : 0; }

int bh8_script_render(bh8_script* s, int frame, void* d_pixels, void* d_class, void* d_key, void* d_steps) {
  if (!s) return BH8_EINVAL;
  bh8_ctx* ctx = s->ctx;
  if (frame < 0 || frame >= s->n_frames) return fail(ctx, BH8_EINVAL, "bh8_script_render: frame out of range");
  if (!d_pixels) return fail(ctx, BH8_EINVAL, "null pixel buffer");
  Device& d = ctx->dev[0];
  Bh8Frame f = s->h_frames[frame];
  return launch_built(ctx, d, f, d_pixels, d_class, d_key, d_steps, d.stream);
}

int bh8_script_render_range(bh8_script* s, int first, int count, void* d_base, size_t frame_stride_bytes) {
  if (!s) return BH8_EINVAL;
  bh8_ctx* ctx = s->ctx;
  if (first < 0 || count < 1 || first + count > s->n_frames)
    return fail(ctx, BH8_EINVAL, "bh8_script_render_range: frames out of range");
  const size_t frame_bytes = static_cast<size_t>(s->width) * s->height * bh8_pixel_bytes(s->prm.pixel_format);
  if (!d_base || frame_stride_bytes < frame_bytes)
    return fail(ctx, BH8_EINVAL, "bh8_script_render_range: null buffer or stride smaller than a frame");
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  if (!s->graph || s->g_first != first || s->g_count != count || s->g_base != d_base || s->g_stride != frame_stride_bytes) {
    if (s->graph) {
      BH8_CUDA(ctx, cudaStreamSynchronize(d.stream));
      cudaGraphExecDestroy(s->graph);
      s->graph = nullptr;
    }
    // Capture what `count` bh8_script_render() calls enqueue -- one kernel per frame, its constants in
    // the node's parameters -- into one graph: the whole range then costs the host a single launch.
    const uint64_t launches_before = ctx->launches;
    BH8_CUDA(ctx, cudaStreamBeginCapture(d.stream, cudaStreamCaptureModeThreadLocal));
    int rc = BH8_OK;
    for (int k = 0; k < count && rc == BH8_OK; ++k)
      rc = bh8_script_render(s, first + k, static_cast<uint8_t*>(d_base) + static_cast<size_t>(k) * frame_stride_bytes,
                             nullptr, nullptr, nullptr);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(d.stream, &graph);
    ctx->launches = launches_before;  // nothing ran yet
    if (rc != BH8_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (ce != cudaSuccess) return fail(ctx, BH8_ECUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&s->graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      s->graph = nullptr;
      return fail(ctx, BH8_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
    }
    s->g_first = first;
    s->g_count = count;
    s->g_base = d_base;
    s->g_stride = frame_stride_bytes;
  }
  BH8_CUDA(ctx, cudaGraphLaunch(s->graph, d.stream));
  ctx->launches += static_cast<uint64_t>(count);
  return BH8_OK;
}

int bh8_script_state(bh8_script* s, int frame, bh8_camera* cam, bh8_object* objs) {
  if (!s) return BH8_EINVAL;
  bh8_ctx* ctx = s->ctx;
  if (frame < 0 || frame >= s->n_frames) return fail(ctx, BH8_EINVAL, "bh8_script_state: frame out of range");
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  if (cam)
    BH8_CUDA(ctx, cudaMemcpyAsync(cam, s->d_cams + frame, sizeof(bh8_camera), cudaMemcpyDeviceToHost, d.stream));
  if (objs)
    BH8_CUDA(ctx, cudaMemcpyAsync(objs, s->d_objs + static_cast<size_t>(frame) * s->n_obj,
                                  s->n_obj * sizeof(bh8_object), cudaMemcpyDeviceToHost, d.stream));
  BH8_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return BH8_OK;
}

int bh8_script_frame_constants(bh8_script* s, int frame, void* out, size_t bytes) {
  if (!s) return BH8_EINVAL;
  bh8_ctx* ctx = s->ctx;
  if (frame < 0 || frame >= s->n_frames || !out || bytes != sizeof(Bh8Frame))
    return fail(ctx, BH8_EINVAL, "bad arguments to bh8_script_frame_constants");
  Device& d = ctx->dev[0];
  BH8_CUDA(ctx, cudaSetDevice(d.ordinal));
  BH8_CUDA(ctx, cudaMemcpyAsync(out, s->d_frames + frame, sizeof(Bh8Frame), cudaMemcpyDeviceToHost, d.stream));
  BH8_CUDA(ctx, cudaStreamSynchronize(d.stream));
  return BH8_OK;
}

void bh8_script_destroy(bh8_script* s) { script_free(s); }

int bh8_draw_text(uint8_t* bgr, int rows, int cols, size_t row_stride_bytes, int x, int y, const char* text, int b,
                  int g, int r) {
  if (!bgr || !text || rows < 1 || cols < 1) return BH8_EINVAL;
  if (row_stride_bytes == 0) row_stride_bytes = static_cast<size_t>(cols) * 3;
  if (row_stride_bytes < static_cast<size_t>(cols) * 3) return BH8_EINVAL;
  bh8_hud_draw(bgr, rows, cols, row_stride_bytes, x, y, text, static_cast<uint8_t>(b), static_cast<uint8_t>(g),
               static_cast<uint8_t>(r));
  return BH8_OK;
}

}  // extern "C"

#endif  // BH8_SCRIPT_H_
