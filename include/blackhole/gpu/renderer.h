// blackhole/gpu/renderer.h -- C++ convenience over the C ABI of include/bh8.h.
//
//   blackhole::gpu::Renderer gpu;                       // device 0; Renderer({0,1,...}) for several
//   gpu.Render(manager, blackhole, camera, &screen);    // replaces the pixel loop of
//                                                       // blackhole_solution_test.cc:161-308
// The frame lands in a cv::Mat CV_8UC3 (BGR), the layout the reference's drivers display and encode,
// so imshow / VideoWriter code stays as it is.  Link with libbh8.so.  Errors throw std::runtime_error
// with bh8_last_error(); there is no CPU fallback.
#ifndef BLACKHOLE_GPU_RENDERER_H_
#define BLACKHOLE_GPU_RENDERER_H_

#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "bh8.h"
#include "blackhole/gpu/snapshot.h"

namespace blackhole {
namespace gpu {

class Renderer {
 public:
  explicit Renderer(const std::vector<int>& devices = {0}) {
    if (bh8_create(&ctx_, devices.data(), static_cast<int>(devices.size())) != BH8_OK)
      throw std::runtime_error(std::string("bh8_create: ") + bh8_last_error(nullptr));
  }
  ~Renderer() {
    bh8_destroy(ctx_);
    for (void* p : pinned_) bh8_host_free(p);
  }
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  // An H x W CV_8UC3 frame in PINNED host memory owned by this Renderer (valid until it is destroyed): the
  // read-back into it is a plain DMA.  A frame in pageable memory -- any cv::Mat the caller allocated -- is
  // staged by the driver: 0.58 instead of 0.21 ms per 1080p frame (profiles/r02bj_pageable.txt).  Render()
  // hands out such frames itself whenever it has to allocate the frame (an empty Mat, another size); a Mat
  // the caller keeps allocating afresh stays pageable after kMaxPinnedFrames of them.
  cv::Mat PinnedFrame(int rows, int cols) {
    void* p = nullptr;
    if (bh8_host_alloc(&p, static_cast<size_t>(rows) * cols * 3) != BH8_OK)
      throw std::runtime_error("bh8_host_alloc: out of pinned host memory");
    pinned_.push_back(p);
    return cv::Mat(rows, cols, CV_8UC3, p);
  }

  bh8_ctx* context() { return ctx_; }
  const bh8_stats& last_stats() const { return stats_; }

  // One frame of the scene as the camera sees it, nstep = the literal of blackhole_solution_test.cc:202.
  template <typename T>
  void Render(const ObjectManager<T>& manager, const StaticBlackhole<T>& blackhole, const Camera<T>& camera,
              cv::Mat* frame, int nstep = 20) {
    const SceneSnapshot snap = Snapshot(manager, blackhole);
    const bh8_camera cam = Snapshot(camera);
    UploadTextures(snap);
    EnsureFrame(cam, frame);
    const bh8_scene scene = snap.view();
    bh8_params prm{};
    prm.nstep = nstep;
    prm.pixel_format = BH8_PIXEL_BGR8;
    Check(bh8_render(ctx_, &scene, &cam, 1, &prm, frame->data, nullptr, nullptr, nullptr, &stats_));
  }

  // Flat space: the pixel loop of ray_tracer_test.cc:140-155 -- RayTracer(camera.focus(), PixelVector)
  // with BasicLinearRayRecurrence, Prograde(manager, dst, steps) (ray_tracer.h:17-35,68-85) and the
  // colour of the hit -- for the whole frame in one call.
  template <typename T>
  void RenderLinear(const ObjectManager<T>& manager, const Camera<T>& camera, cv::Mat* frame, int steps = 10) {
    const SceneSnapshot snap = SnapshotFlat(manager);
    const bh8_camera cam = Snapshot(camera);
    UploadTextures(snap);
    EnsureFrame(cam, frame);
    const bh8_scene scene = snap.view();
    bh8_params prm{};
    prm.tracer = BH8_TRACER_LINEAR;
    prm.linear_steps = steps;
    prm.pixel_format = BH8_PIXEL_BGR8;
    Check(bh8_render(ctx_, &scene, &cam, 1, &prm, frame->data, nullptr, nullptr, nullptr, &stats_));
  }

  // A fly-through: one snapshot per frame (the host replays Camera / Object moves between
  // snapshots, object.h:58-88), rendered in one call; frames are dealt to the context's devices.
  void RenderFrames(const std::vector<SceneSnapshot>& scenes, const std::vector<bh8_camera>& cameras,
                    std::vector<cv::Mat>* frames, int nstep = 20) {
    if (scenes.size() != cameras.size() || scenes.empty()) throw std::runtime_error("scenes / cameras size mismatch");
    const int n = static_cast<int>(scenes.size());
    const size_t bytes = static_cast<size_t>(cameras[0].width) * cameras[0].height * 3;
    std::vector<bh8_scene> views(n);
    for (int i = 0; i < n; ++i) {
      views[i] = scenes[i].view();
      UploadTextures(scenes[i]);  // every scene's textures (normally the same buffers: sent once)
    }
    std::vector<unsigned char> pixels(bytes * n);
    bh8_params prm{};
    prm.nstep = nstep;
    prm.pixel_format = BH8_PIXEL_BGR8;
    Check(bh8_render(ctx_, views.data(), cameras.data(), n, &prm, pixels.data(), nullptr, nullptr, nullptr, &stats_));
    frames->clear();
    for (int i = 0; i < n; ++i) {
      cv::Mat m(cameras[i].height, cameras[i].width, CV_8UC3);
      std::memcpy(m.data, pixels.data() + bytes * i, bytes);
      frames->push_back(m);
    }
  }

 private:
  void Check(int rc) {
    if (rc != BH8_OK) throw std::runtime_error(std::string("bh8: ") + bh8_last_error(ctx_));
  }

  // The frame the kernels' rows are copied into: H x W CV_8UC3 with rows back to back (a ROI or padded Mat is
  // replaced by a fresh one rather than written through with the wrong stride).
  static constexpr size_t kMaxPinnedFrames = 8;
  void EnsureFrame(const bh8_camera& cam, cv::Mat* frame) {
    if (frame->empty() || frame->rows != cam.height || frame->cols != cam.width || frame->type() != CV_8UC3 ||
        !frame->isContinuous())
      *frame = pinned_.size() < kMaxPinnedFrames ? PinnedFrame(cam.height, cam.width)
                                                 : cv::Mat(cam.height, cam.width, CV_8UC3);
  }

  // Textures are re-sent when a slot's image changes: another buffer, another size or row stride.  (A caller
  // that rewrites the pixels of the SAME buffer in place calls InvalidateTextures().)
  struct TexKey {
    const unsigned char* data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
    bool operator==(const TexKey& o) const { return data == o.data && rows == o.rows && cols == o.cols && step == o.step; }
  };
  void UploadTextures(const SceneSnapshot& snap) {
    if (uploaded_.size() < snap.textures.size()) uploaded_.resize(snap.textures.size());
    for (size_t t = 0; t < snap.textures.size(); ++t) {
      const cv::Mat& m = snap.textures[t];
      const TexKey key{m.data, m.rows, m.cols, static_cast<size_t>(m.step)};
      if (uploaded_[t] == key) continue;
      Check(bh8_set_texture(ctx_, static_cast<int>(t), m.data, m.rows, m.cols, static_cast<size_t>(m.step)));
      uploaded_[t] = key;
    }
  }

 public:
  void InvalidateTextures() { uploaded_.clear(); }

 private:
  bh8_ctx* ctx_ = nullptr;
  bh8_stats stats_{};
  std::vector<TexKey> uploaded_;
  std::vector<void*> pinned_;  // PinnedFrame() buffers, freed with the Renderer
  friend class VideoWriter;
  friend class Script;
};

// The HUD of the reference's drivers (blackhole_solution_test.cc:309-326): five lines of green
// FONT_HERSHEY_PLAIN text -- camera position, basis vectors, field of view -- drawn into the frame after
// rendering.  bh8_draw_text reproduces cv::putText's pixels bit for bit, so this is a drop-in for that block
// on frames that came back to the host.  Returns the lines it drew.
// The five strings of that block, formatted as the reference formats them (operator<< of cv::Vec3d).
template <typename T>
std::vector<std::string> HudLines(const Camera<T>& camera) {
  std::vector<std::string> lines;
  const auto add = [&](const char* label, const auto& value, const char* suffix = "") {
    std::stringstream ss;
    ss << value;
    lines.push_back(std::string(label) + ss.str() + suffix);
  };
  add("Position: ", camera.focus());
  add("VectorX: ", camera.vector_x());
  add("VectorY: ", camera.vector_y());
  add("VectorZ: ", camera.vector_z());
  add("FoV: ", camera.fov() * 180.0 / blackhole::pi, " deg");
  return lines;
}

template <typename T>
std::vector<std::string> DrawHud(const Camera<T>& camera, cv::Mat* frame) {
  const std::vector<std::string> lines = HudLines(camera);
  for (size_t k = 0; k < lines.size(); ++k)  // {0, 10}, {0, 25}, {0, 40}, {0, 55}, {0, 70}
    bh8_draw_text(frame->data, frame->rows, frame->cols, static_cast<size_t>(frame->cols) * 3, 0,
                  10 + 15 * static_cast<int>(k), lines[k].c_str(), 0, 255, 0);
  return lines;
}

// A scripted animation: the Move* / Rotate* calls the reference's frame loop makes after
// each frame (key handlers and the disc spin, blackhole_solution_test.cc:346-407), written down and
// replayed ON THE GPU -- bit-identical to calling them on the live objects -- so that a long
// fly-through needs no per-frame host work at all:
//
//   blackhole::gpu::Script fly(&gpu);
//   for (int k = 0; k + 1 < frames; ++k) {
//     fly.Camera(k, BH8_OP_MOVE_X, 10);                          // camera.MoveX(10) after frame k
//     fly.Object(k, id_acc_disc.first, BH8_OP_ROTATE_Z, pi / 180);  // id_acc_disc.second->RotateZ(pi/180)
//   }
//   fly.Compile(manager, blackhole, camera, frames);             // state now = frame 0
//   fly.Render(k, &screen);
class Script {
 public:
  explicit Script(Renderer* gpu) : gpu_(gpu) {}
  ~Script() { Reset(); }
  Script(const Script&) = delete;
  Script& operator=(const Script&) = delete;

  // After frame `frame` is drawn: camera.<op>(amount).  Calls must be recorded in frame order.
  void Camera(int frame, int op, double amount) { actions_.push_back(Make(frame, BH8_TARGET_CAMERA, op, amount)); }
  // After frame `frame` is drawn: the object with this ObjectManager key .<op>(amount).
  void Object(int frame, int key, int op, double amount) { actions_.push_back(Make(frame, -2 - key, op, amount)); }

  template <typename T>
  void Compile(const ObjectManager<T>& manager, const StaticBlackhole<T>& blackhole, const ::blackhole::Camera<T>& camera,
               int n_frames, int nstep = 20) {
    Reset();
    const SceneSnapshot snap = Snapshot(manager, blackhole);
    const bh8_camera cam = Snapshot(camera);
    gpu_->UploadTextures(snap);
    std::vector<bh8_basis> basis;
    manager.ForEach([&](int, const DrawableObject<T>& obj) {
      bh8_basis b{};
      detail::Put3(b.vx, obj.vector_x());
      detail::Put3(b.vy, obj.vector_y());
      detail::Put3(b.vz, obj.vector_z());
      basis.push_back(b);
    });
    std::vector<bh8_action> acts = actions_;
    for (bh8_action& a : acts) {
      if (a.target == BH8_TARGET_CAMERA) continue;
      const int key = -2 - a.target;
      a.target = -2;
      for (size_t j = 0; j < snap.objects.size(); ++j)
        if (snap.objects[j].key == key) a.target = static_cast<int>(j);
      if (a.target < 0) throw std::runtime_error("Script: no object with key " + std::to_string(key));
    }
    const bh8_scene scene = snap.view();
    bh8_params prm{};
    prm.nstep = nstep;
    prm.pixel_format = BH8_PIXEL_BGR8;
    width_ = cam.width;
    height_ = cam.height;
    gpu_->Check(bh8_script_create(gpu_->ctx_, &scene, basis.data(), &cam, &prm, acts.data(), static_cast<int>(acts.size()),
                                  n_frames, &script_));
    gpu_->Check(bh8_frame_alloc(gpu_->ctx_, static_cast<size_t>(width_) * height_ * 3, &d_frame_));
  }

  int frames() const { return bh8_script_frames(script_); }

  // Frame `frame` into a CV_8UC3 (BGR) cv::Mat, like Renderer::Render.
  void Render(int frame, cv::Mat* out) {
    if (!script_) throw std::runtime_error("Script::Render before Compile");
    bh8_camera shape{};
    shape.width = width_;
    shape.height = height_;
    gpu_->EnsureFrame(shape, out);  // a pinned frame of the Renderer's when it has to be allocated
    gpu_->Check(bh8_script_render(script_, frame, d_frame_, nullptr, nullptr, nullptr));
    gpu_->Check(bh8_memcpy_d2h(gpu_->ctx_, out->data, d_frame_, static_cast<size_t>(width_) * height_ * 3));
  }

  // The camera / objects of frame `frame` as the device replay left them (for checks and HUD text).
  bh8_camera CameraAt(int frame) {
    bh8_camera cam{};
    gpu_->Check(bh8_script_state(script_, frame, &cam, nullptr));
    return cam;
  }

 private:
  static bh8_action Make(int frame, int target, int op, double amount) {
    bh8_action a{};
    a.frame = frame;
    a.target = target;
    a.op = op;
    a.amount = amount;
    return a;
  }
  void Reset() {
    if (d_frame_) bh8_frame_free(gpu_->ctx_, d_frame_);
    if (script_) bh8_script_destroy(script_);
    d_frame_ = nullptr;
    script_ = nullptr;
  }
  Renderer* gpu_;
  std::vector<bh8_action> actions_;
  bh8_script* script_ = nullptr;
  void* d_frame_ = nullptr;
  int width_ = 0, height_ = 0;
};

// Stands where the reference's cv::VideoWriter out_capture(save_dir/"video.avi",
// fourcc('M','J','P','G'), 29, size, true) stands (blackhole_solution_test.cc:71-72): a Motion-JPEG
// AVI, but the frames are traced AND JPEG-encoded on the GPU -- Write() replaces the pixel loop
// plus out_capture.write(*buf) (:161-308, :334); only the bitstream leaves the device.
class VideoWriter {
 public:
  VideoWriter(Renderer* gpu, const std::string& path, int width, int height, double fps = 29, int quality = 95)
      : gpu_(gpu) {
    if (bh8_sink_open(gpu->ctx_, path.c_str(), width, height, fps, quality, &sink_) != BH8_OK)
      throw std::runtime_error(std::string("bh8_sink_open: ") + bh8_last_error(gpu->ctx_));
  }
  ~VideoWriter() {
    if (sink_) bh8_sink_close(sink_, nullptr);
  }
  VideoWriter(const VideoWriter&) = delete;
  VideoWriter& operator=(const VideoWriter&) = delete;

  template <typename T>
  void Write(const ObjectManager<T>& manager, const StaticBlackhole<T>& blackhole, const Camera<T>& camera,
             int nstep = 20) {
    const SceneSnapshot snap = Snapshot(manager, blackhole);
    const bh8_camera cam = Snapshot(camera);
    gpu_->UploadTextures(snap);
    const bh8_scene scene = snap.view();
    bh8_params prm{};
    prm.nstep = nstep;
    if (bh8_sink_render(sink_, &scene, &cam, &prm) != BH8_OK)
      throw std::runtime_error(std::string("bh8_sink_render: ") + bh8_sink_last_error(sink_));
  }

  // The reference draws its HUD into the frame BEFORE out_capture.write (blackhole_solution_test.cc:309-334):
  // the frames written from now on carry these lines, drawn on the device before the encode.  Call it again
  // when the camera has moved; HudOff() stops it.
  template <typename T>
  void SetHud(const Camera<T>& camera) {
    const std::vector<std::string> text = HudLines(camera);
    std::vector<bh8_hud_line> lines(text.size());
    for (size_t k = 0; k < text.size(); ++k) {
      bh8_hud_line& ln = lines[k];
      std::memset(&ln, 0, sizeof ln);
      ln.x = 0;
      ln.y = 10 + 15 * static_cast<int>(k);
      ln.g = 255;
      std::strncpy(ln.text, text[k].c_str(), BH8_HUD_MAX_TEXT - 1);
    }
    if (bh8_sink_hud(sink_, lines.data(), static_cast<int>(lines.size())) != BH8_OK)
      throw std::runtime_error(std::string("bh8_sink_hud: ") + bh8_sink_last_error(sink_));
  }
  void HudOff() { bh8_sink_hud(sink_, nullptr, 0); }

  // Pipelined Write(): returns once the frame is queued; frame k+1 is traced while frame k is encoded.
  template <typename T>
  void WriteAsync(const ObjectManager<T>& manager, const StaticBlackhole<T>& blackhole, const Camera<T>& camera,
                  int nstep = 20) {
    const SceneSnapshot snap = Snapshot(manager, blackhole);
    const bh8_camera cam = Snapshot(camera);
    gpu_->UploadTextures(snap);
    const bh8_scene scene = snap.view();
    bh8_params prm{};
    prm.nstep = nstep;
    if (bh8_sink_submit(sink_, &scene, &cam, &prm) != BH8_OK)
      throw std::runtime_error(std::string("bh8_sink_submit: ") + bh8_sink_last_error(sink_));
  }
  void Flush() {
    if (bh8_sink_flush(sink_) != BH8_OK) throw std::runtime_error(std::string("bh8_sink_flush: ") + bh8_sink_last_error(sink_));
  }

  // One video from the per-GPU files of a frame-sharded job, frames back in order (host only).
  static unsigned long long Merge(const std::vector<std::string>& parts, const std::string& out_path) {
    std::vector<const char*> p;
    for (const std::string& s : parts) p.push_back(s.c_str());
    uint64_t frames = 0, bytes = 0;
    if (bh8_sink_merge(p.data(), static_cast<int>(p.size()), out_path.c_str(), &frames, &bytes) != BH8_OK)
      throw std::runtime_error(std::string("bh8_sink_merge: ") + bh8_last_error(nullptr));
    return frames;
  }

  // Finishes the file (index, sizes); returns its size in bytes.
  unsigned long long Release() {
    uint64_t bytes = 0;
    bh8_sink* s = sink_;
    sink_ = nullptr;
    if (s && bh8_sink_close(s, &bytes) != BH8_OK) throw std::runtime_error("bh8_sink_close: write error");
    return bytes;
  }

 private:
  Renderer* gpu_;
  bh8_sink* sink_ = nullptr;
};

}  // namespace gpu
}  // namespace blackhole

#endif  // BLACKHOLE_GPU_RENDERER_H_
