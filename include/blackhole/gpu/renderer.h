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
  ~Renderer() { bh8_destroy(ctx_); }
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  bh8_ctx* context() { return ctx_; }
  const bh8_stats& last_stats() const { return stats_; }

  // One frame of the scene as the camera sees it, nstep = the literal of blackhole_solution_test.cc:202.
  template <typename T>
  void Render(const ObjectManager<T>& manager, const StaticBlackhole<T>& blackhole, const Camera<T>& camera,
              cv::Mat* frame, int nstep = 20) {
    const SceneSnapshot snap = Snapshot(manager, blackhole);
    const bh8_camera cam = Snapshot(camera);
    UploadTextures(snap);
    if (frame->empty() || frame->rows != cam.height || frame->cols != cam.width || frame->type() != CV_8UC3)
      *frame = cv::Mat(cam.height, cam.width, CV_8UC3);
    const bh8_scene scene = snap.view();
    bh8_params prm{};
    prm.nstep = nstep;
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
    for (int i = 0; i < n; ++i) views[i] = scenes[i].view();
    UploadTextures(scenes[0]);
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

  // Textures are re-sent only when a slot's pixel buffer changes.
  void UploadTextures(const SceneSnapshot& snap) {
    if (uploaded_.size() < snap.textures.size()) uploaded_.resize(snap.textures.size(), nullptr);
    for (size_t t = 0; t < snap.textures.size(); ++t) {
      const cv::Mat& m = snap.textures[t];
      if (uploaded_[t] == m.data) continue;
      Check(bh8_set_texture(ctx_, static_cast<int>(t), m.data, m.rows, m.cols, static_cast<size_t>(m.cols) * 3));
      uploaded_[t] = m.data;
    }
  }

  bh8_ctx* ctx_ = nullptr;
  bh8_stats stats_{};
  std::vector<const unsigned char*> uploaded_;
  friend class VideoWriter;
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
