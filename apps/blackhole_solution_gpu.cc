// blackhole_solution_gpu -- GPU counterpart of the reference's blackhole_solution_test driver.
//
// Same scene construction through the same header API (include/blackhole), same frame loop shape
// (render, "Took N ms", hand the frame to the video writer, move the camera / spin the disc,
// blackhole_solution_test.cc:153-408) -- but the per-pixel loop of :161-308 is one call into the
// CUDA renderer.  Headless: runs a scripted number of frames instead of reading keys.
//
//   blackhole_solution_gpu [--cfg N] [--width W] [--height H] [--frames K] [--nstep S]
//                          [--texdir DIR] [--out PREFIX] [--video FILE.avi] [--script] [--hud]
// --hud: the reference's HUD text (:309-326) is drawn into the frames written with --out (on the host) and
// into the frames of --video (on the device, before the encode).
// --script: the frame loop's tail (camera script + disc spin) is recorded as a blackhole::gpu::Script and
// replayed on the GPU; frames are drawn from the device-resident script instead of host snapshots.
// Writes PREFIX_<frame>.bgr (raw: int32 rows, int32 cols, BGR bytes) when --out is given, and a
// Motion-JPEG AVI (the reference's video.avi, :71-72; frames encoded on the GPU) when --video is.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>

#include "blackhole/gpu/renderer.h"
#include "scenes.h"  // apps/scenes.h: the BASELINE scenes, written against the public header API

int main(int argc, char** argv) {
  int cfg = 0, width = 960, height = 540, frames = 1, nstep = -1;
  std::string texdir = "build/textures", out, video;
  bool scripted = false, hud = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--cfg") cfg = std::atoi(next());
    else if (a == "--width") width = std::atoi(next());
    else if (a == "--height") height = std::atoi(next());
    else if (a == "--frames") frames = std::atoi(next());
    else if (a == "--nstep") nstep = std::atoi(next());
    else if (a == "--texdir") texdir = next();
    else if (a == "--out") out = next();
    else if (a == "--video") video = next();
    else if (a == "--script") scripted = true;
    else if (a == "--hud") hud = true;
    else {
      std::fprintf(stderr, "unknown argument %s\n", a.c_str());
      return 2;
    }
  }
  bh8scenes::Scene* scene = bh8scenes::Build(cfg, width, height, 0, texdir);
  if (!scene) {
    std::fprintf(stderr, "unknown cfg %d\n", cfg);
    return 2;
  }
  if (nstep <= 0) nstep = scene->nstep;
  auto& manager = bh8scenes::manager_type::GetInstance();

  try {
    blackhole::gpu::Renderer gpu;
    std::unique_ptr<blackhole::gpu::VideoWriter> out_capture;
    if (!video.empty()) out_capture.reset(new blackhole::gpu::VideoWriter(&gpu, video, width, height, 29));
    cv::Mat screen;
    if (scripted) {  // the same loop tail, written down once and replayed on the device
      if (!scene->disc) throw std::runtime_error("--script needs a scene with the accretion disc");
      blackhole::gpu::Script fly(&gpu);
      int disc_key = -1;
      manager.ForEach([&](int key, const blackhole::DrawableObject<bh8scenes::value_type>& obj) {
        if (&obj == scene->disc) disc_key = key;
      });
      for (int k = 0; k + 1 < frames; ++k) {
        if (cfg == 3) {
          if (k < 120) {
            fly.Camera(k, BH8_OP_MOVE_X, 10);
          } else {
            fly.Camera(k, BH8_OP_ROTATE_Z, blackhole::pi / 1800.0 * 10);
            fly.Camera(k, BH8_OP_MOVE_Y, 10);
          }
        }
        fly.Object(k, disc_key, BH8_OP_ROTATE_Z, blackhole::pi / 180);
      }
      fly.Compile(manager, *scene->blackhole, scene->camera, frames, nstep);
      for (int k = 0; k < frames; ++k) {
        fly.Render(k, &screen);
        if (!out.empty()) cv::imwrite(out + "_" + std::to_string(k), screen);
      }
      const bh8_camera last = fly.CameraAt(frames - 1);
      std::cout << "script: " << frames << " frames, camera ends at (" << last.pos[0] << ", " << last.pos[1] << ", "
                << last.pos[2] << ")\n";
      return 0;
    }
    for (int k = 0; k < frames; ++k) {
      const auto t1 = std::chrono::high_resolution_clock::now();
      if (out_capture) {
        if (hud) out_capture->SetHud(scene->camera);  // :309-326 come before out_capture.write(), :334
        out_capture->Write(manager, *scene->blackhole, scene->camera, nstep);
      }
      if (!out_capture || !out.empty()) gpu.Render(manager, *scene->blackhole, scene->camera, &screen, nstep);
      const auto t2 = std::chrono::high_resolution_clock::now();
      const auto us = std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
      std::cout << "Took " << us / 1000.0 << "ms (kernel " << gpu.last_stats().kernel_ms << "ms, "
                << gpu.last_stats().steps << " geodesic steps)\n";
      if (hud && !out.empty()) blackhole::gpu::DrawHud(scene->camera, &screen);  // :309-326
      if (!out.empty()) cv::imwrite(out + "_" + std::to_string(k), screen);
      // frame loop tail, blackhole_solution_test.cc:346-407: fly-through script + disc spin
      if (cfg == 3) {
        if (k < 120) {
          scene->camera.MoveX(10);
        } else {
          scene->camera.RotateZ(blackhole::pi / 1800.0 * 10);
          scene->camera.MoveY(10);
        }
      }
      if (scene->disc) scene->disc->RotateZ(blackhole::pi / 180);
    }
    if (out_capture) std::cout << "video: " << out_capture->Release() << " bytes\n";
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
