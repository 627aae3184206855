// ray_tracer_gpu -- GPU counterpart of the reference's ray_tracer_test driver (flat space).
//
// Same scene through the same header API (ray_tracer_test.cc:45-98: three textured rectangles, a chess
// floor, a tilted camera), same frame loop shape -- render, "Took N ms", then the "Movement test" that
// moves and turns the rectangles (:237-261) -- but the pixel loop of :140-155
// (RayTracer(camera.focus(), PixelVector).Prograde(manager, dst, 10) per pixel) is one call into the
// CUDA renderer's linear tracer.  Headless: runs a scripted number of frames instead of reading keys.
//
//   ray_tracer_gpu [--width W] [--height H] [--frames K] [--texdir DIR] [--out PREFIX]
// Writes PREFIX_<frame>.bgr (raw: int32 rows, int32 cols, BGR bytes) when --out is given.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "blackhole/gpu/renderer.h"
#include "scenes.h"  // apps/scenes.h: cfg 10 is the ray_tracer_test scene

int main(int argc, char** argv) {
  int width = 800, height = 450, frames = 1;  // ray_tracer_test.cc:36-37
  std::string texdir = "build/textures", out;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--width") width = std::atoi(next());
    else if (a == "--height") height = std::atoi(next());
    else if (a == "--frames") frames = std::atoi(next());
    else if (a == "--texdir") texdir = next();
    else if (a == "--out") out = next();
    else {
      std::fprintf(stderr, "unknown argument %s\n", a.c_str());
      return 2;
    }
  }
  bh8scenes::Scene* scene = bh8scenes::Build(10, width, height, 0, texdir);
  auto& manager = bh8scenes::manager_type::GetInstance();
  try {
    blackhole::gpu::Renderer gpu;
    cv::Mat screen;
    int karina_move_direction = 1;
    for (int k = 0; k < frames; ++k) {
      const auto t1 = std::chrono::high_resolution_clock::now();
      gpu.RenderLinear(manager, scene->camera, &screen, scene->linear_steps);
      const auto t2 = std::chrono::high_resolution_clock::now();
      const auto us = std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
      std::cout << "Took " << us / 1000.0 << "ms (kernel " << gpu.last_stats().kernel_ms << "ms, "
                << gpu.last_stats().steps << " segments)\n";
      if (!out.empty()) cv::imwrite(out + "_" + std::to_string(k), screen);
      // "Movement test", ray_tracer_test.cc:237-261
      scene->movers[1]->MoveX(3 * karina_move_direction);
      if (auto x = scene->movers[1]->position()[0]; x > 100)
        karina_move_direction = -1;
      else if (x < 0)
        karina_move_direction = 1;
      {
        const auto y = -120.0, z = -100.0;
        scene->movers[0]->MoveY(-y / 2);
        scene->movers[0]->MoveZ(-z / 2);
        scene->movers[0]->RotateX(blackhole::pi / 180);
        scene->movers[0]->MoveY(y / 2);
        scene->movers[0]->MoveZ(z / 2);
      }
      {
        const auto x = 100.0, z = -100.0;
        scene->movers[2]->MoveX(-x);
        scene->movers[2]->MoveZ(-z / 2);
        scene->movers[2]->RotateY(blackhole::pi / 120);
        scene->movers[2]->MoveX(x);
        scene->movers[2]->MoveZ(z / 2);
      }
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
