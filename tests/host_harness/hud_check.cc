// TEST-ONLY: blackhole::gpu::DrawHud on a blank frame for a camera that has been moved and turned, no GPU
// involved (bh8_draw_text is host code).  Prints the lines it drew and writes the frame as raw BGR bytes;
// tests/test_hud.py redraws the same lines with cv2.putText and compares.
#include <cstdio>
#include <iostream>

#include "blackhole/gpu/renderer.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  blackhole::Camera<double> camera(640, 360, blackhole::pi / 2);
  camera.MoveTo(-2000, 0, 400);
  for (int k = 0; k < 7; ++k) {
    camera.RotateZ(blackhole::pi / 1800.0 * 10);
    camera.MoveY(10);
  }
  cv::Mat frame(360, 640, CV_8UC3, cv::Scalar(0, 0, 0));
  for (const std::string& line : blackhole::gpu::DrawHud(camera, &frame)) std::cout << line << "\n";
  FILE* f = std::fopen(argv[1], "wb");
  if (!f) return 1;
  std::fwrite(frame.data, 1, 640 * 360 * 3, f);
  std::fclose(f);
  return 0;
}
