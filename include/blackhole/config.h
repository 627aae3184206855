// blackhole/config.h -- compile-time resource / output directories (API of the reference's
// config.h:29-48).  -DBH_RESOURCE_DIR_INPUT=<dir> and -DBH_OUTPUT_DIR_INPUT=<dir> are stringified;
// resource_image() decodes with cv::imread (IMREAD_COLOR by default: 8-bit BGR, alpha dropped), which
// is the pixel data the hot path's texture lookups read.
#ifndef BLACKHOLE_CONFIG_H_
#define BLACKHOLE_CONFIG_H_

#include <chrono>
#include <filesystem>
#include <string>

#include "opencv2/opencv.hpp"

#define BH_STRINGIFY_IMPL(x) #x
#define BH_STRINGIFY(x) BH_STRINGIFY_IMPL(x)

#if defined(BH_RESOURCE_DIR_INPUT)
#define BH_RESOURCE_DIR BH_STRINGIFY(BH_RESOURCE_DIR_INPUT)
#else
#define BH_RESOURCE_DIR
#endif

#if defined(BH_OUTPUT_DIR_INPUT)
#define BH_OUTPUT_DIR BH_STRINGIFY(BH_OUTPUT_DIR_INPUT)
#else
#define BH_OUTPUT_DIR
#endif

namespace blackhole {

inline std::filesystem::path resource_dir() { return BH_RESOURCE_DIR; }
inline std::filesystem::path output_dir() { return BH_OUTPUT_DIR; }

inline cv::Mat resource_image(const std::string& subpath, int flag = cv::IMREAD_COLOR) {
  return cv::imread(resource_dir() / subpath, flag);
}

// output_dir()/<unix seconds>: one directory per run, as the drivers use for video.avi.
inline std::filesystem::path timed_output_dir() {
  const auto now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
  return output_dir() / std::to_string(now);
}

}  // namespace blackhole

#endif  // BLACKHOLE_CONFIG_H_
