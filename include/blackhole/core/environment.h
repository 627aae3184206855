// blackhole/core/environment.h -- key codes and resource / output directories.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/cv_key.h -- cv::waitKeyEx codes the interactive drivers react to (macOS values, as in
// the reference's cv_key.h:10-16).
// blackhole/config.h -- compile-time resource / output directories (API of the reference's
// config.h:29-48).  -DBH_RESOURCE_DIR_INPUT=<dir> and -DBH_OUTPUT_DIR_INPUT=<dir> are stringified;
// resource_image() decodes with cv::imread (IMREAD_COLOR by default: 8-bit BGR, alpha dropped), which
// is the pixel data the hot path's texture lookups read.
#ifndef BLACKHOLE_CORE_ENVIRONMENT_H_
#define BLACKHOLE_CORE_ENVIRONMENT_H_

#include <chrono>
#include <filesystem>
#include <string>

#include "opencv2/opencv.hpp"

namespace blackhole {

enum Key {
  kEscape = 27,
  kUp = 63232,
  kDown = 63233,
  kLeft = 63234,
  kRight = 63235,
};

}  // namespace blackhole

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

#endif  // BLACKHOLE_CORE_ENVIRONMENT_H_
