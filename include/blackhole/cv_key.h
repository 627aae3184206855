// blackhole/cv_key.h -- cv::waitKeyEx codes the interactive drivers react to (macOS values, as in
// the reference's cv_key.h:10-16).
#ifndef BLACKHOLE_CV_KEY_H_
#define BLACKHOLE_CV_KEY_H_

namespace blackhole {

enum Key {
  kEscape = 27,
  kUp = 63232,
  kDown = 63233,
  kLeft = 63234,
  kRight = 63235,
};

}  // namespace blackhole

#endif  // BLACKHOLE_CV_KEY_H_
