// flame/utils/image_utils.h -- small helpers the frontends call
// (fast_abs / fast_roundf: /root/reference/src/flame_nodelet.cc:649,729; applyColorMap:
// /root/reference/src/flame_offline_tum.cc:337-342).
#pragma once
#include <cmath>

#include "flame/types.h"

namespace flame {
namespace utils {

inline float fast_abs(float x) { return x < 0.f ? -x : x; }
inline int fast_roundf(float x) { return (int)(x + (x >= 0.f ? 0.5f : -0.5f)); }

// out(r,c) = fn(in(r,c)) for finite values, black otherwise.
template <typename T, typename Fn>
inline void applyColorMap(const Mat1f& in, Fn fn, Mat3b* out) {
  out->create(in.rows, in.cols);
  for (int r = 0; r < in.rows; ++r)
    for (int c = 0; c < in.cols; ++c) {
      const float v = in(r, c);
      (*out)(r, c) = std::isnan(v) ? Vec3b(0, 0, 0) : fn((T)v);
    }
}

}  // namespace utils
}  // namespace flame
