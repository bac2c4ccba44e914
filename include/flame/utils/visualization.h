// flame/utils/visualization.h -- jet colour map (/root/reference/src/flame_offline_tum.cc:338-342).
#pragma once
#include <algorithm>

#include "flame/types.h"

namespace flame {
namespace utils {

// BGR jet colour of v in [vmin, vmax].
inline Vec3b jet(float v, float vmin, float vmax) {
  float t = (vmax > vmin) ? (v - vmin) / (vmax - vmin) : 0.f;
  t = std::min(1.f, std::max(0.f, t));
  auto ch = [](float x) { return (unsigned char)(255.f * std::min(1.f, std::max(0.f, x))); };
  const float r = 1.5f - std::fabs(4.f * t - 3.f), g = 1.5f - std::fabs(4.f * t - 2.f), b = 1.5f - std::fabs(4.f * t - 1.f);
  return Vec3b(ch(b), ch(g), ch(r));
}

}  // namespace utils
}  // namespace flame
