// flame/utils/triangulator.h -- Delaunay triangulation helper (the `triangulate` stage,
// /root/reference/src/utils.cc:154) backed by libflame_b200's exact-predicate triangulator.
#pragma once
#include <vector>

#include "flame/types.h"
#include "flame_b200.h"

namespace flame {
namespace utils {

// Returns false on degenerate input (fewer than 3 points or all collinear).
inline bool triangulate(const std::vector<Point2f>& pts, std::vector<Triangle>* triangles, std::vector<Edge>* edges) {
  const int n = (int)pts.size();
  std::vector<float> xy(2 * (size_t)n);
  for (int i = 0; i < n; ++i) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
  std::vector<int32_t> t(6 * (size_t)n + 3), e(6 * (size_t)n + 2);
  int32_t nt = 0, ne = 0;
  if (fb_delaunay(n, xy.data(), t.data(), &nt, e.data(), &ne) != FB_OK) return false;
  if (triangles) { triangles->clear(); for (int k = 0; k < nt; ++k) triangles->push_back(Triangle(t[3 * k], t[3 * k + 1], t[3 * k + 2])); }
  if (edges) { edges->clear(); for (int k = 0; k < ne; ++k) edges->push_back(Edge(e[2 * k], e[2 * k + 1])); }
  return true;
}

}  // namespace utils
}  // namespace flame
