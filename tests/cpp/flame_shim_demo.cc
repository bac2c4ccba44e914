// Exercises include/flame/flame.h the way the reference frontends do
// (/root/reference/src/flame_offline_tum.cc:403-420,565-708): construct flame::Flame, call update()
// per frame, read the mesh / depth maps / raw idepths / stats.  Without a GPU the constructor must
// throw (no CPU fallback) -- that is what the CPU test checks.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "flame/flame.h"
#include "flame/utils/load_tracker.h"
#include "flame/utils/triangulator.h"

static float texture(float x, float y) {
  // smooth pseudo-random texture with strong gradients at ~6 px scale
  float v = 0.f;
  v += std::sin(0.91f * x + 1.3f) * std::cos(0.77f * y + 0.2f);
  v += std::sin(0.37f * x - 0.53f * y + 2.1f);
  v += std::cos(0.23f * x + 0.61f * y);
  v += std::sin(1.3f * x + 0.4f * y) * 0.7f;
  return 128.f + 32.f * v;
}

int main() {
  // utilities that never touch the GPU
  flame::utils::StatsTracker st("demo/");
  st.tick("t");
  st.set("k", 3.0);
  FLAME_ASSERT(st.tock("t") >= 0.0 && st.stats("demo/k") == 3.0);
  FLAME_ASSERT(flame::utils::fast_roundf(2.6f) == 3 && flame::utils::fast_abs(-2.f) == 2.f);
  flame::utils::LoadTracker lt(0);
  flame::utils::Load a, b, c;
  lt.get(&a, &b, &c);
  std::vector<flame::Point2f> pts = {{0, 0}, {10, 0}, {0, 10}, {10, 10}, {5, 4}};
  std::vector<flame::Triangle> tris;
  std::vector<flame::Edge> edges;
  FLAME_ASSERT(flame::utils::triangulate(pts, &tris, &edges) && tris.size() == 4 && edges.size() == 8);

  const int W = 320, H = 240;
  flame::Matrix3f K = flame::Matrix3f::Identity();
  K(0, 0) = 260.f; K(1, 1) = 260.f; K(0, 2) = 159.5f; K(1, 2) = 119.5f;
  flame::Params params;
  params.nltgv2_iters = 20;
  params.idepth_var_max_graph = 0.05f;
  std::unique_ptr<flame::Flame> sensor;
  try {
    sensor.reset(new flame::Flame(W, H, K, K, params));
  } catch (const std::exception& e) {
    std::printf("NOGPU %s\n", e.what());
    return 3;
  }
  const float depth = 2.0f;
  int n_ok = 0;
  for (int k = 0; k < 12; ++k) {
    const float tx = 0.02f * k;
    flame::Mat1b img(H, W);
    for (int v = 0; v < H; ++v)
      for (int u = 0; u < W; ++u) {
        // fronto-parallel plane at `depth`: world x of the pixel = (u-cx)/f*depth + tx
        const float X = (u - 159.5f) / 260.f * depth + tx, Y = (v - 119.5f) / 260.f * depth;
        const float val = texture(X * 130.f, Y * 130.f);
        img(v, u) = (unsigned char)std::min(255.f, std::max(0.f, val));
      }
    flame::SE3f pose(flame::compat::Quaternionf(1, 0, 0, 0), flame::Vector3f(tx, 0, 0));
    const bool ok = sensor->update(k / 30.0, (uint32_t)k, pose, img, k % 3 == 0);
    n_ok += ok ? 1 : 0;
  }
  std::vector<flame::Point2f> vtx;
  std::vector<float> idepths, mu, var;
  std::vector<flame::Vector3f> normals;
  std::vector<bool> valid;
  sensor->getInverseDepthMesh(&vtx, &idepths, &normals, &tris, &valid, &edges);
  flame::Mat1f filtered;
  sensor->getFilteredInverseDepthMap(&filtered);
  flame::Mat1f raw = sensor->getInverseDepthMap();
  std::vector<flame::Point2f> fpts;
  sensor->getRawIDepths(&fpts, &mu, &var);
  std::vector<float> sorted = idepths;
  std::sort(sorted.begin(), sorted.end());
  const float med = sorted.empty() ? 0.f : sorted[sorted.size() / 2];
  int covered = 0;
  for (int i = 0; i < W * H; ++i) covered += std::isnan(raw.ptr(0)[i]) ? 0 : 1;
  flame::Mat3b dbg = sensor->getDebugImageInverseDepthMap();
  // the stats keys the wrapper's fillStati / fillStatf look up (/root/reference/src/utils.cc:117-122,143-156)
  const auto& sm = sensor->stats().stats();
  const auto& tm = sensor->stats().timings();
  int keys_ok = 1;
  for (const char* k : {"num_feats", "num_vtx", "num_tris", "num_edges", "coverage", "num_idepth_updates",
                        "num_fail_max_var", "num_fail_max_dropouts", "num_fail_ref_patch_grad",
                        "num_fail_ambiguous_match", "num_fail_max_cost", "nltgv2_total_smoothness_cost",
                        "nltgv2_avg_smoothness_cost", "nltgv2_total_data_cost", "nltgv2_avg_data_cost", "fps_max"})
    if (!sm.count(k)) { std::printf("missing stat %s\n", k); keys_ok = 0; }
  for (const char* k : {"update", "update_locking"})
    if (!tm.count(k)) { std::printf("missing timing %s\n", k); keys_ok = 0; }
  std::printf("{\"keys_ok\": %d, \"num_vtx\": %.0f, \"num_tris\": %.0f, \"num_feats\": %.0f, \"coverage\": %.4f}\n", keys_ok,
              sensor->stats().stats("num_vtx"), sensor->stats().stats("num_tris"), sensor->stats().stats("num_feats"),
              sensor->stats().stats("coverage"));
  if (!keys_ok || sensor->stats().stats("num_vtx") != (double)vtx.size() || sensor->stats().stats("num_tris") != (double)tris.size()) return 2;
  std::printf("{\"updates\": %d, \"vertices\": %zu, \"triangles\": %zu, \"edges\": %zu, \"features\": %zu, "
              "\"median_idepth\": %.4f, \"covered\": %d, \"update_ms\": %.3f, \"num_idepth_updates\": %.0f, \"dbg_rows\": %d}\n",
              n_ok, vtx.size(), tris.size(), edges.size(), fpts.size(), med, covered,
              sensor->stats().timings("update"), sensor->stats().stats("num_idepth_updates"), dbg.rows);
  const bool good = n_ok >= 8 && vtx.size() > 50 && std::fabs(med - 1.0f / depth) < 0.05f && covered > W * H / 4;
  return good ? 0 : 1;
}
