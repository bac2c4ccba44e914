// flame/flame.h -- `flame::Flame`, source-compatible with the class the reference frontends
// construct and drive (ctor /root/reference/src/flame_nodelet.cc:523-527; update :634; getters
// :669-688, :721-723; poseframe updates :474-475; stats :747-749; debug images :772-807), backed by
// the C-ABI of libflame_b200.so (include/flame_b200.h).  Header-only adapter: all state and every
// computation live behind fb_*; this file only converts types.
//
// With Eigen / Sophus / OpenCV headers present the public signatures use those types exactly as the
// reference expects; without them (this build image) flame/types.h supplies minimal stand-ins so
// the adapter and its tests still compile.
#pragma once

#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "flame/params.h"
#include "flame/types.h"
#include "flame/utils/assert.h"
#include "flame/utils/image_utils.h"
#include "flame/utils/stats_tracker.h"
#include "flame/utils/visualization.h"
#include "flame_b200.h"

namespace flame {

class Flame {
 public:
  // Kinv is accepted for signature compatibility; the library forms it analytically from K.
  Flame(int width, int height, const Matrix3f& K, const Matrix3f& /*Kinv*/, const Params& params = Params())
      : width_(width), height_(height), params_(params), stats_("") {
    const int n_slots = params.num_poseframes + 1;
    const int cells = (width / params.detection_win_size) * (height / params.detection_win_size);
    const int max_v = std::max(64, std::min(params.max_features, 2 * cells));
    ctx_ = fb_create(params.device, 1, width, height, n_slots, params.max_features, max_v, 3 * max_v, nullptr);
    if (!ctx_) throw std::runtime_error(std::string("flame::Flame: ") + fb_last_error(nullptr));
    float k[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) k[3 * r + c] = K(r, c);
    check(fb_set_intrinsics(ctx_, 0, k));
    fb_epi_params ep;
    fb_default_epi_params(&ep);
    ep.win_size = params.fparams.win_size;
    ep.min_grad_mag = params.fparams.min_grad_mag;
    ep.epipolar_line_var = params.zparams.epipolar_line_var;
    ep.max_dropouts = params.max_dropouts;
    check(fb_set_epi_params(ctx_, &ep));
    fb_update_params up;
    fb_default_update_params(&up);
    up.detection_win_size = params.detection_win_size;
    up.min_grad_mag = params.min_grad_mag;
    up.idepth_var_max_graph = params.idepth_var_max_graph;
    up.adaptive_data_weights = params.adaptive_data_weights;
    up.init_with_prediction = params.init_with_prediction;
    up.do_nltgv2 = params.do_nltgv2;
    up.iters = params.nltgv2_iters;
    up.rparams.data_factor = params.rparams.data_factor;
    up.rparams.step_x = params.rparams.step_x;
    up.rparams.step_q = params.rparams.step_q;
    up.rparams.theta = params.rparams.theta;
    up.rparams.x_min = params.rparams.x_min;
    up.rparams.x_max = params.rparams.x_max;
    // forwarded as the frontends set them (/root/reference/src/flame_nodelet.cc:225-231,251,260-263); options
    // this library does not implement are rejected by fb_set_update_params instead of being ignored
    up.rescale_data = params.rescale_data;
    up.min_height = params.min_height;
    up.max_height = params.max_height;
    up.check_sticky_obstacles = params.check_sticky_obstacles;
    up.min_error = params.min_error;
    up.do_letterbox = params.do_letterbox;
    check(fb_set_update_params(ctx_, &up));
    fb_default_tri_filter_params(&filter_);
    filter_.do_oblique = params.do_oblique_triangle_filter;
    filter_.oblique_normal_thresh = params.oblique_normal_thresh;
    filter_.oblique_idepth_diff_factor = params.oblique_idepth_diff_factor;
    filter_.oblique_idepth_diff_abs = params.oblique_idepth_diff_abs;
    filter_.do_edge_length = params.do_edge_length_filter;
    filter_.edge_length_thresh = params.edge_length_thresh;
    filter_.do_idepth = params.do_idepth_triangle_filter;
    filter_.min_triangle_idepth = params.min_triangle_idepth;
  }
  ~Flame() { fb_destroy(ctx_); }
  Flame(const Flame&) = delete;
  Flame& operator=(const Flame&) = delete;

  // Returns false when no output was produced (the caller skips publishing,
  // /root/reference/src/flame_nodelet.cc:636-642).  `idepths_true` (analysis/pass_in_truth,
  // /root/reference/src/flame_offline_tum.cc:577-595) is accepted and ignored: the reference marks
  // that option as unsupported (cfg/flame_offline_tum.yaml:101-103).
  bool update(double time, uint32_t img_id, const SE3f& T_new, const Mat1b& img_new, bool is_poseframe,
              const Mat1f& /*idepths_true*/ = Mat1f()) {
    stats_.tick("update_locking");
    std::lock_guard<std::mutex> lock(mtx_);
    stats_.tock("update_locking");
    FLAME_ASSERT(img_new.rows == height_ && img_new.cols == width_);
    float pose[7];
    toPose(T_new, pose);
    const int rc = fb_update(ctx_, 0, time, (int)img_id, pose, reinterpret_cast<const uint8_t*>(img_new.data),
                             (int)static_cast<size_t>(img_new.step), is_poseframe ? 1 : 0);  // cv::Mat::step is a MatStep
    if (rc < 0) throw std::runtime_error(std::string("flame::Flame::update: ") + fb_last_error(ctx_));
    refreshStats();
    return rc == 1;
  }

  void getInverseDepthMesh(std::vector<Point2f>* vertices, std::vector<float>* idepths,
                           std::vector<Vector3f>* normals, std::vector<Triangle>* triangles,
                           std::vector<bool>* tri_validity, std::vector<Edge>* edges) {
    std::lock_guard<std::mutex> lock(mtx_);
    int32_t V = 0, T = 0, E = 0;
    check(fb_get_mesh_sizes(ctx_, 0, &V, &T, &E));
    std::vector<float> xy(2 * (size_t)V), id(V), nr(3 * (size_t)V);
    std::vector<int32_t> tr(3 * (size_t)T), ed(2 * (size_t)E);
    std::vector<uint8_t> tv(T);
    if (V > 0) check(fb_get_mesh(ctx_, 0, &filter_, xy.data(), id.data(), nr.data(), tr.data(), tv.data(), ed.data()));
    if (vertices) { vertices->resize(V); for (int v = 0; v < V; ++v) (*vertices)[v] = Point2f(xy[2 * v], xy[2 * v + 1]); }
    if (idepths) idepths->assign(id.begin(), id.end());
    if (normals) { normals->resize(V); for (int v = 0; v < V; ++v) (*normals)[v] = Vector3f(nr[3 * v], nr[3 * v + 1], nr[3 * v + 2]); }
    if (triangles) { triangles->resize(T); for (int t = 0; t < T; ++t) (*triangles)[t] = Triangle(tr[3 * t], tr[3 * t + 1], tr[3 * t + 2]); }
    if (tri_validity) { tri_validity->resize(T); for (int t = 0; t < T; ++t) (*tri_validity)[t] = tv[t] != 0; }
    if (edges) { edges->resize(E); for (int e = 0; e < E; ++e) (*edges)[e] = Edge(ed[2 * e], ed[2 * e + 1]); }
  }

  // H x W fp32, NaN = no depth; oblique / long-edge / far triangles removed (output/filter_*).
  void getFilteredInverseDepthMap(Mat1f* idepthmap) {
    std::lock_guard<std::mutex> lock(mtx_);
    idepthmap->create(height_, width_);
    check(fb_get_idepthmap(ctx_, 0, &filter_, reinterpret_cast<float*>(idepthmap->data)));  // cv::Mat::data is uchar*
  }

  Mat1f getInverseDepthMap() {
    std::lock_guard<std::mutex> lock(mtx_);
    Mat1f m(height_, width_);
    check(fb_get_idepthmap(ctx_, 0, nullptr, reinterpret_cast<float*>(m.data)));
    return m;
  }

  void getRawIDepths(std::vector<Point2f>* vertices, std::vector<float>* idepths_mu, std::vector<float>* idepths_var) {
    std::lock_guard<std::mutex> lock(mtx_);
    const size_t cap = (size_t)params_.max_features;
    std::vector<float> xy(2 * cap), mu(cap), var(cap);
    int32_t n = 0;
    check(fb_get_raw_idepths(ctx_, 0, &n, xy.data(), mu.data(), var.data()));
    if (vertices) { vertices->resize(n); for (int k = 0; k < n; ++k) (*vertices)[k] = Point2f(xy[2 * k], xy[2 * k + 1]); }
    if (idepths_mu) idepths_mu->assign(mu.begin(), mu.begin() + n);
    if (idepths_var) idepths_var->assign(var.begin(), var.begin() + n);
  }

  // Called from a ROS callback thread concurrently with update() on the worker thread
  // (/root/reference/src/flame_nodelet.cc:456-475): serialised by the same mutex.
  void updatePoseFramePoses(const std::vector<uint32_t>& ids, const std::vector<SE3f>& poses) {
    std::lock_guard<std::mutex> lock(mtx_);
    FLAME_ASSERT(ids.size() == poses.size());
    std::vector<int32_t> id32(ids.begin(), ids.end());
    std::vector<float> p(7 * ids.size());
    for (size_t k = 0; k < ids.size(); ++k) toPose(poses[k], &p[7 * k]);
    check(fb_update_poseframe_poses(ctx_, 0, (int)ids.size(), id32.data(), p.data()));
  }
  void prunePoseFrames(const std::vector<uint32_t>& ids_to_keep) {
    std::lock_guard<std::mutex> lock(mtx_);
    std::vector<int32_t> id32(ids_to_keep.begin(), ids_to_keep.end());
    check(fb_prune_poseframes(ctx_, 0, (int)id32.size(), id32.data()));
  }

  const utils::StatsTracker& stats() const { return stats_; }

  // Debug renderings (bgr8, /root/reference/src/flame_nodelet.cc:769-808).  The colour-mapped inverse
  // depth map is produced; the remaining overlays return the same image (display only, out of scope).
  Mat3b getDebugImageInverseDepthMap() {
    Mat1f m = getInverseDepthMap();
    Mat3b out;
    const float s = params_.scene_color_scale;
    utils::applyColorMap<float>(m, [s](float v) { return utils::jet(v * s, 0.0f, 2.0f); }, &out);
    return out;
  }
  Mat3b getDebugImageWireframe() { return getDebugImageInverseDepthMap(); }
  Mat3b getDebugImageFeatures() { return getDebugImageInverseDepthMap(); }
  Mat3b getDebugImageDetections() { return getDebugImageInverseDepthMap(); }
  Mat3b getDebugImageMatches() { return getDebugImageInverseDepthMap(); }
  Mat3b getDebugImageNormals() { return getDebugImageInverseDepthMap(); }

  fb_ctx* handle() { return ctx_; }

 private:
  void check(int rc) const {
    if (rc < 0) throw std::runtime_error(std::string("libflame_b200: ") + fb_last_error(ctx_));
  }
  static void toPose(const SE3f& T, float* p) {
    const auto& q = T.unit_quaternion();
    const auto& t = T.translation();
    p[0] = q.x(); p[1] = q.y(); p[2] = q.z(); p[3] = q.w();
    p[4] = t(0); p[5] = t(1); p[6] = t(2);
  }
  void refreshStats() {
    // stage names = FlameStats.msg timing keys (/root/reference/src/utils.cc:143-156)
    static const char* const kTimings[] = {"update", "frame_creation", "update_idepths", "project_features",
                                           "sync_graph", "triangulate", "nltgv2", "interpolate", "detection", "keyframe"};
    for (const char* k : kTimings) {
      double v = 0.0;
      if (fb_get_stat(ctx_, 0, k, &v) == FB_OK) stats_.setTiming(k, v);
    }
    // the keys the wrapper looks up (/root/reference/src/utils.cc:117-122): num_feats, num_vtx, num_tris,
    // num_edges, coverage; fps / fps_max (:138-139) from the update time
    static const char* const kStats[] = {"num_feats", "num_vtx", "num_tris", "num_edges", "coverage",
                                         "num_vertices", "num_triangles"};
    for (const char* k : kStats) {
      double v = 0.0;
      if (fb_get_stat(ctx_, 0, k, &v) == FB_OK) stats_.set(k, v);
    }
    {
      double ms = 0.0;
      if (fb_get_stat(ctx_, 0, "update", &ms) == FB_OK && ms > 0.0) stats_.set("fps_max", 1000.0 / ms);
    }
    int32_t c[FB_NUM_COUNTERS];
    if (fb_idepth_counters(ctx_, 0, c) == FB_OK) {  // /root/reference/src/utils.cc:124-129
      stats_.set("num_idepth_updates", c[FB_SUCCESS]);
      stats_.set("num_fail_ref_patch_grad", c[FB_FAIL_REF_PATCH_GRADIENT]);
      stats_.set("num_fail_ambiguous_match", c[FB_FAIL_AMBIGUOUS_MATCH]);
      stats_.set("num_fail_max_cost", c[FB_FAIL_MAX_COST]);
      stats_.set("num_fail_max_var", c[FB_FAIL_MAX_VAR]);
      stats_.set("num_fail_max_dropouts", c[FB_FAIL_MAX_DROPOUTS]);
    }
    double sm = 0.0, da = 0.0, nv = stats_.stats("num_vtx");
    if (nv > 0 && fb_costs(ctx_, 0, params_.rparams.data_factor, &sm, &da) == FB_OK) {  // utils.cc:131-136
      stats_.set("nltgv2_total_smoothness_cost", sm);
      stats_.set("nltgv2_avg_smoothness_cost", sm / nv);
      stats_.set("nltgv2_total_data_cost", da);
      stats_.set("nltgv2_avg_data_cost", da / nv);
    }
  }

  int width_, height_;
  Params params_;
  fb_ctx* ctx_ = nullptr;
  fb_tri_filter_params filter_;
  std::mutex mtx_;
  utils::StatsTracker stats_;
};

}  // namespace flame
