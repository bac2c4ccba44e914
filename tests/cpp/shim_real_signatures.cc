// Compile-only: include/flame/flame.h down its REAL-headers branch (Eigen / Sophus / OpenCV found via
// __has_include; here the stubs of tests/cpp/stubs/ that carry the real signatures), driven with the
// calls the reference frontends make:
//   ctor                               /root/reference/src/flame_nodelet.cc:523-527
//   update(time, id, pose, img, pf)    :634 ; with idepths_true  src/flame_offline_tum.cc:582-594
//   getInverseDepthMesh(...)           :669-676
//   getFilteredInverseDepthMap(&m)     :682-683 ; getInverseDepthMap() bound to const cv::Mat1f&  :687-688
//   getRawIDepths(...)                 :721-723
//   updatePoseFramePoses / prunePoseFrames  :474-475
//   stats().stats() / .timings()       :747-749 ; getDebugImage*() -> cv::Mat3b  :772-807
#include <memory>
#include <unordered_map>
#include <vector>

#include "flame/flame.h"
#include "flame/utils/load_tracker.h"

#ifndef FLAME_B200_HAVE_DEPS
#error "the real-headers branch of flame/types.h was not taken"
#endif

static void publish(const cv::Mat1f& /*idepthmap*/) {}

int drive(int width, int height) {
  Eigen::Matrix3f K = Eigen::Matrix3f::Identity();
  Eigen::Matrix3f Kinv = K.inverse();
  flame::Params params;
  params.omp_num_threads = 4;
  params.do_letterbox = false;
  params.min_grad_mag = 5.0f;
  params.min_error = 100.0f;
  params.detection_win_size = 16;
  params.zparams.win_size = 5;
  params.fparams.win_size = 5;
  params.max_dropouts = 5;
  params.zparams.epipolar_line_var = 4.0f;
  params.do_nltgv2 = true;
  params.adaptive_data_weights = false;
  params.rescale_data = false;
  params.init_with_prediction = true;
  params.idepth_var_max_graph = 0.01f;
  params.rparams.data_factor = 0.15f;
  params.min_height = -1e14f;
  params.max_height = 1e14f;
  params.check_sticky_obstacles = false;
  std::shared_ptr<flame::Flame> sensor = std::make_shared<flame::Flame>(width, height, K, Kinv, params);
  cv::Mat1b img_gray(height, width);
  cv::Mat1f idepths_true;
  Sophus::SE3f pose(Eigen::Quaternionf(1, 0, 0, 0), Eigen::Vector3f(0, 0, 0));
  bool update_success = sensor->update(0.0, 0u, pose, img_gray, true);
  update_success = sensor->update(0.1, 1u, pose, img_gray, false, idepths_true) && update_success;
  std::vector<cv::Point2f> vertices;
  std::vector<float> idepths, idepths_mu, idepths_var;
  std::vector<Eigen::Vector3f> normals;
  std::vector<flame::Triangle> triangles;
  std::vector<flame::Edge> edges;
  std::vector<bool> tri_validity;
  sensor->getInverseDepthMesh(&vertices, &idepths, &normals, &triangles, &tri_validity, &edges);
  cv::Mat1f idepthmap;
  sensor->getFilteredInverseDepthMap(&idepthmap);
  publish(sensor->getInverseDepthMap());
  sensor->getRawIDepths(&vertices, &idepths_mu, &idepths_var);
  std::vector<uint32_t> ids(1, 0u);
  std::vector<Sophus::SE3f> poses(1, pose);
  sensor->updatePoseFramePoses(ids, poses);
  sensor->prunePoseFrames(ids);
  const std::unordered_map<std::string, double>& st = sensor->stats().stats();
  const std::unordered_map<std::string, double>& tm = sensor->stats().timings();
  cv::Mat3b dbg = sensor->getDebugImageWireframe();
  cv::Mat3b dbg2 = sensor->getDebugImageInverseDepthMap();
  const uint32_t first = triangles.empty() ? 0u : triangles[0][2];
  return (int)st.size() + (int)tm.size() + dbg.rows + dbg2.cols + (int)first + (update_success ? 1 : 0);
}
