// Test stub with the signatures of Sophus::SE3f the frontends use (see Eigen/Core in this directory).
#pragma once
#include <Eigen/Core>
namespace Sophus {
template <typename S>
class SE3 {
 public:
  SE3() {}
  SE3(const Eigen::Quaternion<S>& q, const Eigen::Matrix<S, 3, 1>& t) : q_(q), t_(t) {}
  const Eigen::Quaternion<S>& unit_quaternion() const { return q_; }
  const Eigen::Matrix<S, 3, 1>& translation() const { return t_; }
 private:
  Eigen::Quaternion<S> q_;
  Eigen::Matrix<S, 3, 1> t_;
};
typedef SE3<float> SE3f;
}  // namespace Sophus
