// flame/types.h -- mesh element types of the flame:: API (uses: /root/reference/src/utils.cc:224-226,
// /root/reference/src/flame_nodelet.cc:672-673).  Drop-in for the external flame core's header.
#pragma once

#include <stdint.h>

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && __has_include(<sophus/se3.hpp>) && __has_include(<opencv2/core/core.hpp>)
#define FLAME_B200_HAVE_DEPS 1
#endif
#endif

#ifdef FLAME_B200_HAVE_DEPS
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
#include <sophus/se3.hpp>
namespace flame {
typedef Eigen::Matrix3f Matrix3f;
typedef Eigen::Vector3f Vector3f;
typedef Sophus::SE3f SE3f;
typedef cv::Mat1b Mat1b;
typedef cv::Mat1f Mat1f;
typedef cv::Mat3b Mat3b;
typedef cv::Point2f Point2f;
typedef cv::Vec3b Vec3b;
}  // namespace flame
#else
// Eigen / Sophus / OpenCV are not installed in this image: minimal stand-ins with the members the
// frontends touch, so the shim and its tests compile.  With the real headers present the typedefs
// above are used instead and the reference's frontends compile against this header unchanged.
#include <cmath>
#include <cstring>
#include <vector>
namespace flame {
namespace compat {
struct Matrix3f {
  float m[9];  // row-major
  Matrix3f() { std::memset(m, 0, sizeof(m)); }
  float& operator()(int r, int c) { return m[3 * r + c]; }
  float operator()(int r, int c) const { return m[3 * r + c]; }
  static Matrix3f Identity() { Matrix3f k; k(0, 0) = k(1, 1) = k(2, 2) = 1.f; return k; }
};
struct Vector3f {
  float v[3];
  Vector3f() { v[0] = v[1] = v[2] = 0.f; }
  Vector3f(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
  float& operator()(int i) { return v[i]; }
  float operator()(int i) const { return v[i]; }
};
struct Quaternionf {
  float x_, y_, z_, w_;
  Quaternionf() : x_(0), y_(0), z_(0), w_(1) {}
  Quaternionf(float w, float x, float y, float z) : x_(x), y_(y), z_(z), w_(w) {}  // Eigen order
  float x() const { return x_; } float y() const { return y_; } float z() const { return z_; } float w() const { return w_; }
};
struct SE3f {
  Quaternionf q;
  Vector3f t;
  SE3f() {}
  SE3f(const Quaternionf& q_, const Vector3f& t_) : q(q_), t(t_) {}
  const Quaternionf& unit_quaternion() const { return q; }
  const Vector3f& translation() const { return t; }
};
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Vec3b { unsigned char val[3]; Vec3b() { val[0] = val[1] = val[2] = 0; } Vec3b(unsigned char a, unsigned char b, unsigned char c) { val[0] = a; val[1] = b; val[2] = c; } unsigned char& operator[](int i) { return val[i]; } unsigned char operator[](int i) const { return val[i]; } };
template <typename T>
struct Mat_ {
  int rows, cols;
  std::vector<T> store;
  T* data;
  size_t step;  // bytes per row
  Mat_() : rows(0), cols(0), data(nullptr), step(0) {}
  Mat_(int r, int c) { create(r, c); }
  Mat_(int r, int c, const T& v) { create(r, c); for (auto& e : store) e = v; }
  Mat_(const Mat_& o) { *this = o; }
  Mat_& operator=(const Mat_& o) { rows = o.rows; cols = o.cols; store = o.store; data = store.empty() ? nullptr : store.data(); step = sizeof(T) * cols; return *this; }
  void create(int r, int c) { rows = r; cols = c; store.assign((size_t)r * c, T()); data = store.data(); step = sizeof(T) * c; }
  bool empty() const { return rows == 0 || cols == 0; }
  T& operator()(int r, int c) { return store[(size_t)r * cols + c]; }
  const T& operator()(int r, int c) const { return store[(size_t)r * cols + c]; }
  T* ptr(int r = 0) { return store.data() + (size_t)r * cols; }
  const T* ptr(int r = 0) const { return store.data() + (size_t)r * cols; }
  bool isContinuous() const { return true; }
};
}  // namespace compat
typedef compat::Matrix3f Matrix3f;
typedef compat::Vector3f Vector3f;
typedef compat::SE3f SE3f;
typedef compat::Mat_<unsigned char> Mat1b;
typedef compat::Mat_<float> Mat1f;
typedef compat::Mat_<compat::Vec3b> Mat3b;
typedef compat::Point2f Point2f;
typedef compat::Vec3b Vec3b;
}  // namespace flame
#endif

namespace flame {

// Indexable [0..2], element type convertible to uint32 (/root/reference/src/utils.cc:224-226).
struct Triangle {
  uint32_t v[3];
  Triangle() { v[0] = v[1] = v[2] = 0; }
  Triangle(uint32_t a, uint32_t b, uint32_t c) { v[0] = a; v[1] = b; v[2] = c; }
  uint32_t& operator[](int k) { return v[k]; }
  const uint32_t& operator[](int k) const { return v[k]; }
};

struct Edge {
  uint32_t v[2];
  Edge() { v[0] = v[1] = 0; }
  Edge(uint32_t a, uint32_t b) { v[0] = a; v[1] = b; }
  uint32_t& operator[](int k) { return v[k]; }
  const uint32_t& operator[](int k) const { return v[k]; }
};

}  // namespace flame
