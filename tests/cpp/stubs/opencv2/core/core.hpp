// Test stub with the signatures of the OpenCV types the frontends use with flame::Flame
// (/root/reference/src/flame_nodelet.cc:634,669-688,772-807): cv::Mat_<T> whose `data` is uchar*
// and whose `step` is a MatStep (NOT a size_t), ptr<T>(row), create(), clone(); cv::Point2f; cv::Vec3b.
#pragma once
#include <cstddef>
#include <vector>
typedef unsigned char uchar;
namespace cv {
struct MatStep {
  size_t p[2];
  MatStep() { p[0] = p[1] = 0; }
  operator size_t() const { return p[0]; }
  MatStep& operator=(size_t s) { p[0] = s; return *this; }
};
template <typename T, int N>
struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
  Vec(T a, T b, T c) { static_assert(N == 3, "3-vector ctor"); val[0] = a; val[1] = b; val[2] = c; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
class Mat {
 public:
  int rows, cols;
  uchar* data;
  MatStep step;
  Mat() : rows(0), cols(0), data(nullptr), esz_(1) {}
  bool empty() const { return rows == 0 || cols == 0; }
  bool isContinuous() const { return true; }
  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + step.p[0] * (size_t)r); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + step.p[0] * (size_t)r); }
  uchar* ptr(int r = 0) { return data + step.p[0] * (size_t)r; }
  const uchar* ptr(int r = 0) const { return data + step.p[0] * (size_t)r; }
 protected:
  void alloc(int r, int c, size_t esz) {
    rows = r; cols = c; esz_ = esz;
    store_.assign((size_t)r * c * esz, 0);
    data = store_.empty() ? nullptr : store_.data();
    step = (size_t)c * esz;
  }
  std::vector<uchar> store_;
  size_t esz_;
};
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) { alloc(r, c, sizeof(T)); }
  Mat_(int r, int c, const T& v) { alloc(r, c, sizeof(T)); for (int i = 0; i < r * c; ++i) reinterpret_cast<T*>(data)[i] = v; }
  Mat_(const Mat_& o) : Mat() { *this = o; }
  Mat_& operator=(const Mat_& o) { rows = o.rows; cols = o.cols; esz_ = o.esz_; store_ = o.store_; data = store_.empty() ? nullptr : store_.data(); step = o.step; return *this; }
  void create(int r, int c) { alloc(r, c, sizeof(T)); }
  Mat_ clone() const { return *this; }
  T& operator()(int r, int c) { return reinterpret_cast<T*>(data + step.p[0] * (size_t)r)[c]; }
  const T& operator()(int r, int c) const { return reinterpret_cast<const T*>(data + step.p[0] * (size_t)r)[c]; }
};
typedef Mat_<uchar> Mat1b;
typedef Mat_<float> Mat1f;
typedef Mat_<Vec3b> Mat3b;
}  // namespace cv
