// b200reg_compat.hpp -- point-cloud / matrix types for the facade classes.
//
// Where the reference is built (ROS Noetic: PCL 1.10 + Eigen 3.3) the real headers are used and the facades below are
// drop-in for third_party/nano_gicp and third_party/Quatro.  Where they are absent (this repository's CI container has
// neither PCL nor Eigen) a minimal stand-in with the same member names is provided so that the
// call sequence of fast_lio_sam_qn/src/loop_closure.cpp can be compiled and exercised in the tests.
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>

#if defined(__has_include)
#if __has_include(<pcl/point_cloud.h>) && __has_include(<Eigen/Core>)
#define B200REG_HAVE_PCL 1
#endif
#endif

#ifdef B200REG_HAVE_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <Eigen/Core>
#include <Eigen/StdVector>
#else
namespace Eigen {
// column-major 4x4, the subset of Eigen::Matrix<S,4,4> the call sites use
template <typename S>
struct Mat4 {
  S m[16];
  Mat4() {
    for (int i = 0; i < 16; i++) m[i] = S(0);
  }
  static Mat4 Identity() {
    Mat4 r;
    r.m[0] = r.m[5] = r.m[10] = r.m[15] = S(1);
    return r;
  }
  S& operator()(int r, int c) { return m[4 * c + r]; }
  S operator()(int r, int c) const { return m[4 * c + r]; }
  S* data() { return m; }
  const S* data() const { return m; }
  template <typename T>
  Mat4<T> cast() const {
    Mat4<T> r;
    for (int i = 0; i < 16; i++) r.m[i] = static_cast<T>(m[i]);
    return r;
  }
  Mat4 operator*(const Mat4& o) const {
    Mat4 r;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        S s = S(0);
        for (int k = 0; k < 4; k++) s += (*this)(i, k) * o(k, j);
        r(i, j) = s;
      }
    return r;
  }
};
using Matrix4f = Mat4<float>;
using Matrix4d = Mat4<double>;
// column-major R x C, the subset of Eigen::Matrix<S,R,C> getFinalHessian()'s callers use
template <typename S, int R, int C>
struct Matrix {
  S m[R * C];
  Matrix() {
    for (int i = 0; i < R * C; i++) m[i] = S(0);
  }
  static Matrix Identity() {
    Matrix r;
    for (int i = 0; i < (R < C ? R : C); i++) r.m[i * R + i] = S(1);
    return r;
  }
  S& operator()(int r, int c) { return m[R * c + r]; }
  S operator()(int r, int c) const { return m[R * c + r]; }
  S* data() { return m; }
  const S* data() const { return m; }
  static constexpr int rows() { return R; }
  static constexpr int cols() { return C; }
};
template <typename T>
using aligned_allocator = std::allocator<T>;
}  // namespace Eigen

namespace pcl {
struct alignas(16) PointXYZ {
  union {
    float data[4];
    struct {
      float x, y, z;
    };
  };
  PointXYZ() : data{0.f, 0.f, 0.f, 1.f} {}
  PointXYZ(float x_, float y_, float z_) : data{x_, y_, z_, 1.f} {}
};
struct alignas(16) PointXYZI {  // 32 bytes, like PCL's
  union {
    float data[4];
    struct {
      float x, y, z;
    };
  };
  union {
    struct {
      float intensity;
    };
    float data_c[4];
  };
  PointXYZI() : data{0.f, 0.f, 0.f, 1.f}, data_c{0.f, 0.f, 0.f, 0.f} {}
};
template <typename PointT>
class PointCloud {
 public:
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; }
  void reserve(size_t n) { points.reserve(n); }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointT& at(size_t i) { return points.at(i); }
  const PointT& at(size_t i) const { return points.at(i); }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = (uint32_t)points.size();
    return *this;
  }
};
}  // namespace pcl
#endif

#include "../../include/b200reg.h"

namespace b200reg_host {

// One lazily created engine context per host thread (LoopClosure is single-instance and not re-entrant,
// fast_lio_sam_qn/src/fast_lio_sam_qn.cpp:81).  Device = $B200REG_DEVICE or 0.  Failing to get a GPU is fatal:
// there is no CPU fallback behind these classes.
b200reg_ctx* context();

// Eigen (column-major) <-> C ABI (row-major)
template <typename M, typename S>
inline void to_rowmajor(const M& m, S out[16]) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[4 * r + c] = static_cast<S>(m(r, c));
}
template <typename M, typename S>
inline void from_rowmajor(const S in[16], M& m) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = in[4 * r + c];
}

}  // namespace b200reg_host

#ifdef B200REG_HOST_IMPLEMENTATION
#include <cstdio>
#include <cstdlib>
namespace b200reg_host {
b200reg_ctx* context() {
  static thread_local b200reg_ctx* ctx = nullptr;
  if (!ctx) {
    const char* dev = std::getenv("B200REG_DEVICE");
    const int rc = b200reg_ctx_create(dev ? std::atoi(dev) : 0, &ctx);
    if (rc != 0) {
      std::fprintf(stderr, "b200reg: cannot create a GPU context (%d): %s -- this build has no CPU fallback\n", rc, b200reg_last_error());
      std::abort();
    }
  }
  return ctx;
}
}  // namespace b200reg_host
#endif
