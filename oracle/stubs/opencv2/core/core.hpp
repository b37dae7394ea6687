// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <opencv2/core/core.hpp>.
//
// The reference's extras/homography.h:4 includes OpenCV core only because the
// host-side ImproveHomography (extras/homography.cu:280-346 in the file as
// shipped) uses cv::Mat / cv::Scalar / cv::solve on 8x8 and 8x1 fp64 matrices.
// OpenCV's C++ headers are not installed in this image, so the recipe in
// oracle/Makefile puts this directory on the include path when it compiles the
// UNMODIFIED reference sources into oracle/_ref/.  Nothing here is part of the
// product library.
#ifndef ORACLE_STUB_OPENCV_CORE_HPP
#define ORACLE_STUB_OPENCV_CORE_HPP

#include <cmath>
#include <cstring>
#include <vector>

#define CV_64FC1 6

namespace cv {

enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3 };

struct Scalar {
  double v;
  Scalar(double v_ = 0.0) : v(v_) {}
};

// Dense row-major fp64 matrix that always owns its storage.
class Mat {
public:
  int rows, cols;
  std::vector<double> buf;

  Mat() : rows(0), cols(0) {}
  Mat(int r, int c, int /*type*/) : rows(r), cols(c), buf((size_t)r * c, 0.0) {}
  Mat(int r, int c, int /*type*/, void *ext) : rows(r), cols(c), buf((size_t)r * c) {
    std::memcpy(buf.data(), ext, sizeof(double) * buf.size());
  }

  template <typename T> T &at(int i, int j) { return buf[(size_t)i * cols + j]; }
  template <typename T> T &at(int i) { return buf[(size_t)i]; }

  Mat &operator=(const Scalar &s) {
    for (size_t i = 0; i < buf.size(); i++) buf[i] = s.v;
    return *this;
  }
  Mat operator*(double s) const {
    Mat r(*this);
    for (size_t i = 0; i < r.buf.size(); i++) r.buf[i] *= s;
    return r;
  }
  Mat &operator+=(const Mat &o) {
    for (size_t i = 0; i < buf.size(); i++) buf[i] += o.buf[i];
    return *this;
  }
};

// Solve src1 * dst = src2 for a symmetric positive-definite src1 (Cholesky).
inline bool solve(const Mat &src1, const Mat &src2, Mat &dst, int /*flags*/) {
  const int n = src1.rows;
  std::vector<double> L((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j <= i; j++) {
      double s = src1.buf[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[(size_t)i * n + i] = std::sqrt(s);
      } else {
        L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
      }
    }
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) {
    double s = src2.buf[i];
    for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * y[k];
    y[i] = s / L[(size_t)i * n + i];
  }
  if (dst.rows != n || dst.cols != 1) dst = Mat(n, 1, CV_64FC1);
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * dst.buf[k];
    dst.buf[i] = s / L[(size_t)i * n + i];
  }
  return true;
}

}  // namespace cv

#endif
