// Harness compatibility pack: the handful of OpenCV names the reference's main.cpp and
// test/test.cpp use (cv::Mat with rows / cols / data / convertTo, cv::Size, CV_32FC1, imread, imwrite, GaussianBlur),
// so that those files compile and link UNCHANGED against libcusift_b200.so in an image without OpenCV's C++ headers.
// Not part of the product library.  Image I/O understands binary PGM (P5, 8-bit) and raw little-endian float32 files
// named <anything>.<width>x<height>.f32; JPEG needs the real OpenCV.
#ifndef CSB_COMPAT_OPENCV_CORE_HPP
#define CSB_COMPAT_OPENCV_CORE_HPP

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_32FC1 5

namespace cv {

typedef unsigned char uchar;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

// Reference-counted dense single-channel matrix (8-bit or float), like cv::Mat for the uses above.
class Mat {
 public:
  int rows, cols;
  uchar *data;
  Mat() : rows(0), cols(0), data(NULL), type_(CV_8UC1) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    buf_.reset(new std::vector<uchar>((size_t)r * c * elemSize()));
    data = buf_->data();
  }
  int type() const { return type_; }
  size_t elemSize() const { return type_ == CV_32FC1 ? 4 : 1; }
  bool empty() const { return data == NULL || rows == 0 || cols == 0; }
  template <typename T> T &at(int y, int x) { return reinterpret_cast<T *>(data)[(size_t)y * cols + x]; }
  template <typename T> const T &at(int y, int x) const { return reinterpret_cast<const T *>(data)[(size_t)y * cols + x]; }
  // may be called with dst == *this (main.cpp:301 converts a temporary into a fresh Mat)
  void convertTo(Mat &dst, int rtype) const {
    Mat out(rows, cols, rtype);
    const size_t n = (size_t)rows * cols;
    if (type_ == CV_8UC1 && rtype == CV_32FC1) {
      float *o = reinterpret_cast<float *>(out.data);
      for (size_t i = 0; i < n; i++) o[i] = (float)data[i];
    } else if (type_ == CV_32FC1 && rtype == CV_8UC1) {
      const float *s = reinterpret_cast<const float *>(data);
      for (size_t i = 0; i < n; i++) {
        const float v = std::nearbyint(s[i]);
        out.data[i] = (uchar)(v < 0.f ? 0.f : (v > 255.f ? 255.f : v));
      }
    } else if (n) {
      memcpy(out.data, data, n * elemSize());
    }
    dst = out;
  }

 private:
  int type_;
  std::shared_ptr<std::vector<uchar> > buf_;
};

}  // namespace cv
#endif
