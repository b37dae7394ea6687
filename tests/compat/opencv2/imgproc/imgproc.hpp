// Harness compatibility pack: cv::GaussianBlur for CV_32FC1 (main.cpp:308-309 uses Size(3,3), sigma 0.5).
// Kernel = cv::getGaussianKernel(k, sigma, CV_32F); separable; BORDER_REFLECT_101; the two passes are evaluated as
// OpenCV 4's SIMD body does for ksize 3 (row: fma(c, k0, (l + r) k1); column: fma(u + d, k1, c k0)), which is what
// csb_ingest_u8's device pre-blur restates as well.  See core/core.hpp.
#ifndef CSB_COMPAT_OPENCV_IMGPROC_HPP
#define CSB_COMPAT_OPENCV_IMGPROC_HPP

#include "opencv2/core/core.hpp"

namespace cv {

inline int compat_reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

inline void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY = 0) {
  (void)sigmaY;
  const int k = ksize.width, r = k / 2, w = src.cols, h = src.rows;
  if (src.type() != CV_32FC1 || k != ksize.height || (k & 1) == 0 || src.empty()) {
    fprintf(stderr, "GaussianBlur (compat): only square odd kernels on CV_32FC1 images\n");
    dst = src;
    return;
  }
  std::vector<float> taps(k);
  {
    const double scale2x = -0.5 / (sigmaX * sigmaX);
    double sum = 0.0;
    for (int i = 0; i < k; i++) { taps[i] = (float)std::exp(scale2x * (i - r) * (i - r)); sum += taps[i]; }
    for (int i = 0; i < k; i++) taps[i] = (float)(taps[i] * (1.0 / sum));
  }
  const float *s = reinterpret_cast<const float *>(src.data);
  std::vector<float> tmp((size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const float *row = s + (size_t)y * w;
      float acc = 0.0f;                                        // symmetric pairs first, centre last via fma
      for (int i = 1; i <= r; i++) acc = std::fmaf(row[compat_reflect101(x - i, w)] + row[compat_reflect101(x + i, w)], taps[r + i], acc);
      tmp[(size_t)y * w + x] = std::fmaf(row[x], taps[r], acc);
    }
  Mat out(h, w, CV_32FC1);
  float *o = reinterpret_cast<float *>(out.data);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float acc = tmp[(size_t)y * w + x] * taps[r];
      for (int i = 1; i <= r; i++)
        acc = std::fmaf(tmp[(size_t)compat_reflect101(y - i, h) * w + x] + tmp[(size_t)compat_reflect101(y + i, h) * w + x], taps[r + i], acc);
      o[(size_t)y * w + x] = acc;
    }
  dst = out;
}

}  // namespace cv
#endif
