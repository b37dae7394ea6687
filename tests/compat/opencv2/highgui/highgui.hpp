// Harness compatibility pack: cv::imread / cv::imwrite for PGM (P5) and raw float32 files.  See core/core.hpp.
#ifndef CSB_COMPAT_OPENCV_HIGHGUI_HPP
#define CSB_COMPAT_OPENCV_HIGHGUI_HPP

#include "opencv2/core/core.hpp"

namespace cv {

enum { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };

// flags 0 = grey.  PGM: 8-bit Mat (like OpenCV).  "<name>.<w>x<h>.f32": CV_32FC1 Mat.
inline Mat imread(const std::string &path, int /*flags*/ = 1) {
  Mat m;
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) {
    fprintf(stderr, "imread (compat): cannot open %s\n", path.c_str());
    return m;
  }
  int w = 0, h = 0;
  if (path.size() > 4 && path.compare(path.size() - 4, 4, ".f32") == 0) {
    const size_t dot = path.rfind('.', path.size() - 5);
    if (dot != std::string::npos && sscanf(path.c_str() + dot + 1, "%dx%d", &w, &h) == 2 && w > 0 && h > 0) {
      m.create(h, w, CV_32FC1);
      if (fread(m.data, 4, (size_t)w * h, fp) != (size_t)w * h) m = Mat();
    }
  } else {
    char magic[3] = {0, 0, 0};
    int maxv = 0;
    if (fscanf(fp, "%2s", magic) == 1 && strcmp(magic, "P5") == 0) {
      int c = fgetc(fp);
      while (c == '#' || c == '\n' || c == ' ' || c == '\r' || c == '\t') {      // comments / whitespace
        if (c == '#') while (c != '\n' && c != EOF) c = fgetc(fp);
        c = fgetc(fp);
      }
      ungetc(c, fp);
      if (fscanf(fp, "%d %d %d", &w, &h, &maxv) == 3 && maxv == 255) {
        fgetc(fp);                                                              // the single whitespace before the raster
        m.create(h, w, CV_8UC1);
        if (fread(m.data, 1, (size_t)w * h, fp) != (size_t)w * h) m = Mat();
      }
    } else {
      fprintf(stderr, "imread (compat): %s is neither PGM (P5) nor <name>.<w>x<h>.f32; JPEG/PNG need the real OpenCV\n", path.c_str());
    }
  }
  fclose(fp);
  return m;
}

// float images are saturated to 8 bits like OpenCV does; always writes binary PGM.
inline bool imwrite(const std::string &path, const Mat &img) {
  Mat u8;
  if (img.type() == CV_32FC1) img.convertTo(u8, CV_8UC1);
  else u8 = img;
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  fprintf(fp, "P5\n%d %d\n255\n", u8.cols, u8.rows);
  fwrite(u8.data, 1, (size_t)u8.rows * u8.cols, fp);
  fclose(fp);
  return true;
}

}  // namespace cv
#endif
