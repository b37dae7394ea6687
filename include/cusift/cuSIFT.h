// cusift_b200 — reference-compatible SIFT API.
// SiftPoint (588 B) and SiftData (56 B) keep the exact field order of
// danielsuo/cuSIFT cuSIFT.h:10-30 and :32-74 because callers read them directly
// (main.cpp:343,358-366,400-446; test/detector.cpp:55-60).  Both API generations
// of the reference are provided: the class methods of HEAD (cuSIFT.h:57-73) and
// the legacy free functions main.cpp is written against (bodies preserved as
// comments in cuSIFT.cu:122-134,272-303).  Everything forwards to the C ABI in
// cusift_b200.h; there is no CPU fallback.
#ifndef CUSIFT_B200_CUSIFT_H
#define CUSIFT_B200_CUSIFT_H

#include <iostream>

#include "cuImage.h"
#include "cutils.h"

using namespace std;

class SiftPoint {
public:
  float coords2D[2];
  float scale;
  float sharpness;
  float edgeness;
  float orientation;   // degrees
  float score;
  float ambiguity;
  int match;
  float match_xpos;
  float match_ypos;
  float match_error;
  float subsampling;
  float empty[3];
  float data[128];
  float coords3D[3];
};

class SiftData {
public:
  int numPts;   // points available
  int maxPts;   // points allocated
  SiftPoint *h_data;   // host copy   (malloc'ed, page-locked while owned by this object)
  SiftPoint *d_data;   // device copy (cudaMalloc'ed)

  // extraction parameters — NOT set by the constructor (as in the reference)
  int numOctaves;
  int numScales;           // ignored: 5 scales per octave are compiled in (NUM_SCALES)
  double initBlur;
  float initSubsampling;   // ignored (the reference never reads it either)
  float peakThresh;
  float edgeThresh;
  float lowestScale;

  SiftData(int maxPts = 1024, bool host = false, bool dev = false);
  ~SiftData();

  void Synchronize();   // device -> host copy of numPts points
  // Uploads the dense host frame and extracts with the parameter fields above.
  void Extract(float *im, int width, int height, float subsampling = 1.0f);
  double ConvertSiftToRootSift();
};

// 2x downsample with the reference's 5x5 kernel; `variance` must be 0.5f (the only
// value the reference uses, cuSIFT.cu:185).
double ScaleDown(cuImage &res, cuImage &src, float variance);

// ---- legacy free-function API (main.cpp:99-103,173-174,324-328,348-349) ----
void InitSiftData(SiftData &data, int num = 1024, bool host = false, bool dev = false);
void FreeSiftData(SiftData &data);
// The frame must already be on the device (img.HostToDevice(), main.cpp:317-318).
// edgeThresh is fixed at 10.0f, the only value the reference ever uses.
void ExtractSift(SiftData &siftData, cuImage &img, int numOctaves, double initBlur, float thresh,
                 float lowestScale = 0.0f, float subsampling = 1.0f);
void ExtractRootSift(SiftData &siftData, cuImage &img, int numOctaves, double initBlur, float thresh,
                     float lowestScale = 0.0f, float subsampling = 1.0f);

#endif
