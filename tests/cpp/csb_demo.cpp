// C++ consumer written against the reference-style API exactly the way the
// reference's own main.cpp:285-350 (demo) uses it — cuImage::Allocate / HostToDevice,
// InitSiftData, ExtractSift, MatchSiftData, FindHomography, ImproveHomography,
// FreeSiftData — minus OpenCV (frames come from raw float32 files).  It also runs the
// HEAD-generation call (SiftData::Extract, test/detector.cpp:41-49) on the first frame.
// Used by tests/test_gpu_cxx_api.py; prints one JSON line.
//
//   csb_demo left.f32 right.f32 width height [devNum]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuImage.h"
#include "cuSIFT.h"
#include "extras/homography.h"
#include "extras/matching.h"

static std::vector<float> readRaw(const char *path, int w, int h) {
  std::vector<float> v((size_t)w * h);
  FILE *fp = fopen(path, "rb");
  if (!fp || fread(v.data(), sizeof(float), v.size(), fp) != v.size()) {
    fprintf(stderr, "cannot read %s\n", path);
    exit(2);
  }
  fclose(fp);
  return v;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: csb_demo left.f32 right.f32 width height [devNum]\n");
    return 1;
  }
  const int w = atoi(argv[3]), h = atoi(argv[4]);
  const int devNum = argc > 5 ? atoi(argv[5]) : 0;
  std::vector<float> limg = readRaw(argv[1], w, h), rimg = readRaw(argv[2], w, h);

  InitCuda(devNum);
  CudaImage img1, img2;                       // CudaImage == cuImage
  img1.Allocate(w, h, iAlignUp(w, 128), false, NULL, limg.data());
  img2.Allocate(w, h, iAlignUp(w, 128), false, NULL, rimg.data());
  img1.HostToDevice();
  img2.HostToDevice();

  SiftData siftData1, siftData2;
  float initBlur = 0.0f;
  float thresh = 0.1f;
  InitSiftData(siftData1, 4096, true, true);
  InitSiftData(siftData2, 4096, true, true);
  ExtractSift(siftData1, img1, 6, initBlur, thresh, 0.0f);
  ExtractSift(siftData2, img2, 6, initBlur, thresh, 0.0f);

  vector<SiftMatch *> matches = MatchSiftData(siftData1, siftData2);
  const size_t numAll = matches.size();
  for (SiftMatch *m : matches) delete m;
  matches = MatchSiftData(siftData1, siftData2, MatchSiftDistanceL2, 1000, 0.6);
  const size_t numRatio = matches.size();
  bool ptrsOk = true;
  for (SiftMatch *m : matches) {
    ptrsOk = ptrsOk && m->pt1 >= siftData1.h_data && m->pt1 < siftData1.h_data + siftData1.numPts &&
             m->pt2 == &siftData2.h_data[m->pt1->match];
    delete m;
  }

  float homography[9];
  int numMatches = 0;
  FindHomography(siftData1, homography, &numMatches, 10000, 0.00f, 0.80f, 5.0);
  int numFit = ImproveHomography(siftData1, homography, 5, 0.00f, 0.80f, 3.0);

  // HEAD API generation on the same frame, unsaturated
  SiftData *sd = new SiftData(32768, true, true);
  sd->numOctaves = 6;
  sd->initBlur = initBlur;
  sd->peakThresh = thresh;
  sd->edgeThresh = 10.0f;
  sd->lowestScale = 0.0f;
  sd->initSubsampling = 1.0f;
  sd->Extract(limg.data(), w, h);
  const int headPts = sd->numPts;
  double norm0 = 0.0;
  for (int i = 0; i < 128; i++) norm0 += sd->h_data[0].data[i] * sd->h_data[0].data[i];
  sd->ConvertSiftToRootSift();
  sd->Synchronize();
  double root0 = 0.0;
  for (int i = 0; i < 128; i++) root0 += sd->h_data[0].data[i] * sd->h_data[0].data[i];
  delete sd;

  // ScaleDown free function
  cuImage half;
  half.Allocate(w / 2, h / 2, iAlignUp(w / 2, 128), true);
  ScaleDown(half, img1, 0.5f);
  half.DeviceToHost();
  const float half00 = half.h_data[0];

  printf("{\"numPts1\": %d, \"numPts2\": %d, \"matches_all\": %zu, \"matches_ratio\": %zu, \"ptrs_ok\": %d, "
         "\"numMatches\": %d, \"numFit\": %d, \"H\": [%g, %g, %g, %g, %g, %g, %g, %g, %g], \"head_pts\": %d, "
         "\"desc_norm2\": %.6f, \"rootsift_norm2\": %.6f, \"half00\": %.6f}\n",
         siftData1.numPts, siftData2.numPts, numAll, numRatio, ptrsOk ? 1 : 0, numMatches, numFit, homography[0],
         homography[1], homography[2], homography[3], homography[4], homography[5], homography[6], homography[7],
         homography[8], headPts, norm0, root0, half00);

  FreeSiftData(siftData1);
  FreeSiftData(siftData2);
  return 0;
}
