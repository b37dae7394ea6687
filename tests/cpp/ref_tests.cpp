// The reference's own tests, restated without gtest / OpenCV / VLFeat on top of the
// reference-compatible headers (include/cusift/) and the harness compatibility pack
// (include/cusift/extras/debug.h):
//   Matching.MatchingTest        test/test.cpp:25-46   326 MATLAB nearest-neighbour pairs
//   Matching.MatchingRatioTest   test/test.cpp:48-56   exactly 340 ratio-test matches
//   Detector.DetectorCUSIFTTest  test/detector.cpp:18-88   count + "found" test against cusift1_check
//   RigidTransform.RANSACWithIndices / RANSACWithRandom  test/test.cpp:58-127  (the reference only prints its result
//                                next to the MATLAB Rt stored in the file; here the two are compared)
// usage: csb_ref_tests <golden dir> <gray1 640x480 float32 raw>; prints one JSON line, exit code =
// number of failed expectations.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cuImage.h"
#include "cuSIFT.h"
#include "extras/debug.h"
#include "extras/homography.h"
#include "extras/matching.h"
#include "extras/rigidTransform.h"

static int g_fail = 0;
#define EXPECT(cond)                                                       \
  do {                                                                     \
    if (!(cond)) {                                                         \
      g_fail++;                                                            \
      fprintf(stderr, "EXPECT failed at line %d: %s\n", __LINE__, #cond);  \
    }                                                                      \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <golden dir> <gray1.f32>\n", argv[0]);
    return 100;
  }
  const std::string dir = argv[1];
  InitCuda(0);

  // ---- Matching.MatchingTest ----
  int nn_checked = 0, nn_equal = 0;
  {
    SiftData s1(1024, true, true), s2(1024, true, true);   // HEAD's default ctor allocates nothing (cuSIFT.cu:13-33)
    ReadVLFeatSiftData(s1, (dir + "/sift1.bin").c_str());
    ReadVLFeatSiftData(s2, (dir + "/sift2.bin").c_str());
    int n = ReadMATLABMatchIndices((dir + "/match_indices1_2.bin").c_str());
    std::vector<uint32_t> ii(n), jj(n);
    n = ReadMATLABMatchIndices((dir + "/match_indices1_2.bin").c_str(), ii.data(), jj.data());
    vector<SiftMatch *> matches = MatchSiftData(s1, s2, MatchSiftDistanceL2);
    EXPECT((int)matches.size() == s1.numPts);
    for (int i = 0; i < n; i++) {
      nn_checked++;
      const bool same = (int)jj[i] == matches[ii[i] - 1]->pt1->match + 1;
      nn_equal += same;
      EXPECT(same);
    }
    for (SiftMatch *m : matches) delete m;
  }
  // ---- Matching.MatchingRatioTest ----
  int ratio_matches = 0;
  {
    SiftData s1(1024, true, true), s2(1024, true, true);
    ReadVLFeatSiftData(s1, (dir + "/sift1.bin").c_str());
    ReadVLFeatSiftData(s2, (dir + "/sift2.bin").c_str());
    vector<SiftMatch *> matches = MatchSiftData(s1, s2, MatchSiftDistanceL2, 1000, 0.6);
    ratio_matches = (int)matches.size();
    EXPECT(ratio_matches == 340);
    for (SiftMatch *m : matches) delete m;
  }
  // ---- Detector.DetectorCUSIFTTest ----
  int det_pts = 0, det_found = 0, det_found_abs = 0, det_rows = 0;
  {
    const int w = 640, h = 480;
    std::vector<float> im((size_t)w * h);
    FILE *fp = fopen(argv[2], "rb");
    if (!fp || fread(im.data(), sizeof(float), im.size(), fp) != im.size()) {
      fprintf(stderr, "cannot read %s\n", argv[2]);
      return 101;
    }
    fclose(fp);
    SiftData sd(4096, true, true);
    sd.numOctaves = 6;
    sd.initBlur = 0.0f;
    sd.peakThresh = 0.1f;
    sd.edgeThresh = 10.0f;
    sd.lowestScale = 0.0f;
    sd.initSubsampling = 1.0f;
    sd.Extract(im.data(), w, h);
    det_pts = sd.numPts;
    // unsaturated run for the symmetric (|d| < 0.1) variant of the reference's one-sided "found" test
    SiftData big(32768, true, true);
    big.numOctaves = 6; big.initBlur = 0.0f; big.peakThresh = 0.1f; big.edgeThresh = 10.0f; big.lowestScale = 0.0f;
    big.initSubsampling = 1.0f;
    big.Extract(im.data(), w, h);
    fp = fopen((dir + "/cusift1_check.bin").c_str(), "rb");
    uint32_t n = 0;
    if (!fp || fread(&n, sizeof(uint32_t), 1, fp) != 1) return 102;
    EXPECT((int)n == sd.numPts);                          // detector.cpp:68 ASSERT_EQ(numPts, siftData->numPts)
    det_rows = (int)n;
    for (uint32_t i = 0; i < n; i++) {
      float d[4];
      if (fread(d, sizeof(float), 4, fp) != 4) return 103;
      bool found = false, found_abs = false;
      // detector.cpp:76-80, one-sided as written.  WHICH 4096 of the ~9500 keypoints survive maxPts is
      // decided by atomics (in the reference too), so the test is evaluated on the unsaturated set.
      for (int j = 0; j < big.numPts; j++) {
        const SiftPoint &p = big.h_data[j];
        if (p.coords2D[0] - d[0] < 0.1 && p.coords2D[1] - d[1] < 0.1 && p.scale - d[2] < 0.1 && p.orientation - d[3] < 0.1)
          found = true;
      }
      for (int j = 0; j < big.numPts && !found_abs; j++) {
        const SiftPoint &p = big.h_data[j];
        found_abs = fabsf(p.coords2D[0] - d[0]) < 0.1f && fabsf(p.coords2D[1] - d[1]) < 0.1f && fabsf(p.scale - d[2]) < 0.1f &&
                    fabsf(p.orientation - d[3]) < 0.1f;
      }
      det_found += found;
      det_found_abs += found_abs;
      EXPECT(found);
      EXPECT(found_abs);
    }
    fclose(fp);
  }
  // ---- RigidTransform.RANSACWithIndices / RANSACWithRandom ----
  int rt_inliers = 0, rt_inliers_random = 0;
  double rt_maxdiff = 0.0, rt_maxdiff_random = 0.0;
  {
    vector<int> indices;
    float Rt[12], Rt_test[12];
    vector<SiftMatch *> matches = ReadMATLABRANSAC((dir + "/rigid_ransac.bin").c_str(), indices, Rt);
    EXPECT(matches.size() == 120 && indices.size() == 30);
    const int numLoops = (int)indices.size() / 3;
    std::vector<float> h_coord(6 * matches.size());
    for (size_t i = 0; i < matches.size(); i++) {
      memcpy(&h_coord[6 * i], matches[i]->pt1->coords3D, sizeof(float) * 3);
      memcpy(&h_coord[6 * i + 3], matches[i]->pt2->coords3D, sizeof(float) * 3);
    }
    std::vector<char> h_inliers(matches.size());
    EstimateRigidTransformH(h_coord.data(), Rt_test, &rt_inliers, numLoops, (int)matches.size(), 0.05f * 0.05f,
                            RigidTransformType3D, indices.data(), h_inliers.data());
    for (int i = 0; i < 12; i++) rt_maxdiff = fmax(rt_maxdiff, fabs((double)Rt_test[i] - (double)Rt[i]));
    EXPECT(rt_inliers == 114);
    EXPECT(rt_maxdiff < 1e-5);
    EstimateRigidTransform(matches, Rt_test, &rt_inliers_random, 4096, 0.05f, RigidTransformType3D);   // test.cpp:112-127
    for (int i = 0; i < 12; i++) rt_maxdiff_random = fmax(rt_maxdiff_random, fabs((double)Rt_test[i] - (double)Rt[i]));
    EXPECT(rt_inliers_random >= 114);
    EXPECT(rt_maxdiff_random < 2e-2);   // a larger consensus set (116 of 120) than the 10 stored loops find: its refit differs slightly
    for (SiftMatch *m : matches) {
      delete m->pt1;
      delete m->pt2;
      delete m;
    }
  }
  printf("{\"rt_inliers\": %d, \"rt_maxdiff\": %.3g, \"rt_inliers_random\": %d, \"rt_maxdiff_random\": %.3g, ", rt_inliers,
         rt_maxdiff, rt_inliers_random, rt_maxdiff_random);
  printf("\"nn_checked\": %d, \"nn_equal\": %d, \"ratio_matches\": %d, \"det_pts\": %d, \"det_rows\": %d, \"det_found\": %d, "
         "\"det_found_abs\": %d, \"failed\": %d}\n",
         nn_checked, nn_equal, ratio_matches, det_pts, det_rows, det_found, det_found_abs, g_fail);
  return g_fail > 255 ? 255 : g_fail;
}
