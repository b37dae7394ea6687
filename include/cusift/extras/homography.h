// cusift_b200 — reference-compatible RANSAC homography API
// (danielsuo/cuSIFT extras/homography.h:8; ImproveHomography is declared ad hoc by
// main.cpp:19 and defined in extras/homography.cu:280).  No OpenCV dependency.
#ifndef CUSIFT_B200_HOMOGRAPHY_H
#define CUSIFT_B200_HOMOGRAPHY_H

#include "cuSIFT.h"
#include "cutils.h"

// RANSAC over the matches stored in `data` (after MatchSiftData).  Samples are
// drawn with libc rand() exactly like the reference.  Returns elapsed ms.
double FindHomography(SiftData &data, float *homography, int *numMatches, int numLoops = 1000,
                      float minScore = 0.85f, float maxAmbiguity = 0.95f, float thresh = 5.0f);

// Iteratively re-weighted least squares on the host copy; returns the number of
// points with error < thresh^2 and writes match_error per point.
int ImproveHomography(SiftData &data, float *homography, int numLoops, float minScore, float maxAmbiguity,
                      float thresh);

#endif
