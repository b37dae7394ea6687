// cusift_b200 — reference-compatible brute-force matcher API
// (danielsuo/cuSIFT extras/matching.h:10-39).
#ifndef CUSIFT_B200_MATCHING_H
#define CUSIFT_B200_MATCHING_H

#include <vector>

#include "cuSIFT.h"
#include "cutils.h"

using namespace std;

typedef enum { MatchSiftDistanceDotProduct, MatchSiftDistanceL2 } MatchSiftDistance;

typedef enum { MatchType2D, MatchType3D } MatchType;

typedef struct {
  SiftPoint *pt1;   // into data1.h_data
  SiftPoint *pt2;   // into data2.h_data
  float score;
  float ambiguity;
  float error;
} SiftMatch;

// For every point of data1: best and second-best point of data2.  Fills score,
// ambiguity, match, match_xpos, match_ypos of data1 (device and host) and returns
// heap-allocated SiftMatch records (caller deletes) for the points passing
// score < scoreThreshold^2 && ambiguity < ambiguityThreshold^2.
vector<SiftMatch *> MatchSiftData(SiftData &data1, SiftData &data2,
                                  MatchSiftDistance distance = MatchSiftDistanceL2, float scoreThreshold = 999.0,
                                  float ambiguityThreshold = 1.0, MatchType type = MatchType2D);

#endif
