// cusift_b200 — reference-compatible rigid-transform RANSAC API
// (danielsuo/cuSIFT extras/rigidTransform.h:17-34).
#ifndef CUSIFT_B200_RIGIDTRANSFORM_H
#define CUSIFT_B200_RIGIDTRANSFORM_H

#include <vector>

#include "cuSIFT.h"
#include "cutils.h"
#include "extras/matching.h"

typedef enum { RigidTransformType2D, RigidTransformType3D } RigidTransformType;

// h_coord: numPts x 6 (reference xyz, moving xyz).  Writes the 3 x 4 transform that maps moving-frame
// points into the reference frame and the inlier count of the best of numLoops hypotheses; h_indices:
// numLoops x 3 point indices (NULL: drawn on the device), h_inliers: numPts flags of the winner (may be NULL).
void EstimateRigidTransformH(const float *h_coord, float *Rt_relative, int *numInliers, int numLoops, int numPts,
                             float thresh2, RigidTransformType type = RigidTransformType2D, int *h_indices = NULL,
                             char *h_inliers = NULL);

// Same on SiftMatch records (pt1->coords3D = reference, pt2->coords3D = moving); thresh is a distance.
void EstimateRigidTransform(vector<SiftMatch *> matches, float *Rt_relative, int *numInliers, int numLoops, float thresh,
                            RigidTransformType type, int *h_indices = NULL, char *h_inliers = NULL);

#endif
