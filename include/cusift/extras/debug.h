// cusift_b200 — harness compatibility pack: the OpenCV-free part of the reference's
// debug helpers (danielsuo/cuSIFT extras/debug.h:18-34, extras/debug.cpp:27-454), which its
// tests (test/test.cpp, test/detector.cpp) use to load golden data.  The cv::Mat
// overloads (writeMatToFile, PrintMatchData, ReadMATLABMatchData(cv::Mat&, ...)) draw or
// convert with OpenCV and are not part of this library.
#ifndef CUSIFT_B200_DEBUG_H
#define CUSIFT_B200_DEBUG_H

#include <stdint.h>

#include <vector>

#include "cuSIFT.h"
#include "extras/matching.h"

using namespace std;

// human-readable dump of every point (debug.cpp:27-70); downloads d_data if there is no host copy
void PrintSiftData(SiftData &data);
// tab-separated x y match_x match_y ind ind2 per point (debug.cpp:97-118)
void PrintMatchSiftData(SiftData &siftData1, const char *filename, int imgw);

// vl_sift_tofile.m format: uint32 n | float[4n] (x, y, scale, orientation[rad]) | float[128n] (debug.cpp:120-165)
void ReadVLFeatSiftData(SiftData &siftData, const char *filename);
// uint32 n | uint32[n] i | uint32[n] j, 1-based (debug.cpp:167-181); returns n, fills the arrays when both are given
int ReadMATLABMatchIndices(const char *indices_filename, uint32_t *indices_i = NULL, uint32_t *indices_j = NULL);
// uint32 n | n x (double[3], double[3]) -> pt1/pt2->coords3D (debug.cpp:244-279); caller deletes pt1, pt2 and the match
vector<SiftMatch *> ReadMATLABMatchData(const char *filename);
// uint32 n | uint32[n] | uint32[n] | float[128n] | float[128n] -> descriptors of pt1/pt2 (debug.cpp:282-316)
vector<SiftMatch *> ReadMATLABMatchDataBeforeRANSAC(const char *filename);
// uint32 nMatches, nLoops | float[3n] | float[3n] | int[3 nLoops] (1-based) | float[12] Rt (debug.cpp:318-372)
vector<SiftMatch *> ReadMATLABRANSAC(const char *filename, vector<int> &indices, float *Rt);
// uint32 nPairs | nPairs x (uint32, uint32), 1-based -> flat 0-based vector (debug.cpp:374-392)
vector<int> ReadMATLABIndices(const char *filename);
// double[12] (debug.cpp:394-407)
void ReadMATLABRt(double *Rt_relative, const char *filename);

// appends host points to a SiftData, doubling its buffers as needed (debug.cpp:413-454)
void AddSiftData(SiftData &data, SiftPoint *h_data, int numPts);

#endif
