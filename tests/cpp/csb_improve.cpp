// Calls the drop-in ImproveHomography (include/cusift/extras/homography.h; reference: extras/homography.cu:271-337,
// declared ad hoc in main.cpp:19) on a SiftPoint file that already carries match fields, starting from a given H.
// Used by tests/test_gpu_parity.py to compare the shim with the unmodified reference (ref_driver homography).
//
//   csb_improve in.sift out.sift numLoops minScore maxAmbiguity thresh h0 .. h8     -> prints {"numFit": n, "H": [...]}
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuSIFT.h"
#include "extras/homography.h"

int main(int argc, char **argv) {
  if (argc < 16) {
    fprintf(stderr, "usage: csb_improve in.sift out.sift numLoops minScore maxAmbiguity thresh h0..h8\n");
    return 1;
  }
  FILE *fp = fopen(argv[1], "rb");
  uint32_t n = 0;
  if (!fp || fread(&n, sizeof(n), 1, fp) != 1) return 2;
  SiftData data;
  InitSiftData(data, (int)n, true, false);        // host data only: all the reference's ImproveHomography touches
  if (fread(data.h_data, sizeof(SiftPoint), n, fp) != n) return 2;
  fclose(fp);
  data.numPts = (int)n;
  float H[9];
  for (int i = 0; i < 9; i++) H[i] = (float)atof(argv[7 + i]);
  const int numFit = ImproveHomography(data, H, atoi(argv[3]), (float)atof(argv[4]), (float)atof(argv[5]), (float)atof(argv[6]));
  fp = fopen(argv[2], "wb");
  if (!fp) return 3;
  fwrite(&n, sizeof(n), 1, fp);
  fwrite(data.h_data, sizeof(SiftPoint), n, fp);
  fclose(fp);
  printf("{\"numFit\": %d, \"H\": [%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g]}\n", numFit, H[0], H[1], H[2], H[3],
         H[4], H[5], H[6], H[7], H[8]);
  FreeSiftData(data);
  return 0;
}
