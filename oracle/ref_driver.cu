// TEST INFRASTRUCTURE ONLY (oracle/): command-line driver around the UNMODIFIED
// reference (danielsuo/cuSIFT) compiled from the sources where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/ref_driver.
//
// It exists so that the GPU parity tests and the `--impl reference` bench arm can
// run the reference's own kernels on the same raw inputs as the product library
// and compare files.  Only tests/, __graft_entry__.smoke() and bench.py execute
// it; the product library never links or calls anything here.
//
// This translation unit textually includes the reference's cuSIFT.cu (which in
// turn includes cuSIFT_D.cu, cuSIFT.cu:11) so that the `stages`/`extract_safe`
// commands can reach the file-scope __device__ counters (cuSIFT_D.cu:16-18) the
// same way SiftData::Extract does (cuSIFT.cu:69-70).
//
// File formats (little endian):
//   image  : raw float32[h][w]
//   .sift  : u32 n, then n x SiftPoint (588 B, cuSIFT.h:10-30)
//   stages : i32 nOct, nOct x {i32 w, i32 h}, then per octave base f32[h][w],
//            dog f32[7][h][w]
#include "cuSIFT.cu"  // reference: class impl + kernels (cuSIFT.cu, cuSIFT_D.cu)

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "extras/homography.h"
#include "extras/matching.h"
#include "extras/rigidTransform.h"

int ImproveHomography(SiftData &data, float *homography, int numLoops, float minScore,
                      float maxAmbiguity, float thresh);  // extras/homography.cu:280

static std::vector<float> readImage(const char *path, int w, int h) {
  std::vector<float> img((size_t)w * h);
  FILE *fp = fopen(path, "rb");
  if (!fp) { fprintf(stderr, "ref_driver: cannot open %s\n", path); exit(2); }
  size_t got = fread(img.data(), sizeof(float), img.size(), fp);
  fclose(fp);
  if (got != img.size()) { fprintf(stderr, "ref_driver: short read %s\n", path); exit(2); }
  return img;
}

static void writeSift(const char *path, const SiftPoint *pts, int n) {
  FILE *fp = fopen(path, "wb");
  if (!fp) { fprintf(stderr, "ref_driver: cannot write %s\n", path); exit(2); }
  uint32_t nn = (uint32_t)n;
  fwrite(&nn, sizeof(nn), 1, fp);
  fwrite(pts, sizeof(SiftPoint), n, fp);
  fclose(fp);
}

static void loadSift(const char *path, SiftData &d) {
  FILE *fp = fopen(path, "rb");
  if (!fp) { fprintf(stderr, "ref_driver: cannot open %s\n", path); exit(2); }
  uint32_t n = 0;
  if (fread(&n, sizeof(n), 1, fp) != 1) exit(2);
  if ((int)n > d.maxPts) { fprintf(stderr, "ref_driver: %u > maxPts\n", n); exit(2); }
  if (fread(d.h_data, sizeof(SiftPoint), n, fp) != n) exit(2);
  fclose(fp);
  d.numPts = (int)n;
  safeCall(cudaMemcpy(d.d_data, d.h_data, sizeof(SiftPoint) * n, cudaMemcpyHostToDevice));
}

static void setParams(SiftData &sd, char **a) {
  sd.numOctaves = atoi(a[0]);
  sd.numScales = NUM_SCALES;
  sd.initBlur = atof(a[1]);
  sd.initSubsampling = 1.0f;
  sd.peakThresh = (float)atof(a[2]);
  sd.edgeThresh = (float)atof(a[3]);
  sd.lowestScale = (float)atof(a[4]);
}

// Width the reference's ScaleDown_D actually stores per row (cuSIFT.cu:346,
// cuSIFT_D.cu:64,123): 80 columns per 160-wide block regardless of the
// destination width.  Used to size race-free destinations in the *_safe paths.
static int scaleDownStoreCols(int srcW) { return (SCALEDOWN_W / 2) * iDivUp(srcW, SCALEDOWN_W); }
static int scaleDownStoreRows(int srcH) { return (SCALEDOWN_H / 2) * iDivUp(srcH, SCALEDOWN_H); }

struct SafeSub {
  float *d;
  int w, h, pitch;
};

// Allocate a sub-image big enough for every (unguarded) store ScaleDown_D makes.
static SafeSub allocSafeSub(int srcW, int srcH) {
  SafeSub s;
  s.w = srcW / 2;
  s.h = srcH / 2;
  s.pitch = iAlignUp(std::max(s.w, scaleDownStoreCols(srcW)), 128);
  int rows = std::max(s.h, scaleDownStoreRows(srcH)) + 1;
  safeCall(cudaMalloc((void **)&s.d, sizeof(float) * (size_t)s.pitch * rows));
  safeCall(cudaMemset(s.d, 0, sizeof(float) * (size_t)s.pitch * rows));
  return s;
}

static cudaTextureObject_t makeTexture(cuImage &img) {  // same descriptor as cuSIFT.cu:218-236
  cudaResourceDesc resDesc;
  memset(&resDesc, 0, sizeof(resDesc));
  resDesc.resType = cudaResourceTypePitch2D;
  resDesc.res.pitch2D.devPtr = img.d_data;
  resDesc.res.pitch2D.width = img.width;
  resDesc.res.pitch2D.height = img.height;
  resDesc.res.pitch2D.pitchInBytes = img.pitch * sizeof(float);
  resDesc.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  cudaTextureDesc texDesc;
  memset(&texDesc, 0, sizeof(texDesc));
  texDesc.addressMode[0] = cudaAddressModeClamp;
  texDesc.addressMode[1] = cudaAddressModeClamp;
  texDesc.filterMode = cudaFilterModeLinear;
  texDesc.readMode = cudaReadModeElementType;
  texDesc.normalizedCoords = 0;
  cudaTextureObject_t tex = 0;
  safeCall(cudaCreateTextureObject(&tex, &resDesc, &texDesc, NULL));
  return tex;
}

static std::vector<float> download(const float *d, int w, int h, int pitch) {
  std::vector<float> out((size_t)w * h);
  safeCall(cudaMemcpy2D(out.data(), sizeof(float) * w, d, sizeof(float) * pitch, sizeof(float) * w, h,
                        cudaMemcpyDeviceToHost));
  return out;
}

// stages: run the reference's ScaleDown (cuSIFT.cu:313) and LaplaceMulti
// (cuSIFT.cu:399) octave by octave and dump every intermediate image.
static int cmdStages(int argc, char **argv) {
  if (argc < 6) return 1;
  int w = atoi(argv[1]), h = atoi(argv[2]), nOct = atoi(argv[3]);
  double initBlur = atof(argv[4]);
  std::vector<float> h_img = readImage(argv[0], w, h);
  FILE *fp = fopen(argv[5], "wb");
  if (!fp) return 2;
  fwrite(&nOct, sizeof(int), 1, fp);
  { int ww = w, hh = h; for (int o = 0; o < nOct; o++) { fwrite(&ww, 4, 1, fp); fwrite(&hh, 4, 1, fp); ww /= 2; hh /= 2; } }

  cuImage *img = new cuImage();
  img->Allocate(w, h, iAlignUp(w, 128), false, NULL, h_img.data());
  img->HostToDevice();
  SiftData sd(16, false, false);
  std::vector<float *> owned;
  for (int o = 0; o < nOct; o++) {
    int cw = img->width, ch = img->height, cp = img->pitch;
    std::vector<float> base = download(img->d_data, cw, ch, cp);
    fwrite(base.data(), sizeof(float), base.size(), fp);
    float *d_dog = NULL;
    safeCall(cudaMalloc((void **)&d_dog, sizeof(float) * (size_t)cp * ch * (NUM_SCALES + 3)));
    cuImage diffImg[NUM_SCALES + 3];
    for (int i = 0; i < NUM_SCALES + 2; i++) diffImg[i].Allocate(cw, ch, cp, false, d_dog + (size_t)i * cp * ch);
    cudaTextureObject_t tex = makeTexture(*img);
    float baseBlur = pow(2.0f, -1.0f / NUM_SCALES);   // cuSIFT.cu:239
    float diffScale = pow(2.0f, 1.0f / NUM_SCALES);   // cuSIFT.cu:240
    sd.LaplaceMulti(tex, diffImg, baseBlur, diffScale, initBlur);
    safeCall(cudaDeviceSynchronize());
    for (int i = 0; i < NUM_SCALES + 2; i++) {
      std::vector<float> pl = download(d_dog + (size_t)i * cp * ch, cw, ch, cp);
      fwrite(pl.data(), sizeof(float), pl.size(), fp);
    }
    safeCall(cudaDestroyTextureObject(tex));
    safeCall(cudaFree(d_dog));
    if (o + 1 < nOct) {
      SafeSub s = allocSafeSub(cw, ch);
      owned.push_back(s.d);
      cuImage *sub = new cuImage();
      sub->Allocate(s.w, s.h, s.pitch, false, s.d);
      ScaleDown(*sub, *img, 0.5f);                                           // cuSIFT.cu:185
      safeCall(cudaDeviceSynchronize());
      initBlur = (float)sqrt(initBlur * initBlur + 0.5f * 0.5f) / 2.0f;     // cuSIFT.cu:188
      delete img;
      img = sub;
    }
  }
  delete img;
  for (float *p : owned) cudaFree(p);
  fclose(fp);
  return 0;
}

// Same recursion as SiftData::ExtractSiftLoop (cuSIFT.cu:175-202) but every
// sub-image is allocated large enough for ScaleDown_D's unguarded stores, so the
// result does not depend on what lies behind a too-small destination (the
// 240->120 px step of a 1080p pyramid otherwise races on columns 0..31).
static void safeLoop(SiftData &sd, cuImage &img, int numOctaves, double initBlur, float subsampling,
                     float *memoryTmp) {
  if (numOctaves > 1) {
    SafeSub s = allocSafeSub(img.width, img.height);
    cuImage sub;
    sub.Allocate(s.w, s.h, s.pitch, false, s.d);
    ScaleDown(sub, img, 0.5f);
    float totInitBlur = (float)sqrt(initBlur * initBlur + 0.5f * 0.5f) / 2.0f;
    safeLoop(sd, sub, numOctaves - 1, totInitBlur, subsampling * 2.0f, memoryTmp);
    safeCall(cudaDeviceSynchronize());
    safeCall(cudaFree(s.d));
  }
  if (sd.lowestScale < subsampling * 2.0f) sd.ExtractSiftOctave(img, initBlur, subsampling, memoryTmp);
}

static void extractSafe(SiftData &sd, float *h_img, int w, int h) {
  cuImage img;
  img.Allocate(w, h, iAlignUp(w, 128), false, NULL, h_img);
  img.HostToDevice();
  int zero = 0;
  safeCall(cudaMemcpyToSymbol(d_PointCounter, &zero, sizeof(int)));        // cuSIFT.cu:69
  safeCall(cudaMemcpyToSymbol(d_MaxNumPoints, &sd.maxPts, sizeof(int)));   // cuSIFT.cu:70
  float *memoryTmp = NULL;
  size_t planes = (size_t)(NUM_SCALES + 3) * img.pitch * (h + 1);
  safeCall(cudaMalloc((void **)&memoryTmp, sizeof(float) * planes));
  safeLoop(sd, img, sd.numOctaves, sd.initBlur, 1.0f, memoryTmp);
  safeCall(cudaMemcpyFromSymbol(&sd.numPts, d_PointCounter, sizeof(int)));
  sd.numPts = (sd.numPts < sd.maxPts ? sd.numPts : sd.maxPts);
  safeCall(cudaFree(memoryTmp));
  sd.Synchronize();
}

// extract / extract_safe <img> w h nOct initBlur peak edge lowest maxPts rootsift out.sift
static int cmdExtract(int argc, char **argv, bool safe) {
  if (argc < 11) return 1;
  int w = atoi(argv[1]), h = atoi(argv[2]);
  std::vector<float> h_img = readImage(argv[0], w, h);
  int maxPts = atoi(argv[8]);
  int rootsift = atoi(argv[9]);
  SiftData sd(maxPts, true, true);
  setParams(sd, argv + 3);
  if (safe) extractSafe(sd, h_img.data(), w, h);
  else sd.Extract(h_img.data(), w, h);
  if (rootsift) {                       // legacy ExtractRootSift body, cuSIFT.cu:122-134
    sd.ConvertSiftToRootSift();
    sd.Synchronize();
  }
  writeSift(argv[10], sd.h_data, sd.numPts);
  fprintf(stderr, "ref_driver: %d keypoints\n", sd.numPts);
  return 0;
}

// match a.sift b.sift distance(0 dot,1 L2) scoreThr ambThr out.sift
static int cmdMatch(int argc, char **argv) {
  if (argc < 6) return 1;
  SiftData a(1 << 17, true, true), b(1 << 17, true, true);
  loadSift(argv[0], a);
  loadSift(argv[1], b);
  MatchSiftDistance dist = atoi(argv[2]) ? MatchSiftDistanceL2 : MatchSiftDistanceDotProduct;
  std::vector<SiftMatch *> m = MatchSiftData(a, b, dist, (float)atof(argv[3]), (float)atof(argv[4]));
  printf("{\"matches\": %zu}\n", m.size());
  for (SiftMatch *p : m) delete p;
  writeSift(argv[5], a.h_data, a.numPts);
  return 0;
}

// homography a.sift numLoops minScore maxAmb thresh improveLoops improveThresh out.txt
// a.sift must already carry match fields (output of `match`).
static int cmdHomography(int argc, char **argv) {
  if (argc < 8) return 1;
  SiftData a(1 << 17, true, true);
  loadSift(argv[0], a);
  float H[9];
  int numMatches = 0;
  FindHomography(a, H, &numMatches, atoi(argv[1]), (float)atof(argv[2]), (float)atof(argv[3]),
                 (float)atof(argv[4]));
  FILE *fp = fopen(argv[7], "w");
  fprintf(fp, "%d", numMatches);
  for (int i = 0; i < 9; i++) fprintf(fp, " %.9g", H[i]);
  fprintf(fp, "\n");
  int improveLoops = atoi(argv[5]);
  if (improveLoops > 0) {
    int numFit = ImproveHomography(a, H, improveLoops, (float)atof(argv[2]), (float)atof(argv[3]),
                                   (float)atof(argv[6]));
    fprintf(fp, "%d", numFit);
    for (int i = 0; i < 9; i++) fprintf(fp, " %.9g", H[i]);
    fprintf(fp, "\n");
  }
  fclose(fp);
  return 0;
}

// bench <img> w h nOct initBlur peak edge lowest maxPts warmup iters [safe]
// Times SiftData::Extract (host image in, SiftPoints in h_data out; cuSIFT.cu:61-120).
static int cmdBench(int argc, char **argv) {
  if (argc < 11) return 1;
  int w = atoi(argv[1]), h = atoi(argv[2]);
  std::vector<float> h_img = readImage(argv[0], w, h);
  int maxPts = atoi(argv[8]);
  int warmup = atoi(argv[9]), iters = atoi(argv[10]);
  bool safe = argc > 11 && atoi(argv[11]);
  SiftData sd(maxPts, true, true);
  setParams(sd, argv + 3);
  // silence the per-call "Total time incl memory" printf (cuSIFT.cu:117-119)
  FILE *saved = stdout;
  stdout = fopen("/dev/null", "w");
  for (int i = 0; i < warmup; i++) { if (safe) extractSafe(sd, h_img.data(), w, h); else sd.Extract(h_img.data(), w, h); }
  safeCall(cudaDeviceSynchronize());
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++) { if (safe) extractSafe(sd, h_img.data(), w, h); else sd.Extract(h_img.data(), w, h); }
  safeCall(cudaDeviceSynchronize());
  auto t1 = std::chrono::steady_clock::now();
  fclose(stdout);
  stdout = saved;
  double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  printf("{\"iters\": %d, \"total_ms\": %.4f, \"ms_per_frame\": %.5f, \"keypoints\": %d}\n", iters, ms, ms / iters,
         sd.numPts);
  return 0;
}

// benchmatch a.sift b.sift warmup iters : times MatchSiftData (L2 defaults, matching.cu:272)
static int cmdBenchMatch(int argc, char **argv) {
  if (argc < 4) return 1;
  SiftData a(1 << 17, true, true), b(1 << 17, true, true);
  loadSift(argv[0], a);
  loadSift(argv[1], b);
  int warmup = atoi(argv[2]), iters = atoi(argv[3]);
  for (int i = 0; i < warmup; i++) { auto m = MatchSiftData(a, b); for (auto *p : m) delete p; }
  safeCall(cudaDeviceSynchronize());
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; i++) { auto m = MatchSiftData(a, b); for (auto *p : m) delete p; }
  safeCall(cudaDeviceSynchronize());
  auto t1 = std::chrono::steady_clock::now();
  double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  printf("{\"iters\": %d, \"ms_per_pair\": %.5f, \"n1\": %d, \"n2\": %d}\n", iters, ms / iters, a.numPts, b.numPts);
  return 0;
}

// rigid coord.f32(numPts x 6) numPts indices.i32(numLoops x 3) numLoops thresh2 type(0|1) out.txt
// -> unmodified EstimateRigidTransformH (extras/rigidTransform.cu:387-520) with the given indices
static int cmdRigid(int argc, char **argv) {
  if (argc < 7) return 2;
  const int numPts = atoi(argv[1]), numLoops = atoi(argv[3]);
  const float thresh2 = (float)atof(argv[4]);
  const int type = atoi(argv[5]);
  std::vector<float> coord((size_t)6 * numPts);
  std::vector<int> idx((size_t)3 * numLoops);
  FILE *f = fopen(argv[0], "rb");
  if (!f || fread(coord.data(), 4, coord.size(), f) != coord.size()) return 3;
  fclose(f);
  f = fopen(argv[2], "rb");
  if (!f || fread(idx.data(), 4, idx.size(), f) != idx.size()) return 3;
  fclose(f);
  float Rt[12];
  int numInliers = 0;
  std::vector<char> inl(numPts, 0);
  EstimateRigidTransformH(coord.data(), Rt, &numInliers, numLoops, numPts, thresh2,
                          type ? RigidTransformType3D : RigidTransformType2D, idx.data(), inl.data());
  f = fopen(argv[6], "w");
  if (!f) return 4;
  fprintf(f, "%d\n", numInliers);
  for (int i = 0; i < 12; i++) fprintf(f, "%.9g\n", Rt[i]);
  for (int i = 0; i < numPts; i++) fprintf(f, "%d\n", (int)inl[i]);
  fclose(f);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: ref_driver {extract|extract_safe|stages|match|homography|rigid|bench|benchmatch} ...\n");
    return 1;
  }
  InitCuda(0);
  std::string cmd = argv[1];
  int rc = 1;
  if (cmd == "extract") rc = cmdExtract(argc - 2, argv + 2, false);
  else if (cmd == "extract_safe") rc = cmdExtract(argc - 2, argv + 2, true);
  else if (cmd == "stages") rc = cmdStages(argc - 2, argv + 2);
  else if (cmd == "match") rc = cmdMatch(argc - 2, argv + 2);
  else if (cmd == "homography") rc = cmdHomography(argc - 2, argv + 2);
  else if (cmd == "rigid") rc = cmdRigid(argc - 2, argv + 2);
  else if (cmd == "bench") rc = cmdBench(argc - 2, argv + 2);
  else if (cmd == "benchmatch") rc = cmdBenchMatch(argc - 2, argv + 2);
  if (rc == 1) fprintf(stderr, "ref_driver: bad arguments for %s\n", cmd.c_str());
  return rc;
}
