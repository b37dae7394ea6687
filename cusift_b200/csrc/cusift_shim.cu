// Reference-compatible C++ API (include/cusift/*.h) implemented on top of the C ABI
// (include/cusift_b200.h).  This is the host-side mirror of the reference's
// operator interface: same names, argument meaning, ownership and error
// behaviour (CUDA failure -> message on stderr + exit(-1), cutils.h:24-48; soft
// failures -> 0.0 / empty vector / identity H).
//
//   cuImage            <- cuImage.cu:11-117
//   SiftData, Extract  <- cuSIFT.cu:13-120, legacy wrappers cuSIFT.cu:122-134,272-303
//   ScaleDown          <- cuSIFT.cu:313-353
//   MatchSiftData      <- extras/matching.cu:272-402 (host part)
//   FindHomography     <- extras/homography.cu:191-278 (host part)
//   ImproveHomography  <- extras/homography.cu:280-346 (OpenCV replaced by a local 8x8 Cholesky)
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "cusift_b200.h"
#include "cusift/cuImage.h"
#include "cusift/cuSIFT.h"
#include "cusift/extras/homography.h"
#include "cusift/extras/matching.h"
#include "cusift/extras/rigidTransform.h"

static_assert(sizeof(SiftPoint) == sizeof(csb_sift_point), "SiftPoint layout");
static_assert(sizeof(SiftPoint) == 588, "SiftPoint must be 588 bytes (cuSIFT.h:10-30)");
static_assert(sizeof(SiftData) == 56, "SiftData must be 56 bytes (cuSIFT.h:32-74)");
static_assert(sizeof(cuImage) == 48, "cuImage must be 48 bytes (cuImage.h:8-26)");

namespace {

// One lazily created context per device, shared by every shim object of the process.
std::mutex &shim_mu() { static std::mutex mu; return mu; }
std::map<int, csb_ctx *> &shim_ctxs() { static std::map<int, csb_ctx *> ctxs; return ctxs; }
csb_ctx *shim_ctx_if_exists() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(shim_mu());
  auto it = shim_ctxs().find(dev);
  return it != shim_ctxs().end() ? it->second : nullptr;
}
csb_ctx *shim_ctx() {
  std::mutex &mu = shim_mu();
  std::map<int, csb_ctx *> &ctxs = shim_ctxs();
  int dev = 0;
  safeCall(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = ctxs.find(dev);
  if (it != ctxs.end()) return it->second;
  csb_ctx *ctx = nullptr;
  int rc = csb_ctx_create(dev, 2, &ctx);
  if (rc != 0 || !ctx) {
    fprintf(stderr, "cusift_b200: cannot create context on device %d (status %d)\n", dev, rc);
    exit(-1);
  }
  ctxs[dev] = ctx;
  return ctx;
}

void shim_check(csb_ctx *ctx, int rc, const char *what) {
  if (rc == 0) return;
  fprintf(stderr, "cusift_b200: %s failed (status %d): %s\n", what, rc, csb_last_error(ctx));
  exit(-1);
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ------------------------------------------------------------------ cuImage ---
cuImage::cuImage()
    : width(0), height(0), pitch(0), h_data(NULL), d_data(NULL), t_data(NULL), d_internalAlloc(false),
      h_internalAlloc(false) {}

cuImage::cuImage(int width_, int height_, float *h_data_, bool download)
    : width(0), height(0), pitch(0), h_data(NULL), d_data(NULL), t_data(NULL), d_internalAlloc(false),
      h_internalAlloc(false) {
  AllocateWithHostMemory(width_, height_, h_data_);
  if (download) HostToDevice();
}

cuImage::~cuImage() {
  // the context caches a texture object / TMA descriptor per frame it has seen: drop them before the memory goes away
  if (d_data != NULL)
    if (csb_ctx *ctx = shim_ctx_if_exists()) csb_forget_image(ctx, d_data);
  if (d_internalAlloc && d_data != NULL) safeCall(cudaFree(d_data));
  d_data = NULL;
  if (h_internalAlloc && h_data != NULL) free(h_data);
  h_data = NULL;
  t_data = NULL;
}

void cuImage::AllocateWithHostMemory(int width_, int height_, float *h_data_) {
  Allocate(width_, height_, iAlignUp(width_, 128), false, NULL, h_data_);
}

void cuImage::Allocate(int width_, int height_, int pitch_, bool withHost, float *d_data_, float *h_data_) {
  width = width_;
  height = height_;
  pitch = pitch_;
  d_data = d_data_;
  h_data = h_data_;
  t_data = NULL;
  if (d_data == NULL) {
    safeCall(cudaMalloc((void **)&d_data, sizeof(float) * (size_t)pitch * height));
    d_internalAlloc = true;
  }
  if (withHost && h_data == NULL) {
    h_data = (float *)malloc(sizeof(float) * (size_t)pitch * height);
    h_internalAlloc = true;
  }
}

double cuImage::HostToDevice() {
  const double t0 = now_ms();
  if (d_data != NULL && h_data != NULL)
    safeCall(cudaMemcpy2D(d_data, sizeof(float) * pitch, h_data, sizeof(float) * width, sizeof(float) * width, height,
                          cudaMemcpyHostToDevice));
  return now_ms() - t0;
}

double cuImage::DeviceToHost() {
  const double t0 = now_ms();
  if (d_data != NULL && h_data != NULL)
    safeCall(cudaMemcpy2D(h_data, sizeof(float) * width, d_data, sizeof(float) * pitch, sizeof(float) * width, height,
                          cudaMemcpyDeviceToHost));
  return now_ms() - t0;
}

// ----------------------------------------------------------------- SiftData ---
void InitSiftData(SiftData &data, int num, bool host, bool dev) {
  data.numPts = 0;
  data.maxPts = num;
  const size_t bytes = sizeof(SiftPoint) * (size_t)num;
  data.h_data = NULL;
  if (host) {
    data.h_data = (SiftPoint *)malloc(bytes);
    // page-lock so the GPU can deliver results straight into it; still free()-able
    if (data.h_data && cudaHostRegister(data.h_data, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) !=
                           cudaSuccess)
      cudaGetLastError();
  }
  data.d_data = NULL;
  if (dev) safeCall(cudaMalloc((void **)&data.d_data, bytes));
}

void FreeSiftData(SiftData &data) {
  if (data.d_data != NULL) safeCall(cudaFree(data.d_data));
  data.d_data = NULL;
  if (data.h_data != NULL) {
    if (cudaHostUnregister(data.h_data) != cudaSuccess) cudaGetLastError();
    free(data.h_data);
  }
  data.h_data = NULL;
  data.numPts = 0;
  data.maxPts = 0;
}

SiftData::SiftData(int maxPts_, bool host, bool dev) { InitSiftData(*this, maxPts_, host, dev); }
SiftData::~SiftData() { FreeSiftData(*this); }

void SiftData::Synchronize() {
  if (h_data && d_data && numPts > 0)
    safeCall(cudaMemcpy(h_data, d_data, sizeof(SiftPoint) * (size_t)numPts, cudaMemcpyDeviceToHost));
}

static csb_params params_of(const SiftData &d, float subsampling, int rootsift) {
  csb_params p;
  p.num_octaves = d.numOctaves;
  p.init_blur = d.initBlur;
  p.peak_thresh = d.peakThresh;
  p.edge_thresh = d.edgeThresh;
  p.lowest_scale = d.lowestScale;
  p.subsampling = subsampling;
  p.rootsift = rootsift;
  return p;
}

void SiftData::Extract(float *im, int width, int height, float subsampling) {
  csb_ctx *ctx = shim_ctx();
  if (d_data == NULL) {
    fprintf(stderr, "SiftData::Extract: no device storage (construct with dev = true)\n");
    exit(-1);
  }
  csb_params p = params_of(*this, subsampling, 0);
  int n = 0;
  shim_check(ctx, csb_extract_host(ctx, im, width, height, &p, d_data, maxPts, h_data, &n), "SiftData::Extract");
  numPts = n;
}

double SiftData::ConvertSiftToRootSift() {
  csb_ctx *ctx = shim_ctx();
  if (d_data != NULL && numPts > 0) shim_check(ctx, csb_rootsift(ctx, d_data, numPts), "ConvertSiftToRootSift");
  return 0.0;
}

static void extract_legacy(SiftData &sd, cuImage &img, int numOctaves, double initBlur, float thresh,
                           float lowestScale, float subsampling, int rootsift) {
  csb_ctx *ctx = shim_ctx();
  if (sd.d_data == NULL || img.d_data == NULL) {
    fprintf(stderr, "ExtractSift: missing device data\n");
    exit(-1);
  }
  sd.numOctaves = numOctaves;
  sd.numScales = CSB_NUM_SCALES;
  sd.initBlur = initBlur;
  sd.initSubsampling = subsampling;
  sd.peakThresh = thresh;
  sd.edgeThresh = 10.0f;
  sd.lowestScale = lowestScale;
  csb_params p = params_of(sd, subsampling, rootsift);
  int n = 0;
  shim_check(ctx, csb_extract(ctx, img.d_data, img.width, img.height, img.pitch, &p, sd.d_data, sd.maxPts, sd.h_data, &n),
             "ExtractSift");
  sd.numPts = n;
}

void ExtractSift(SiftData &siftData, cuImage &img, int numOctaves, double initBlur, float thresh, float lowestScale,
                 float subsampling) {
  extract_legacy(siftData, img, numOctaves, initBlur, thresh, lowestScale, subsampling, 0);
}

void ExtractRootSift(SiftData &siftData, cuImage &img, int numOctaves, double initBlur, float thresh,
                     float lowestScale, float subsampling) {
  extract_legacy(siftData, img, numOctaves, initBlur, thresh, lowestScale, subsampling, 1);
}

double ScaleDown(cuImage &res, cuImage &src, float variance) {
  if (res.d_data == NULL || src.d_data == NULL) {
    printf("ScaleDown: missing data\n");
    return 0.0;
  }
  csb_ctx *ctx = shim_ctx();
  shim_check(ctx, csb_scale_down_var(ctx, src.d_data, src.width, src.height, src.pitch, res.d_data, res.pitch, variance),
             "ScaleDown");
  return 0.0;
}

// ----------------------------------------------------------------- matching ---
vector<SiftMatch *> MatchSiftData(SiftData &data1, SiftData &data2, MatchSiftDistance distance, float scoreThreshold,
                                  float ambiguityThreshold, MatchType type) {
  vector<SiftMatch *> matches;
  if (!data1.numPts || !data2.numPts) return matches;
  if (data1.d_data == NULL || data2.d_data == NULL) return matches;
  csb_ctx *ctx = shim_ctx();
  shim_check(ctx,
             csb_match(ctx, data1.d_data, data1.numPts, data2.d_data, data2.numPts,
                       distance == MatchSiftDistanceL2 ? 1 : 0, data1.h_data),
             "MatchSiftData");
  if (data1.h_data == NULL) return matches;
  const float thresh2 = scoreThreshold * scoreThreshold;
  const float athresh2 = ambiguityThreshold * ambiguityThreshold;
  for (int i = 0; i < data1.numPts; i++) {
    SiftPoint &a = data1.h_data[i];
    if (!(a.score < thresh2 && a.ambiguity < athresh2)) continue;
    if (type == MatchType3D) {
      if (data2.h_data == NULL || a.coords3D[2] == 0 || data2.h_data[a.match].coords3D[2] == 0) continue;
    }
    SiftMatch *m = new SiftMatch();
    m->pt1 = &a;
    m->pt2 = data2.h_data ? &data2.h_data[a.match] : NULL;
    m->score = a.score;
    m->ambiguity = a.ambiguity;
    matches.push_back(m);
  }
  return matches;
}

// --------------------------------------------------------------- homography ---
double FindHomography(SiftData &data, float *homography, int *numMatches, int numLoops, float minScore,
                      float maxAmbiguity, float thresh) {
  *numMatches = 0;
  homography[0] = homography[4] = homography[8] = 1.0f;
  homography[1] = homography[2] = homography[3] = 0.0f;
  homography[5] = homography[6] = homography[7] = 0.0f;
  if (data.d_data == NULL) return 0.0;
  const double t0 = now_ms();
  numLoops = iDivUp(numLoops, 16) * 16;
  const int numPts = data.numPts;
  if (numPts < 8) return 0.0;
  const SiftPoint *d_sift = data.d_data;
  std::vector<float> scores(numPts), ambiguities(numPts);
  safeCall(cudaMemcpy2D(scores.data(), sizeof(float), &d_sift[0].score, sizeof(SiftPoint), sizeof(float), numPts,
                        cudaMemcpyDeviceToHost));
  safeCall(cudaMemcpy2D(ambiguities.data(), sizeof(float), &d_sift[0].ambiguity, sizeof(SiftPoint), sizeof(float),
                        numPts, cudaMemcpyDeviceToHost));
  std::vector<int> validPts;
  validPts.reserve(numPts);
  for (int i = 0; i < numPts; i++)
    if (scores[i] > minScore && ambiguities[i] < maxAmbiguity) validPts.push_back(i);
  const int numValid = (int)validPts.size();
  if (numValid >= 8) {
    std::vector<int> randPts((size_t)4 * numLoops);
    for (int i = 0; i < numLoops; i++) {   // same draw order as homography.cu:232-244
      int p1 = rand() % numValid;
      int p2 = rand() % numValid;
      int p3 = rand() % numValid;
      int p4 = rand() % numValid;
      while (p2 == p1) p2 = rand() % numValid;
      while (p3 == p1 || p3 == p2) p3 = rand() % numValid;
      while (p4 == p1 || p4 == p2 || p4 == p3) p4 = rand() % numValid;
      randPts[i + 0 * (size_t)numLoops] = validPts[p1];
      randPts[i + 1 * (size_t)numLoops] = validPts[p2];
      randPts[i + 2 * (size_t)numLoops] = validPts[p3];
      randPts[i + 3 * (size_t)numLoops] = validPts[p4];
    }
    csb_ctx *ctx = shim_ctx();
    shim_check(ctx, csb_find_homography(ctx, data.d_data, numPts, randPts.data(), numLoops, thresh, homography, numMatches),
               "FindHomography");
  }
  return now_ms() - t0;
}

namespace {
// Solves the symmetric positive-definite 8x8 system M A = X (Cholesky); false if not SPD.
bool chol_solve8(const double M[8][8], const double X[8], double A[8]) {
  double L[8][8];
  memset(L, 0, sizeof(L));
  for (int i = 0; i < 8; i++)
    for (int j = 0; j <= i; j++) {
      double s = M[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) {
        if (!(s > 0.0)) return false;
        L[i][i] = sqrt(s);
      } else {
        L[i][j] = s / L[j][j];
      }
    }
  double y[8];
  for (int i = 0; i < 8; i++) {
    double s = X[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = 7; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < 8; k++) s -= L[k][i] * A[k];
    A[i] = s / L[i][i];
  }
  return true;
}
}  // namespace

int ImproveHomography(SiftData &data, float *homography, int numLoops, float minScore, float maxAmbiguity,
                      float thresh) {
  if (data.h_data == NULL) return 0;
  SiftPoint *mpts = data.h_data;
  const float limit = thresh * thresh;
  const int numPts = data.numPts;
  double A[8], M[8][8], X[8], Y[8];
  for (int i = 0; i < 8; i++) A[i] = homography[i] / homography[8];
  for (int loop = 0; loop < numLoops; loop++) {
    memset(M, 0, sizeof(M));
    memset(X, 0, sizeof(X));
    for (int i = 0; i < numPts; i++) {
      SiftPoint &pt = mpts[i];
      if (pt.score < minScore || pt.ambiguity > maxAmbiguity) continue;
      float den = A[6] * pt.coords2D[0] + A[7] * pt.coords2D[1] + 1.0f;
      float dx = (A[0] * pt.coords2D[0] + A[1] * pt.coords2D[1] + A[2]) / den - pt.match_xpos;
      float dy = (A[3] * pt.coords2D[0] + A[4] * pt.coords2D[1] + A[5]) / den - pt.match_ypos;
      float err = dx * dx + dy * dy;
      float wei = limit / (err + limit);
      Y[0] = pt.coords2D[0]; Y[1] = pt.coords2D[1]; Y[2] = 1.0;
      Y[3] = Y[4] = Y[5] = 0.0;
      Y[6] = -pt.coords2D[0] * pt.match_xpos; Y[7] = -pt.coords2D[1] * pt.match_xpos;
      for (int c = 0; c < 8; c++)
        for (int r = 0; r < 8; r++) M[r][c] += (Y[c] * Y[r] * wei);
      for (int r = 0; r < 8; r++) X[r] += Y[r] * ((double)pt.match_xpos * (double)wei);
      Y[0] = Y[1] = Y[2] = 0.0;
      Y[3] = pt.coords2D[0]; Y[4] = pt.coords2D[1]; Y[5] = 1.0;
      Y[6] = -pt.coords2D[0] * pt.match_ypos; Y[7] = -pt.coords2D[1] * pt.match_ypos;
      for (int c = 0; c < 8; c++)
        for (int r = 0; r < 8; r++) M[r][c] += (Y[c] * Y[r] * wei);
      for (int r = 0; r < 8; r++) X[r] += Y[r] * ((double)pt.match_ypos * (double)wei);
    }
    chol_solve8(M, X, A);
  }
  int numfit = 0;
  for (int i = 0; i < numPts; i++) {
    SiftPoint &pt = mpts[i];
    float den = A[6] * pt.coords2D[0] + A[7] * pt.coords2D[1] + 1.0;
    float dx = (A[0] * pt.coords2D[0] + A[1] * pt.coords2D[1] + A[2]) / den - pt.match_xpos;
    float dy = (A[3] * pt.coords2D[0] + A[4] * pt.coords2D[1] + A[5]) / den - pt.match_ypos;
    float err = dx * dx + dy * dy;
    if (err < limit) numfit++;
    pt.match_error = sqrt(err);
  }
  for (int i = 0; i < 8; i++) homography[i] = A[i];
  homography[8] = 1.0f;
  return numfit;
}

// ---------------------------------------------------------- rigid transform ---
void EstimateRigidTransformH(const float *h_coord, float *Rt_relative, int *numInliers, int numLoops, int numPts,
                             float thresh2, RigidTransformType type, int *h_indices, char *h_inliers) {
  csb_ctx *ctx = shim_ctx();
  static unsigned int call = 0;   // the reference seeds cuRAND with time(0): a different draw per call
  shim_check(ctx, csb_rigid_transform(ctx, h_coord, numPts, type == RigidTransformType3D ? 1 : 0, h_indices, numLoops, thresh2,
                                      0x5bd1e995u + 7919u * call++, Rt_relative, numInliers, h_inliers),
             "EstimateRigidTransformH");
}

void EstimateRigidTransform(vector<SiftMatch *> matches, float *Rt_relative, int *numInliers, int numLoops, float thresh,
                            RigidTransformType type, int *h_indices, char *h_inliers) {
  std::vector<float> coord(6 * matches.size());
  for (size_t i = 0; i < matches.size(); i++) {
    memcpy(&coord[6 * i], matches[i]->pt1->coords3D, sizeof(float) * 3);
    memcpy(&coord[6 * i + 3], matches[i]->pt2->coords3D, sizeof(float) * 3);
  }
  EstimateRigidTransformH(coord.data(), Rt_relative, numInliers, numLoops, (int)matches.size(), thresh * thresh, type,
                          h_indices, h_inliers);
}
