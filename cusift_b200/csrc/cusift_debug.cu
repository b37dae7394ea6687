// Harness compatibility pack (include/cusift/extras/debug.h): loaders for the golden files of the
// reference's tests and SiftData helpers, without OpenCV.  File formats: extras/debug.cpp:120-407 of
// danielsuo/cuSIFT.  Unlike the reference these readers check every read; a missing or short file
// prints a message and yields an empty result instead of reading uninitialised memory.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "extras/debug.h"

namespace {

// Sequential reader over one binary file.
class BinFile {
 public:
  explicit BinFile(const char *path) : fp_(fopen(path, "rb")), path_(path), ok_(fp_ != NULL) {
    if (!ok_) fprintf(stderr, "cusift_b200: cannot open %s\n", path);
  }
  ~BinFile() {
    if (fp_) fclose(fp_);
  }
  template <typename T>
  bool read(T *dst, size_t n) {
    if (!ok_) return false;
    if (n && fread((void *)dst, sizeof(T), n, fp_) != n) {
      fprintf(stderr, "cusift_b200: %s is truncated\n", path_);
      ok_ = false;
    }
    return ok_;
  }
  uint32_t u32() {
    uint32_t v = 0;
    read(&v, 1);
    return ok_ ? v : 0;
  }
  bool ok() const { return ok_; }

 private:
  FILE *fp_;
  const char *path_;
  bool ok_;
};

SiftMatch *new_match_pair() {
  SiftMatch *m = new SiftMatch();
  m->pt1 = new SiftPoint();
  m->pt2 = new SiftPoint();
  memset(m->pt1, 0, sizeof(SiftPoint));
  memset(m->pt2, 0, sizeof(SiftPoint));
  m->score = m->ambiguity = m->error = 0.0f;
  return m;
}

// (re)allocates the page-locked host array the way InitSiftData does
SiftPoint *host_points(size_t n) {
  SiftPoint *p = (SiftPoint *)malloc(sizeof(SiftPoint) * n);
  if (p && cudaHostRegister(p, sizeof(SiftPoint) * n, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess)
    cudaGetLastError();
  return p;
}

}  // namespace

void PrintSiftData(SiftData &data) {
  SiftPoint *h = data.h_data;
  bool temp = false;
  if (h == NULL && data.d_data != NULL && data.numPts > 0) {   // the reference keeps this buffer; we free it
    h = (SiftPoint *)malloc(sizeof(SiftPoint) * (size_t)data.numPts);
    safeCall(cudaMemcpy(h, data.d_data, sizeof(SiftPoint) * (size_t)data.numPts, cudaMemcpyDeviceToHost));
    temp = true;
  }
  for (int i = 0; h != NULL && i < data.numPts; i++) {
    const SiftPoint &p = h[i];
    printf("xpos         = %.2f\n", p.coords2D[0]);
    printf("ypos         = %.2f\n", p.coords2D[1]);
    printf("scale        = %.2f\n", p.scale);
    printf("sharpness    = %.2f\n", p.sharpness);
    printf("edgeness     = %.2f\n", p.edgeness);
    printf("orientation  = %.2f\n", p.orientation);
    printf("score        = %.2f\n", p.score);
    for (int row = 0; row < 8; row++) {
      printf(row == 0 ? "data = " : "       ");
      for (int k = 0; k < 16; k++) {
        const float v = p.data[row * 16 + k];
        if (v < 0.01) printf(" .   ");
        else printf("%.2f ", v);
      }
      printf("\n");
    }
  }
  printf("Number of available points: %d\n", data.numPts);
  printf("Number of allocated points: %d\n", data.maxPts);
  if (temp) free(h);
}

void PrintMatchSiftData(SiftData &siftData1, const char *filename, int imgw) {
  std::ofstream out(filename);
  if (!out) {
    std::cout << "File Not Opened" << std::endl;
    return;
  }
  for (int i = 0; i < siftData1.numPts && siftData1.h_data != NULL; i++) {
    const SiftPoint &p = siftData1.h_data[i];
    const int ind = (int)p.coords2D[0] + (int)p.coords2D[1] * imgw;
    const int ind2 = (int)p.match_xpos + (int)p.match_ypos * imgw;
    out << p.coords2D[0] << "\t" << p.coords2D[1] << "\t" << p.match_xpos << "\t" << p.match_ypos << "\t" << ind << "\t"
        << ind2 << "\t" << std::endl;
  }
}

void ReadVLFeatSiftData(SiftData &siftData, const char *filename) {
  fprintf(stderr, "Reading vlfeat data from %s", filename);
  BinFile f(filename);
  const uint32_t n = f.u32();
  std::vector<float> frames((size_t)4 * n), desc((size_t)128 * n);
  if (!f.read(frames.data(), frames.size()) || !f.read(desc.data(), desc.size())) {
    fprintf(stderr, " ... failed\n");
    return;
  }
  fprintf(stderr, " ... and got %u points\n", n);
  std::vector<SiftPoint> pts(n);
  if (n) memset(pts.data(), 0, sizeof(SiftPoint) * n);
  for (uint32_t i = 0; i < n; i++) {
    pts[i].coords2D[0] = frames[4 * i + 0];
    pts[i].coords2D[1] = frames[4 * i + 1];
    pts[i].scale = frames[4 * i + 2];
    pts[i].orientation = frames[4 * i + 3];
    memcpy(pts[i].data, &desc[(size_t)128 * i], sizeof(float) * 128);
  }
  AddSiftData(siftData, pts.data(), (int)n);
}

int ReadMATLABMatchIndices(const char *indices_filename, uint32_t *indices_i, uint32_t *indices_j) {
  fprintf(stderr, "Reading match indices data from %s\n", indices_filename);
  BinFile f(indices_filename);
  const uint32_t n = f.u32();
  if (indices_i != NULL && indices_j != NULL) {
    f.read(indices_i, n);
    f.read(indices_j, n);
  }
  return f.ok() ? (int)n : 0;
}

vector<SiftMatch *> ReadMATLABMatchData(const char *filename) {
  fprintf(stderr, "Reading MATLAB match data from %s\n", filename);
  vector<SiftMatch *> matches;
  BinFile f(filename);
  const uint32_t n = f.u32();
  for (uint32_t i = 0; i < n; i++) {
    double c[6];
    if (!f.read(c, 6)) break;
    SiftMatch *m = new_match_pair();
    for (int k = 0; k < 3; k++) {
      m->pt1->coords3D[k] = (float)c[k];
      m->pt2->coords3D[k] = (float)c[3 + k];
    }
    matches.push_back(m);
  }
  return matches;
}

vector<SiftMatch *> ReadMATLABMatchDataBeforeRANSAC(const char *filename) {
  fprintf(stderr, "Reading MATLAB match data before ransac from %s\n", filename);
  vector<SiftMatch *> matches;
  BinFile f(filename);
  const uint32_t n = f.u32();
  std::vector<uint32_t> id_i(n), id_j(n);
  std::vector<float> des_i((size_t)128 * n), des_j((size_t)128 * n);
  if (!f.read(id_i.data(), n) || !f.read(id_j.data(), n) || !f.read(des_i.data(), des_i.size()) ||
      !f.read(des_j.data(), des_j.size()))
    return matches;
  for (uint32_t i = 0; i < n; i++) {
    // the reference leaks its two points here (match->pt1/pt2 are never set, debug.cpp:303-312);
    // attaching them is what its callers would need
    SiftMatch *m = new_match_pair();
    memcpy(m->pt1->data, &des_i[(size_t)128 * i], sizeof(float) * 128);
    memcpy(m->pt2->data, &des_j[(size_t)128 * i], sizeof(float) * 128);
    matches.push_back(m);
  }
  return matches;
}

vector<SiftMatch *> ReadMATLABRANSAC(const char *filename, vector<int> &indices, float *Rt) {
  fprintf(stderr, "Reading MATLAB RANSAC data produced using DEBUG_ransactfitRt.m from %s\n", filename);
  vector<SiftMatch *> matches;
  BinFile f(filename);
  const uint32_t n = f.u32(), loops = f.u32();
  fprintf(stderr, "Read %u matches and %u loop indices\n", n, loops);
  std::vector<float> ci((size_t)3 * n), cj((size_t)3 * n);
  std::vector<int> idx((size_t)3 * loops);
  if (!f.read(ci.data(), ci.size()) || !f.read(cj.data(), cj.size()) || !f.read(idx.data(), idx.size()) || !f.read(Rt, 12))
    return matches;
  for (uint32_t i = 0; i < n; i++) {
    SiftMatch *m = new_match_pair();
    memcpy(m->pt1->coords3D, &ci[(size_t)3 * i], sizeof(float) * 3);
    memcpy(m->pt2->coords3D, &cj[(size_t)3 * i], sizeof(float) * 3);
    matches.push_back(m);
  }
  for (size_t i = 0; i < idx.size(); i++) indices.push_back(idx[i] - 1);
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 4; c++) fprintf(stderr, "%0.4f ", Rt[r * 4 + c]);
    fprintf(stderr, "\n");
  }
  return matches;
}

vector<int> ReadMATLABIndices(const char *filename) {
  fprintf(stderr, "Reading MATLAB indices data from %s\n", filename);
  vector<int> out;
  BinFile f(filename);
  const uint32_t pairs = f.u32();
  std::vector<uint32_t> raw((size_t)2 * pairs);
  if (!f.read(raw.data(), raw.size())) return out;
  for (size_t i = 0; i < raw.size(); i++) out.push_back((int)raw[i] - 1);
  return out;
}

void ReadMATLABRt(double *Rt_relative, const char *filename) {
  fprintf(stderr, "Reading MATLAB Rt data from %s\n", filename);
  BinFile f(filename);
  if (!f.read(Rt_relative, 12)) return;
  fprintf(stderr, "MATLAB Rt: ");
  for (int i = 0; i < 12; i++) fprintf(stderr, "%0.4f ", Rt_relative[i]);
  fprintf(stderr, "\n");
}

void AddSiftData(SiftData &data, SiftPoint *h_data, int numPts) {
  if (numPts <= 0 || h_data == NULL) return;
  const int total = data.numPts + numPts;
  if (data.h_data == NULL && data.d_data == NULL) {
    // A default-constructed SiftData owns no storage (cuSIFT.h:57: host = dev = false), and the InitSiftData call the
    // reference's reader used to make is commented out (debug.cpp:128) - at the reference's HEAD its own
    // test/test.cpp:26-40 therefore reads nothing and indexes an empty match vector.  Deviation:
    // give such an object host + device storage here, so that the reference's tests run as written.
    int cap = data.maxPts > 0 ? data.maxPts : 1024;
    while (cap < total) cap *= 2;
    data.numPts = 0;
    data.h_data = host_points((size_t)cap);
    safeCall(cudaMalloc((void **)&data.d_data, sizeof(SiftPoint) * (size_t)cap));
    data.maxPts = cap;
  }
  if (data.maxPts < total) {
    int cap = data.maxPts > 0 ? 2 * data.maxPts : 1024;
    while (cap < total) cap *= 2;
    if (data.h_data != NULL) {
      SiftPoint *grown = host_points((size_t)cap);
      memcpy(grown, data.h_data, sizeof(SiftPoint) * (size_t)data.numPts);
      if (cudaHostUnregister(data.h_data) != cudaSuccess) cudaGetLastError();
      free(data.h_data);
      data.h_data = grown;
    }
    if (data.d_data != NULL) {
      SiftPoint *grown = NULL;
      safeCall(cudaMalloc((void **)&grown, sizeof(SiftPoint) * (size_t)cap));
      safeCall(cudaMemcpy(grown, data.d_data, sizeof(SiftPoint) * (size_t)data.numPts, cudaMemcpyDeviceToDevice));
      safeCall(cudaFree(data.d_data));
      data.d_data = grown;
    }
    data.maxPts = cap;
  }
  if (data.h_data != NULL) memcpy(data.h_data + data.numPts, h_data, sizeof(SiftPoint) * (size_t)numPts);
  if (data.d_data != NULL)
    safeCall(cudaMemcpy(data.d_data + data.numPts, h_data, sizeof(SiftPoint) * (size_t)numPts, cudaMemcpyHostToDevice));
  data.numPts = total;
}
