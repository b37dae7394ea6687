// cusift_b200 — reference-compatible CUDA utility header.
// Same names and semantics as danielsuo/cuSIFT cutils.h (iDivUp/iAlignUp :15-18,
// safeCall/checkMsg :20-48 -> message on stderr + exit(-1), InitCuda :71-92,
// TimerGPU :94-114, TimerCPU :116-141); written from scratch for this library.
#ifndef CUSIFT_B200_CUTILS_H
#define CUSIFT_B200_CUTILS_H

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "cuda_runtime_api.h"

inline int iDivUp(int a, int b) { return (a + b - 1) / b; }
inline int iDivDown(int a, int b) { return a / b; }
inline int iAlignUp(int a, int b) { return iDivUp(a, b) * b; }
inline int iAlignDown(int a, int b) { return (a / b) * b; }

#define safeCall(err) cusift_b200_safe_call((err), __FILE__, __LINE__)
#define safeThreadSync() cusift_b200_safe_call(cudaDeviceSynchronize(), __FILE__, __LINE__)
#define checkMsg(msg) cusift_b200_check_msg((msg), __FILE__, __LINE__)

inline void cusift_b200_safe_call(cudaError_t err, const char *file, int line) {
  if (err == cudaSuccess) return;
  fprintf(stderr, "safeCall() Runtime API error in file <%s>, line %i : %s.\n", file, line, cudaGetErrorString(err));
  exit(-1);
}

inline void cusift_b200_check_msg(const char *what, const char *file, int line) {
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) return;
  fprintf(stderr, "checkMsg() CUDA error: %s in file <%s>, line %i : %s.\n", what, file, line, cudaGetErrorString(err));
  exit(-1);
}

// Selects device `dev`, clamped to the devices present; false when there is none.
inline bool deviceInit(int dev) {
  int count = 0;
  safeCall(cudaGetDeviceCount(&count));
  if (count == 0) {
    fprintf(stderr, "CUDA error: no devices supporting CUDA.\n");
    return false;
  }
  dev = std::max(0, std::min(dev, count - 1));
  safeCall(cudaSetDevice(dev));
  return true;
}

inline void InitCuda(int devNum) {
  int count = 0;
  cudaGetDeviceCount(&count);
  if (count == 0) {
    std::cerr << "No CUDA devices available" << std::endl;
    return;
  }
  deviceInit(std::min(count - 1, devNum));
}

// Event timer: starts at construction, read() returns elapsed ms (and synchronises).
class TimerGPU {
public:
  cudaEvent_t start, stop;
  cudaStream_t stream;
  TimerGPU(cudaStream_t stream_ = 0) : stream(stream_) {
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    cudaEventRecord(start, stream);
  }
  ~TimerGPU() {
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
  }
  float read() {
    cudaEventRecord(stop, stream);
    cudaEventSynchronize(stop);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, start, stop);
    return ms;
  }
};

// Wall-clock timer; `freq` (MHz in the reference's rdtsc version) is accepted and ignored.
class TimerCPU {
public:
  std::chrono::steady_clock::time_point beg;
  float freq;
  TimerCPU(float freq_) : beg(std::chrono::steady_clock::now()), freq(freq_) {}
  float read() {
    return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - beg).count();
  }
};

#endif
