// Development micro-benchmark: issue throughput of scalar vs packed fp32 FMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, int iters) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, blockIdx.x * 0.002f - i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) {          // scalar FFMA: 2 per element pair
        x[i].x = __fmaf_rn(x[i].x, a, b);
        x[i].y = __fmaf_rn(x[i].y, a, b);
      } else if (MODE == 1) {   // packed FFMA2
        x[i] = __ffma2_rn(x[i], a2, b2);
      } else if (MODE == 2) {   // packed FADD2
        x[i] = __fadd2_rn(x[i], a2);
      } else {                  // FMNMX3-like
        x[i].x = fmaxf(fmaxf(x[i].x, a), x[i].y);
        x[i].y = fminf(fminf(x[i].y, b), x[i].x);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, int flop_per_iter_per_thread) {
  float *d;
  const int blocks = 148 * 8, iters = 4096;
  cudaMalloc(&d, blocks * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f, 16);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, 1.0001f, 0.5f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * 256 * iters * flop_per_iter_per_thread;
  printf("%-10s %8.3f ms  %8.2f Gop/s (lane-ops)  = %.1f lane-ops/clk/SM @1.965GHz\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(d);
}
int main() {
  run<0>("FFMA", 16);    // 16 scalar fma per iter
  run<1>("FFMA2", 16);   // 8 packed = 16 lane-fma
  run<2>("FADD2", 16);
  run<3>("FMNMX", 16);
  return 0;
}
