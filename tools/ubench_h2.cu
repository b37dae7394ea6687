// Development micro-benchmark: issue throughput of the packed-half instructions considered for the
// extrema prefilter (VHMNMX 3-input half2 max, F2FP pack, HSET2 compares, SHFL) on sm_100a.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned hmax2(unsigned a, unsigned b) {
  unsigned d;
  asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned *out, unsigned a, unsigned b, float fa, int iters) {
  unsigned x[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    x[i] = threadIdx.x * 77u + i * 1315423911u + blockIdx.x;
    f[i] = threadIdx.x * 0.01f + i;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) x[i] = hmax2(hmax2(x[i], a), x[(i + 1) & 7]);                 // VHMNMX (3-input)
      else if (MODE == 1) x[i] = hmax2(x[i], x[(i + 1) & 7] ^ a);                 // 2-input + LOP
      else if (MODE == 2) x[i] = __vimax3_s16x2(x[i], a, x[(i + 1) & 7]);          // VIMNMX3.S16x2
      else if (MODE == 3) {                                                       // F2FP pack
        __half2 p = __floats2half2_rn(f[i], f[(i + 1) & 7]);
        x[i] ^= *reinterpret_cast<unsigned *>(&p);
        f[i] += fa;
      } else if (MODE == 4) {                                                     // HSET2
        x[i] = __heq2_mask(*reinterpret_cast<__half2 *>(&x[i]), *reinterpret_cast<__half2 *>(&x[(i + 1) & 7])) ^ a;
      } else if (MODE == 5) x[i] = __shfl_down_sync(0xffffffffu, x[i], 1) + a;    // SHFL
      else if (MODE == 7) x[i] = hmax2(x[i], x[(i + 3) & 7]);                     // 2-input alone
      else if (MODE == 8) { __half2 h = __hmax2(*reinterpret_cast<__half2 *>(&x[i]), *reinterpret_cast<__half2 *>(&x[(i + 3) & 7])); x[i] = *reinterpret_cast<unsigned *>(&h); }
      else if (MODE == 9) { __half2 h = __hfma2(*reinterpret_cast<__half2 *>(&x[i]), *reinterpret_cast<__half2 *>(&x[(i + 3) & 7]), *reinterpret_cast<__half2 *>(&a)); x[i] = *reinterpret_cast<unsigned *>(&h); }
      else if (MODE == 10) { f[i] = -f[i] * fa; __half2 p = __floats2half2_rn(f[i], f[(i + 1) & 7]); x[i] = *reinterpret_cast<unsigned *>(&p); }   // F2FP + FMUL
      else if (MODE == 6) f[i] = fmaxf(fmaxf(f[i], fa), f[(i + 1) & 7]);           // FMNMX3 fp32
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name) {
  unsigned *d;
  const int blocks = 148 * 8, iters = 4096;
  cudaMalloc(&d, blocks * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(d, 0x3c003c00u, 5u, 0.5f, 16);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, 0x3c003c00u, 5u, 0.5f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * 256 * iters * 8;
  printf("%-28s %8.3f ms  %.1f thread-instr/clk/SM @1.965GHz\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(d);
}
int main() {
  run<0>("VHMNMX (3-in half2 max)");
  run<1>("HMNMX2 + LOP3");
  run<2>("VIMNMX3.S16x2");
  run<3>("F2FP pack + LOP3 + FADD");
  run<4>("HSET2.EQ + LOP3");
  run<5>("SHFL + IADD");
  run<6>("FMNMX3 fp32");
  run<7>("max.f16x2 2-input");
  run<8>("__hmax2");
  run<9>("HFMA2");
  run<10>("F2FP + FMUL");
  return 0;
}
