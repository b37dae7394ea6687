// Development micro-benchmark: what do pure-read, pure-write and copy kernels reach on this GPU?
// (context for roofline.frac: MEASURED_PEAKS.json's hbm_gbs is a COPY figure, bytes read + bytes written)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_read(const float4 *__restrict__ p, size_t n, float *out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = p[i];
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) out[0] = acc;
}
__global__ void k_write(float4 *__restrict__ p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void k_copy(const float4 *__restrict__ a, float4 *__restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
int main() {
  const size_t sizes[] = {78u << 20, 512u << 20, 2048ull << 20};
  for (size_t bytes : sizes) {
    float4 *a, *b;
    float *o;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    const size_t n = bytes / 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int grid = 148 * 16, reps = 20;
    // flush L2 between measurements with a large write
    auto time = [&](int which) {
      k_write<<<grid, 256>>>(b, n);   // evicts `a` when bytes > L2; for 78 MB also measure the L2-resident case
      cudaEventRecord(e0);
      for (int r = 0; r < reps; r++) {
        if (which == 0) k_read<<<grid, 256>>>(a, n, o);
        else if (which == 1) k_write<<<grid, 256>>>(a, n);
        else k_copy<<<grid, 256>>>(a, b, n);
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      return ms / reps;
    };
    const float tr = time(0), tw = time(1), tc = time(2);
    printf("%5zu MB: read %.2f TB/s  write %.2f TB/s  copy %.2f TB/s (read+write bytes)  [%.1f / %.1f / %.1f us]\n", bytes >> 20,
           bytes / tr / 1e9, bytes / tw / 1e9, 2.0 * bytes / tc / 1e9, tr * 1e3, tw * 1e3, tc * 1e3);
    cudaFree(a); cudaFree(b); cudaFree(o);
  }
  return 0;
}
