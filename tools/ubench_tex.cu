// Development micro-benchmark: peak rate of bilinear tex2D<float> fetches on sm_100a with the access pattern of
// k_orient_desc (a warp samples a 16x16-ish patch around one keypoint; unnormalised coordinates, linear filter,
// clamp addressing, pitched 2-D fp32 texture as in cuSIFT.cu:218-236).  Prints G fetches/s and fetches/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tools/ubench_tex tools/ubench_tex.cu
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(128) k(cudaTextureObject_t tex, float *out, int iters, int w, int h, float step) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // one "keypoint" per warp, spread over the image
  const float cx = 20.0f + (float)((warp * 97) % (w - 40)), cy = 20.0f + (float)((warp * 57) % (h - 40));
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = 0.f;
  float px = cx + step * (float)(lane & 15) - 8.0f * step, py = cy + step * (float)(lane >> 4) - 8.0f * step;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] += tex2D<float>(tex, px + 0.37f * i, py + step * 2.0f * (float)(it & 7));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  const int w = 1920, h = 1080, pitch = 1920;
  float *d_img, *d_out;
  cudaMalloc(&d_img, sizeof(float) * pitch * h);
  std::vector<float> img((size_t)pitch * h);
  for (size_t i = 0; i < img.size(); i++) img[i] = (float)(i % 251);
  cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
  cudaResourceDesc res; memset(&res, 0, sizeof(res));
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = d_img; res.res.pitch2D.width = w; res.res.pitch2D.height = h;
  res.res.pitch2D.pitchInBytes = pitch * 4; res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  cudaTextureDesc td; memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &res, &td, nullptr);
  const int blocks = 148 * 16, iters = 512;
  cudaMalloc(&d_out, blocks * 128 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (float step : {0.75f, 1.5f, 3.0f}) {
    k<8><<<blocks, 128>>>(tex, d_out, 8, w, h, step);
    cudaEventRecord(e0);
    k<8><<<blocks, 128>>>(tex, d_out, iters, w, h, step);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)blocks * 128 * iters * 8;
    printf("{\"pattern\": \"16x2 patch, step %.2f px\", \"gfetch_per_s\": %.1f, \"fetch_per_clk_per_sm\": %.2f}\n", step,
           n / (ms * 1e-3) / 1e9, n / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
