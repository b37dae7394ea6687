// Development probe: characterises the texture unit's bilinear filter (linear filter mode,
// clamp addressing, unnormalised coordinates — the descriptor used by the hot path) so that the
// CPU oracle's restatement of it (oracle/oracle.c tex2d) can be pinned to the hardware.
// Writes gpurun_out/tex_probe.bin:
//   i32 W, H, N ; f32 tex[H][W] ; f32 xs[N] ; f32 ys[N] ; f32 out[N]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void sample(cudaTextureObject_t tex, const float *xs, const float *ys, float *out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex2D<float>(tex, xs[i], ys[i]);
}

int main(int argc, char **argv) {
  const int W = 2048, H = 16;
  const char *path = argc > 1 ? argv[1] : "gpurun_out/tex_probe.bin";
  std::vector<float> tex((size_t)W * H);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
  for (auto &v : tex) v = rnd() * 255.0f;
  // impulse rows: row 2 is all zero except texel 4 and texel 1000 = 1
  for (int i = 0; i < W; i++) { tex[2 * W + i] = 0.f; tex[1 * W + i] = 0.f; tex[3 * W + i] = 0.f; }
  tex[2 * W + 4] = 1.0f;
  tex[2 * W + 1000] = 1.0f;
  std::vector<float> xs, ys;
  // (a) fine sweep of the x fraction around texel 4 and texel 1000 at the row centre
  for (int k = -8192; k <= 8192; k++) { xs.push_back(4.5f + k / 8192.0f); ys.push_back(2.5f); }
  for (int k = -8192; k <= 8192; k++) { xs.push_back(1000.5f + k / 8192.0f); ys.push_back(2.5f); }
  // (b) 2-D fractions on the impulse
  for (int a = -64; a <= 64; a++)
    for (int b = -64; b <= 64; b++) { xs.push_back(4.5f + a / 64.0f + 1.0f / 1024); ys.push_back(2.5f + b / 64.0f + 3.0f / 2048); }
  // (c) random coordinates over the random part (rows 4..15), incl. slightly outside (clamp)
  for (int k = 0; k < 200000; k++) { xs.push_back(-2.0f + rnd() * (W + 4)); ys.push_back(4.0f + rnd() * 12.5f); }
  const int N = (int)xs.size();

  float *d_tex, *d_x, *d_y, *d_o;
  size_t pitch = W * sizeof(float);
  cudaMalloc(&d_tex, pitch * H);
  cudaMemcpy(d_tex, tex.data(), pitch * H, cudaMemcpyHostToDevice);
  cudaMalloc(&d_x, N * 4); cudaMalloc(&d_y, N * 4); cudaMalloc(&d_o, N * 4);
  cudaMemcpy(d_x, xs.data(), N * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_y, ys.data(), N * 4, cudaMemcpyHostToDevice);
  cudaResourceDesc res = {};
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = d_tex;
  res.res.pitch2D.width = W;
  res.res.pitch2D.height = H;
  res.res.pitch2D.pitchInBytes = pitch;
  res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t t = 0;
  if (cudaCreateTextureObject(&t, &res, &td, nullptr) != cudaSuccess) { fprintf(stderr, "tex create failed\n"); return 1; }
  sample<<<(N + 255) / 256, 256>>>(t, d_x, d_y, d_o, N);
  std::vector<float> out(N);
  if (cudaMemcpy(out.data(), d_o, N * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
  FILE *fp = fopen(path, "wb");
  int hdr[3] = {W, H, N};
  fwrite(hdr, 4, 3, fp);
  fwrite(tex.data(), 4, tex.size(), fp);
  fwrite(xs.data(), 4, N, fp);
  fwrite(ys.data(), 4, N, fp);
  fwrite(out.data(), 4, N, fp);
  fclose(fp);
  printf("tex_probe: %d samples written to %s\n", N, path);
  return 0;
}
