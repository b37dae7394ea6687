// Frame ingest: 8-bit grey frame -> fp32 octave-0 image, optionally with the 3x3 sigma = 0.5
// Gaussian pre-blur the reference's demo applies on the host before uploading
// (danielsuo/cuSIFT main.cpp:301-309: imread(.,0).convertTo(CV_32FC1); GaussianBlur(Size(3,3), 0.5)).
// Uploading the 8-bit frame moves 4x fewer bytes over PCIe than uploading the float image.
//
// The blur restates cv::GaussianBlur for CV_32F, ksize 3, BORDER_DEFAULT (reflect-101) bit for bit as
// OpenCV 4's SIMD loop body evaluates it (verified against cv2 in tests/; OpenCV's own scalar row tails
// round differently, so rows whose length is not a multiple of its vector width agree within 1 ulp there): kernel from
// getGaussianKernel(3, 0.5, CV_32F) (host side, csb_api.cu); row pass  fma(c, k0, (l + r) * k1);
// column pass  fma(u + d, k1, c * k0)  on the rounded row results.
#include "csb_internal.h"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) return -i;
  if (i >= n) return 2 * n - 2 - i;
  return i;
}

template <bool kBlur>
__global__ void __launch_bounds__(256) k_ingest_u8(const unsigned char *__restrict__ src, int stride, int w, int h,
                                                   float *__restrict__ dst, int pitch, float k0, float k1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  if (!kBlur) {
    dst[(size_t)y * pitch + x] = (float)src[(size_t)y * stride + x];
    return;
  }
  const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
  float row[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const unsigned char *r = src + (size_t)reflect101(y - 1 + j, h) * stride;
    const float l = (float)r[xl], c = (float)r[x], rr = (float)r[xr];
    row[j] = __fmaf_rn(c, k0, __fmul_rn(__fadd_rn(l, rr), k1));
  }
  dst[(size_t)y * pitch + x] = __fmaf_rn(__fadd_rn(row[0], row[2]), k1, __fmul_rn(row[1], k0));
}

// Measurement aid: keeps the stream busy for `ns` nanoseconds (bounded spin on the global timer) so that the launches
// queued behind it are all resident in the stream before the first one starts; per-launch CUDA events then bracket
// back-to-back device execution instead of host enqueue gaps.  Only launched when csb_profile_enable is on.
__global__ void k_delay(unsigned long long ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < ns && t1 - t0 < 2000000ull);
}

}  // namespace

void launch_delay(unsigned long long ns, cudaStream_t st) { k_delay<<<1, 1, 0, st>>>(ns); }

void launch_ingest_u8(const unsigned char *d_src, int stride, int w, int h, float *d_dst, int pitch, int preblur, float k0,
                      float k1, cudaStream_t st) {
  dim3 blk(64, 4), grd((w + 63) / 64, (h + 3) / 4);
  if (preblur) k_ingest_u8<true><<<grd, blk, 0, st>>>(d_src, stride, w, h, d_dst, pitch, k0, k1);
  else k_ingest_u8<false><<<grd, blk, 0, st>>>(d_src, stride, w, h, d_dst, pitch, k0, k1);
}
