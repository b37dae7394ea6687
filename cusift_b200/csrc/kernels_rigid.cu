// Rigid-transform RANSAC: replaces EstimateRigidTransformD / testRigidTransform
// (danielsuo/cuSIFT extras/rigidTransform.cu:292-374).  The reference runs one thread per hypothesis
// that allocates with device-side new/delete, loops over every correspondence and writes a
// numLoops x numPts inlier mask; here hypotheses are estimated allocation-free (rigid_math.h), scored
// by (hypothesis, point-slice) CTAs from shared memory, and only the winner's mask is materialised.
#include "csb_internal.h"
#include "rigid_math.h"

namespace {

__host__ __device__ inline unsigned int rt_hash(unsigned int seed, unsigned int loop, unsigned int k, unsigned int attempt) {
  unsigned int x = seed ^ (loop * 0x9E3779B9u) ^ (k * 0x85EBCA6Bu) ^ (attempt * 0xC2B2AE35u);
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

// thread = one hypothesis: (optionally draw 3 distinct points,) estimate Rt
__global__ void __launch_bounds__(64) k_rt_estimate(const float *__restrict__ coord, int num_pts, int *__restrict__ indices,
                                                    int draw, unsigned int seed, int type3d, int num_loops,
                                                    float *__restrict__ Rt, int *__restrict__ counts) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= num_loops) return;
  counts[l] = 0;
  int p[3];
  if (draw) {   // rigidTransform.cu:340-355 (curand there; a counter-based hash here)
    for (int k = 0; k < 3; k++) {
      unsigned int attempt = 0;
      for (;;) {
        const int c = (int)(rt_hash(seed, (unsigned int)l, (unsigned int)k, attempt++) % (unsigned int)num_pts);
        bool dup = false;
        for (int q = 0; q < k; q++) dup = dup || (p[q] == c);
        if (!dup) { p[k] = c; break; }
      }
      indices[3 * l + k] = p[k];
    }
  } else {
    for (int k = 0; k < 3; k++) p[k] = indices[3 * l + k];
  }
  float r[12];
  if (type3d) csb_rigid3d(coord, p, 3, r);
  else csb_rigid2d(coord, p[0], p[1], r);
  for (int i = 0; i < 12; i++) Rt[12 * l + i] = r[i];
}

#define RT_HYP 128
#define RT_PTS 256
// thread = one hypothesis (its 12 coefficients in registers), CTA = 128 hypotheses x 256 points
__global__ void __launch_bounds__(RT_HYP) k_rt_score(const float *__restrict__ coord, int num_pts, const float *__restrict__ Rt,
                                                     int num_loops, float thresh2, int *__restrict__ counts) {
  __shared__ float pts[RT_PTS * 6];
  const int p0 = blockIdx.y * RT_PTS;
  const int np = min(RT_PTS, num_pts - p0);
  for (int i = threadIdx.x; i < np * 6; i += RT_HYP) pts[i] = coord[(size_t)p0 * 6 + i];
  __syncthreads();
  const int l = blockIdx.x * RT_HYP + threadIdx.x;
  if (l >= num_loops) return;
  float r[12];
#pragma unroll
  for (int i = 0; i < 12; i++) r[i] = Rt[12 * l + i];
  int cnt = 0;
  for (int i = 0; i < np; i++) cnt += csb_rigid_inlier(r, pts + 6 * i, thresh2) ? 1 : 0;
  if (cnt) atomicAdd(counts + l, cnt);
}

__global__ void k_rt_mask(const float *__restrict__ coord, int num_pts, const float *__restrict__ Rt, int loop, float thresh2,
                          char *__restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_pts) return;
  float r[12];
#pragma unroll
  for (int k = 0; k < 12; k++) r[k] = Rt[12 * loop + k];
  mask[i] = csb_rigid_inlier(r, coord + 6 * (size_t)i, thresh2) ? 1 : 0;
}

}  // namespace

void launch_rigid_hypotheses(const float *d_coord, int num_pts, int *d_indices, int draw, unsigned int seed, int type3d,
                             int num_loops, float thresh2, float *d_Rt, int *d_counts, cudaStream_t st) {
  k_rt_estimate<<<(num_loops + 63) / 64, 64, 0, st>>>(d_coord, num_pts, d_indices, draw, seed, type3d, num_loops, d_Rt, d_counts);
  k_rt_score<<<dim3((num_loops + RT_HYP - 1) / RT_HYP, (num_pts + RT_PTS - 1) / RT_PTS), RT_HYP, 0, st>>>(
      d_coord, num_pts, d_Rt, num_loops, thresh2, d_counts);
}

void launch_rigid_mask(const float *d_coord, int num_pts, const float *d_Rt, int loop, float thresh2, char *d_mask,
                       cudaStream_t st) {
  k_rt_mask<<<(num_pts + 255) / 256, 256, 0, st>>>(d_coord, num_pts, d_Rt, loop, thresh2, d_mask);
}

void rigid_refit_host(const float *h_coord, const int *idx, int n, float *Rt) { csb_rigid3d(h_coord, idx, n, Rt); }
unsigned int rigid_hash_host(unsigned int seed, unsigned int loop, unsigned int k, unsigned int attempt) {
  return rt_hash(seed, loop, k, attempt);
}
