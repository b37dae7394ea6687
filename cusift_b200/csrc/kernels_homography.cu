// RANSAC homography: hypothesis generation and scoring.
//
// Replaces ComputeHomographies + InvertMatrix<8> (reference
// extras/homography.cu:98-139, 12-96) and TestHomographies (:144-187), plus the
// four strided cudaMemcpy2D gathers of FindHomography (:246-249), with three
// launches on one stream and no intermediate host synchronisation.
//   k_gather_coords : AoS SiftPoint -> SoA x1,y1,x2,y2 (pad slots zeroed; the
//                     reference leaves them uninitialised, homography.cu:215)
//   k_hypotheses    : one thread per 4-point sample, 8x8 DLT system solved by
//                     the same Crout LU / implicit pivoting as the reference
//   k_score         : one warp per hypothesis, inlier test with the reference's
//                     round-toward-zero products (__fmul_rz), shuffle sum
#include "csb_internal.h"

namespace {

__global__ void k_gather_coords(const csb_sift_point *__restrict__ d_sift, int n, int n_up, float *__restrict__ coord,
                                int *__restrict__ counts, int num_loops) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < num_loops) counts[i] = 0;             // k_score accumulates with atomicAdd
  if (i >= n_up) return;
  float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
  if (i < n) {
    x1 = d_sift[i].coords2D[0];
    y1 = d_sift[i].coords2D[1];
    x2 = d_sift[i].match_xpos;
    y2 = d_sift[i].match_ypos;
  }
  coord[i + 0 * n_up] = x1;
  coord[i + 1 * n_up] = y1;
  coord[i + 2 * n_up] = x2;
  coord[i + 3 * n_up] = y2;
}

// Numerical-Recipes style LU inverse, restated from homography.cu:12-96.
__device__ void invert8(float elem[8][8], float res[8][8]) {
  const int size = 8;
  int indx[8];
  float b[8], vv[8];
  for (int i = 0; i < size; i++) indx[i] = 0;
  int imax = 0;
  for (int i = 0; i < size; i++) {
    float big = 0.0f;
    for (int j = 0; j < size; j++) {
      const float temp = fabsf(elem[i][j]);
      if (temp > big) big = temp;
    }
    if (big > 0.0f) vv[i] = (float)(1.0 / (double)big);
    else vv[i] = 1e16f;
  }
  for (int j = 0; j < size; j++) {
    for (int i = 0; i < j; i++) {
      float sum = elem[i][j];
      for (int k = 0; k < i; k++) sum -= elem[i][k] * elem[k][j];
      elem[i][j] = sum;
    }
    float big = 0.0f;
    for (int i = j; i < size; i++) {
      float sum = elem[i][j];
      for (int k = 0; k < j; k++) sum -= elem[i][k] * elem[k][j];
      elem[i][j] = sum;
      const float dum = vv[i] * fabsf(sum);
      if (dum >= big) {
        big = dum;
        imax = i;
      }
    }
    if (j != imax) {
      for (int k = 0; k < size; k++) {
        const float dum = elem[imax][k];
        elem[imax][k] = elem[j][k];
        elem[j][k] = dum;
      }
      vv[imax] = vv[j];
    }
    indx[j] = imax;
    if (elem[j][j] == 0.0f) elem[j][j] = 1e-16f;
    if (j != (size - 1)) {
      const float dum = (float)(1.0 / (double)elem[j][j]);
      for (int i = j + 1; i < size; i++) elem[i][j] *= dum;
    }
  }
  for (int j = 0; j < size; j++) {
    for (int k = 0; k < size; k++) b[k] = 0.0f;
    b[j] = 1.0f;
    int ii = -1;
    for (int i = 0; i < size; i++) {
      const int ip = indx[i];
      float sum = b[ip];
      b[ip] = b[i];
      if (ii != -1) {
        for (int jj = ii; jj < i; jj++) sum -= elem[i][jj] * b[jj];
      } else if (sum != 0.0f) {
        ii = i;
      }
      b[i] = sum;
    }
    for (int i = size - 1; i >= 0; i--) {
      float sum = b[i];
      for (int jj = i + 1; jj < size; jj++) sum -= elem[i][jj] * b[jj];
      b[i] = sum / elem[i][i];
    }
    for (int i = 0; i < size; i++) res[i][j] = b[i];
  }
}

__global__ void __launch_bounds__(64) k_hypotheses(const float *__restrict__ coord, const int *__restrict__ randPts,
                                                   float *__restrict__ homo, int numPts, int numLoops) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= numLoops) return;
  float a[8][8], ia[8][8], b[8];
  for (int i = 0; i < 4; i++) {
    const int pt = randPts[i * numLoops + idx];
    const float x1 = coord[pt + 0 * numPts], y1 = coord[pt + 1 * numPts];
    const float x2 = coord[pt + 2 * numPts], y2 = coord[pt + 3 * numPts];
    float *row1 = a[2 * i + 0];
    row1[0] = x1; row1[1] = y1; row1[2] = 1.0f;
    row1[3] = row1[4] = row1[5] = 0.0f;
    row1[6] = -x2 * x1; row1[7] = -x2 * y1;
    float *row2 = a[2 * i + 1];
    row2[0] = row2[1] = row2[2] = 0.0f;
    row2[3] = x1; row2[4] = y1; row2[5] = 1.0f;
    row2[6] = -y2 * x1; row2[7] = -y2 * y1;
    b[2 * i + 0] = x2;
    b[2 * i + 1] = y2;
  }
  invert8(a, ia);
  for (int j = 0; j < 8; j++) {
    float sum = 0.0f;
    for (int i = 0; i < 8; i++) sum += ia[j][i] * b[i];
    homo[j * numLoops + idx] = sum;
  }
}

// TestHomographies (homography.cu:139-192): inliers of every hypothesis over ALL numPts points.
// Thread = one hypothesis (its 8 coefficients in registers), CTA = 128 hypotheses x a slice of
// SCORE_PTS points read from shared memory as broadcasts; the integer partial counts are summed
// with atomicAdd, so the result does not depend on the slicing.
#define SCORE_HYP 128
#define SCORE_PTS 256
__global__ void __launch_bounds__(SCORE_HYP) k_score(const float *__restrict__ coord, const float *__restrict__ homo,
                                                     int *__restrict__ counts, int numPts, int numLoops, float thresh2) {
  __shared__ float4 pts[SCORE_PTS];
  const int p0 = blockIdx.y * SCORE_PTS;
  const int np = min(SCORE_PTS, numPts - p0);
  for (int i = threadIdx.x; i < np; i += SCORE_HYP)
    pts[i] = make_float4(coord[p0 + i + 0 * numPts], coord[p0 + i + 1 * numPts], coord[p0 + i + 2 * numPts],
                         coord[p0 + i + 3 * numPts]);
  __syncthreads();
  const int loop = blockIdx.x * SCORE_HYP + threadIdx.x;
  if (loop >= numLoops) return;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = homo[loop + i * numLoops];
  int cnt = 0;
#pragma unroll 4
  for (int i = 0; i < np; i++) {
    const float4 p = pts[i];
    const float x1 = p.x, y1 = p.y, x2 = p.z, y2 = p.w;
    // homography.cu:165-171, every product rounded toward zero
    const float nomx = __fadd_rn(__fadd_rn(__fmul_rz(a[0], x1), __fmul_rz(a[1], y1)), a[2]);
    const float nomy = __fadd_rn(__fadd_rn(__fmul_rz(a[3], x1), __fmul_rz(a[4], y1)), a[5]);
    const float deno = __fadd_rn(__fadd_rn(__fmul_rz(a[6], x1), __fmul_rz(a[7], y1)), 1.0f);
    const float errx = __fsub_rn(__fmul_rz(x2, deno), nomx);
    const float erry = __fsub_rn(__fmul_rz(y2, deno), nomy);
    const float err2 = __fadd_rn(__fmul_rz(errx, errx), __fmul_rz(erry, erry));
    if (err2 < __fmul_rz(thresh2, __fmul_rz(deno, deno))) cnt++;
  }
  if (cnt) atomicAdd(counts + loop, cnt);
}

// ---- batched (all-pairs) RANSAC support: everything stays on the device ----------------------
// Valid points (score > min_score && ambiguity < max_ambiguity, homography.cu:225-228) in increasing
// index order, like the reference's validPts.  One CTA, ordered block scan.
__global__ void __launch_bounds__(1024) k_valid_compact(const csb_sift_point *__restrict__ d_sift, int n, float min_score,
                                                        float max_amb, int *__restrict__ valid, int *__restrict__ n_valid) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // thread t owns the contiguous range [t*ipt, (t+1)*ipt): one block-wide exclusive scan orders everything
  const int ipt = (n + 1023) / 1024;
  const int i0 = threadIdx.x * ipt, i1 = min(n, i0 + ipt);
  unsigned int mask = 0;
  int cnt = 0;
  for (int i = i0; i < i1; i++) {
    const bool ok = d_sift[i].score > min_score && d_sift[i].ambiguity < max_amb;
    if (ok) {
      if (i - i0 < 32) mask |= 1u << (i - i0);
      cnt++;
    }
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    warp_sums[lane] = w;                       // inclusive over warps
  }
  __syncthreads();
  int pos = incl - cnt + (warp > 0 ? warp_sums[warp - 1] : 0);
  for (int i = i0; i < i1; i++) {
    const bool ok = (i - i0 < 32) ? ((mask >> (i - i0)) & 1u) != 0
                                  : (d_sift[i].score > min_score && d_sift[i].ambiguity < max_amb);
    if (ok) valid[pos++] = i;
  }
  if (threadIdx.x == 0) *n_valid = warp_sums[31];
}

// Counter-based generator for the 4-point samples (the reference draws them with libc rand(),
// homography.cu:232-244 — unseeded and continuing across calls, so no particular sequence is part
// of its contract).  csb_sample_hash is restated in Python by the tests.
__host__ __device__ inline unsigned int csb_hash5(unsigned int seed, unsigned int pair, unsigned int loop,
                                                        unsigned int k, unsigned int attempt) {
  unsigned int x = seed ^ (pair * 0x9E3779B9u) ^ (loop * 0x85EBCA6Bu) ^ (k * 0xC2B2AE35u) ^ (attempt * 0x27D4EB2Fu);
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

__global__ void k_make_samples(const int *__restrict__ valid, const int *__restrict__ n_valid, int num_loops,
                               unsigned int seed, unsigned int pair, int *__restrict__ rand_pts) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= num_loops) return;
  const int nv = *n_valid;
  int p[4] = {0, 0, 0, 0};
  if (nv >= 8) {
    for (int k = 0; k < 4; k++) {
      unsigned int attempt = 0;
      for (;;) {
        const int c = (int)(csb_hash5(seed, pair, (unsigned int)l, (unsigned int)k, attempt++) % (unsigned int)nv);
        bool dup = false;
        for (int q = 0; q < k; q++) dup = dup || (p[q] == c);
        if (!dup) { p[k] = c; break; }
      }
    }
    for (int k = 0; k < 4; k++) rand_pts[k * num_loops + l] = valid[p[k]];
  } else {
    for (int k = 0; k < 4; k++) rand_pts[k * num_loops + l] = 0;
  }
}

// First-maximum hypothesis (homography.cu:259-264) -> result record {H[9], inliers, n_valid} on the device.
__global__ void __launch_bounds__(256) k_pick_best(const int *__restrict__ counts, const float *__restrict__ homo,
                                                   int num_loops, const int *__restrict__ n_valid, int n_pts,
                                                   float *__restrict__ H_out, int *__restrict__ inl_out,
                                                   int *__restrict__ nvalid_out) {
  __shared__ int s_cnt[256], s_idx[256];
  int best = -1, bidx = 0x7fffffff;
  for (int i = threadIdx.x; i < num_loops; i += 256) {
    const int c = counts[i];
    if (c > best) { best = c; bidx = i; }
  }
  s_cnt[threadIdx.x] = best;
  s_idx[threadIdx.x] = bidx;
  __syncthreads();
  for (int len = 128; len > 0; len >>= 1) {
    if (threadIdx.x < len) {
      const int oc = s_cnt[threadIdx.x + len], oi = s_idx[threadIdx.x + len];
      if (oc > s_cnt[threadIdx.x] || (oc == s_cnt[threadIdx.x] && oi < s_idx[threadIdx.x])) {
        s_cnt[threadIdx.x] = oc;
        s_idx[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  const int nv = *n_valid;
  const bool ok = nv >= 8 && n_pts >= 8;       // homography.cu:207,231: otherwise identity, 0 matches
  if (threadIdx.x < 9) {
    float v = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.0f : 0.0f;
    if (ok && threadIdx.x < 8) v = homo[threadIdx.x * num_loops + s_idx[0]];
    H_out[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) {
    *inl_out = ok ? s_cnt[0] : 0;
    *nvalid_out = nv;
  }
}

}  // namespace

unsigned int csb_sample_hash_host(unsigned int seed, unsigned int pair, unsigned int loop, unsigned int k,
                                  unsigned int attempt) {
  return csb_hash5(seed, pair, loop, k, attempt);
}

void launch_pair_ransac(const csb_sift_point *d_sift, int n, int n_up, float min_score, float max_amb, int *d_valid,
                        int *d_nvalid, float *d_coord, int *d_rand, float *d_homo, int *d_counts, int num_loops,
                        float thresh2, unsigned int seed, unsigned int pair, float *H_out, int *inl_out, int *nvalid_out,
                        cudaStream_t st) {
  k_valid_compact<<<1, 1024, 0, st>>>(d_sift, n, min_score, max_amb, d_valid, d_nvalid);
  k_make_samples<<<(num_loops + 127) / 128, 128, 0, st>>>(d_valid, d_nvalid, num_loops, seed, pair, d_rand);
  k_gather_coords<<<((n_up > num_loops ? n_up : num_loops) + 255) / 256, 256, 0, st>>>(d_sift, n, n_up, d_coord, d_counts, num_loops);
  k_hypotheses<<<(num_loops + 63) / 64, 64, 0, st>>>(d_coord, d_rand, d_homo, n_up, num_loops);
  k_score<<<dim3((num_loops + SCORE_HYP - 1) / SCORE_HYP, (n_up + SCORE_PTS - 1) / SCORE_PTS), SCORE_HYP, 0, st>>>(d_coord, d_homo, d_counts, n_up, num_loops, thresh2);
  k_pick_best<<<1, 256, 0, st>>>(d_counts, d_homo, num_loops, d_nvalid, n, H_out, inl_out, nvalid_out);
}

void launch_homography(const csb_sift_point *d_sift, int n, int n_up, float *d_coord, const int *d_rand, float *d_homo,
                       int *d_counts, int num_loops, float thresh2, cudaStream_t st) {
  k_gather_coords<<<((n_up > num_loops ? n_up : num_loops) + 255) / 256, 256, 0, st>>>(d_sift, n, n_up, d_coord, d_counts, num_loops);
  k_hypotheses<<<(num_loops + 63) / 64, 64, 0, st>>>(d_coord, d_rand, d_homo, n_up, num_loops);
  k_score<<<dim3((num_loops + SCORE_HYP - 1) / SCORE_HYP, (n_up + SCORE_PTS - 1) / SCORE_PTS), SCORE_HYP, 0, st>>>(d_coord, d_homo, d_counts, n_up, num_loops, thresh2);
}
