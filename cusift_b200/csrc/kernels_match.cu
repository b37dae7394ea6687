// Brute-force nearest / second-nearest neighbour over two SiftPoint sets — exact
// fp32 path.
//
// Replaces ComputeDistance (reference extras/matching.cu:52-98), ComputeL2Distance
// (:103-114) and FindMinCorr / FindMaxCorr (:194-270 / :116-192) with ONE kernel
// that never materialises the n1 x n2 score matrix (the reference allocates and
// frees it on every call, matching.cu:295-302,351).
//
// Bit-exactness: thread (tx,ty) of a CTA scores query ty against candidate
// column p2 = 16*chunk + tx with the reference's rotated k order
// (k = (i+tx)&127, an FFMA chain from 0) and keeps the running (best, second,
// argbest) over the columns p2 = tx (mod 16) — exactly the per-lane scan of
// FindMinCorr — followed by the same 16-lane tree (pairs t/t+8, t/t+4, t/t+2,
// t/t+1, strict comparisons), so score, ambiguity and match equal the
// reference's including its tie-breaking.
#include "csb_internal.h"

namespace {

template <bool kL2>
__device__ __forceinline__ bool better(float a, float b) {
  return kL2 ? (a < b) : (a > b);
}

template <bool kL2>
__global__ void __launch_bounds__(256) k_match(csb_sift_point *__restrict__ s1, int n1,
                                               const csb_sift_point *__restrict__ s2, int n2, int chunks,
                                               const int *__restrict__ block_list, const int *__restrict__ block_count) {
  __shared__ float A[16][128];
  __shared__ float B[16][128];
  const int tx = threadIdx.x, ty = threadIdx.y;
  // optional indirection: only the listed blocks of 16 queries (redo pass of the tensor-core matcher)
  if (block_list != nullptr && (int)blockIdx.x >= *block_count) return;
  const int qb = block_list != nullptr ? block_list[blockIdx.x] : (int)blockIdx.x;

  {
    const float *ptr1 = s1[min(n1 - 1, qb * 16 + ty)].data;
#pragma unroll
    for (int i = 0; i < 8; i++) A[ty][16 * i + tx] = ptr1[16 * i + tx];
  }
  const float init = kL2 ? 999.0f : -1.0f;   // FLT_MAX is #defined to 999.0 in matching.cu:43
  float best = init, second = init;
  int idx = -1;

  for (int c = 0; c < chunks; c++) {
    __syncthreads();
    {
      const float *ptr2 = s2[min(n2 - 1, c * 16 + ty)].data;
#pragma unroll
      for (int i = 0; i < 8; i++) B[ty][16 * i + tx] = ptr2[16 * i + tx];
    }
    __syncthreads();
    float sum = 0.0f;
#pragma unroll 16
    for (int i = 0; i < 128; i++) {
      const int k = (i + tx) & 127;
      sum = __fmaf_rn(A[ty][k], B[tx][k], sum);
    }
    const int p2 = c * 16 + tx;
    float val = (p2 < n2) ? sum : -1.0f;
    if (kL2) val = (val > -1.0f) ? __fsub_rn(2.0f, __fadd_rn(val, val)) : 999.0f;
    if (better<kL2>(val, best)) {
      second = best;
      best = val;
      idx = p2;
    } else if (better<kL2>(val, second)) {
      second = val;
    }
  }

  // 16-lane tree; a query row is one half-warp (lanes tx + 16*(ty&1))
#pragma unroll
  for (int len = 8; len > 0; len >>= 1) {
    const float ob = __shfl_down_sync(0xffffffffu, best, len, 16);
    const int oi = __shfl_down_sync(0xffffffffu, idx, len, 16);
    const float os = __shfl_down_sync(0xffffffffu, second, len, 16);
    if (better<kL2>(ob, best)) {
      second = best;
      best = ob;
      idx = oi;
    } else if (better<kL2>(ob, second)) {
      second = ob;
    }
    if (better<kL2>(os, second)) second = os;
  }

  const int p1 = qb * 16 + ty;
  if (tx == 0 && p1 < n1) {
    csb_sift_point *o = s1 + p1;
    o->score = best;
    if (kL2) o->ambiguity = (float)((double)best / ((double)second + 1e-6));
    else o->ambiguity = (float)((double)__fsub_rn(1.0f, best) / ((double)__fsub_rn(1.0f, second) + 1e-6));
    o->match = idx;
    if (idx >= 0) {
      o->match_xpos = s2[idx].coords2D[0];
      o->match_ypos = s2[idx].coords2D[1];
    }
  }
}

}  // namespace

void launch_match(csb_sift_point *d_sift1, int n1, const csb_sift_point *d_sift2, int n2, int distance,
                  cudaStream_t st) {
  if (n1 <= 0 || n2 <= 0) return;
  dim3 blk(16, 16), grd((n1 + 15) / 16);
  const int chunks = (n2 + 15) / 16;
  if (distance == 1) k_match<true><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, nullptr, nullptr);
  else k_match<false><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, nullptr, nullptr);
}

// Exact pass restricted to the 16-query blocks in block_list[0 .. *block_count).
void launch_match_blocks(csb_sift_point *d_sift1, int n1, const csb_sift_point *d_sift2, int n2, int distance,
                         const int *block_list, const int *block_count, cudaStream_t st) {
  if (n1 <= 0 || n2 <= 0) return;
  dim3 blk(16, 16), grd((n1 + 15) / 16);
  const int chunks = (n2 + 15) / 16;
  if (distance == 1) k_match<true><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, block_list, block_count);
  else k_match<false><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, block_list, block_count);
}
