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

// score / ambiguity / match / match_xpos / match_ypos of one query (matching.cu:236-246, 352-356)
template <bool kL2>
__device__ __forceinline__ void write_match(csb_sift_point *o, const csb_sift_point *__restrict__ s2, float best,
                                            float second, int idx) {
  o->score = best;
  if (kL2) o->ambiguity = (float)((double)best / ((double)second + 1e-6));
  else o->ambiguity = (float)((double)__fsub_rn(1.0f, best) / ((double)__fsub_rn(1.0f, second) + 1e-6));
  o->match = idx;
  if (idx >= 0) {
    o->match_xpos = s2[idx].coords2D[0];
    o->match_ypos = s2[idx].coords2D[1];
  }
}

// kPartial: redo pass of the tensor-core matcher.  CTA (x, y) scores the x-th listed block of 16
// queries against candidate chunks [y*cps, (y+1)*cps) and stores the slice's (best, second, argbest)
// in part[]; k_match_finish merges the slices.  FindMinCorr's winner is the minimum of the total order
// (score, bitrev4(col % 16), col / 16) and its runner-up the best of the remaining scores, so the
// merge over slices gives the same result as the one long scan.
struct MatchPart {
  float best, second;
  int idx;
};

template <bool kL2, bool kPartial>
__global__ void __launch_bounds__(256) k_match(csb_sift_point *__restrict__ s1, int n1,
                                               const csb_sift_point *__restrict__ s2, int n2, int chunks,
                                               const int *__restrict__ block_list, const int *__restrict__ block_count,
                                               int cps, MatchPart *__restrict__ part) {
  // MG candidate chunks (of 16) are scored together: thread (tx, ty) runs MG independent chains - candidates tx,
  // tx + 16, ... of the super-chunk share the rotation tx, hence the query operand - which hides the FFMA latency
  // and quarters the barriers (one chain at a time cost 3 us per chunk: 50 us to redo a single block of 16 queries).
  constexpr int MG = 4;
  __shared__ float A[16][144];   // row stride 144: the two query rows of a warp (ty, ty + 1) read disjoint halves of the banks
  __shared__ float B[16 * MG][128];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n_blk = kPartial ? *block_count : (int)gridDim.x;
  for (int bi = blockIdx.x; bi < n_blk; bi += gridDim.x) {
  const int qb = kPartial ? block_list[bi] : bi;
  const int c_begin = kPartial ? (int)blockIdx.y * cps : 0;
  const int c_end = kPartial ? min(chunks, c_begin + cps) : chunks;

  if (kPartial) __syncthreads();   // A is rewritten per listed block
  {
    const float *ptr1 = s1[min(n1 - 1, qb * 16 + ty)].data;
#pragma unroll
    for (int i = 0; i < 8; i++) A[ty][16 * i + tx] = ptr1[16 * i + tx];
  }
  const float init = kL2 ? 999.0f : -1.0f;   // FLT_MAX is #defined to 999.0 in matching.cu:43
  float best = init, second = init;
  int idx = -1;

  float pre[MG][8];                               // next super-chunk of candidates, in flight during the dot products
  auto fetch = [&](int c) {                       // rows 16 (c + g) + ty, clamped (chunks past the end are never scored)
#pragma unroll
    for (int g = 0; g < MG; g++) {
      const float *ptr2 = s2[min(n2 - 1, (c + g) * 16 + ty)].data;
#pragma unroll
      for (int i = 0; i < 8; i++) pre[g][i] = ptr2[16 * i + tx];
    }
  };
  if (c_begin < c_end) fetch(c_begin);
  for (int c = c_begin; c < c_end; c += MG) {
    __syncthreads();
#pragma unroll
    for (int g = 0; g < MG; g++)
#pragma unroll
      for (int i = 0; i < 8; i++) B[16 * g + ty][16 * i + tx] = pre[g][i];
    __syncthreads();
    if (c + MG < c_end) fetch(c + MG);
    float sum[MG];
#pragma unroll
    for (int g = 0; g < MG; g++) sum[g] = 0.0f;
#pragma unroll 8
    for (int i = 0; i < 128; i++) {
      const int k = (i + tx) & 127;
      const float a = A[ty][k];
#pragma unroll
      for (int g = 0; g < MG; g++) sum[g] = __fmaf_rn(a, B[16 * g + tx][k], sum[g]);
    }
#pragma unroll
    for (int g = 0; g < MG; g++) {               // ascending columns: the running scan of FindMinCorr's lane
      if (c + g < c_end) {
        const int p2 = (c + g) * 16 + tx;
        float val = (p2 < n2) ? sum[g] : -1.0f;
        if (kL2) val = (val > -1.0f) ? __fsub_rn(2.0f, __fadd_rn(val, val)) : 999.0f;
        if (better<kL2>(val, best)) {
          second = best;
          best = val;
          idx = p2;
        } else if (better<kL2>(val, second)) {
          second = val;
        }
      }
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
  if (kPartial) {
    if (tx == 0) part[((size_t)bi * gridDim.y + blockIdx.y) * 16 + ty] = MatchPart{best, second, idx};
  } else if (tx == 0 && p1 < n1) {
    write_match<kL2>(s1 + p1, s2, best, second, idx);
  }
  }
}

__device__ __forceinline__ int bitrev4(int x) { return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3); }

// Merges the candidate slices of the redo pass: WARP = one query of one listed block, lane = slice (the first
// version walked the 32 slices in one thread: 32 dependent L2 round trips, 22 us for a handful of blocks).  The
// merge - winner = minimum of the total order (score, bitrev4(col % 16), col / 16), runner-up = best of all the
// other scores - is associative and commutative, so a shuffle tree gives the result of the serial scan.
template <bool kL2>
__global__ void __launch_bounds__(128) k_match_finish(csb_sift_point *__restrict__ s1, int n1, const csb_sift_point *__restrict__ s2,
                                                      const int *__restrict__ block_list, const int *__restrict__ block_count,
                                                      int n_slices, const MatchPart *__restrict__ part) {
  static_assert(CSB_REDO_SLICES <= 32, "one lane per slice");
  const int lane = threadIdx.x & 31;
  const int n_q = *block_count * 16;
  for (int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wq < n_q; wq += (gridDim.x * blockDim.x) >> 5) {
  const int b = wq >> 4, ty = wq & 15;
  const float init = kL2 ? 999.0f : -1.0f;
  float best = init, second = init;
  int idx = -1;
  if (lane < n_slices) {
    const MatchPart p = part[((size_t)b * n_slices + lane) * 16 + ty];
    if (p.idx >= 0) best = p.best, second = p.second, idx = p.idx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (oi < 0) continue;                            // the other side is empty (warp-uniform per pair of lanes: no shuffle follows)
    bool take;                                       // is the other side's winner preferred to ours ?
    if (idx < 0) take = true;
    else if (ob != best) take = better<kL2>(ob, best);
    else {
      const int ra = bitrev4(oi & 15), rb = bitrev4(idx & 15);
      take = ra != rb ? ra < rb : oi < idx;
    }
    float loser;
    if (take) {
      loser = best;
      best = ob;
      idx = oi;
    } else {
      loser = ob;
    }
    if (better<kL2>(loser, second)) second = loser;
    if (better<kL2>(os, second)) second = os;
  }
  const int p1 = block_list[b] * 16 + ty;
  if (lane == 0 && p1 < n1) write_match<kL2>(s1 + p1, s2, best, second, idx);
  }
}

}  // namespace

void launch_match(csb_sift_point *d_sift1, int n1, const csb_sift_point *d_sift2, int n2, int distance,
                  cudaStream_t st) {
  if (n1 <= 0 || n2 <= 0) return;
  dim3 blk(16, 16), grd((n1 + 15) / 16);
  const int chunks = (n2 + 15) / 16;
  if (distance == 1) k_match<true, false><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, nullptr, nullptr, 0, nullptr);
  else k_match<false, false><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, nullptr, nullptr, 0, nullptr);
}

// Exact pass restricted to the first min(*block_count, max_blocks) 16-query blocks of block_list,
// candidates cut into CSB_REDO_SLICES slices so that a handful of blocks does not serialise a whole
// scan on one SM.  part: match_redo_scratch_bytes(max_blocks) bytes of scratch.
size_t match_redo_scratch_bytes(int max_blocks) { return (size_t)max_blocks * CSB_REDO_SLICES * 16 * sizeof(MatchPart); }

void launch_match_blocks(csb_sift_point *d_sift1, int n1, const csb_sift_point *d_sift2, int n2, int distance,
                         const int *block_list, const int *block_count, int max_blocks, void *part, cudaStream_t st) {
  if (n1 <= 0 || n2 <= 0 || max_blocks <= 0) return;
  const int chunks = (n2 + 15) / 16;
  const int cps = (chunks + CSB_REDO_SLICES - 1) / CSB_REDO_SLICES;
  const int slices = (chunks + cps - 1) / cps;
  dim3 blk(16, 16), grd(max_blocks < 64 ? max_blocks : 64, slices);
  MatchPart *mp = (MatchPart *)part;
  const int fin_blocks = min((max_blocks * 16 + 3) / 4, 148 * 2);   // one warp per query, grid-stride
  if (distance == 1) {
    k_match<true, true><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, block_list, block_count, cps, mp);
    k_match_finish<true><<<fin_blocks, 128, 0, st>>>(d_sift1, n1, d_sift2, block_list, block_count, slices, mp);
  } else {
    k_match<false, true><<<grd, blk, 0, st>>>(d_sift1, n1, d_sift2, n2, chunks, block_list, block_count, cps, mp);
    k_match_finish<false><<<fin_blocks, 128, 0, st>>>(d_sift1, n1, d_sift2, block_list, block_count, slices, mp);
  }
}
