// RANSAC homography: hypothesis generation and scoring.
//
// Replaces ComputeHomographies + InvertMatrix<8> (reference
// extras/homography.cu:98-139, 12-96) and TestHomographies (:144-187), plus the
// four strided cudaMemcpy2D gathers of FindHomography (:246-249), with three
// launches on one stream and no intermediate host synchronisation.
//   k_gather_coords : AoS SiftPoint -> SoA x1,y1,x2,y2 (pad slots zeroed; the
//                     reference leaves them uninitialised, homography.cu:215)
//   k_hypotheses    : eight lanes per 4-point sample, the 8x8 DLT system in registers,
//                     Crout LU / implicit pivoting in the reference's operation order
//   k_score         : thread per hypothesis over a shared-memory slice of the points, inlier
//                     test with the reference's round-toward-zero products (__fmul_rz)
// The batched all-pairs path (everything stays on the device) needs three launches per pair as well:
//   k_ransac_prep   : valid-point list and hash-drawn samples (CTA 0), coordinates (the other CTAs)
//   k_hypotheses
//   k_score<true>   : the CTA that finishes last also picks the first-maximum hypothesis
// and k_improve_homography (ImproveHomography, homography.cu:271-337) runs as one 8-CTA cluster per pair.
#include <cooperative_groups.h>
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

// ComputeHomographies + InvertMatrix<8> (homography.cu:98-139, 12-96) as a warp-cooperative kernel:
// EIGHT LANES PER HYPOTHESIS, lane r owning row r of the 8x8 DLT system in eight registers (the reference
// gives every hypothesis one thread and two 8x8 matrices in local memory).  All loops are unrolled, every
// register index is static, rows travel by shuffles: no local memory at all.
//   * Crout LU, column by column: the finished upper entry of row k is broadcast and rows below it subtract
//     their multiple - the same fused multiply-adds, in the same order (k ascending), as the reference's
//     inner loops, so the factors are bit-identical;
//   * implicit-scaling pivot search = 3-step shuffle arg-max over the group with the reference's tie rule
//     (">=" while scanning upwards: the LAST maximal row wins; no candidate at all, e.g. NaNs, keeps the
//     previous column's pivot row); the row swap is one shuffle per register;
//   * the inverse is never formed row by row: lane j solves L U x = P e_j for ITS column of the inverse (the
//     reference's "skip leading zeros" rule included), reading the factors by broadcast;
//   * h = A^-1 b accumulates in the reference's order (i ascending).
// Numerics that are part of the contract: reciprocals through double ((float)(1.0 / (double)x)), IEEE division
// in the back substitution, a zero pivot replaced by 1e-16.
__device__ __forceinline__ float recip_via_double(float x) { return (float)(1.0 / (double)x); }

__global__ void __launch_bounds__(128) k_hypotheses(const float *__restrict__ coord, const int *__restrict__ randPts,
                                                    float *__restrict__ homo, int numPts, int numLoops) {
  constexpr unsigned FULLM = 0xffffffffu;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int idx = gtid >> 3;                        // hypothesis
  const int r = threadIdx.x & 7;                    // row owned by this lane
  const bool live = idx < numLoops;                 // whole 8-lane groups are live or not; dead groups compute on zeros
  float e[8];                                       // row r of the system / of its LU factors
  float rhs = 0.0f;
  {
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    if (live) {
      const int pt = randPts[(r >> 1) * numLoops + idx];
      x1 = coord[pt + 0 * numPts]; y1 = coord[pt + 1 * numPts];
      x2 = coord[pt + 2 * numPts]; y2 = coord[pt + 3 * numPts];
    }
    const bool even = (r & 1) == 0;                 // homography.cu:107-129
    const float t = even ? x2 : y2;
    e[0] = even ? x1 : 0.0f; e[1] = even ? y1 : 0.0f; e[2] = even ? 1.0f : 0.0f;
    e[3] = even ? 0.0f : x1; e[4] = even ? 0.0f : y1; e[5] = even ? 0.0f : 1.0f;
    e[6] = __fmul_rn(-t, x1);
    e[7] = __fmul_rn(-t, y1);
    rhs = t;
  }
  // implicit scaling: vv = 1 / max |row|
  float vv;
  {
    float big = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float temp = fabsf(e[j]);
      if (temp > big) big = temp;
    }
    vv = big > 0.0f ? recip_via_double(big) : 1e16f;
  }
  int indx[8];
  int imax = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    // column j: rows below the finished entry of row k subtract (their entry in column k) * (that entry)
#pragma unroll
    for (int k = 0; k < j; k++) {
      const float u = __shfl_sync(FULLM, e[j], k, 8);
      if (r > k) e[j] = __fmaf_rn(-e[k], u, e[j]);
    }
    // pivot: last row i >= j with the largest vv[i] * |a[i][j]|
    {
      float d = vv * fabsf(e[j]);
      int di = r;
      if (!(r >= j && d >= 0.0f)) { d = -1.0f; di = -1; }     // not a candidate (row above the diagonal, or NaN)
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(FULLM, d, o, 8);
        const int oi = __shfl_xor_sync(FULLM, di, o, 8);
        if (od > d || (od == d && oi > di)) { d = od; di = oi; }
      }
      if (di >= 0) imax = di;
    }
    {
      // row swap j <-> imax.  imax differs between the four hypotheses of a warp, so the shuffles are executed
      // unconditionally (a group that needs no swap reads its own rows back)
      const int src = (r == j) ? imax : ((r == imax) ? j : r);
#pragma unroll
      for (int c = 0; c < 8; c++) e[c] = __shfl_sync(FULLM, e[c], src, 8);
      const float vj = __shfl_sync(FULLM, vv, j, 8);
      if (r == imax) vv = vj;                        // vv[imax] = vv[j] (vv[j] itself is not needed again)
    }
    indx[j] = imax;
    if (r == j && e[j] == 0.0f) e[j] = 1e-16f;
    if (j != 7) {
      const float piv = __shfl_sync(FULLM, e[j], j, 8);
      const float dum = recip_via_double(piv);
      if (r > j) e[j] = __fmul_rn(e[j], dum);
    }
  }
  // lane r solves for column r of the inverse: b = e_r, permuted by the recorded row swaps
  int pos = r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int ip = indx[i];
    if (pos == i) pos = ip;
    else if (pos == ip) pos = i;
  }
  float y[8];
  int ii = -1;
#pragma unroll
  for (int i = 0; i < 8; i++) {                      // forward substitution (unit lower factor)
    float sum = (pos == i) ? 1.0f : 0.0f;
#pragma unroll
    for (int jj = 0; jj < i; jj++) {
      const float l = __shfl_sync(FULLM, e[jj], i, 8);            // a[i][jj]
      if (ii != -1 && jj >= ii) sum = __fmaf_rn(-l, y[jj], sum);
    }
    if (ii == -1 && sum != 0.0f) ii = i;
    y[i] = sum;
  }
#pragma unroll
  for (int i = 7; i >= 0; i--) {                     // back substitution
    float sum = y[i];
#pragma unroll
    for (int jj = i + 1; jj < 8; jj++) {
      const float u = __shfl_sync(FULLM, e[jj], i, 8);            // a[i][jj]
      sum = __fmaf_rn(-u, y[jj], sum);
    }
    const float diag = __shfl_sync(FULLM, e[i], i, 8);
    y[i] = __fdiv_rn(sum, diag);
  }
  // y[i] = inverse[i][r].  h[j] = sum_i inverse[j][i] * b[i], i ascending (homography.cu:133-138)
#pragma unroll
  for (int j = 0; j < 8; j++) {
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float ia = __shfl_sync(FULLM, y[j], i, 8);
      const float bi = __shfl_sync(FULLM, rhs, i, 8);
      sum = __fmaf_rn(ia, bi, sum);
    }
    if (live && r == j) homo[j * numLoops + idx] = sum;
  }
}

// ImproveHomography (homography.cu:280-346) on the device, one cluster of CTAs per homography: iteratively re-weighted
// least squares, weights limit / (err + limit), the 8x8 normal equations accumulated in fp64 and solved by
// Cholesky, then the inlier count and match_error of every point.  Per-point arithmetic follows the reference
// (projection in double rounded to float, float error and weight); the fp64 sums are formed in a different
// order (per-thread partial sums, then a tree), so H agrees with the host version to ~1e-12 relative, not bit
// for bit.  The reference runs this on the host with OpenCV (cv::solve, DECOMP_CHOLESKY).
struct ImproveJob {
  csb_sift_point *pts;
  int n;
  const float *H_in;        // 9 floats (device)
  float *H_out;             // 9 floats (device)
  int *numfit_out;
};
#define IH_NT 256
#define IH_CL 8             // CTAs per job: one thread-block cluster
#define IH_PPT 4            // points per thread kept in registers
#define IH_NACC 44          // 36 upper-triangle entries of M + 8 of X
__device__ __forceinline__ int ih_tri(int r, int c) { return r * 8 - r * (r - 1) / 2 + (c - r); }   // r <= c

// One CLUSTER of IH_CL CTAs per job (the first version ran a job on one CTA: 270 us for 8192 points and 5 loops, all of it
// latency - 32 points per thread and loop, each a chain of fp64 divisions and multiply-adds).  Every CTA accumulates the
// normal equations over its share of the points; the partial sums meet in CTA 0 through distributed shared memory
// (fixed order: results do not depend on timing), thread 0 there solves the 8 x 8 system, and the new H travels back
// the same way.  Two cluster barriers per loop.
__global__ void __cluster_dims__(IH_CL, 1, 1) __launch_bounds__(IH_NT)
    k_improve_homography(const ImproveJob *__restrict__ jobs, int num_loops, float min_score, float max_amb, float limit) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s_part[IH_NT / 32][IH_NACC];
  __shared__ double s_cta[IH_NACC];      // this CTA's sums (read by CTA 0)
  __shared__ double s_tot[IH_NACC];      // CTA 0: sums over the cluster
  __shared__ double s_new[8];            // CTA 0: the solution of this loop (read by everybody)
  __shared__ double s_A[8];
  __shared__ int s_cnt[IH_NT / 32];
  __shared__ int s_fit;
  const ImproveJob J = jobs[blockIdx.x / IH_CL];
  const int rank = (int)cluster.block_rank();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 8) s_A[threadIdx.x] = (double)(J.H_in[threadIdx.x] / J.H_in[8]);   // float division, as the reference
  // the thread's first IH_PPT points stay in registers for all loops (8192 points = 4 per thread): the loops then
  // run without memory latency and the points' division chains overlap
  float px[IH_PPT], py[IH_PPT], pmx[IH_PPT], pmy[IH_PPT];
  bool pok[IH_PPT];
#pragma unroll
  for (int k = 0; k < IH_PPT; k++) {
    const int i = rank * IH_NT + threadIdx.x + k * IH_CL * IH_NT;
    pok[k] = false;
    px[k] = py[k] = pmx[k] = pmy[k] = 0.0f;
    if (i < J.n) {
      const csb_sift_point &pt = J.pts[i];
      pok[k] = !(pt.score < min_score || pt.ambiguity > max_amb);
      px[k] = pt.coords2D[0], py[k] = pt.coords2D[1], pmx[k] = pt.match_xpos, pmy[k] = pt.match_ypos;
    }
  }
  __syncthreads();
  for (int loop = 0; loop < num_loops; loop++) {
    double A[8];
#pragma unroll
    for (int i = 0; i < 8; i++) A[i] = s_A[i];
    double acc[IH_NACC];
#pragma unroll
    for (int i = 0; i < IH_NACC; i++) acc[i] = 0.0;
    auto add_point = [&](float x, float y, float mx, float my) {
      const float den = (float)(A[6] * x + A[7] * y + 1.0f);
      const float dx = (float)((A[0] * x + A[1] * y + A[2]) / den - mx);
      const float dy = (float)((A[3] * x + A[4] * y + A[5]) / den - my);
      const float err = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      const double wei = (double)__fdiv_rn(limit, __fadd_rn(err, limit));
      // two rows of the design matrix: Yx = [x y 1 0 0 0 -x*mx -y*mx], Yy = [0 0 0 x y 1 -x*my -y*my]
      const double Yx[8] = {x, y, 1.0, 0.0, 0.0, 0.0, (double)__fmul_rn(-x, mx), (double)__fmul_rn(-y, mx)};
      const double Yy[8] = {0.0, 0.0, 0.0, x, y, 1.0, (double)__fmul_rn(-x, my), (double)__fmul_rn(-y, my)};
      const double ax = (double)mx * wei, ay = (double)my * wei;
      // M += (Y Y^T) wei for both rows, upper triangle only; products with a structural zero are skipped (they add 0)
#pragma unroll
      for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int c = r; c < 8; c++) {
          const bool in_x = (r < 3 || r >= 6) && (c < 3 || c >= 6), in_y = r >= 3 && c >= 3;   // compile-time after unrolling
          if (in_x) acc[ih_tri(r, c)] += (Yx[c] * Yx[r]) * wei;
          if (in_y) acc[ih_tri(r, c)] += (Yy[c] * Yy[r]) * wei;
        }
        if (r < 3 || r >= 6) acc[36 + r] += Yx[r] * ax;
        if (r >= 3) acc[36 + r] += Yy[r] * ay;
      }
    };
#pragma unroll
    for (int k = 0; k < IH_PPT; k++)
      if (pok[k]) add_point(px[k], py[k], pmx[k], pmy[k]);
    for (int i = rank * IH_NT + threadIdx.x + IH_PPT * IH_CL * IH_NT; i < J.n; i += IH_CL * IH_NT) {   // sets beyond 8192 points
      const csb_sift_point &pt = J.pts[i];
      if (pt.score < min_score || pt.ambiguity > max_amb) continue;
      add_point(pt.coords2D[0], pt.coords2D[1], pt.match_xpos, pt.match_ypos);
    }
#pragma unroll
    for (int i = 0; i < IH_NACC; i++) {
      double v = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_part[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < IH_NACC) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < IH_NT / 32; w++) v += s_part[w][threadIdx.x];
      s_cta[threadIdx.x] = v;
    }
    cluster.sync();                                  // every CTA's s_cta is complete
    if (rank == 0) {
      if (threadIdx.x < IH_NACC) {
        double v = 0.0;
        for (int r = 0; r < IH_CL; r++) v += cluster.map_shared_rank(s_cta, r)[threadIdx.x];
        s_tot[threadIdx.x] = v;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        // Cholesky M = L L^T with reciprocal diagonals (one rsqrt per column instead of a square root and up to
        // seven fp64 divisions, which were 8 us of serial latency per loop), then the two triangular solves.  Fully
        // unrolled: L lives in registers, M and X are read from shared memory where they are needed.
        double L[8][8], inv[8];
        bool spd = true;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          double d = s_tot[ih_tri(j, j)];
#pragma unroll
          for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
          spd = spd && (d > 0.0);
          inv[j] = rsqrt(d);
          L[j][j] = d * inv[j];
#pragma unroll
          for (int i = j + 1; i < 8; i++) {
            double sacc = s_tot[ih_tri(j, i)];
#pragma unroll
            for (int k = 0; k < j; k++) sacc -= L[i][k] * L[j][k];
            L[i][j] = sacc * inv[j];
          }
        }
        double Av[8], yv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          double sacc = s_tot[36 + i];
#pragma unroll
          for (int k = 0; k < i; k++) sacc -= L[i][k] * yv[k];
          yv[i] = sacc * inv[i];
        }
#pragma unroll
        for (int i = 7; i >= 0; i--) {
          double sacc = yv[i];
#pragma unroll
          for (int k = i + 1; k < 8; k++) sacc -= L[k][i] * Av[k];
          Av[i] = sacc * inv[i];
        }
        if (!spd) {                                  // not positive definite (NaNs above are discarded): H stays as it is
#pragma unroll
          for (int i = 0; i < 8; i++) Av[i] = s_A[i];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) s_new[i] = Av[i];
      }
    }
    cluster.sync();                                  // CTA 0's s_new is complete; nobody reads s_cta any more
    if (threadIdx.x < 8) s_A[threadIdx.x] = cluster.map_shared_rank(s_new, 0)[threadIdx.x];
    __syncthreads();
  }
  double A[8];
#pragma unroll
  for (int i = 0; i < 8; i++) A[i] = s_A[i];
  int numfit = 0;
  for (int i = rank * IH_NT + threadIdx.x; i < J.n; i += IH_CL * IH_NT) {
    csb_sift_point &pt = J.pts[i];
    const float x = pt.coords2D[0], y = pt.coords2D[1];
    const float den = (float)(A[6] * x + A[7] * y + 1.0);
    const float dx = (float)((A[0] * x + A[1] * y + A[2]) / den - pt.match_xpos);
    const float dy = (float)((A[3] * x + A[4] * y + A[5]) / den - pt.match_ypos);
    const float err = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (err < limit) numfit++;
    pt.match_error = (float)sqrt((double)err);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) numfit += __shfl_xor_sync(0xffffffffu, numfit, o);
  if (lane == 0) s_cnt[warp] = numfit;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < IH_NT / 32; w++) tot += s_cnt[w];
    s_fit = tot;
  }
  cluster.sync();
  if (rank == 0) {
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int r = 0; r < IH_CL; r++) tot += *cluster.map_shared_rank(&s_fit, r);
      *J.numfit_out = tot;
    }
    if (threadIdx.x < 8) J.H_out[threadIdx.x] = (float)A[threadIdx.x];
    if (threadIdx.x == 8) J.H_out[8] = 1.0f;
  }
  cluster.sync();                                    // remote shared memory stays valid until CTA 0 has read it
}

// TestHomographies (homography.cu:139-192): inliers of every hypothesis over ALL numPts points.
// Thread = one hypothesis (its 8 coefficients in registers), CTA = 128 hypotheses x a slice of
// SCORE_PTS points read from shared memory as broadcasts; the integer partial counts are summed
// with atomicAdd, so the result does not depend on the slicing.
#define SCORE_HYP 128
#define SCORE_PTS 64
// kPick (batched all-pairs path): the CTA that finishes last - a ticket counter in pick.ticket, reset for the next pair -
// also selects the first-maximum hypothesis (homography.cu:259-264) and writes the pair's result record
// {H[9], inliers, n_valid}: one launch less per pair than a separate arg-max kernel.
struct PickArgs {
  int *ticket;
  const int *n_valid;
  int n_pts;
  float *H_out;
  int *inl_out, *nvalid_out;
};
template <bool kPick>
__global__ void __launch_bounds__(SCORE_HYP) k_score(const float *__restrict__ coord, const float *__restrict__ homo,
                                                     int *__restrict__ counts, int numPts, int numLoops, float thresh2,
                                                     PickArgs pick) {
  __shared__ float4 pts[SCORE_PTS];
  const int p0 = blockIdx.y * SCORE_PTS;
  const int np = min(SCORE_PTS, numPts - p0);
  for (int i = threadIdx.x; i < np; i += SCORE_HYP)
    pts[i] = make_float4(coord[p0 + i + 0 * numPts], coord[p0 + i + 1 * numPts], coord[p0 + i + 2 * numPts],
                         coord[p0 + i + 3 * numPts]);
  __syncthreads();
  const int loop = blockIdx.x * SCORE_HYP + threadIdx.x;
  if (loop < numLoops) {
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
  if constexpr (kPick) {
    __shared__ int s_last;
    __shared__ int s_cnt[SCORE_HYP], s_idx[SCORE_HYP];
    __threadfence();                                 // this CTA's counts are visible before its ticket is
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(pick.ticket, 1) == (int)(gridDim.x * gridDim.y) - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    int best = -1, bidx = 0x7fffffff;
    for (int i = threadIdx.x; i < numLoops; i += SCORE_HYP) {
      const int c = __ldcg(counts + i);              // other CTAs' atomics: read at L2
      if (c > best) { best = c; bidx = i; }
    }
    s_cnt[threadIdx.x] = best;
    s_idx[threadIdx.x] = bidx;
    __syncthreads();
    for (int len = SCORE_HYP / 2; len > 0; len >>= 1) {
      if (threadIdx.x < len) {
        const int oc = s_cnt[threadIdx.x + len], oi = s_idx[threadIdx.x + len];
        if (oc > s_cnt[threadIdx.x] || (oc == s_cnt[threadIdx.x] && oi < s_idx[threadIdx.x])) {
          s_cnt[threadIdx.x] = oc;
          s_idx[threadIdx.x] = oi;
        }
      }
      __syncthreads();
    }
    const int nv = *pick.n_valid;
    const bool ok = nv >= 8 && pick.n_pts >= 8;      // homography.cu:207,231: otherwise identity, 0 matches
    if (threadIdx.x < 9) {
      float v = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.0f : 0.0f;
      if (ok && threadIdx.x < 8) v = homo[threadIdx.x * numLoops + s_idx[0]];
      pick.H_out[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) {
      *pick.inl_out = ok ? s_cnt[0] : 0;
      *pick.nvalid_out = nv;
      *pick.ticket = 0;
    }
  }
}

// ---- batched (all-pairs) RANSAC support: everything stays on the device ----------------------
// Counter-based generator for the 4-point samples (the reference draws them with libc rand(),
// homography.cu:232-244 — unseeded and continuing across calls, so no particular sequence is part
// of its contract).  csb_sample_hash is restated in Python by the tests.
__host__ __device__ inline unsigned int csb_hash5(unsigned int seed, unsigned int pair, unsigned int loop,
                                                        unsigned int k, unsigned int attempt) {
  unsigned int x = seed ^ (pair * 0x9E3779B9u) ^ (loop * 0x85EBCA6Bu) ^ (k * 0xC2B2AE35u) ^ (attempt * 0x27D4EB2Fu);
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

// ONE kernel prepares a pair's RANSAC (three launches in the first version); CTA 0 does steps 1 and 3, the other CTAs step 2:
//   1. valid points (score > min_score && ambiguity < max_ambiguity, homography.cu:225-228) in increasing index
//      order, like the reference's validPts: thread t owns a contiguous range, one block-wide exclusive scan orders
//      everything; the range's loads are all issued before the first is used;
//   2. AoS SiftPoint -> SoA x1, y1, x2, y2 (pad slots zeroed), hypothesis counters and the arg-max ticket zeroed;
//   3. the 4-point samples of every hypothesis, drawn from the valid list with the counter-based generator.
constexpr int PREP_IPT = 8;                          // points per thread whose loads are batched (n <= 8192)
__global__ void __launch_bounds__(1024) k_ransac_prep(const csb_sift_point *__restrict__ d_sift, int n, int n_up, float min_score,
                                                      float max_amb, int *__restrict__ valid, int *__restrict__ n_valid,
                                                      float *__restrict__ coord, int *__restrict__ counts, int num_loops,
                                                      unsigned int seed, unsigned int pair, int *__restrict__ rand_pts) {
  __shared__ int warp_sums[32];
  if (blockIdx.x > 0) {
    // CTAs 1 .. : coordinates and counters (no dependency on the valid list)
    const int nt = (gridDim.x - 1) * 1024, t = (blockIdx.x - 1) * 1024 + threadIdx.x;
    for (int i = t; i < num_loops; i += nt) counts[i] = 0;
    for (int i = t; i < n_up; i += nt) {
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
    return;
  }
  // CTA 0: valid list, then the samples
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ipt = (n + 1023) / 1024;
  const int i0 = threadIdx.x * ipt, i1 = min(n, i0 + ipt);
  unsigned int mask = 0;
  int cnt = 0;
  if (ipt <= PREP_IPT) {
    float sc[PREP_IPT], am[PREP_IPT];
#pragma unroll
    for (int k = 0; k < PREP_IPT; k++) {
      sc[k] = am[k] = 0.0f;
      if (i0 + k < i1) {
        sc[k] = d_sift[i0 + k].score;
        am[k] = d_sift[i0 + k].ambiguity;
      }
    }
#pragma unroll
    for (int k = 0; k < PREP_IPT; k++)
      if (i0 + k < i1 && sc[k] > min_score && am[k] < max_amb) {
        mask |= 1u << k;
        cnt++;
      }
  } else {
    for (int i = i0; i < i1; i++) {
      const bool ok = d_sift[i].score > min_score && d_sift[i].ambiguity < max_amb;
      if (ok) {
        if (i - i0 < 32) mask |= 1u << (i - i0);
        cnt++;
      }
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
  const int nv = warp_sums[31];
  if (threadIdx.x == 0) {
    n_valid[0] = nv;
    n_valid[1] = 0;                            // k_score's ticket
  }
  __syncthreads();                             // valid[] (written by this CTA) is complete
  // 3. samples
  for (int l = threadIdx.x; l < num_loops; l += 1024) {
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
  // d_nvalid: [0] = valid points, [1] = k_score's ticket (256 bytes of scratch)
  k_ransac_prep<<<1 + 8, 1024, 0, st>>>(d_sift, n, n_up, min_score, max_amb, d_valid, d_nvalid, d_coord, d_counts, num_loops, seed, pair,
                                    d_rand);
  k_hypotheses<<<(num_loops * 8 + 127) / 128, 128, 0, st>>>(d_coord, d_rand, d_homo, n_up, num_loops);
  const PickArgs pick{d_nvalid + 1, d_nvalid, n, H_out, inl_out, nvalid_out};
  k_score<true><<<dim3((num_loops + SCORE_HYP - 1) / SCORE_HYP, (n_up + SCORE_PTS - 1) / SCORE_PTS), SCORE_HYP, 0, st>>>(
      d_coord, d_homo, d_counts, n_up, num_loops, thresh2, pick);
}

void launch_homography(const csb_sift_point *d_sift, int n, int n_up, float *d_coord, const int *d_rand, float *d_homo,
                       int *d_counts, int num_loops, float thresh2, cudaStream_t st) {
  k_gather_coords<<<((n_up > num_loops ? n_up : num_loops) + 255) / 256, 256, 0, st>>>(d_sift, n, n_up, d_coord, d_counts, num_loops);
  k_hypotheses<<<(num_loops * 8 + 127) / 128, 128, 0, st>>>(d_coord, d_rand, d_homo, n_up, num_loops);
  k_score<false><<<dim3((num_loops + SCORE_HYP - 1) / SCORE_HYP, (n_up + SCORE_PTS - 1) / SCORE_PTS), SCORE_HYP, 0, st>>>(
      d_coord, d_homo, d_counts, n_up, num_loops, thresh2, PickArgs{});
}

void launch_improve_homography(const void *d_jobs, int n_jobs, int num_loops, float min_score, float max_amb, float limit,
                               cudaStream_t st) {
  if (n_jobs <= 0) return;
  k_improve_homography<<<n_jobs * IH_CL, IH_NT, 0, st>>>(reinterpret_cast<const ImproveJob *>(d_jobs), num_loops, min_score, max_amb,
                                                 limit);
}
size_t improve_job_bytes() { return sizeof(ImproveJob); }
void improve_job_fill(void *h_job, void *d_pts, int n, const float *d_H_in, float *d_H_out, int *d_numfit) {
  ImproveJob *j = reinterpret_cast<ImproveJob *>(h_job);
  j->pts = reinterpret_cast<csb_sift_point *>(d_pts);
  j->n = n;
  j->H_in = d_H_in;
  j->H_out = d_H_out;
  j->numfit_out = d_numfit;
}
