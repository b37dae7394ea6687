// Gaussian scale-space: 8 blur levels -> 7 DoG planes per octave, fused with the 2x downsample that
// seeds the next octave, so an octave base image is read from HBM once.
//
// Replaces (reference, danielsuo/cuSIFT):
//   LaplaceMulti_D  cuSIFT_D.cu:525-553  (8 x separable 9-tap on texture fetches + DoG)
//   ScaleDown_D     cuSIFT_D.cu:37-182   (5x5 separable blur + decimation)
//
// Results are bit-identical to the reference's: every multiply-add below is pinned with
// __fmul_rn/__fmaf_rn/__fadd_rn (or their packed f32x2 forms, each half an IEEE fp32 operation) in the
// order the reference's own sm_100a SASS evaluates it (FMUL k3*(a1+b1); FFMA c*k4; FFMA k2; k1; k0).
//
// Kernels in this file:
//   k_pyramid      the production kernel: ONE launch covers any set of octaves (a CTA looks its octave and
//                  tile up in a table).  The octave base reaches shared memory through 2-D tiled TMA loads
//                  (cp.async.bulk.tensor.2d -> UTMALDG, completion on mbarriers); the four warps of a CTA then
//                  run vertical pass, shared-memory exchange and horizontal pass + DoG autonomously.
//   k_down_chain   octave bases k+1 .. k+3 from base k in one launch (tile-local recomputation of the
//                  intermediate levels), so that all coarse octaves can then go through ONE k_pyramid launch.
//   k_blur_dog     scalar fallback (source image not 16-byte aligned / pitch not a multiple of 4 floats:
//                  TMA cannot address it), also selected by CSB_NO_FUSE=1.
//   k_scale_down   stand-alone ScaleDown (cuSIFT.h:76).
// Algorithmic HBM traffic of the pyramid: 4 B read + 28 B (+1 B next octave) written per pixel.
#include <cstdlib>

#include "csb_internal.h"
#include "tma_util.h"

namespace {

constexpr int TW = 120;       // output columns per strip
constexpr int NT = 128;       // threads = TW + 2*4 halo columns
constexpr int BATCH = 4;      // rows per shared-memory batch (= warps per CTA)
constexpr int ROWS = 32;      // output rows per CTA of the scalar kernel
constexpr int NLEV = CSB_NUM_LEVELS;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// cuSIFT_D.cu:536-540 / 544-548 in the reference's evaluation order.
__device__ __forceinline__ float tap9(const float (&k)[5], float c, float s1, float s2, float s3, float s4) {
  float t = __fmul_rn(k[3], s1);
  t = __fmaf_rn(c, k[4], t);
  t = __fmaf_rn(k[2], s2, t);
  t = __fmaf_rn(k[1], s3, t);
  t = __fmaf_rn(k[0], s4, t);
  return t;
}

// ScaleDown_D row filter (cuSIFT_D.cu:111-113): k0*(c-2+c+2) + k1*(c-1+c+1) + k2*c0.
__device__ __forceinline__ float down_h(float c0, float c1, float c2, float c3, float c4, float k0, float k1, float k2) {
  float t = __fmul_rn(__fadd_rn(c1, c3), k1);
  t = __fmaf_rn(__fadd_rn(c0, c4), k0, t);
  t = __fmaf_rn(c2, k2, t);
  return t;
}
// ScaleDown_D column filter, fork behaviour (cuSIFT_D.cu:123-125 and the four
// rotated copies): k2*r[2j] + k0*(r[2j+2]+r[2j+3]) + k1*(r[2j-1]+r[2j+1]).
__device__ __forceinline__ float down_v(float rm1, float r0, float r1, float r2, float r3, float k0, float k1, float k2) {
  float t = __fmul_rn(__fadd_rn(r2, r3), k0);
  t = __fmaf_rn(r0, k2, t);
  t = __fmaf_rn(__fadd_rn(rm1, r1), k1, t);
  return t;
}

// ---------------------------------------------------------------------------------
// Scalar fallback: a CTA owns 120 output columns x ROWS rows; thread t streams source column x0-4+t
// (clamped) downwards with the 9-row window in registers, the 8 vertically blurred levels of 4 rows go
// through shared memory, warp b filters row b horizontally.  Source pitch and DoG pitch are independent.
template <bool kDown>
__global__ void __launch_bounds__(NT) k_blur_dog(const float *__restrict__ src, int w, int h, int spitch,
                                                 float *__restrict__ dog, int dpitch, const __grid_constant__ DogWeights W,
                                                 float *__restrict__ next, int npitch, DownK dk) {
  __shared__ __align__(16) float V[BATCH][NLEV][NT];
  __shared__ float Raw[kDown ? BATCH : 1][NT];

  const int t = threadIdx.x;
  const int x0 = blockIdx.x * TW;
  const int y0 = blockIdx.y * ROWS;
  const int cx = clampi(x0 + t - 4, 0, w - 1);
  const size_t plane = CSB_DOG_PS(dpitch, h), drow = CSB_DOG_RS(dpitch);
  const int warp = t >> 5, lane = t & 31;

  float win[9];
#pragma unroll
  for (int i = 0; i < 9; i++) win[i] = 0.0f;
  float hw[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // down-sample window (threads < TW/2)

  float pre[BATCH];
#pragma unroll
  for (int b = 0; b < BATCH; b++) pre[b] = src[(size_t)clampi(y0 - 4 + b, 0, h - 1) * spitch + cx];

  constexpr int NB = (ROWS + 8) / BATCH;
  for (int nb = 0; nb < NB; nb++) {
    const int r0 = y0 - 4 + nb * BATCH;      // first source row of this batch
    float cur[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; b++) cur[b] = pre[b];
    if (nb + 1 < NB) {
#pragma unroll
      for (int b = 0; b < BATCH; b++) pre[b] = src[(size_t)clampi(r0 + BATCH + b, 0, h - 1) * spitch + cx];
    }
    const bool hasOut = nb >= 2;             // block-uniform
#pragma unroll
    for (int b = 0; b < BATCH; b++) {
#pragma unroll
      for (int i = 0; i < 8; i++) win[i] = win[i + 1];
      win[8] = cur[b];
      if constexpr (kDown) Raw[b][t] = cur[b];
      if (hasOut) {
        const float c = win[4];
        const float s1 = __fadd_rn(win[3], win[5]);
        const float s2 = __fadd_rn(win[2], win[6]);
        const float s3 = __fadd_rn(win[1], win[7]);
        const float s4 = __fadd_rn(win[0], win[8]);
#pragma unroll
        for (int s = 0; s < NLEV; s++) V[b][s][t] = tap9(W.k[s], c, s1, s2, s3, s4);
      }
    }
    __syncthreads();

    if (hasOut) {
      const int y = r0 - 4 + warp;           // output row of V[warp]
      const int xo = x0 + 4 * lane;
      if (lane < TW / 4 && y < h && xo < w) {
        float prev[4];
#pragma unroll
        for (int s = 0; s < NLEV; s++) {
          const float4 a = *reinterpret_cast<const float4 *>(&V[warp][s][4 * lane]);
          const float4 bq = *reinterpret_cast<const float4 *>(&V[warp][s][4 * lane + 4]);
          const float4 cq = *reinterpret_cast<const float4 *>(&V[warp][s][4 * lane + 8]);
          const float v[12] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w, cq.x, cq.y, cq.z, cq.w};
          float L[4];
#pragma unroll
          for (int j = 0; j < 4; j++)
            L[j] = tap9(W.k[s], v[j + 4], __fadd_rn(v[j + 3], v[j + 5]), __fadd_rn(v[j + 2], v[j + 6]),
                        __fadd_rn(v[j + 1], v[j + 7]), __fadd_rn(v[j], v[j + 8]));
          if (s > 0) {
            float *o = dog + (size_t)(s - 1) * plane + (size_t)y * drow + xo;
            const float d0 = __fsub_rn(prev[0], L[0]), d1 = __fsub_rn(prev[1], L[1]);
            const float d2 = __fsub_rn(prev[2], L[2]), d3 = __fsub_rn(prev[3], L[3]);
            if (xo + 3 < w) {
              *reinterpret_cast<float4 *>(o) = make_float4(d0, d1, d2, d3);
            } else {
              o[0] = d0;
              if (xo + 1 < w) o[1] = d1;
              if (xo + 2 < w) o[2] = d2;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; j++) prev[j] = L[j];
        }
      }
    }

    if constexpr (kDown) {
      if (t < TW / 2) {
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
          const int r = r0 + b;
          const float hv = down_h(Raw[b][2 * t + 2], Raw[b][2 * t + 3], Raw[b][2 * t + 4], Raw[b][2 * t + 5],
                                  Raw[b][2 * t + 6], dk.k0, dk.k1, dk.k2);
#pragma unroll
          for (int i = 0; i < 4; i++) hw[i] = hw[i + 1];
          hw[4] = hv;
          const int twoj = r - 3;            // window = rows r-4..r = 2j-1..2j+3
          if (twoj >= y0 && twoj < y0 + ROWS && !(twoj & 1)) {
            const int j = twoj >> 1, i = (x0 >> 1) + t;
            if (j < (h >> 1) && i < (w >> 1))
              next[(size_t)j * npitch + i] = down_v(hw[0], hw[1], hw[2], hw[3], hw[4], dk.k0, dk.k1, dk.k2);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------
// k_pyramid.  Blackwell issues two fp32 multiply-adds per instruction with FFMA2/FADD2/FMUL2
// (fma.rn.f32x2; measured 124 lane-ops/clk/SM, the same as scalar FFMA, in half the issue slots): each
// half is an ordinary IEEE single-precision operation, so results stay bit-identical.  A CTA therefore
// processes TWO adjacent 120-column strips in lock-step: every value is a float2 {strip A, strip B}.
//
// Source staging: the CTA's (rows+8) x 248 source rectangle arrives as 4-row chunks (one TMA box of
// 248 x 4 floats each) in a 3-slot ring; the three first chunks are requested up front, chunk n+3 when
// batch n has been consumed.  Thread = strip position: it reads its (clamped) column of both strips
// from the chunk, keeps a 12-row register window that rotates by renaming (the batch loop is unrolled
// three ways), writes the 8 vertically blurred levels of 4 rows to shared memory, then warp b filters
// row b horizontally (4 outputs x 2 strips per lane) and stores the 7 DoG rows as 128-bit words.
//
// Shared-memory layout of the vertically blurred rows: a row of 128 float2 positions = 32 quads of 32 B,
// no padding.  The horizontal pass reads one 16-byte chunk per lane with lane stride one quad (32 B),
// so lanes q and q+4 of a quarter-warp would share a bank group; the two 16-byte halves of every quad
// whose index has bit 2 set are therefore swapped (an XOR swizzle): 128-bit reads and the 64-bit writes
// of the vertical pass are both conflict-free.
#ifndef K1_MINB
#define K1_MINB 4             // resident CTAs per SM the register budget is sized for
#endif
constexpr int V2_QUAD = 4;                      // float2 slots per quad
constexpr int V2_ROW = (NT / 4) * V2_QUAD;      // float2 slots per (batch row, level)
constexpr int PY_SRC_COLS = 2 * TW + 8;         // staged source columns (both strips + halo)
constexpr int PY_SLOTS = 3;                     // ring of 4-row chunks
constexpr uint32_t PY_CHUNK_BYTES = PY_SRC_COLS * BATCH * sizeof(float);   // 3968 = 31 * 128
constexpr size_t PY_SMEM = sizeof(float2) * BATCH * NLEV * V2_ROW + PY_SLOTS * PY_CHUNK_BYTES + 64;

// The taps are SCALARS: written as {k, k} they compile to the scalar-broadcast operand form of the packed instructions
// (FFMA2 Rd, Ra.F32x2, URb.F32, Rc.F32x2), i.e. all 40 taps live in uniform registers and every packed operation reads
// two register pairs + one uniform register.  (Taps passed as pre-duplicated float2 pairs - the first version - occupy
// 80 uniform registers, more than there are, so half of them ended up in ordinary registers: a packed operation with
// three register-pair sources reads three registers from one bank and issues every 3 cycles instead of every 2.)
__device__ __forceinline__ float2 tap9x2(const float (&k)[5], float2 c, float2 s1, float2 s2, float2 s3, float2 s4) {
  float2 t = __fmul2_rn(make_float2(k[3], k[3]), s1);
  t = __ffma2_rn(c, make_float2(k[4], k[4]), t);
  t = __ffma2_rn(make_float2(k[2], k[2]), s2, t);
  t = __ffma2_rn(make_float2(k[1], k[1]), s3, t);
  t = __ffma2_rn(make_float2(k[0], k[0]), s4, t);
  return t;
}
__device__ __forceinline__ float2 down_h2(float2 c0, float2 c1, float2 c2, float2 c3, float2 c4, float2 k0, float2 k1,
                                          float2 k2) {
  float2 t = __fmul2_rn(__fadd2_rn(c1, c3), k1);
  t = __ffma2_rn(__fadd2_rn(c0, c4), k0, t);
  t = __ffma2_rn(c2, k2, t);
  return t;
}
__device__ __forceinline__ float2 down_v2(float2 rm1, float2 r0, float2 r1, float2 r2, float2 r3, float2 k0, float2 k1,
                                          float2 k2) {
  float2 t = __fmul2_rn(__fadd2_rn(r2, r3), k0);
  t = __ffma2_rn(r0, k2, t);
  t = __ffma2_rn(__fadd2_rn(rm1, r1), k1, t);
  return t;
}

// Vertical pass of one batch.  PH = batch index mod 3 fixes where the 12-row window starts in P[], so the
// rotation costs no register moves.  The batch holds source rows r0 .. r0+3; the window then covers rows
// r0-8 .. r0+3 and yields output rows r0-4 .. r0-1.
template <int PH, bool kDown>
__device__ __forceinline__ void py_vertical(float2 (&P)[12], float2 &last, const float *__restrict__ chunk,
                                            const float *__restrict__ row0, float *chunk_w, float *row0_w, int r0, int h,
                                            int colA, int colB, bool patchA, bool patchB, int pos, bool hasOut,
                                            const DogWeights &W, float2 *__restrict__ V, int vslot) {
#define PYWIN(i) P[((i) + 4 * PH) % 12]
  if (r0 >= 0 && r0 + BATCH - 1 <= h - 1) {    // CTA-uniform: all four rows inside the image (all but the first / last batch
                                               // of tiles on the top / bottom border): no row clamping
#pragma unroll
    for (int b = 0; b < BATCH; b++) {
      const float *src = chunk + b * PY_SRC_COLS;
      const float2 v = make_float2(src[colA], src[colB]);
      if constexpr (kDown) {
        if (patchA) chunk_w[b * PY_SRC_COLS + pos] = v.x;
        if (patchB) chunk_w[b * PY_SRC_COLS + pos + TW] = v.y;
      }
      PYWIN(8 + b) = v;
    }
    last = PYWIN(8 + BATCH - 1);
  } else
#pragma unroll
  for (int b = 0; b < BATCH; b++) {
    const int r = r0 + b;
    if (r <= h - 1) {                          // CTA-uniform; rows past the image repeat the last one (clamp)
      const float *src = (r < 0) ? row0 : chunk + b * PY_SRC_COLS;
      last = make_float2(src[colA], src[colB]);
      if constexpr (kDown) {
        // positions outside the image get the clamped value written back, so that the downsample below can
        // read its 5 taps without clamping (the TMA unit had zero-filled them)
        float *dst = (r < 0) ? row0_w : chunk_w + b * PY_SRC_COLS;
        if (patchA) dst[pos] = last.x;
        if (patchB) dst[pos + TW] = last.y;
      }
    }
    PYWIN(8 + b) = last;
  }
  if (hasOut) {
#pragma unroll
    for (int b = 0; b < BATCH; b++) {
      const float2 c = PYWIN(b + 4);
      const float2 s1 = __fadd2_rn(PYWIN(b + 3), PYWIN(b + 5));
      const float2 s2 = __fadd2_rn(PYWIN(b + 2), PYWIN(b + 6));
      const float2 s3 = __fadd2_rn(PYWIN(b + 1), PYWIN(b + 7));
      const float2 s4 = __fadd2_rn(PYWIN(b), PYWIN(b + 8));
#pragma unroll
      for (int s = 0; s < NLEV; s++) V[(b * NLEV + s) * V2_ROW + vslot] = tap9x2(W.k[s], c, s1, s2, s3, s4);
    }
  }
#undef PYWIN
}

// kMulti = false: the launch holds exactly one octave (index 0 is a compile-time constant, so the 40 filter
// taps are constant-bank operands at fixed offsets instead of indexed loads).
template <bool kDown, bool kMulti>
__global__ void __launch_bounds__(NT, K1_MINB) k_pyramid(const __grid_constant__ PyramidParams P,
                                                         const __grid_constant__ PyramidMaps TM) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *src_ring = reinterpret_cast<float *>(smem_raw);                                  // [PY_SLOTS][BATCH][PY_SRC_COLS]
  float2 *V = reinterpret_cast<float2 *>(smem_raw + PY_SLOTS * PY_CHUNK_BYTES);           // [BATCH][NLEV][V2_ROW]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + PY_SLOTS * PY_CHUNK_BYTES + sizeof(float2) * BATCH * NLEV * V2_ROW);

  // which octave does this CTA belong to?
  int oi = 0;
  if constexpr (kMulti) {
#pragma unroll 1
    for (int i = 1; i < P.n_oct; i++)
      if ((int)blockIdx.x >= P.oct[i].cta_begin) oi = i;
  }
  const PyramidOctave &O = P.oct[oi];
  const DogWeights &W = P.W[oi];
  const int w = O.w, h = O.h, rows = O.rows;
  const int local = (int)blockIdx.x - O.cta_begin;
  const int bx = local % O.tiles_x, by = local / O.tiles_x;

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int pos = t;                                          // vertical phase: thread -> strip position
  // float2 slot of this position in a V row: quad * 4 + (index in quad, halves swapped when quad bit 2 is set)
  const int vslot = (pos & ~3) + ((pos & 3) ^ (((pos >> 4) & 1) << 1));
  const int xA = bx * (2 * TW), xB = xA + TW;                 // first output column of strip A / B
  const int xs = xA - 4;                                      // image column of staged column 0
  const int y0 = by * rows;
  const int colA = clampi(xs + pos, 0, w - 1) - xs, colB = clampi(xB - 4 + pos, 0, w - 1) - xs;
  const bool patchA = colA != pos, patchB = colB != pos + TW;
  float *dog = O.dog;
  const size_t plane = CSB_DOG_PS(O.dpitch, h), drow = CSB_DOG_RS(O.dpitch);
  const int NB = rows / BATCH + 2;

  if (t == 0) {
#pragma unroll
    for (int s = 0; s < PY_SLOTS; s++) tma::mbar_init(full + s, 1);
    tma::mbar_fence_init();
#pragma unroll
    for (int s = 0; s < PY_SLOTS; s++) {                      // NB >= 3 always
      tma::mbar_expect_tx(full + s, PY_CHUNK_BYTES);
      tma::load_2d(src_ring + s * (BATCH * PY_SRC_COLS), &TM.m[oi], xs, y0 - 4 + s * BATCH, full + s);
    }
  }
  __syncthreads();

  float2 win[12];
#pragma unroll
  for (int i = 0; i < 12; i++) win[i] = make_float2(0.f, 0.f);
  float2 last = make_float2(0.f, 0.f);
  float2 hc[3];                                               // downsample: row-filtered rows r0-3 .. r0-1
#pragma unroll
  for (int i = 0; i < 3; i++) hc[i] = make_float2(0.f, 0.f);
  float2 hlast = make_float2(0.f, 0.f);
  const float2 dk0 = make_float2(P.dk.k0, P.dk.k0), dk1 = make_float2(P.dk.k1, P.dk.k1), dk2 = make_float2(P.dk.k2, P.dk.k2);
  float *row0 = src_ring + 1 * (BATCH * PY_SRC_COLS);         // image row 0 of a top-border tile: first row of chunk 1

  for (int nb = 0; nb < NB; nb++) {
    const int r0 = y0 - 4 + nb * BATCH;                       // first source row of this batch
    const int slot = nb % PY_SLOTS;
    float *chunk = src_ring + slot * (BATCH * PY_SRC_COLS);
    tma::mbar_wait(full + slot, (nb / PY_SLOTS) & 1);
    if (nb == 0 && y0 == 0) tma::mbar_wait(full + 1, 0);      // rows < 0 read image row 0, which is in chunk 1
    const bool hasOut = nb >= 2;                              // block-uniform: window full
    if (slot == 0)
      py_vertical<0, kDown>(win, last, chunk, row0, chunk, row0, r0, h, colA, colB, patchA, patchB, pos, hasOut, W, V, vslot);
    else if (slot == 1)
      py_vertical<1, kDown>(win, last, chunk, row0, chunk, row0, r0, h, colA, colB, patchA, patchB, pos, hasOut, W, V, vslot);
    else
      py_vertical<2, kDown>(win, last, chunk, row0, chunk, row0, r0, h, colA, colB, patchA, patchB, pos, hasOut, W, V, vslot);
    __syncthreads();

    if (hasOut) {
      // warp `warp` filters batch row `warp` horizontally; lane q -> outputs 4q..4q+3 of both strips
      const int y = r0 - 4 + warp;
      if (lane < TW / 4 && y < h) {
        const int xoA = xA + 4 * lane, xoB = xB + 4 * lane;
        const float4 *q4 = reinterpret_cast<const float4 *>(V + (size_t)(warp * NLEV) * V2_ROW);
        int ch[6];                                          // swizzled 16-byte chunk indices of quads lane .. lane+2
#pragma unroll
        for (int i = 0; i < 6; i++) ch[i] = 2 * (lane + (i >> 1)) + ((i & 1) ^ (((lane + (i >> 1)) >> 2) & 1));
        constexpr int LSTRIDE = V2_ROW / 2;                 // float4 units between levels
        float *oA = dog + (size_t)y * drow + xoA;           // plane 0; advanced by `plane` per level
        const bool fullA = xoA + 3 < w, fullB = xoB + 3 < w;
        const bool fast = __all_sync(__activemask(), fullA && fullB);
        float4 ld[6];                                       // software-pipelined shared loads (next level)
#pragma unroll
        for (int i = 0; i < 6; i++) ld[i] = q4[ch[i]];
        float2 prev[4];
#pragma unroll
        for (int s = 0; s < NLEV; s++) {
          float2 v[12];
#pragma unroll
          for (int i = 0; i < 6; i++) {
            v[2 * i + 0] = make_float2(ld[i].x, ld[i].y);
            v[2 * i + 1] = make_float2(ld[i].z, ld[i].w);
          }
          if (s + 1 < NLEV) {
#pragma unroll
            for (int i = 0; i < 6; i++) ld[i] = q4[(s + 1) * LSTRIDE + ch[i]];
          }
          float2 L[4];
#pragma unroll
          for (int j = 0; j < 4; j++)
            L[j] = tap9x2(W.k[s], v[j + 4], __fadd2_rn(v[j + 3], v[j + 5]), __fadd2_rn(v[j + 2], v[j + 6]),
                          __fadd2_rn(v[j + 1], v[j + 7]), __fadd2_rn(v[j], v[j + 8]));
          if (s > 0) {
            const float2 m1 = make_float2(-1.0f, -1.0f);
            float2 d[4];
#pragma unroll
            for (int j = 0; j < 4; j++) d[j] = __ffma2_rn(L[j], m1, prev[j]);   // prev - L, exactly rounded
            float *oB = oA + TW;
            if (fast) {
              *reinterpret_cast<float4 *>(oA) = make_float4(d[0].x, d[1].x, d[2].x, d[3].x);
              *reinterpret_cast<float4 *>(oB) = make_float4(d[0].y, d[1].y, d[2].y, d[3].y);
            } else {
              if (fullA) {
                *reinterpret_cast<float4 *>(oA) = make_float4(d[0].x, d[1].x, d[2].x, d[3].x);
              } else if (xoA < w) {
                oA[0] = d[0].x;
                if (xoA + 1 < w) oA[1] = d[1].x;
                if (xoA + 2 < w) oA[2] = d[2].x;
              }
              if (fullB) {
                *reinterpret_cast<float4 *>(oB) = make_float4(d[0].y, d[1].y, d[2].y, d[3].y);
              } else if (xoB < w) {
                oB[0] = d[0].y;
                if (xoB + 1 < w) oB[1] = d[1].y;
                if (xoB + 2 < w) oB[2] = d[2].y;
              }
            }
            oA += plane;
          }
#pragma unroll
          for (int j = 0; j < 4; j++) prev[j] = L[j];
        }
      }
    }

    if constexpr (kDown) {
      // threads 0..59: next-octave columns xA/2+t and xB/2+t.  The batch's rows r0 .. r0+3 are row-filtered
      // (5 taps read from the staged chunk, whose out-of-image positions the vertical pass patched); with the
      // three rows carried over, output rows (r0-2)/2 [rows r0-3..r0+1] and r0/2 [rows r0-1..r0+3] follow.
      if (t < TW / 2) {
        float2 hv[BATCH];
        if (r0 >= 0 && r0 + BATCH - 1 <= h - 1) {
#pragma unroll
          for (int b = 0; b < BATCH; b++) {
            const float *rw = chunk + b * PY_SRC_COLS + 2 * t + 2;
            hv[b] = down_h2(make_float2(rw[0], rw[TW]), make_float2(rw[1], rw[TW + 1]), make_float2(rw[2], rw[TW + 2]),
                            make_float2(rw[3], rw[TW + 3]), make_float2(rw[4], rw[TW + 4]), dk0, dk1, dk2);
          }
          hlast = hv[BATCH - 1];
        } else
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
          const int r = r0 + b;
          if (r <= h - 1) {
            const float *rw = ((r < 0) ? row0 : chunk + b * PY_SRC_COLS) + 2 * t + 2;
            hlast = down_h2(make_float2(rw[0], rw[TW]), make_float2(rw[1], rw[TW + 1]), make_float2(rw[2], rw[TW + 2]),
                            make_float2(rw[3], rw[TW + 3]), make_float2(rw[4], rw[TW + 4]), dk0, dk1, dk2);
          }
          hv[b] = hlast;
        }
        const int iA = (xA >> 1) + t, iB = (xB >> 1) + t;
        const int ja = (r0 - 2) >> 1, jb = r0 >> 1;
        if (nb >= 2 && ja < (h >> 1)) {                       // 2*ja = r0-2 in [y0, y0+rows)
          const float2 o = down_v2(hc[0], hc[1], hc[2], hv[0], hv[1], dk0, dk1, dk2);
          if (iA < (w >> 1)) O.next[(size_t)ja * O.npitch + iA] = o.x;
          if (iB < (w >> 1)) O.next[(size_t)ja * O.npitch + iB] = o.y;
        }
        if (nb >= 1 && nb + 1 < NB && jb < (h >> 1)) {        // 2*jb = r0 in [y0, y0+rows)
          const float2 o = down_v2(hc[2], hv[0], hv[1], hv[2], hv[3], dk0, dk1, dk2);
          if (iA < (w >> 1)) O.next[(size_t)jb * O.npitch + iA] = o.x;
          if (iB < (w >> 1)) O.next[(size_t)jb * O.npitch + iB] = o.y;
        }
        hc[0] = hv[1];
        hc[1] = hv[2];
        hc[2] = hv[3];
      }
    }
    __syncthreads();
    // the chunk has been consumed by everybody: refill its slot with the chunk three batches ahead
    if (t == 0 && nb + PY_SLOTS < NB) {
      tma::mbar_expect_tx(full + slot, PY_CHUNK_BYTES);
      tma::load_2d(chunk, &TM.m[oi], xs, r0 + PY_SLOTS * BATCH, full + slot);
    }
  }
}

// ---------------------------------------------------------------------------------
// k_down_chain<N>: octave bases k+1 .. k+N (N <= 3) from base k in ONE launch.  A CTA owns a 4x4 tile of the
// last level, the matching 8x8 / 16x16 tiles of the levels in between, and recomputes the halo each level
// needs from the level below inside shared memory (2T+3 -> 4T+9 -> 8T+21 pixels per side), so there is no
// grid-wide dependency between the levels.  Same arithmetic as ScaleDown_D (row filter, then the fork's
// column filter); out-of-image coordinates are clamped at every level like the reference's loads
// (cuSIFT_D.cu:65-67,78-79).
constexpr int DC_T = 4;
constexpr int DC_MAXL = 3;
constexpr int DC_NT = 256;
__host__ __device__ constexpr int dc_side(int levels_above) {   // region side at a level with `levels_above` coarser levels still to feed
  return levels_above == 0 ? DC_T : 2 * dc_side(levels_above - 1) + 3;
}

struct ChainLevel {
  float *ptr;
  int w, h, pitch;
};
struct ChainParams {
  const float *src;
  int sw, sh, spitch;
  ChainLevel lv[DC_MAXL];
  int tiles_x;
  DownK dk;
};

// one level: `in` holds the SI x SI region of level l-1 whose origin is (ixo, iyo); produces the SO x SO region of
// level l with origin (oxo, oyo) = ((ixo+2)/2, (iyo+1)/2) into `out` and stores the owned part.  kClamp = false:
// every region of the tile lies inside its image (all tiles but those on the image border), so region entry
// (ry, cx) of level l reads columns 2cx .. 2cx+4 and rows 2ry .. 2ry+4 of the region below, no clamping.
template <int SI, int SO, bool kClamp>
__device__ __forceinline__ void dc_level(const float *__restrict__ in, float *__restrict__ mid, float *__restrict__ out,
                                         int ixo, int iyo, int pw, int ph, int oxo, int oyo, const ChainLevel &L, int own,
                                         int own_x, int own_y, const DownK &dk) {
  const int cw = L.w, chh = L.h;
  for (int e = threadIdx.x; e < SI * SO; e += DC_NT) {
    const int ry = e / SO, cxo = e - ry * SO;
    const float *r = in + ry * SI;
    if constexpr (kClamp) {
      // out-of-image coordinates are evaluated at their clamped coordinate.  Region-relative indices are clamped to
      // the region too: tiles that exist only to cover an odd last column of a finer level hold no valid pixel of
      // the coarser ones, whose values are then unused.
      const int i = clampi(oxo + cxo, 0, cw - 1);
      auto rc = [&](int c) { return clampi(clampi(c, 0, pw - 1) - ixo, 0, SI - 1); };
      mid[e] = down_h(r[rc(2 * i - 2)], r[rc(2 * i - 1)], r[rc(2 * i)], r[rc(2 * i + 1)], r[rc(2 * i + 2)], dk.k0, dk.k1, dk.k2);
    } else {
      r += 2 * cxo;
      mid[e] = down_h(r[0], r[1], r[2], r[3], r[4], dk.k0, dk.k1, dk.k2);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < SO * SO; e += DC_NT) {
    const int ryo = e / SO, cxo = e - ryo * SO;
    float v;
    if constexpr (kClamp) {
      const int j = clampi(oyo + ryo, 0, chh - 1);
      const float *c = mid + cxo;
      auto rr = [&](int r) { return clampi(clampi(r, 0, ph - 1) - iyo, 0, SI - 1) * SO; };
      v = down_v(c[rr(2 * j - 1)], c[rr(2 * j)], c[rr(2 * j + 1)], c[rr(2 * j + 2)], c[rr(2 * j + 3)], dk.k0, dk.k1, dk.k2);
    } else {
      const float *c = mid + cxo + 2 * ryo * SO;
      v = down_v(c[0], c[SO], c[2 * SO], c[3 * SO], c[4 * SO], dk.k0, dk.k1, dk.k2);
    }
    out[e] = v;
    const int gx = oxo + cxo, gy = oyo + ryo;
    if (gx >= own_x && gx < own_x + own && gy >= own_y && gy < own_y + own && gx < cw && gy < chh)
      L.ptr[(size_t)gy * L.pitch + gx] = v;
  }
  __syncthreads();
}

template <int N, bool kClamp>
__device__ __forceinline__ void dc_levels(float *bufA, float *bufB, float *bufC, const int *ox, const int *oy,
                                          const ChainParams &C, int tx, int ty) {
  constexpr int S0 = dc_side(N), S1 = dc_side(N - 1), S2 = N >= 2 ? dc_side(N - 2) : 1, S3 = N >= 3 ? dc_side(N - 3) : 1;
  dc_level<S0, S1, kClamp>(bufA, bufB, bufC, ox[0], oy[0], C.sw, C.sh, ox[1], oy[1], C.lv[0], DC_T << (N - 1),
                           tx * (DC_T << (N - 1)), ty * (DC_T << (N - 1)), C.dk);
  if constexpr (N >= 2)
    dc_level<S1, S2, kClamp>(bufC, bufB, bufA, ox[1], oy[1], C.lv[0].w, C.lv[0].h, ox[2], oy[2], C.lv[1], DC_T << (N - 2),
                             tx * (DC_T << (N - 2)), ty * (DC_T << (N - 2)), C.dk);
  if constexpr (N >= 3)
    dc_level<S2, S3, kClamp>(bufA, bufB, bufC, ox[2], oy[2], C.lv[1].w, C.lv[1].h, ox[3], oy[3], C.lv[2], DC_T, tx * DC_T,
                             ty * DC_T, C.dk);
}

template <int N>
__global__ void __launch_bounds__(DC_NT) k_down_chain(const __grid_constant__ ChainParams C) {
  constexpr int S0 = dc_side(N), S1 = dc_side(N - 1);
  __shared__ float bufA[S0 * S0];          // level 0 region, later level 2
  __shared__ float bufB[S0 * S1];          // row-filtered intermediate [in_rows][out_cols]
  __shared__ float bufC[S1 * S1];          // level 1 region, later level 3
  const int tx = blockIdx.x % C.tiles_x, ty = blockIdx.x / C.tiles_x;
  // origin (column, row) of the region held at every level, from the last level backwards:
  // level l-1 needs columns 2i-2 .. 2i+2 and rows 2j-1 .. 2j+3 of what level l holds
  int ox[DC_MAXL + 1], oy[DC_MAXL + 1];
  ox[N] = tx * DC_T; oy[N] = ty * DC_T;
#pragma unroll
  for (int l = N - 1; l >= 0; l--) {
    ox[l] = 2 * ox[l + 1] - 2;
    oy[l] = 2 * oy[l + 1] - 1;
  }
  for (int e = threadIdx.x; e < S0 * S0; e += DC_NT) {
    const int ry = e / S0, rx = e - ry * S0;
    bufA[e] = C.src[(size_t)clampi(oy[0] + ry, 0, C.sh - 1) * C.spitch + clampi(ox[0] + rx, 0, C.sw - 1)];
  }
  __syncthreads();
  // interior tile: the region of every level (incl. the source) lies inside its image
  bool inside = ox[0] >= 0 && oy[0] >= 0 && ox[0] + S0 <= C.sw && oy[0] + S0 <= C.sh;
#pragma unroll
  for (int l = 1; l <= N; l++)
    inside = inside && ox[l] >= 0 && oy[l] >= 0 && ox[l] + dc_side(N - l) <= C.lv[l - 1].w && oy[l] + dc_side(N - l) <= C.lv[l - 1].h;
  if (inside) dc_levels<N, false>(bufA, bufB, bufC, ox, oy, C, tx, ty);
  else dc_levels<N, true>(bufA, bufB, bufC, ox, oy, C, tx, ty);
}

// Stand-alone ScaleDown (cuSIFT.h:76): one thread per destination pixel.
__global__ void k_scale_down(const float *__restrict__ src, int w, int h, int spitch, float *__restrict__ dst,
                             int dpitch, DownK dk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= (w >> 1) || j >= (h >> 1)) return;
  float hr[5];
#pragma unroll
  for (int d = 0; d < 5; d++) {
    const float *r = src + (size_t)clampi(2 * j - 1 + d, 0, h - 1) * spitch;
    hr[d] = down_h(r[clampi(2 * i - 2, 0, w - 1)], r[clampi(2 * i - 1, 0, w - 1)], r[clampi(2 * i, 0, w - 1)],
                   r[clampi(2 * i + 1, 0, w - 1)], r[clampi(2 * i + 2, 0, w - 1)], dk.k0, dk.k1, dk.k2);
  }
  dst[(size_t)j * dpitch + i] = down_v(hr[0], hr[1], hr[2], hr[3], hr[4], dk.k0, dk.k1, dk.k2);
}

}  // namespace

void launch_scale_down(const float *src, int w, int h, int spitch, float *dst, int dpitch, const float k[3],
                       cudaStream_t st) {
  dim3 blk(32, 8), grd(((w >> 1) + 31) / 32, ((h >> 1) + 7) / 8);
  if (grd.x == 0 || grd.y == 0) return;
  DownK dk{k[0], k[1], k[2]};
  k_scale_down<<<grd, blk, 0, st>>>(src, w, h, spitch, dst, dpitch, dk);
}

void launch_blur_dog(const float *base, int w, int h, int spitch, float *dog, int dpitch, const DogWeights &wts,
                     cudaStream_t st) {
  DownK dk{0.f, 0.f, 0.f};
  dim3 grd((w + TW - 1) / TW, (h + ROWS - 1) / ROWS);
  k_blur_dog<false><<<grd, NT, 0, st>>>(base, w, h, spitch, dog, dpitch, wts, nullptr, 0, dk);
}

void launch_blur_dog_down(const float *base, int w, int h, int spitch, float *dog, int dpitch, const DogWeights &wts,
                          float *next, int npitch, const float k[3], cudaStream_t st) {
  DownK dk{k[0], k[1], k[2]};
  dim3 grd((w + TW - 1) / TW, (h + ROWS - 1) / ROWS);
  k_blur_dog<true><<<grd, NT, 0, st>>>(base, w, h, spitch, dog, dpitch, wts, next, npitch, dk);
}

// ---- k_pyramid host side ------------------------------------------------------------------------
int pyramid_source_map(CUtensorMap *out, const float *base, int w, int h, int pitch) {
  return csb_tmap_2d_f32(out, base, (uint64_t)w, (uint64_t)h, (uint64_t)pitch * sizeof(float), PY_SRC_COLS, BATCH);
}

bool pyramid_tma_ok(const float *base, int pitch) {
  return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (pitch & 3) == 0;
}

void pyramid_set_weights(PyramidParams *pp, int idx, const DogWeights &wts) { pp->W[idx] = wts; }

// Rows per CTA for every octave of the launch (multiple of 4, <= 16): as large as possible while the whole
// launch still fits in one wave of K1_MINB CTAs per SM, so that a lone frame fills the GPU and a CTA's
// serial row loop is as short as the frame allows.
int plan_pyramid(PyramidParams *pp, int sm_count) {
  const long long slots = (long long)sm_count * K1_MINB;
  static int forced = -1;
  if (forced < 0) {
    const char *e = getenv("CSB_K1_ROWS");
    forced = e ? atoi(e) : 0;
  }
  int rows = 16;
  for (; rows > 4; rows -= 4) {
    long long ctas = 0;
    for (int i = 0; i < pp->n_oct; i++)
      ctas += (long long)((pp->oct[i].w + 2 * TW - 1) / (2 * TW)) * ((pp->oct[i].h + rows - 1) / rows);
    if (ctas * 2 > slots) break;          // at least half a wave: do not shrink the tiles any further
  }
  if (forced >= 4 && forced <= 16 && forced % 4 == 0) rows = forced;
  int ctas = 0;
  for (int i = 0; i < pp->n_oct; i++) {
    PyramidOctave &o = pp->oct[i];
    o.rows = rows;
    o.tiles_x = (o.w + 2 * TW - 1) / (2 * TW);
    o.cta_begin = ctas;
    ctas += o.tiles_x * ((o.h + rows - 1) / rows);
  }
  return ctas;
}

void launch_pyramid(const PyramidParams &pp, const PyramidMaps &maps, int n_ctas, bool down, cudaStream_t st) {
  if (n_ctas <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_pyramid<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PY_SMEM);
    cudaFuncSetAttribute(k_pyramid<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PY_SMEM);
    cudaFuncSetAttribute(k_pyramid<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PY_SMEM);
    attr_set = true;
  }
  if (pp.n_oct > 1) k_pyramid<false, true><<<n_ctas, NT, PY_SMEM, st>>>(pp, maps);     // (never combined with `down`)
  else if (down) k_pyramid<true, false><<<n_ctas, NT, PY_SMEM, st>>>(pp, maps);
  else k_pyramid<false, false><<<n_ctas, NT, PY_SMEM, st>>>(pp, maps);
}

// dst[0..n-1] = successive half-size levels below `src`; at most 3 levels per launch (the caller's list is cut up)
void launch_down_chain(const float *src, int sw, int sh, int spitch, float *const *dst, const int *dw, const int *dh,
                       const int *dpitch, int n, const float k[3], cudaStream_t st) {
  while (n > 0) {
    const int m = n < DC_MAXL ? n : DC_MAXL;
    ChainParams C;
    C.src = src; C.sw = sw; C.sh = sh; C.spitch = spitch;
    for (int l = 0; l < DC_MAXL; l++) C.lv[l] = l < m ? ChainLevel{dst[l], dw[l], dh[l], dpitch[l]} : ChainLevel{nullptr, 0, 0, 0};
    C.dk = DownK{k[0], k[1], k[2]};
    // tiles of the last level; a finer level with an odd size has one more pixel than twice the next one, so the
    // grid must also cover every finer level's extent
    int txn = 1, tyn = 1;
    for (int l = 0; l < m; l++) {
      const int own = DC_T << (m - 1 - l);
      txn = txn > (dw[l] + own - 1) / own ? txn : (dw[l] + own - 1) / own;
      tyn = tyn > (dh[l] + own - 1) / own ? tyn : (dh[l] + own - 1) / own;
    }
    C.tiles_x = txn;
    if (m == 1) k_down_chain<1><<<txn * tyn, DC_NT, 0, st>>>(C);
    else if (m == 2) k_down_chain<2><<<txn * tyn, DC_NT, 0, st>>>(C);
    else k_down_chain<3><<<txn * tyn, DC_NT, 0, st>>>(C);
    src = dst[m - 1]; sw = dw[m - 1]; sh = dh[m - 1]; spitch = dpitch[m - 1];
    dst += m; dw += m; dh += m; dpitch += m;
    n -= m;
  }
}
