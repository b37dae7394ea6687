// Gaussian scale-space for one octave: 8 blur levels -> 7 DoG planes, fused with
// the 2x downsample that seeds the next octave, so the octave base image is read
// from HBM once.
//
// Replaces (reference, danielsuo/cuSIFT):
//   LaplaceMulti_D  cuSIFT_D.cu:525-553  (8 x separable 9-tap on texture fetches + DoG)
//   ScaleDown_D     cuSIFT_D.cu:37-182   (5x5 separable blur + decimation)
//
// Results are bit-identical to the reference's: every multiply-add below is
// pinned with __fmul_rn/__fmaf_rn/__fadd_rn in the order the reference's own
// sm_100a SASS evaluates it (FMUL k3*(a1+b1); FFMA c*k4; FFMA k2; k1; k0).
//
// Data movement (scalar kernel k_blur_dog; the packed kernel k_blur_dog2 further down is the one the
// pipeline uses and processes two such strips as float2): a CTA owns a strip of 120 output columns x
// ROWS output rows.  Thread t streams source column x0-4+t (clamped) downwards with coalesced 512-B
// row reads, keeping the 9-row vertical window in registers.  Per batch of 4 rows the 8
// vertically-blurred levels go to shared memory ([4][8][128] floats), then warp b filters row b
// horizontally, 4 adjacent outputs per lane from three 128-bit shared loads per level, and stores the
// 7 DoG rows as 128-bit words (DoG layout: csb_internal.h, CSB_DOG_PS / CSB_DOG_RS).
// Algorithmic HBM traffic: 4 B read + 28 B (+1 B next octave) written per pixel.
#include <cstdlib>

#include "csb_internal.h"

namespace {

constexpr int TW = 120;       // output columns per CTA
constexpr int NT = 128;       // threads = TW + 2*4 halo columns
constexpr int BATCH = 4;      // rows per shared-memory batch (= warps per CTA)
constexpr int ROWS = 32;      // output rows per CTA
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

struct DownK {
  float k0, k1, k2;
};

template <bool kDown>
__global__ void __launch_bounds__(NT) k_blur_dog(const float *__restrict__ src, int w, int h, int pitch,
                                                 float *__restrict__ dog, const __grid_constant__ DogWeights W,
                                                 float *__restrict__ next, int npitch, DownK dk) {
  __shared__ __align__(16) float V[BATCH][NLEV][NT];
  __shared__ float Raw[kDown ? BATCH : 1][NT];

  const int t = threadIdx.x;
  const int x0 = blockIdx.x * TW;
  const int y0 = blockIdx.y * ROWS;
  const int cx = clampi(x0 + t - 4, 0, w - 1);
  const size_t plane = CSB_DOG_PS(pitch, h), drow = CSB_DOG_RS(pitch);
  const int warp = t >> 5, lane = t & 31;

  float win[9];
#pragma unroll
  for (int i = 0; i < 9; i++) win[i] = 0.0f;
  float hw[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // down-sample window (threads < TW/2)

  // source rows y0-4 .. y0+ROWS+3 stream through in batches of 4; the window is
  // full (an output row exists) from the third batch on.
  float pre[BATCH];
#pragma unroll
  for (int b = 0; b < BATCH; b++) pre[b] = src[(size_t)clampi(y0 - 4 + b, 0, h - 1) * pitch + cx];

  constexpr int NB = (ROWS + 8) / BATCH;
  for (int nb = 0; nb < NB; nb++) {
    const int r0 = y0 - 4 + nb * BATCH;      // first source row of this batch
    float cur[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; b++) cur[b] = pre[b];
    if (nb + 1 < NB) {
#pragma unroll
      for (int b = 0; b < BATCH; b++) pre[b] = src[(size_t)clampi(r0 + BATCH + b, 0, h - 1) * pitch + cx];
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
      // warp `warp` filters batch row `warp` horizontally; lane q -> outputs 4q..4q+3
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
      // threads 0..59: next-octave column x0/2+t, fed by source rows r0..r0+3
      if (t < TW / 2) {
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
          const int r = r0 + b;
          const float hv = down_h(Raw[b][2 * t + 2], Raw[b][2 * t + 3], Raw[b][2 * t + 4], Raw[b][2 * t + 5],
                                  Raw[b][2 * t + 6], dk.k0, dk.k1, dk.k2);
#pragma unroll
          for (int i = 0; i < 4; i++) hw[i] = hw[i + 1];
          hw[4] = hv;
          // window now holds rows r-4..r; output row j needs rows 2j-1..2j+3 = r-4..r
          const int twoj = r - 3;
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
// Packed-fp32 version (the one the pipeline uses).  Blackwell issues two fp32
// multiply-adds per instruction with FFMA2/FADD2/FMUL2 (fma.rn.f32x2): each half is an
// ordinary IEEE single-precision operation, so results stay bit-identical while the
// number of floating-point issue slots halves.  A CTA therefore processes TWO adjacent
// 120-column strips in lock-step: every value is a float2 {strip A, strip B}.
//
// Shared-memory layout of the vertically blurred rows: a row of 128 float2 positions = 32 quads of 32 B,
// no padding.  The horizontal pass reads one 16-byte chunk per lane with lane stride one quad (32 B),
// so lanes q and q+4 of a quarter-warp would share a bank group; the two 16-byte halves of every quad
// whose index has bit 2 set are therefore swapped (an XOR swizzle): 128-bit reads and the 64-bit writes
// of the vertical pass (16 consecutive positions = 128 contiguous bytes) are both conflict-free, and the
// tile is 32 KB instead of 48 KB, which lets more CTAs share an SM.
#ifndef K1_MINB
#define K1_MINB 4           // resident CTAs per SM the register budget is sized for
#endif
constexpr int V2_QUAD = 4;                      // float2 slots per quad
constexpr int V2_ROW = (NT / 4) * V2_QUAD;      // float2 slots per (batch row, level)
#ifndef K1_ROWS
#define K1_ROWS 16
#endif
#ifndef K1_ROWS_SMALL
#define K1_ROWS_SMALL 4
#endif
constexpr int ROWS2 = K1_ROWS;                  // output rows per CTA (large octaves)
constexpr int ROWS2_SMALL = K1_ROWS_SMALL;                  // ... when the octave would not fill the GPU otherwise: the
                                                // row loop is what a small octave's launch waits for
constexpr size_t K1V2_SMEM = sizeof(float2) * (BATCH * NLEV * V2_ROW + BATCH * NT);

struct DogWeights2 {
  float2 k[NLEV][5];   // each tap duplicated {k, k}
};

__device__ __forceinline__ float2 tap9x2(const float2 (&k)[5], float2 c, float2 s1, float2 s2, float2 s3, float2 s4) {
  float2 t = __fmul2_rn(k[3], s1);
  t = __ffma2_rn(c, k[4], t);
  t = __ffma2_rn(k[2], s2, t);
  t = __ffma2_rn(k[1], s3, t);
  t = __ffma2_rn(k[0], s4, t);
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

template <bool kDown, int kRows>
__global__ void __launch_bounds__(NT, K1_MINB) k_blur_dog2(const float *__restrict__ src, int w, int h, int pitch,
                                                  float *__restrict__ dog, const __grid_constant__ DogWeights2 W,
                                                  float *__restrict__ next, int npitch, DownK dk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2 *V = reinterpret_cast<float2 *>(smem_raw);          // [BATCH][NLEV][V2_ROW]
  float2 *Raw = V + BATCH * NLEV * V2_ROW;                    // [BATCH][NT], plain position order

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int pos = t;                                          // vertical phase: thread -> strip position
  // float2 slot of this position in a V row: quad * 4 + (index in quad, halves swapped when quad bit 2 is set)
  const int vslot = (pos & ~3) + ((pos & 3) ^ (((pos >> 4) & 1) << 1));
  const int xA = blockIdx.x * (2 * TW), xB = xA + TW;         // first output column of strip A / B
  const int y0 = blockIdx.y * kRows;
  const int cA = clampi(xA + pos - 4, 0, w - 1), cB = clampi(xB + pos - 4, 0, w - 1);
  const size_t plane = CSB_DOG_PS(pitch, h), drow = CSB_DOG_RS(pitch);

  float2 win[8 + BATCH];
#pragma unroll
  for (int i = 0; i < 8 + BATCH; i++) win[i] = make_float2(0.f, 0.f);
  float2 hw[5];
#pragma unroll
  for (int i = 0; i < 5; i++) hw[i] = make_float2(0.f, 0.f);
  const float2 dk0 = make_float2(dk.k0, dk.k0), dk1 = make_float2(dk.k1, dk.k1), dk2 = make_float2(dk.k2, dk.k2);

  // two batches of source rows are kept in flight ahead of the window (load latency >> batch time)
  float2 pre[BATCH], pre2[BATCH];
#pragma unroll
  for (int b = 0; b < BATCH; b++) {
    const float *r = src + (size_t)clampi(y0 - 4 + b, 0, h - 1) * pitch;
    pre[b] = make_float2(r[cA], r[cB]);
  }
#pragma unroll
  for (int b = 0; b < BATCH; b++) {
    const float *r = src + (size_t)clampi(y0 + b, 0, h - 1) * pitch;
    pre2[b] = make_float2(r[cA], r[cB]);
  }

  constexpr int NB = (kRows + 8) / BATCH;
  for (int nb = 0; nb < NB; nb++) {
    const int r0 = y0 - 4 + nb * BATCH;      // first source row of this batch
    // window holds source rows r0-8 .. r0+3 after this: shift by BATCH, append the batch
#pragma unroll
    for (int i = 0; i < 8; i++) win[i] = win[i + BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; b++) {
      win[8 + b] = pre[b];
      pre[b] = pre2[b];
    }
    if (nb + 2 < NB) {
#pragma unroll
      for (int b = 0; b < BATCH; b++) {
        const float *r = src + (size_t)clampi(r0 + 2 * BATCH + b, 0, h - 1) * pitch;
        pre2[b] = make_float2(r[cA], r[cB]);
      }
    }
    const bool hasOut = nb >= 2;             // block-uniform: window full
    if constexpr (kDown) {
#pragma unroll
      for (int b = 0; b < BATCH; b++) Raw[b * NT + pos] = win[8 + b];
    }
    if (hasOut) {
#pragma unroll
      for (int b = 0; b < BATCH; b++) {
        // output row r0+b-4: window rows b .. b+8, centre b+4
        const float2 c = win[b + 4];
        const float2 s1 = __fadd2_rn(win[b + 3], win[b + 5]);
        const float2 s2 = __fadd2_rn(win[b + 2], win[b + 6]);
        const float2 s3 = __fadd2_rn(win[b + 1], win[b + 7]);
        const float2 s4 = __fadd2_rn(win[b], win[b + 8]);
#pragma unroll
        for (int s = 0; s < NLEV; s++) V[(b * NLEV + s) * V2_ROW + vslot] = tap9x2(W.k[s], c, s1, s2, s3, s4);
      }
    }
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
      // threads 0..59: next-octave columns xA/2+t and xB/2+t, fed by source rows r0..r0+3
      if (t < TW / 2) {
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
          const int r = r0 + b;
          const float2 *rw = Raw + b * NT + 2 * t;
          const float2 hv = down_h2(rw[2], rw[3], rw[4], rw[5], rw[6], dk0, dk1, dk2);
#pragma unroll
          for (int i = 0; i < 4; i++) hw[i] = hw[i + 1];
          hw[4] = hv;
          const int twoj = r - 3;   // window = rows r-4..r = 2j-1..2j+3
          if (twoj >= y0 && twoj < y0 + kRows && !(twoj & 1)) {
            const int j = twoj >> 1;
            if (j < (h >> 1)) {
              const float2 o = down_v2(hw[0], hw[1], hw[2], hw[3], hw[4], dk0, dk1, dk2);
              const int iA = (xA >> 1) + t, iB = (xB >> 1) + t;
              if (iA < (w >> 1)) next[(size_t)j * npitch + iA] = o.x;
              if (iB < (w >> 1)) next[(size_t)j * npitch + iB] = o.y;
            }
          }
        }
      }
    }
    __syncthreads();
  }
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

// CSB_K1_SCALAR=1 selects the scalar-fp32 kernels above (debugging aid; same results).
static bool csb_use_scalar_pyramid() {
  static const bool v = [] {
    const char *e = getenv("CSB_K1_SCALAR");
    return e && e[0] == '1';
  }();
  return v;
}

namespace {
DogWeights2 dup_weights(const DogWeights &wts) {
  DogWeights2 w2;
  for (int s = 0; s < NLEV; s++)
    for (int j = 0; j < 5; j++) w2.k[s][j] = make_float2(wts.k[s][j], wts.k[s][j]);
  return w2;
}
template <bool kDown, int kRows>
void launch_blur_dog2(const float *base, int w, int h, int pitch, float *dog, const DogWeights &wts, float *next, int npitch,
                      DownK dk, cudaStream_t st) {
  static_assert(K1V2_SMEM <= 48 * 1024, "needs cudaFuncAttributeMaxDynamicSharedMemorySize above 48 KB");
  dim3 grd((w + 2 * TW - 1) / (2 * TW), (h + kRows - 1) / kRows);
  k_blur_dog2<kDown, kRows><<<grd, NT, K1V2_SMEM, st>>>(base, w, h, pitch, dog, dup_weights(wts), next, npitch, dk);
}
// fewer than two CTAs per SM with 16-row tiles: use 4-row tiles (the frame's latency, not its SM time)
inline bool small_octave(int w, int h) { return ((w + 2 * TW - 1) / (2 * TW)) * ((h + ROWS2 - 1) / ROWS2) < 2 * 148; }
}  // namespace

void launch_blur_dog(const float *base, int w, int h, int pitch, float *dog, const DogWeights &wts, cudaStream_t st) {
  DownK dk{0.f, 0.f, 0.f};
  if (csb_use_scalar_pyramid()) {
    dim3 grd((w + TW - 1) / TW, (h + ROWS - 1) / ROWS);
    k_blur_dog<false><<<grd, NT, 0, st>>>(base, w, h, pitch, dog, wts, nullptr, 0, dk);
    return;
  }
  if (small_octave(w, h)) launch_blur_dog2<false, ROWS2_SMALL>(base, w, h, pitch, dog, wts, nullptr, 0, dk, st);
  else launch_blur_dog2<false, ROWS2>(base, w, h, pitch, dog, wts, nullptr, 0, dk, st);
}

void launch_blur_dog_down(const float *base, int w, int h, int pitch, float *dog, const DogWeights &wts, float *next,
                          int npitch, const float k[3], cudaStream_t st) {
  DownK dk{k[0], k[1], k[2]};
  if (csb_use_scalar_pyramid()) {
    dim3 grd((w + TW - 1) / TW, (h + ROWS - 1) / ROWS);
    k_blur_dog<true><<<grd, NT, 0, st>>>(base, w, h, pitch, dog, wts, next, npitch, dk);
    return;
  }
  if (small_octave(w, h)) launch_blur_dog2<true, ROWS2_SMALL>(base, w, h, pitch, dog, wts, next, npitch, dk, st);
  else launch_blur_dog2<true, ROWS2>(base, w, h, pitch, dog, wts, next, npitch, dk, st);
}
