// Orientation assignment + 4x4x8 descriptor (+ optional RootSIFT), one warp per
// keypoint, fused into a single persistent launch that reads the keypoint count
// from device memory (the reference reads it back to the host three times per
// octave to size two launches, cuSIFT.cu:243,251,255).
//
// Replaces (reference, danielsuo/cuSIFT):
//   ComputeOrientations_D     cuSIFT_D.cu:319-396   (128-thread block per keypoint)
//   ExtractSiftDescriptors_D  cuSIFT_D.cu:184-297   (16x8-thread block per keypoint)
//   ConvertSiftToRootSift_D   cuSIFT_D.cu:299-317
//
// Sampling goes through the texture unit exactly like the reference (bilinear,
// clamp, unnormalised coordinates, raw — not +0.5 — coordinates), because the
// hardware filter's 1.8 fixed-point weights are part of the reference's results.
// The sample-coordinate and weight arithmetic follows the reference's sm_100a
// SASS (which products are fused, which are not).  The histogram sums are the one
// place where the reference is nondeterministic (shared-memory float atomicAdd,
// cuSIFT_D.cu:234-253,343 — a CAS loop on sm_100a that serialises on every bin
// collision).  Here they are accumulated WITHOUT atomics and in a fixed order:
//   orientation: every lane owns a private 32-bin column in shared memory
//                (hist[bin][lane], bank = lane), reduced by a skewed read;
//   descriptor : lane = (half, cell_y, cell_x) samples its own 4x4 cell, so at each
//                of the 8 trilinear vote sites the 16 lanes of a half hit 16
//                different cells; the two halves own separate 128-bin copies.
// Results are run-to-run deterministic and within the reference's own spread.
#include <cstdlib>

#include <cuda_fp16.h>

#include "csb_internal.h"

namespace {

constexpr int WARPS = 4;
#ifndef K3_MINB
#define K3_MINB 8          // resident CTAs per SM the register budget is sized for
#endif
#ifndef K3_GRIDMUL
#define K3_GRIDMUL 1
#endif
#ifndef K3_BATCH
#define K3_BATCH 2         // descriptor samples whose texture fetches are in flight together
#endif
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// Sum of squares of 128 values held 4 per lane (element i = lane + 32 j), in the
// reference's tree order (cuSIFT_D.cu:260-271): s[i] = b[i]^2 + b[i+64]^2 for
// i<64, then +32, +16, +8, +4, then ((s0+s1)+s2)+s3.
__device__ __forceinline__ float sumsq_tree(const float (&b)[4], int lane) {
  float sA = __fmaf_rn(b[0], b[0], __fmul_rn(b[2], b[2]));   // idx = lane
  float sB = __fmaf_rn(b[1], b[1], __fmul_rn(b[3], b[3]));   // idx = lane + 32
  float s = __fadd_rn(sA, sB);                               // sums[idx] += sums[idx+32]
  s = __fadd_rn(s, __shfl_down_sync(FULL, s, 16));           // idx < 16
  s = __fadd_rn(s, __shfl_down_sync(FULL, s, 8));            // idx < 8
  s = __fadd_rn(s, __shfl_down_sync(FULL, s, 4));            // idx < 4
  const float s0 = __shfl_sync(FULL, s, 0), s1 = __shfl_sync(FULL, s, 1);
  const float s2 = __shfl_sync(FULL, s, 2), s3 = __shfl_sync(FULL, s, 3);
  return __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
}

// Descriptor accumulator of one warp: buf[row][bank], 8 rows x 32 banks: row = angular bin, bank 16 h + c =
// cell c of half-warp h.
// At one vote site the 32 lanes address 32 different (half, cell) pairs, hence 32 different banks
// whatever their angular bins are (the cell-major layout measured 72 % conflict wavefronts).
constexpr int DBUF = 8 * 32;

// The two angular shares of one spatial share as plain read-modify-writes: q0 != q1 (a lane's a1 and
// angp differ; a spilled a1 is handled by the atomic pass instead), and at one vote site the 16 lanes of a
// half hit 16 different cells, so no two lanes of the warp touch the same address in one call and both
// read-modify-writes can be in flight together.
__device__ __forceinline__ void vote2(bool ok0, float *q0, float v0, bool ok1, float *q1, float v1) {
  // predicated, not branched: a share without a destination (outside the 4x4 grid, or spilled) simply
  // makes no access, so the 32 lanes never meet in a bank
  float o0 = 0.f, o1 = 0.f;
  if (ok0) o0 = *q0;
  if (ok1) o1 = *q1;
  if (ok0) *q0 = __fadd_rn(o0, v0);
  if (ok1) *q1 = __fadd_rn(o1, v1);
  __syncwarp();
}

struct Grad {
  float dx, dy;
};

struct OctaveSubs {
  float v[CSB_MAX_OCTAVES];
};

__global__ void __launch_bounds__(WARPS * 32, K3_MINB) k_orient_desc(const __grid_constant__ OctaveTexSet T,
                                                            const __grid_constant__ OctaveSubs S,
                                                            const KpStage *__restrict__ d_stage,
                                                            csb_sift_point *__restrict__ d_sift,
                                                            unsigned int *__restrict__ counter, int max_pts,
                                                            int rootsift) {
  __shared__ __align__(16) float s_hist[WARPS][32 * 32];   // orientation: [bin][lane] private columns
  __shared__ float s_sm[WARPS][64];                         // reduced + smoothed orientation histogram
  __shared__ float s_gauss[WARPS][16];
  __shared__ float s_buf[WARPS][DBUF];                      // descriptor accumulator, see DBUF

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *hist = s_hist[warp], *sm = s_sm[warp], *gauss = s_gauss[warp], *buf = s_buf[warp];
  // descriptor sampling pattern: lane = half*16 + cell_y*4 + cell_x
  const int half = lane >> 4, cellx = lane & 3, celly = (lane >> 2) & 3;
  float *copy = buf + half * 16;                            // + cell + 32 * bin

  // k_find_points leaves one list per octave (counter[1 + o] entries, capped at max_pts); blockIdx.y
  // selects the octave.  That makes the texture handle a function of a special register, i.e.
  // provably warp-uniform: a handle looked up per keypoint costs a divergence ("waterfall") loop
  // around every one of the 48 texture fetches of a keypoint and keeps ptxas from batching them.
  // Output slot of entry i of octave o = (entries of all coarser octaves) + i: the reference's order,
  // coarse octaves first (cuSIFT.cu:181-196), and the order in which max_pts truncates.
  const int o = blockIdx.y;
  unsigned int before = 0, total = 0;
  for (int c = CSB_MAX_OCTAVES - 1; c >= 0; c--) {
    const unsigned int nc = min(counter[1 + c], (unsigned int)max_pts);
    if (c > o) before += nc;
    total += nc;
  }
  if (blockIdx.x == 0 && o == 0 && threadIdx.x == 0) counter[0] = total;   // keypoints found (host caps at max_pts)
  const int n_o = (int)min(counter[1 + o], (unsigned int)max_pts);
  const KpStage *__restrict__ stage = d_stage + (size_t)o * max_pts;
  const float psub = S.v[o];
  const cudaTextureObject_t tex = T.tex[o];
  {
#pragma unroll 1
  for (int i = blockIdx.x * WARPS + warp; i < n_o; i += gridDim.x * WARPS) {
    const unsigned int k = before + (unsigned int)i;
    if (k >= (unsigned int)max_pts) break;                  // later entries of this warp are beyond the cap too
    csb_sift_point *pt = d_sift + k;
    const KpStage kp = stage[i];
    const float px = kp.x, py = kp.y, pscale = kp.scale;

    // ---------------- orientation (cuSIFT_D.cu:319-396) ----------------
    const float i2sigma2 = __fdiv_rn(-1.0f, __fmul_rn(__fmul_rn(pscale, 4.5f), pscale));
    if (lane < 11) {
      const float d = (float)(lane - 5);
      gauss[lane] = expf(__fmul_rn(__fmul_rn(d, i2sigma2), d));
    }
    {
      float4 *z = reinterpret_cast<float4 *>(hist);
#pragma unroll
      for (int j = 0; j < 8; j++) z[j * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const float xp = __fsub_rn(px, 5.0f), yp = __fsub_rn(py, 5.0f);
    // all 16 texture fetches of the lane's (up to) 4 window samples are issued before the first is used
    float odx[4], ody[4];
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int s = lane + 32 * it;
      if (s < 121) {
        const int yd = s / 11, xd = s - yd * 11;
        const float xf = __fadd_rn(xp, (float)xd), yf = __fadd_rn(yp, (float)yd);
        odx[it] = __fsub_rn(tex2D<float>(tex, __fadd_rn(xf, 1.0f), yf), tex2D<float>(tex, __fsub_rn(xf, 1.0f), yf));
        ody[it] = __fsub_rn(tex2D<float>(tex, xf, __fadd_rn(yf, 1.0f)), tex2D<float>(tex, xf, __fsub_rn(yf, 1.0f)));
      }
    }
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int s = lane + 32 * it;
      if (s < 121) {
        const int yd = s / 11, xd = s - yd * 11;
        const float dx = odx[it], dy = ody[it];
        int bin = (int)__fadd_rn(__fdiv_rn(__fmul_rn(16.0f, atan2f(dy, dx)), 3.1416f), 16.5f);
        if (bin > 31) bin = 0;
        const float grad = sqrtf(__fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        float *q = hist + bin * 32 + lane;          // private column: no other lane touches it
        *q = __fadd_rn(*q, __fmul_rn(__fmul_rn(grad, gauss[xd]), gauss[yd]));
      }
    }
    __syncwarp();
    {
      // lane b sums bin b over the 32 private columns, skewed so that banks differ
      float acc = 0.0f;
#pragma unroll 8
      for (int l = 0; l < 32; l++) acc = __fadd_rn(acc, hist[lane * 32 + ((l + lane) & 31)]);
      sm[lane] = acc;
    }
    __syncwarp();
    {
      const float h0 = sm[lane];
      const float h1 = __fadd_rn(sm[(lane + 31) & 31], sm[(lane + 1) & 31]);
      const float h2 = __fadd_rn(sm[(lane + 30) & 31], sm[(lane + 2) & 31]);
      sm[32 + lane] = __fadd_rn(__fmaf_rn(h0, 6.0f, __fmul_rn(h1, 4.0f)), h2);
    }
    __syncwarp();
    float orient;
    {
      const float v = sm[32 + lane];
      const float pk = (v > sm[32 + ((lane + 31) & 31)] && v >= sm[32 + ((lane + 1) & 31)]) ? v : 0.0f;
      // serial scan of the reference (strict >, first maximum wins) == lowest lane holding the maximum
      const float maxval1 = warp_max(pk);
      const unsigned int who = __ballot_sync(FULL, pk == maxval1);
      const int i1 = (maxval1 > 0.0f) ? (__ffs(who) - 1) : -1;
      const float val1 = sm[32 + ((i1 + 1) & 31)];
      const float val2 = sm[32 + ((i1 + 31) & 31)];
      const float peak = __fadd_rn(
          (float)i1, __fdiv_rn(__fmul_rn(0.5f, __fsub_rn(val1, val2)),
                               __fsub_rn(__fsub_rn(__fadd_rn(maxval1, maxval1), val1), val2)));
      orient = __fmul_rn(11.25f, (peak < 0.0f ? __fadd_rn(peak, 32.0f) : peak));
    }
    __syncwarp();

    // ---------------- descriptor (cuSIFT_D.cu:184-297) ----------------
    if (lane < 16) {
      const float d = __fsub_rn((float)lane, 7.5f);
      gauss[lane] = expf(__fdiv_rn(__fmul_rn(-d, d), 128.0f));
    }
    for (int j = lane; j < DBUF; j += 32) buf[j] = 0.0f;
    __syncwarp();
    const float theta = __fmul_rn(2.0f * 3.1415f / 360.0f, orient);
    const float sina = sinf(theta), cosa = cosf(theta);
    const float sc = __fmul_rn(pscale, 0.75f);
    const float ssina = __fmul_rn(sina, sc), scosa = __fmul_rn(cosa, sc);
    // sample `it` of this lane's cell
    auto fetch = [&](int it) -> Grad {
      const int j = it * 2 + half;
      const int tx = 4 * cellx + (j & 3), y = 4 * celly + (j >> 2);
      const float ftx = __fsub_rn((float)tx, 7.5f), fy = __fsub_rn((float)y, 7.5f);
      const float xpos = __fmaf_rn(-ssina, fy, __fadd_rn(__fmul_rn(ftx, scosa), px));
      const float ypos = __fmaf_rn(scosa, fy, __fadd_rn(__fmul_rn(ftx, ssina), py));
      Grad g;
      g.dx = __fsub_rn(tex2D<float>(tex, __fadd_rn(xpos, cosa), __fadd_rn(ypos, sina)),
                       tex2D<float>(tex, __fsub_rn(xpos, cosa), __fsub_rn(ypos, sina)));
      g.dy = __fsub_rn(tex2D<float>(tex, __fsub_rn(xpos, sina), __fadd_rn(ypos, cosa)),
                       tex2D<float>(tex, __fadd_rn(xpos, sina), __fsub_rn(ypos, cosa)));
      return g;
    };
#pragma unroll 1
    for (int it0 = 0; it0 < 8; it0 += K3_BATCH) {
    Grad gs[K3_BATCH];                             // 4 K3_BATCH texture fetches of the lane in flight at once
#pragma unroll
    for (int i = 0; i < K3_BATCH; i++) gs[i] = fetch(it0 + i);
#pragma unroll
    for (int i = 0; i < K3_BATCH; i++) {
      const int it = it0 + i;
      const Grad g = gs[i];
      const int j = it * 2 + half;                 // sample within the lane's 4x4 cell
      const int tx = 4 * cellx + (j & 3), y = 4 * celly + (j >> 2);
      const float dx = g.dx, dy = g.dy;
      const float grad = __fmul_rn(__fmul_rn(gauss[y], gauss[tx]), sqrtf(__fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
      float angf = __fmaf_rn(atan2f(dy, dx), 4.0f / 3.1415f, 4.0f);
      const int hori = (tx + 2) / 4 - 1;
      const float horf = __fmaf_rn(__fsub_rn((float)tx, 1.5f), 0.25f, -(float)hori);
      const float ihorf = __fsub_rn(1.0f, horf);
      const int veri = (y + 2) / 4 - 1;
      const float verf = __fmaf_rn(__fsub_rn((float)y, 1.5f), 0.25f, -(float)veri);
      const float iverf = __fsub_rn(1.0f, verf);
      const int angi = (int)angf;
      const int angp = (angi < 7 ? angi + 1 : 0);
      angf = __fsub_rn(angf, (float)angi);
      const float iangf = __fsub_rn(1.0f, angf);
      const int cell = 4 * veri + hori;              // the UL cell; DL = +4, UR = +1, DR = +5 (may be -5 .. 20)
      // the four spatial shares (guards of cuSIFT_D.cu:230-255; `tx<=14` sic).  A share whose cell
      // lies outside buffer[128] is dropped (in the reference it lands beyond its last shared array);
      // invalid shares make no access.
      const bool gl = tx >= 2, gr = tx <= 14, gu = y >= 2, gd = y <= 13;
      const float gUL = __fmul_rn(iverf, __fmul_rn(ihorf, grad)), gDL = __fmul_rn(verf, __fmul_rn(ihorf, grad));
      const float gUR = __fmul_rn(iverf, __fmul_rn(horf, grad)), gDR = __fmul_rn(verf, __fmul_rn(horf, grad));
      const bool vUL = gl && gu, vDL = gl && gd && (cell + 4 < 16);
      const bool vUR = gr && gu && (cell + 1 < 16), vDR = gr && gd && (cell + 5 < 16);
      float *cUL = copy + cell, *cDL = copy + cell + 4, *cUR = copy + cell + 1, *cDR = copy + cell + 5;
      // angi == 8 (atan2f >= 3.1415, e.g. dy == +0, dx < 0) makes p1 point at bin 0 of the NEXT cell,
      // which another lane may be voting into at the same site: those votes go through an atomic pass.
      const bool spill = angi >= 8;
      const int a1 = spill ? 0 : 32 * angi, a2 = 32 * angp;
      vote2(vUL && !spill, cUL + a1, __fmul_rn(iangf, gUL), vUL, cUL + a2, __fmul_rn(angf, gUL));
      vote2(vDL && !spill, cDL + a1, __fmul_rn(iangf, gDL), vDL, cDL + a2, __fmul_rn(angf, gDL));
      vote2(vUR && !spill, cUR + a1, __fmul_rn(iangf, gUR), vUR, cUR + a2, __fmul_rn(angf, gUR));
      vote2(vDR && !spill, cDR + a1, __fmul_rn(iangf, gDR), vDR, cDR + a2, __fmul_rn(angf, gDR));
      if (__any_sync(FULL, spill)) {
        if (spill) {
          if (vUL && cell + 1 < 16) atomicAdd(copy + cell + 1, __fmul_rn(iangf, gUL));
          if (vDL && cell + 5 < 16) atomicAdd(copy + cell + 5, __fmul_rn(iangf, gDL));
          if (vUR && cell + 2 < 16) atomicAdd(copy + cell + 2, __fmul_rn(iangf, gUR));
          if (vDR && cell + 6 < 16) atomicAdd(copy + cell + 6, __fmul_rn(iangf, gDR));
        }
        __syncwarp();
      }
    }
    }
    __syncwarp();

    // normalise, clamp at 0.2, normalise (cuSIFT_D.cu:259-291)
    float b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {                    // element lane + 32 j = cell 4 j + lane / 8, bin lane % 8
      const float *e = buf + (lane & 7) * 32 + 4 * j + (lane >> 3);
      b[j] = __fadd_rn(e[0], e[16]);
    }
    const float r1 = rsqrtf(sumsq_tree(b, lane));
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float v = __fmul_rn(b[j], r1);
      b[j] = (v > 0.2f) ? 0.2f : v;
    }
    const float r2 = rsqrtf(sumsq_tree(b, lane));
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = __fmul_rn(b[j], r2);

    if (rootsift) {
      // ConvertSiftToRootSift_D (cuSIFT_D.cu:299-317): serial fp32 sum over i=0..127,
      // then sqrtf((float)(max(0.0,(double)d) / (double)sum)).
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; j++) buf[lane + 32 * j] = b[j];
      __syncwarp();
      float sum = 0.0f;
      for (int i = 0; i < 128; i++) sum = __fadd_rn(sum, buf[i]);
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sqrtf((float)(fmax(0.0, (double)b[j]) / (double)sum));
    }
#pragma unroll
    for (int j = 0; j < 4; j++) pt->data[lane + 32 * j] = b[j];
    if (lane < 16) {   // the 76 header + 12 trailer bytes of the record, one word per lane
      float v = 0.0f;                                        // score, ambiguity, match (int 0), match_*, empty[], coords3D[]
      if (lane == 0) v = __fmul_rn(px, psub);                // coords2D, scale: octave -> frame pixels (cuSIFT_D.cu:292-296)
      else if (lane == 1) v = __fmul_rn(py, psub);
      else if (lane == 2) v = __fmul_rn(pscale, psub);
      else if (lane == 3) v = kp.sharp;
      else if (lane == 4) v = kp.edge;
      else if (lane == 5) v = orient;
      else if (lane == 12) v = psub;
      reinterpret_cast<float *>(pt)[lane] = v;               // words 0..15 = fields up to empty[2]
      if (lane < 3) pt->coords3D[lane] = 0.0f;
    }
    __syncwarp();
  }
  }
}

// Stand-alone SiftData::ConvertSiftToRootSift (cuSIFT.cu:383-395): one warp per point.
__global__ void __launch_bounds__(128) k_rootsift(csb_sift_point *__restrict__ d_sift, int n) {
  __shared__ float s_buf[4][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *buf = s_buf[warp];
  for (int k = blockIdx.x * 4 + warp; k < n; k += gridDim.x * 4) {
    float b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      b[j] = d_sift[k].data[lane + 32 * j];
      buf[lane + 32 * j] = b[j];
    }
    __syncwarp();
    float sum = 0.0f;
    for (int i = 0; i < 128; i++) sum = __fadd_rn(sum, buf[i]);
#pragma unroll
    for (int j = 0; j < 4; j++) d_sift[k].data[lane + 32 * j] = sqrtf((float)(fmax(0.0, (double)b[j]) / (double)sum));
    __syncwarp();
  }
}

// Opt-in compact result records (csb_extract_batch_compact): the seven header fields a consumer of keypoints needs +
// the descriptor rounded to fp16 = 288 bytes instead of 588.  One warp per point; the count is read on the device
// (the frame's own counter), so no host round trip precedes the launch.
__global__ void __launch_bounds__(128) k_compact(const csb_sift_point *__restrict__ d_sift, const unsigned int *__restrict__ count,
                                                 int max_pts, csb_compact_point *__restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = (int)min(*count, (unsigned int)max_pts);
  for (int k = blockIdx.x * 4 + warp; k < n; k += gridDim.x * 4) {
    const csb_sift_point &p = d_sift[k];
    csb_compact_point &o = out[k];
    if (lane < 8) {
      const float v = lane == 0 ? p.coords2D[0] : lane == 1 ? p.coords2D[1] : lane == 2 ? p.scale : lane == 3 ? p.orientation
                    : lane == 4 ? p.sharpness : lane == 5 ? p.edgeness : lane == 6 ? p.subsampling : 0.0f;
      reinterpret_cast<float *>(&o)[lane] = v;
    }
    // (588-byte records: data[] is only 4-byte aligned, so scalar loads; the warp still reads 512 contiguous bytes)
    const float4 d = make_float4(p.data[4 * lane], p.data[4 * lane + 1], p.data[4 * lane + 2], p.data[4 * lane + 3]);
    uint2 h;
    h.x = (unsigned int)__half_as_ushort(__float2half_rn(d.x)) | ((unsigned int)__half_as_ushort(__float2half_rn(d.y)) << 16);
    h.y = (unsigned int)__half_as_ushort(__float2half_rn(d.z)) | ((unsigned int)__half_as_ushort(__float2half_rn(d.w)) << 16);
    *reinterpret_cast<uint2 *>(o.data + 4 * lane) = h;
  }
}

}  // namespace

void launch_compact(const csb_sift_point *d_sift, const unsigned int *d_count, int max_pts, csb_compact_point *d_out,
                    cudaStream_t st) {
  k_compact<<<148 * 4, 128, 0, st>>>(d_sift, d_count, max_pts, d_out);
}

void launch_orient_desc(const OctaveTexSet &texs, int n_oct, const float *subs, const KpStage *d_stage, csb_sift_point *d_sift,
                        unsigned int *d_counter, int max_pts, int rootsift, int sm_count, cudaStream_t st) {
  // K3_GRIDMUL x the resident capacity: with about one keypoint per warp the hardware CTA scheduler
  // balances the load (CTAs start as others retire) instead of a static two-keypoints-or-one split
  int blocks = sm_count * K3_MINB * K3_GRIDMUL;
  const int need = (max_pts + WARPS - 1) / WARPS;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  OctaveSubs S;
  for (int o = 0; o < CSB_MAX_OCTAVES; o++) S.v[o] = o < n_oct ? subs[o] : 0.0f;
  // grid.y = octave, finest first: its list is by far the longest, the short ones fill the tail
  k_orient_desc<<<dim3(blocks, n_oct), WARPS * 32, 0, st>>>(texs, S, d_stage, d_sift, d_counter, max_pts, rootsift);
}

void launch_rootsift(csb_sift_point *d_sift, int n, cudaStream_t st) {
  if (n <= 0) return;
  int blocks = (n + 3) / 4;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_rootsift<<<blocks, 128, 0, st>>>(d_sift, n);
}
