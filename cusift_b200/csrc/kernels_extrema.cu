// DoG extrema detection, sub-pixel refinement, edge rejection and compaction.
//
// Replaces FindPointsMulti_D (reference cuSIFT_D.cu:402-523, host side
// cuSIFT.cu:424-455).  Differences in structure, not in results:
//   * one pass over the octave (each thread owns a pixel column of the 5 centre
//     planes) instead of 5 overlapping scale-blocks; a DoG value is fetched once
//     for the |v|>thresh gate and neighbours are only touched behind that gate;
//   * no 32-entry block-local list (the reference's wraps, cuSIFT_D.cu:455,465):
//     keypoints are compacted with warp ballots, one atomicAdd per warp;
//   * the refinement is evaluated in the exact multiply-add order of the
//     reference's sm_100a SASS, so x, y, scale, sharpness and edgeness are
//     bit-identical to the reference's for the same DoG input.
#include "csb_internal.h"

namespace {

__device__ __forceinline__ bool strict_extremum(const float *__restrict__ dog, size_t plane, int pitch, int sc, int x,
                                                int y, float v, bool isMax) {
  // 26 neighbours on planes sc, sc+1, sc+2 (cuSIFT_D.cu:430-470); interior pixels only.
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const float *q = dog + (size_t)(sc + p) * plane + (size_t)y * pitch + x;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        if (p == 1 && dy == 0 && dx == 0) continue;
        const float u = q[dy * pitch + dx];
        if (isMax ? !(v > u) : !(v < u)) return false;
      }
    }
  }
  return true;
}

__global__ void __launch_bounds__(256) k_find_points(const float *__restrict__ dog, int w, int h, int pitch,
                                                     const __grid_constant__ ExtremaParams P,
                                                     csb_sift_point *__restrict__ d_sift, int *__restrict__ d_oct,
                                                     unsigned int *__restrict__ counter, int max_pts) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  const int lane = threadIdx.x;   // blockDim.x == 32: one warp per tile row
  // image-border pixels can never be strict extrema (their clamped neighbours
  // include the pixel itself, cuSIFT_D.cu:416,427-429)
  const bool inside = (x >= 1 && y >= 1 && x < w - 1 && y < h - 1);
  const size_t plane = (size_t)pitch * h;

#pragma unroll 1
  for (int sc = 0; sc < CSB_NUM_SCALES; sc++) {
    bool emit = false;
    float ox = 0.f, oy = 0.f, oscale = 0.f, osharp = 0.f, oedge = 0.f;
    if (inside) {
      const float *p1 = dog + (size_t)(sc + 1) * plane + (size_t)y * pitch + x;
      const float val = p1[0];
      const bool isMax = val > P.thresh, isMin = val < -P.thresh;
      if ((isMax || isMin) && strict_extremum(dog, plane, pitch, sc, x, y, val, isMax)) {
        // ---- refinement, cuSIFT_D.cu:474-522 ----
        const float *p0 = p1 - plane, *p2 = p1 + plane;
        const float two = __fadd_rn(val, val);
        const float dxx = __fsub_rn(__fsub_rn(two, p1[-1]), p1[1]);
        const float dyy = __fsub_rn(__fsub_rn(two, p1[-pitch]), p1[pitch]);
        const float dxy = __fmul_rn(
            0.25f, __fsub_rn(__fsub_rn(__fadd_rn(p1[pitch + 1], p1[-pitch - 1]), p1[-pitch + 1]), p1[pitch - 1]));
        const float tra = __fadd_rn(dxx, dyy);
        const float det = __fmaf_rn(dxx, dyy, -__fmul_rn(dxy, dxy));
        const float tra2 = __fmul_rn(tra, tra);
        if (tra2 < __fmul_rn(det, P.edge_limit)) {
          const float edge = __fdividef(tra2, det);
          const float dx = __fmul_rn(0.5f, __fsub_rn(p1[1], p1[-1]));
          const float dy = __fmul_rn(0.5f, __fsub_rn(p1[pitch], p1[-pitch]));
          const float ds = __fmul_rn(0.5f, __fsub_rn(p0[0], p2[0]));
          const float dss = __fsub_rn(__fsub_rn(two, p2[0]), p0[0]);
          const float dxs = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(p2[1], p0[-1]), p0[1]), p2[-1]));
          const float dys =
              __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(p2[pitch], p0[-pitch]), p2[-pitch]), p0[pitch]));
          const float idxx = __fmaf_rn(dyy, dss, -__fmul_rn(dys, dys));
          const float idxy = __fmaf_rn(dxs, dys, -__fmul_rn(dxy, dss));
          const float idxs = __fmaf_rn(dxy, dys, -__fmul_rn(dyy, dxs));
          const float den = __fmaf_rn(dxs, idxs, __fmaf_rn(dxx, idxx, __fmul_rn(dxy, idxy)));
          const float idet = __fdividef(1.0f, den);
          const float idyy = __fmaf_rn(dxx, dss, -__fmul_rn(dxs, dxs));
          const float idys = __fmaf_rn(dxy, dxs, -__fmul_rn(dxx, dys));
          const float idss = det;   // dxx*dyy - dxy*dxy, same value (CSE'd in the reference too)
          float pdx = __fmul_rn(idet, __fmaf_rn(ds, idxs, __fmaf_rn(dx, idxx, __fmul_rn(dy, idxy))));
          float pdy = __fmul_rn(idet, __fmaf_rn(ds, idys, __fmaf_rn(dy, idyy, __fmul_rn(dx, idxy))));
          float pds = __fmul_rn(idet, __fmaf_rn(idss, ds, __fmaf_rn(dx, idxs, __fmul_rn(dy, idys))));
          if (pdx < -0.5f || pdx > 0.5f || pdy < -0.5f || pdy > 0.5f || pds < -0.5f || pds > 0.5f) {
            pdx = __fdividef(dx, dxx);
            pdy = __fdividef(dy, dyy);
            pds = __fdividef(ds, dss);
          }
          const float dval = __fmaf_rn(ds, pds, __fmaf_rn(dx, pdx, __fmul_rn(dy, pdy)));
          ox = __fadd_rn((float)x, pdx);
          oy = __fadd_rn((float)y, pdy);
          oscale = __fmul_rn(P.scales[sc], exp2f(__fmul_rn(pds, P.factor)));
          osharp = __fmaf_rn(dval, 0.5f, val);
          oedge = edge;
          emit = true;
        }
      }
    }
    // warp-ballot compaction: one atomic per warp, slots in lane order
    const unsigned int m = __ballot_sync(0xffffffffu, emit);
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned int base = 0;
      if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (emit) {
        const unsigned int idx = base + __popc(m & ((1u << lane) - 1u));
        if (idx < (unsigned int)max_pts) {
          csb_sift_point *o = d_sift + idx;
          o->coords2D[0] = ox;
          o->coords2D[1] = oy;
          o->scale = oscale;
          o->sharpness = osharp;
          o->edgeness = oedge;
          o->orientation = 0.f;
          o->score = 0.f;
          o->ambiguity = 0.f;
          o->match = 0;
          o->match_xpos = 0.f;
          o->match_ypos = 0.f;
          o->match_error = 0.f;
          o->subsampling = P.subsampling;
          o->empty[0] = o->empty[1] = o->empty[2] = 0.f;
          o->coords3D[0] = o->coords3D[1] = o->coords3D[2] = 0.f;
          d_oct[idx] = P.octave;
        }
      }
    }
  }
}

}  // namespace

void launch_find_points(const float *dog, int w, int h, int pitch, const ExtremaParams &ep, csb_sift_point *d_sift,
                        int *d_oct, unsigned int *d_counter, int max_pts, cudaStream_t st) {
  dim3 blk(32, 8), grd((w + 31) / 32, (h + 7) / 8);
  k_find_points<<<grd, blk, 0, st>>>(dog, w, h, pitch, ep, d_sift, d_oct, d_counter, max_pts);
}
