// DoG extrema detection, sub-pixel refinement, edge rejection and compaction.
//
// Replaces FindPointsMulti_D (reference cuSIFT_D.cu:402-523, host side
// cuSIFT.cu:424-455).  Same results, different structure:
//   * ONE launch for all octaves, ONE pass over the 7 DoG planes (28 B/pixel, each value fetched
//     once) instead of 5 overlapping scale-blocks per octave that re-read every plane ~3x.
//   * warp-specialised CTA: a loader warp feeds a 3-stage shared-memory ring with one 2-D tiled TMA
//     load per stage (3 rows x 7 planes x 128 columns; the DoG buffer is stored row-interleaved, so
//     that box is one rectangle of a (pitch x 7h) matrix), mbarriers in both directions; four
//     scanning warps own 30 columns each (+1 halo lane each side) and stream the rows out of the
//     ring: per plane a 3-row window lives in registers, horizontal neighbours come from shuffles.
//   * the scan is a conservative PREFILTER in packed fp16: every value travels as the
//     half2 (rn(v), rn(-v)), so ONE 3-input packed max (VHMNMX) advances the maximum
//     and the minimum network together (min/max issue at half rate on sm_100a and were
//     the bound of the fp32 scan).  The 3x3x3 maximum is taken rows first, planes second,
//     columns last, so only the 5 per-scale partial maxima go through shuffles.  Rounding is
//     monotone, so a true fp32 extremum above the threshold always satisfies rn(v) == M and
//     rn(|v|) >= rn(thresh); the scan flags those pixels (plus fp16 ties).
//   * flagged pixels go to a per-CTA list; after the scan the whole CTA applies the
//     reference's STRICT fp32 comparison against each of the 26 neighbours
//     (cuSIFT_D.cu:450-470) and its |v| > thresh test, refines the survivors and compacts
//     them with warp ballots into the octave's keypoint list, one global atomicAdd per warp.
//     (The reference's 32-entry list silently wraps, cuSIFT_D.cu:455,465; here a tile whose
//     flagged pixels exceed the list is re-examined pixel by pixel.)
//   * the refinement is evaluated in the exact multiply-add order of the
//     reference's sm_100a SASS, so x, y, scale, sharpness and edgeness are
//     bit-identical to the reference's for the same DoG input.
#include <cuda_fp16.h>

#include <cstdlib>

#include <type_traits>

#include "csb_internal.h"
#include "tma_util.h"

namespace {

constexpr int XT_COLS = 30;          // output columns per warp
#ifndef K2_WARPS
#define K2_WARPS 4          // scanning warps per CTA (4: 120-column tiles, 512-byte TMA rows; 8: 240 columns, 1 KB rows)
#endif
constexpr int XT_WARPS = K2_WARPS;
constexpr int XT_TW = XT_COLS * XT_WARPS;   // output columns per CTA
#ifndef K2_MINB
#define K2_MINB (K2_WARPS == 4 ? 6 : 3)     // resident CTAs per SM (35 KB / 67 KB of shared memory each)
#endif
#ifndef K2_WAVES
#define K2_WAVES 1          // CTAs launched per resident slot
#endif
constexpr int XT_MAX_ROWS = 36;      // output rows per CTA: a launch parameter, multiple of 3, <= 63
constexpr int NPL = CSB_NUM_DOG;     // 7 planes
constexpr int XT_CAP = 1024;         // flagged pixels per CTA held for the dense second phase
constexpr int ST_ROWS = 3;           // source rows per pipeline stage (one turn of the 3-row register window)
#ifndef K2_STAGES
#define K2_STAGES 3
#endif
constexpr int ST_N = K2_STAGES;      // stages: 9 rows x 7 planes of loads in flight per CTA
constexpr int ST_COLS = (XT_TW + 8 + 31) / 32 * 32;   // columns staged per row: output columns + halo, rounded so that a stage
                                                      // is a multiple of 128 bytes (TMA destination alignment): 128 or 256
constexpr uint32_t ST_BYTES = ST_ROWS * NPL * ST_COLS * sizeof(float);
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
using tma::mbar_arrive;
using tma::mbar_expect_tx;
using tma::mbar_init;
using tma::mbar_wait;
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int x, int yp, uint64_t *bar) {
  tma::load_2d(dst_smem, map, x, yp, bar);
}

// packed fp16 helpers: p = (rn(v), rn(-v)); ptxas fuses the two max.f16x2 into one 3-input VHMNMX
__device__ __forceinline__ unsigned int pack_pm(float v) {
  const __half2 p = __floats2half2_rn(v, -v);
  return *reinterpret_cast<const unsigned int *>(&p);
}
__device__ __forceinline__ unsigned int hmax3(unsigned int a, unsigned int b, unsigned int c) {
  unsigned int t, d;
  asm("max.f16x2 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(b));
  asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(t), "r"(c));
  return d;
}
__device__ __forceinline__ unsigned int eq_pm(unsigned int p, unsigned int m) {
  return __heq2_mask(*reinterpret_cast<const __half2 *>(&p), *reinterpret_cast<const __half2 *>(&m));
}
__device__ __forceinline__ unsigned int ge_pm(unsigned int p, unsigned int t) {
  return __hge2_mask(*reinterpret_cast<const __half2 *>(&p), *reinterpret_cast<const __half2 *>(&t));
}
// 0xffff in each half where p == m and p >= t
__device__ __forceinline__ unsigned int flag_pm(unsigned int p, unsigned int m, unsigned int t) {
  const __half2 hp = *reinterpret_cast<const __half2 *>(&p);
  return __heq2_mask(hp, *reinterpret_cast<const __half2 *>(&m)) & __hge2_mask(hp, *reinterpret_cast<const __half2 *>(&t));
}

struct Refined {
  float x, y, scale, sharp, edge;
};

// Strict 26-neighbour test (cuSIFT_D.cu:450-470) + refinement (cuSIFT_D.cu:474-522) of one flagged pixel.
// All 27 values are loaded up front as independent loads (ONE memory latency; a short-circuit chain of
// compare-then-load cost up to 26 dependent L2 round trips and was 30 % of the kernel's stall samples),
// then both steps run from registers in the reference's evaluation order.
__device__ __forceinline__ bool verify_refine(const float *__restrict__ dog, size_t plane, int pitch,
                                              const ExtremaParams &P, int x, int y, int sc, Refined &o) {
  const float *c = dog + (size_t)(sc + 1) * plane + (size_t)y * pitch + x;
  float n[3][3][3];   // [plane - sc][dy + 1][dx + 1]
#pragma unroll
  for (int p = 0; p < 3; p++)
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
      for (int dx = 0; dx < 3; dx++) n[p][dy][dx] = c[(ptrdiff_t)(p - 1) * (ptrdiff_t)plane + (dy - 1) * pitch + (dx - 1)];
  const float val = n[1][1][1];
  const bool isMax = val > P.thresh, isMin = val < -P.thresh;
  bool ok = isMax || isMin;
#pragma unroll
  for (int p = 0; p < 3; p++)
#pragma unroll
    for (int dy = 0; dy < 3; dy++)
#pragma unroll
      for (int dx = 0; dx < 3; dx++)
        if (!(p == 1 && dy == 1 && dx == 1)) ok &= isMax ? (val > n[p][dy][dx]) : (val < n[p][dy][dx]);
  if (!ok) return false;
  // names as in refine(): p0 = plane below (sc), p1 = centre plane, p2 = plane above
#define N0(dy, dx) n[0][(dy) + 1][(dx) + 1]
#define N1(dy, dx) n[1][(dy) + 1][(dx) + 1]
#define N2(dy, dx) n[2][(dy) + 1][(dx) + 1]
  const float two = __fadd_rn(val, val);
  const float dxx = __fsub_rn(__fsub_rn(two, N1(0, -1)), N1(0, 1));
  const float dyy = __fsub_rn(__fsub_rn(two, N1(-1, 0)), N1(1, 0));
  const float dxy = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(N1(1, 1), N1(-1, -1)), N1(-1, 1)), N1(1, -1)));
  const float tra = __fadd_rn(dxx, dyy);
  const float det = __fmaf_rn(dxx, dyy, -__fmul_rn(dxy, dxy));
  const float tra2 = __fmul_rn(tra, tra);
  if (!(tra2 < __fmul_rn(det, P.edge_limit))) return false;
  const float edge = __fdividef(tra2, det);
  const float dx = __fmul_rn(0.5f, __fsub_rn(N1(0, 1), N1(0, -1)));
  const float dy = __fmul_rn(0.5f, __fsub_rn(N1(1, 0), N1(-1, 0)));
  const float ds = __fmul_rn(0.5f, __fsub_rn(N0(0, 0), N2(0, 0)));
  const float dss = __fsub_rn(__fsub_rn(two, N2(0, 0)), N0(0, 0));
  const float dxs = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(N2(0, 1), N0(0, -1)), N0(0, 1)), N2(0, -1)));
  const float dys = __fmul_rn(0.25f, __fsub_rn(__fsub_rn(__fadd_rn(N2(1, 0), N0(-1, 0)), N2(-1, 0)), N0(1, 0)));
#undef N0
#undef N1
#undef N2
  const float idxx = __fmaf_rn(dyy, dss, -__fmul_rn(dys, dys));
  const float idxy = __fmaf_rn(dxs, dys, -__fmul_rn(dxy, dss));
  const float idxs = __fmaf_rn(dxy, dys, -__fmul_rn(dyy, dxs));
  const float den = __fmaf_rn(dxs, idxs, __fmaf_rn(dxx, idxx, __fmul_rn(dxy, idxy)));
  const float idet = __fdividef(1.0f, den);
  const float idyy = __fmaf_rn(dxx, dss, -__fmul_rn(dxs, dxs));
  const float idys = __fmaf_rn(dxy, dxs, -__fmul_rn(dxx, dys));
  const float idss = det;   // dxx*dyy - dxy*dxy: same value (CSE'd in the reference too)
  float pdx = __fmul_rn(idet, __fmaf_rn(ds, idxs, __fmaf_rn(dx, idxx, __fmul_rn(dy, idxy))));
  float pdy = __fmul_rn(idet, __fmaf_rn(ds, idys, __fmaf_rn(dy, idyy, __fmul_rn(dx, idxy))));
  float pds = __fmul_rn(idet, __fmaf_rn(idss, ds, __fmaf_rn(dx, idxs, __fmul_rn(dy, idys))));
  if (pdx < -0.5f || pdx > 0.5f || pdy < -0.5f || pdy > 0.5f || pds < -0.5f || pds > 0.5f) {
    pdx = __fdividef(dx, dxx);
    pdy = __fdividef(dy, dyy);
    pds = __fdividef(ds, dss);
  }
  const float dval = __fmaf_rn(ds, pds, __fmaf_rn(dx, pdx, __fmul_rn(dy, pdy)));
  o.x = __fadd_rn((float)x, pdx);
  o.y = __fadd_rn((float)y, pdy);
  o.scale = __fmul_rn(P.scales[sc], exp2f(__fmul_rn(pds, P.factor)));
  o.sharp = __fmaf_rn(dval, 0.5f, val);
  o.edge = edge;
  return true;
}

// Warp-ballot compaction of `emit` lanes into the octave's list (whole warp must call):
// one global atomicAdd per warp on the octave's counter.
__device__ __forceinline__ void emit_warp(bool emit, const Refined &r, KpStage *__restrict__ stage,
                                          unsigned int *__restrict__ oct_counter, int max_pts, int lane) {
  const unsigned int m = __ballot_sync(FULL, emit);
  if (!m) return;
  const int leader = __ffs(m) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(oct_counter, (unsigned int)__popc(m));
  base = __shfl_sync(FULL, base, leader);
  if (emit) {
    const unsigned int idx = base + __popc(m & ((1u << lane) - 1u));
    if (idx < (unsigned int)max_pts) stage[idx] = KpStage{r.x, r.y, r.scale, r.sharp, r.edge};
  }
}

__global__ void __launch_bounds__((XT_WARPS + 1) * 32, K2_MINB) k_find_points(const __grid_constant__ ExtremaParams P,
                                                                              const __grid_constant__ ExtremaMaps TM,
                                                                              KpStage *__restrict__ d_stage,
                                                                              unsigned int *__restrict__ counter,
                                                                              int max_pts, int cap) {
  // Warp-specialised: warp 4 is the LOADER (one elected lane issues ONE 2-D tiled TMA load per stage —
  // 128 columns x (3 rows x 7 planes), 10.5 KB — into a 3-stage shared-memory ring, completion on
  // mbarriers); warps 0-3 scan.  Out-of-image rows / columns of a box are zero-filled by the TMA unit:
  // they only ever feed border pixels, which cannot be extrema.
  // With loads issued by the scanning warps themselves the kernel ran at 3.2 TB/s although the same access
  // pattern alone reaches about 5 TB/s: the memory pipeline only moved when the compute warps got round to it.
  extern __shared__ __align__(128) unsigned char xt_smem[];   // the ring: ST_N stages of ST_BYTES (dynamic: > 48 KB for wide tiles)
  float(*s_tile)[ST_ROWS][NPL][ST_COLS] = reinterpret_cast<float(*)[ST_ROWS][NPL][ST_COLS]>(xt_smem);
  __shared__ __align__(8) uint64_t s_full[ST_N], s_empty[ST_N];
  __shared__ unsigned int s_cnt;
  __shared__ unsigned int s_list[XT_CAP];     // local column | local row << 8 | scale << 14

  // which octave does this CTA belong to?  (octave 0 owns the first, and by far the most, CTAs)
  int oi = 0;
#pragma unroll 1
  for (int i = 1; i < P.n_oct; i++)
    if ((int)blockIdx.x >= P.oct[i].cta_begin) oi = i;
  const ExtremaOctave &O = P.oct[oi];
  const float *__restrict__ dog = O.dog;
  const int w = O.w, h = O.h, pitch = O.pitch, rows = P.rows;
  const int local = (int)blockIdx.x - O.cta_begin;
  const int bx = local % O.tiles_x, by = local / O.tiles_x;
  KpStage *stage = d_stage + (size_t)O.octave * max_pts;
  unsigned int *oct_counter = counter + 1 + O.octave;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int y0 = by * rows;
  const size_t plane = CSB_DOG_PS(pitch, h);
  const int drow = (int)CSB_DOG_RS(pitch);                    // elements between consecutive rows of one plane
  // staged columns [s0, s0 + 128): the tile's 120 output columns with a 4-column halo
  const int s0 = bx * XT_TW - 4;
  const int n_stages = (rows + 2 + ST_ROWS - 1) / ST_ROWS;    // source rows y0-1 .. y0+rows
  if (threadIdx.x == 0) {
    s_cnt = 0;
    for (int s = 0; s < ST_N; s++) {
      mbar_init(s_full + s, 1);
      mbar_init(s_empty + s, XT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == XT_WARPS) {
    // ===== loader =====
    if (lane == 0) {
      for (int st = 0; st < n_stages; st++) {
        const int s = st % ST_N;
        if (st >= ST_N) mbar_wait(s_empty + s, ((st / ST_N) - 1) & 1);
        mbar_expect_tx(s_full + s, ST_BYTES);
        tma_load_2d(&s_tile[s][0][0][0], &TM.m[oi], s0, (y0 - 1 + st * ST_ROWS) * NPL, s_full + s);
      }
    }
  } else {
  // ===== scanning warps =====
  const int x = bx * XT_TW + warp * XT_COLS - 1 + lane;
  const int tcol = clampi(x, 0, w - 1) - s0;                 // this lane's column inside the staged tile (0 .. 127)
  // image-border pixels can never be strict extrema (their clamped neighbours include
  // the pixel itself, cuSIFT_D.cu:416,427-429); lanes 0 and 31 are halo columns
  const bool colOK = (lane >= 1) && (lane <= XT_COLS) && (x >= 1) && (x <= w - 2);

  // per plane: the packed values of three consecutive rows (slots rotate)
  unsigned int c3[NPL][3];
  const unsigned int tpk = pack_pm(P.thresh) & 0xffffu, tp = tpk | (tpk << 16);   // (rn(t), rn(t))

  auto place = [&](auto SLOT, const float (*rowp)[ST_COLS]) {   // rowp: the 7 planes of one staged source row
    constexpr int S = decltype(SLOT)::value;
#pragma unroll
    for (int p = 0; p < NPL; p++) c3[p][S] = pack_pm(rowp[p][tcol]);
  };
  // The 3x3x3 maximum is separable; taking the column (3 rows) and the plane (3 planes) maxima first and
  // the horizontal one last means only the 5 per-scale partial maxima travel through shuffles.
  auto test = [&](auto SM, int y) {   // output row y = slot SM; the other two slots are rows y-1 / y+1
    constexpr int M = decltype(SM)::value;
    unsigned int vx[NPL];
#pragma unroll
    for (int p = 0; p < NPL; p++) vx[p] = hmax3(c3[p][0], c3[p][1], c3[p][2]);
    unsigned int mx[CSB_NUM_SCALES];
#pragma unroll
    for (int sc = 0; sc < CSB_NUM_SCALES; sc++) {
      const unsigned int wv = hmax3(vx[sc], vx[sc + 1], vx[sc + 2]);
      const unsigned int l = __shfl_up_sync(FULL, wv, 1), r = __shfl_down_sync(FULL, wv, 1);
      mx[sc] = hmax3(l, wv, r);
    }
    // flagged: packed value equals the 27-neighbourhood maximum at some scale AND some scale of this
    // pixel reaches the threshold (a superset of "at the same scale", refined in the rare path below)
    const unsigned int eq = (eq_pm(c3[1][M], mx[0]) | eq_pm(c3[2][M], mx[1]) | eq_pm(c3[3][M], mx[2])) |
                            (eq_pm(c3[4][M], mx[3]) | eq_pm(c3[5][M], mx[4]));
    const unsigned int big = hmax3(hmax3(c3[1][M], c3[2][M], c3[3][M]), c3[4][M], c3[5][M]);
    const bool rowOK = colOK && (y >= y0) && (y >= 1) && (y <= h - 2) && (y < y0 + rows);
    if (rowOK && (eq & ge_pm(big, tp))) {                    // rare
      const unsigned int loc = (unsigned int)(x - bx * XT_TW) | ((unsigned int)(y - y0) << 8);
#pragma unroll
      for (int sc = 0; sc < CSB_NUM_SCALES; sc++)
        if (flag_pm(c3[sc + 1][M], mx[sc], tp)) {
          const unsigned int slot = atomicAdd(&s_cnt, 1u);
          if (slot < (unsigned int)cap) s_list[slot] = loc | ((unsigned int)sc << 14);
        }
    }
  };
  using I0 = std::integral_constant<int, 0>;
  using I1 = std::integral_constant<int, 1>;
  using I2 = std::integral_constant<int, 2>;
  // stage st holds source rows r = y0 - 1 + 3 st + {0, 1, 2}; row r goes to window slot (r - y0 + 1) % 3 = j,
  // and once rows r-2 .. r are in the window, output row r - 1 (slot j - 1) is tested
  for (int st = 0; st < n_stages; st++) {
    const int s = st % ST_N;
    mbar_wait(s_full + s, (st / ST_N) & 1);
    const int r = y0 - 1 + st * ST_ROWS;
    place(I0{}, s_tile[s][0]);
    if (st > 0) test(I2{}, r - 1);
    place(I1{}, s_tile[s][1]);
    if (st > 0) test(I0{}, r);
    place(I2{}, s_tile[s][2]);
    test(I1{}, r + 1);
    __syncwarp();
    if (lane == 0) mbar_arrive(s_empty + s);
  }
  }
  __syncthreads();

  // dense second phase: strict 26-neighbour test, refinement, compaction (whole CTA)
  auto drain = [&]() {
    const unsigned int flagged = s_cnt;
    // more flagged pixels than the list holds (pathological input): forget the list and put every
    // pixel of the tile through the strict test instead
    const bool dense = flagged > (unsigned int)cap;
    const int tile_w = min(XT_TW, w - bx * XT_TW), tile_h = min(rows, h - y0);
    const unsigned int n = dense ? (unsigned int)(tile_w * tile_h * CSB_NUM_SCALES) : flagged;
    for (unsigned int base = 0; base < n; base += (XT_WARPS + 1) * 32) {
      const unsigned int i = base + threadIdx.x;
      bool emit = false;
      Refined r;
      if (i < n) {
        int ex, ey, es;
        if (dense) {
          es = (int)(i % CSB_NUM_SCALES);
          ex = bx * XT_TW + (int)((i / CSB_NUM_SCALES) % (unsigned int)tile_w);
          ey = y0 + (int)((i / CSB_NUM_SCALES) / (unsigned int)tile_w);
        } else {
          const unsigned int en = s_list[i];
          ex = bx * XT_TW + (int)(en & 0xffu), ey = y0 + (int)((en >> 8) & 0x3fu), es = (int)(en >> 14);
        }
        const bool inner = ex >= 1 && ex <= w - 2 && ey >= 1 && ey <= h - 2;
        if (inner) emit = verify_refine(dog, plane, drow, P, ex, ey, es, r);
      }
      emit_warp(emit, r, stage, oct_counter, max_pts, lane);
    }
  };
  drain();
}

}  // namespace

int plan_find_points(ExtremaParams *ep, int sm_count) {
  // rows per CTA: about one full wave of CTAs over ALL octaves (K2_MINB resident per SM), so that the
  // serial row loop of a CTA is as short as the frame allows and every CTA carries the same work;
  // multiple of 3 (the register window rotates in threes)
  long long row_tiles = 0;
  for (int i = 0; i < ep->n_oct; i++) {
    ep->oct[i].tiles_x = (ep->oct[i].w + XT_TW - 1) / XT_TW;
    row_tiles += (long long)ep->oct[i].h * ep->oct[i].tiles_x;
  }
  const long long slots = (long long)sm_count * K2_MINB * K2_WAVES;
  int rows = (int)((row_tiles + slots - 1) / slots);
  rows = ((rows + 2) / 3) * 3;
  if (rows < 6) rows = 6;
  if (rows > XT_MAX_ROWS) rows = XT_MAX_ROWS;
  ep->rows = rows;
  int ctas = 0;
  for (int i = 0; i < ep->n_oct; i++) {
    ep->oct[i].cta_begin = ctas;
    ctas += ep->oct[i].tiles_x * ((ep->oct[i].h + rows - 1) / rows);
  }
  return ctas;
}

int make_dog_tensor_map(CUtensorMap *out, const float *dog, int h, int pitch) {
  // x, interleaved (row, plane): a (pitch x 7h) matrix
  return csb_tmap_2d_f32(out, dog, (uint64_t)pitch, (uint64_t)h * NPL, (uint64_t)pitch * sizeof(float), ST_COLS, ST_ROWS * NPL);
}

void launch_find_points(const ExtremaParams &ep, const ExtremaMaps &maps, int n_ctas, KpStage *d_stage,
                        unsigned int *d_counter, int max_pts, cudaStream_t st) {
  if (n_ctas <= 0) return;
  // CSB_XT_CAP shrinks the per-CTA list (tests of the dense fallback); read on every launch so that a test can set it
  // for one context without needing a fresh process
  int cap = XT_CAP;
  if (const char *e = getenv("CSB_XT_CAP")) {
    cap = atoi(e);
    if (cap < 1 || cap > XT_CAP) cap = XT_CAP;
  }
  constexpr size_t ring_bytes = (size_t)ST_N * ST_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_find_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes);
    attr_set = true;
  }
  k_find_points<<<n_ctas, (XT_WARPS + 1) * 32, ring_bytes, st>>>(ep, maps, d_stage, d_counter, max_pts, cap);
}
