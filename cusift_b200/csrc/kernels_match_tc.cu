// Brute-force matching on the 5th-generation tensor cores (tcgen05 / TMEM), with
// exact fp32 rescoring so that the results equal the reference's.
//
// MatchSiftData is the one dense contraction of the hot path: S = D1 * D2^T with
// K = 128 (reference: ComputeDistance, extras/matching.cu:52-98, a 16x16-tile fp32
// CUDA-core kernel that writes the n1 x n2 matrix to HBM, then FindMinCorr/FindMaxCorr
// :116-270 reads it back).  Here:
//
//  1. k_pack_f16     descriptors (fp32, 588-byte AoS records) -> fp16, K-major,
//                    128-byte-swizzled operand tiles in global memory, laid out so that
//                    any block of rows of one 64-wide K half is ONE contiguous range.
//  2. k_match_tc     persistent-style CTA per (256 queries, slice of the candidates):
//                      warp 0   producer: 1-D bulk copies (cp.async.bulk -> UBLKCP) of the
//                               query tiles (once) and of the 128-candidate tiles (4 stages),
//                               completion on mbarriers (expect_tx);
//                      warp 1   one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                               M128 x N128 x K16, 8 per accumulator, operands straight from
//                               shared memory (smem descriptors, SWIZZLE_128B), fp32
//                               accumulators in TMEM: FOUR 128x128 accumulators (512 columns),
//                               two per 128-query half, so the MMAs of tile i+1 run while the
//                               epilogue still reads tile i (with one 128x256 accumulator per half
//                               the tensor pipe idled for the whole arithmetic of the epilogue:
//                               measured 25 us of MMA + 15 us of exposed epilogue per 8192^2);
//                               tcgen05.commit releases the smem stage / publishes the accumulator;
//                      warps 2-9  epilogue: thread = one query row (one TMEM lane), tcgen05.ld 32
//                               columns at a time.
//                               sweep 1 runs over the FIRST QUARTER of the slice only and keeps
//                               m1 >= m2, the largest approximate dot product and a lower bound of
//                               the second largest (any subset of the columns gives a valid lower
//                               bound);
//                               sweep 2 runs over the whole slice, the tiles sweep 1 has not seen
//                               first: it lists every candidate with approximate dot >= m2 - 2 eps
//                               (chunk-max filter, so the divergent append path is rare) and keeps
//                               raising m2 from the chunk maxima of the unseen tiles, so that the
//                               threshold has its final value when the seen quarter comes round
//                               again.  The threshold never decreases, so a candidate at or above
//                               the FINAL threshold is always listed; the expected over-listing is
//                               2 ln 4 - 1.5 ~ 1.3 entries per query and split.
//                               Any candidate left out has exact dot < m2 - eps <= the exact dots
//                               of at least two listed ones, so the exact best and second best
//                               are always in the list.  1.25 passes of MMA instead of 2.
//     What bounds it (measured on the B200, 8192 x 8192, ablation builds): with the epilogue arithmetic removed the
//     kernel takes 23 us for 1.25 sweeps - the TMEM READ PORT: 64 B/clk per SM, i.e. 14.4 us for one pass over the
//     8192^2 fp32 scores on 148 SMs, whatever the MMA shape (N = 256 with two accumulators: 70 B/clk, N = 128 with four:
//     57 B/clk, 16-column loads from sixteen warps: 64 B/clk).  The MMAs themselves need 15.6 us.  With the arithmetic
//     the kernel takes 31 us: one epilogue warp turns over ~6 B/clk (load -> wait -> dependent max chains -> rare
//     divergent append), eight of them 45 B/clk.  Tried on top of this design and measured slower or equal: sixteen
//     epilogue warps on column halves (more threads per row = more hits: each thread's threshold only knows its own
//     columns), thresholds shared between the CTAs of a query tile through global memory (the hits a shared bound
//     would save happen in the first tiles of sweep 2, before anything published can arrive), a common sample of tiles
//     as sweep 1 in every CTA (1.5 sweeps of TMEM reads), sweep 1 over an eighth of the slice (raw lists overflow).
//  3. k_rescore      eight queries per warp: the splits' (m1, m2) give a threshold over ALL candidates that drops
//                    most listed entries; exact fp32 scores of the survivors in the
//                    reference's rotated k order (bit-identical to ComputeDistance), then the
//                    reference's best / second-best rule incl. its tie-breaking
//                    (FindMinCorr/FindMaxCorr).  A query whose list overflowed (> 16 entries per
//                    split: massive near-ties) is appended to the redo list and recomputed by the exact
//                    fp32 kernel (kernels_match.cu).
//
// fp16 inputs and the error bound eps.  With q^ = fp16(q), dq = q - q^ (same for c):
//   |q.c - q^.c^| = |dq.c + q^.dc| <= |dq| |c| + |q^| |dc|            (Cauchy-Schwarz, any signs)
// k_pack_f16 MEASURES |dq| of every row while it packs (the rounding errors are known exactly there) and keeps the
// maximum over the set, E; it also checks that every element is finite in fp16 and that no squared norm exceeds
// 1.002.  A pair of sets (i, j) then uses eps = 1.002 (E_i + E_j) + 2e-5 (the last term covers the tensor core's
// fp32 accumulation of 128 products <= 1): a rigorous bound, ~6e-4 for unit SIFT / RootSIFT descriptors instead of
// the worst case 2^-10 = 9.8e-4.  Sets outside the checked domain (un-normalised, 0..255, huge values) raise a flag
// and csb_match / the all-pairs path route the call to the exact fp32 kernel, so arbitrary SiftPoint.data still gets
// the reference's exact result.
#include <cuda_fp16.h>
#include <cuda_pipeline.h>

#include "csb_internal.h"

namespace {

constexpr int TC_QT = 256;        // queries per CTA (two 128-row accumulators)
constexpr int TC_CT = 128;        // candidates per tile (UMMA N)
constexpr int TC_STAGES = 4;
#ifndef TC_S1_DIV
#define TC_S1_DIV 4               // sweep 1 covers the last 1/TC_S1_DIV of the candidate slice
#endif
constexpr int TC_TOPK = 16;        // listed candidates per query and split handed to the rescoring kernel (8 made clusters of
                                   // near-identical descriptors overflow: one to five 16-query blocks per 8192^2 match went to the exact kernel)
constexpr int TC_RAW = 16;         // candidates per query a CTA can hold while the threshold is still rising
constexpr int KHALF_BYTES_PER_ROW = 128;          // 64 fp16
constexpr uint32_t A_HALF_BYTES = 128 * KHALF_BYTES_PER_ROW;      // one 128-row x 64-K operand block: 16 KB
constexpr uint32_t B_HALF_BYTES = TC_CT * KHALF_BYTES_PER_ROW;    // 16 KB
constexpr uint32_t SMEM_A = 2 /*query halves*/ * 2 /*K halves*/ * A_HALF_BYTES;   // 64 KB
constexpr uint32_t SMEM_B_STAGE = 2 * B_HALF_BYTES;                                 // 32 KB
constexpr uint32_t SMEM_RAW = TC_RAW * TC_QT * 8;                                   // raw list: value + index per entry, 32 KB
constexpr uint32_t SMEM_TC = SMEM_A + TC_STAGES * SMEM_B_STAGE + SMEM_RAW + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 32 * (2 + 8);
constexpr float TC_MAX_NORM2 = 1.002f;   // squared-norm limit of the domain in which the error bound is derived

// ---- PTX wrappers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: 8-row x 128-byte atoms
// (1024 B), consecutive atoms along M/N 1024 B apart (SBO); LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address >> 4, bits [0,14)
  d |= (uint64_t)0 << 16;                                // leading byte offset >> 4 (ignored)
  d |= (uint64_t)(1024u >> 4) << 32;                     // stride byte offset >> 4, bits [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: fp16 x fp16 -> fp32, both operands K-major, M=128, N=TC_CT.
constexpr uint32_t IDESC_F16_M128 = (1u << 4)                 // D format = F32
                                         | (0u << 7) | (0u << 10)   // A, B format = F16
                                         | (0u << 15) | (0u << 16)  // A, B K-major
                                         | ((uint32_t)(TC_CT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// ---- 1. pack ----------------------------------------------------------------------
// Packed layout (per set): two K-halves; half kb is a [n_pad][64] fp16 matrix (128-byte rows) in
// which the 16-byte chunk c of row r is stored at chunk position c ^ (r & 7) (128B swizzle).
// Also validates the domain of the error bound and measures its ingredients (see the header): info[0] is set when a
// row has a non-finite / fp16-overflowing element or a squared norm above TC_MAX_NORM2, or when a set that needs
// padding rows (n not a multiple of 256) has a negative element (the scan treats the padding's score 0 as a lower
// bound of real scores, which only holds for non-negative descriptors - every SIFT / RootSIFT descriptor is); info[1] receives (as float
// bits) the largest squared rounding-error norm |q - fp16(q)|^2 over the rows.  The caller zeroes info[0..1].
__global__ void __launch_bounds__(256) k_pack_f16(const csb_sift_point *__restrict__ pts, int n, int n_pad,
                                                  __half *__restrict__ packed, int *__restrict__ info) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (row, 16-byte chunk): 16 chunks/row
  const int r = gid >> 4, c16 = gid & 15;
  const bool live = r < n_pad;                             // n_pad * 16 is a multiple of 32: whole warps are live or not
  const int kb = c16 >> 3, c = c16 & 7;
  __align__(16) __half v[8];
  float ss = 0.0f, es = 0.0f;
  bool bad = false;
  if (live && r < n) {
    const float *d = pts[r].data + 8 * c16;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float f = d[i];
      v[i] = __float2half_rn(f);
      ss = __fmaf_rn(f, f, ss);
      const float de = f - __half2float(v[i]);             // exact (Sterbenz / representable difference)
      es = __fmaf_rn(de, de, es);
      bad = bad || !(fabsf(f) <= 65504.0f);               // NaN, inf, or beyond the fp16 range
      bad = bad || (f < 0.0f && n < n_pad);               // padding rows score 0: only below every real score if those are >= 0
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __float2half_rn(0.0f);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {                        // the row's 16 threads are adjacent lanes
    ss += __shfl_xor_sync(0xffffffffu, ss, o, 16);
    es += __shfl_xor_sync(0xffffffffu, es, o, 16);
  }
  if (bad || !(ss <= TC_MAX_NORM2)) info[0] = 1;
  else if (c16 == 0 && es > 0.0f) atomicMax(info + 1, __float_as_int(es * 1.0001f));   // non-negative floats order like ints
  if (!live) return;
  char *dst = reinterpret_cast<char *>(packed) + (size_t)kb * n_pad * KHALF_BYTES_PER_ROW +
              (size_t)r * KHALF_BYTES_PER_ROW + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(v);
}

// ---- 2. tensor-core scan --------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) k_match_tc(const __half *__restrict__ q_packed, int nq_pad,
                                                            const __half *__restrict__ c_packed, int nc, int nc_pad,
                                                            int tiles_per_split, float *__restrict__ out_val,
                                                            int *__restrict__ out_idx, int n_splits,
                                                            const int *__restrict__ q_info, const int *__restrict__ c_info,
                                                            int *__restrict__ redo_flags, int n_redo_flags,
                                                            int *__restrict__ redo_count) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the swizzle atoms
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = base;                          // [qhalf][khalf][128 rows][128 B]
  unsigned char *sB = base + SMEM_A;                 // [stage][khalf][256 rows][128 B]
  float2 *s_raw = reinterpret_cast<float2 *>(base + SMEM_A + TC_STAGES * SMEM_B_STAGE);   // [TC_RAW][TC_QT] {value, index}
  uint64_t *bars = reinterpret_cast<uint64_t *>(base + SMEM_A + TC_STAGES * SMEM_B_STAGE + SMEM_RAW);
  uint64_t *a_full = bars + 0;
  uint64_t *b_full = bars + 1;                       // [TC_STAGES]
  uint64_t *b_empty = bars + 1 + TC_STAGES;          // [TC_STAGES]
  uint64_t *acc_full = bars + 1 + 2 * TC_STAGES;     // [4]: accumulator 2 qh + (iteration & 1)
  uint64_t *acc_empty = bars + 5 + 2 * TC_STAGES;    // [4]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9 + 2 * TC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x, split = blockIdx.y;
  const int n_tiles_total = nc_pad / TC_CT;
  const int t0 = split * tiles_per_split;
  const int t1 = min(t0 + tiles_per_split, n_tiles_total);
  const int n_tiles = max(t1 - t0, 0);
  // iteration schedule (producer, MMA issuer and epilogue all follow it): sweep 1 = the first n1 tiles of the
  // slice (padding, if any, is at the end), then sweep 2 = the other tiles followed by those n1 tiles again
  const int n1 = (n_tiles + TC_S1_DIV - 1) / TC_S1_DIV;
  const int n_iter = n1 + n_tiles;
  auto tile_of = [&](int it) { return t0 + (it < n_tiles ? it : it - n_tiles); };

  if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < TC_STAGES; s++) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 1);
    }
    for (int a = 0; a < 4; a++) {
      mbar_init(acc_full + a, 1);
      mbar_init(acc_empty + a, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: all 512 columns (four 128 x 128 fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the rescoring kernel that follows appends the 16-query blocks it cannot prove to this list: reset it here
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int i = threadIdx.x; i < n_redo_flags; i += blockDim.x) redo_flags[i] = 0;
    if (threadIdx.x == 0) *redo_count = 0;
  }

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      mbar_expect_tx(a_full, SMEM_A);
      for (int qh = 0; qh < 2; qh++)
        for (int kb = 0; kb < 2; kb++)
          bulk_g2s(sA + (qh * 2 + kb) * A_HALF_BYTES,
                   reinterpret_cast<const char *>(q_packed) + (size_t)kb * nq_pad * KHALF_BYTES_PER_ROW +
                       (size_t)(qtile * TC_QT + qh * 128) * KHALF_BYTES_PER_ROW,
                   A_HALF_BYTES, a_full);
      for (int i = 0; i < n_iter; i++) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        const int tile = tile_of(i);
        mbar_wait(b_empty + s, ph ^ 1);
        mbar_expect_tx(b_full + s, SMEM_B_STAGE);
        for (int kb = 0; kb < 2; kb++)
          bulk_g2s(sB + s * SMEM_B_STAGE + kb * B_HALF_BYTES,
                   reinterpret_cast<const char *>(c_packed) + (size_t)kb * nc_pad * KHALF_BYTES_PER_ROW +
                       (size_t)tile * TC_CT * KHALF_BYTES_PER_ROW,
                   B_HALF_BYTES, b_full + s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      mbar_wait(a_full, 0);
      tc_fence_after();
      for (int i = 0; i < n_iter; i++) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(b_full + s, ph);
        tc_fence_after();
        for (int qh = 0; qh < 2; qh++) {
          // accumulator (qh, i & 1) is reused every other iteration: wait until the epilogue drained its previous use
          const int a = qh * 2 + (i & 1);
          mbar_wait(acc_empty + a, ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * TC_CT);
#pragma unroll
          for (int kb = 0; kb < 2; kb++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint64_t ad = smem_desc_k128(smem_u32(sA + (qh * 2 + kb) * A_HALF_BYTES) + k * 32);
              const uint64_t bd = smem_desc_k128(smem_u32(sB + s * SMEM_B_STAGE + kb * B_HALF_BYTES) + k * 32);
              tc_mma_f16(d_tmem, ad, bd, IDESC_F16_M128, (kb | k) ? 1u : 0u);
            }
          }
          tc_commit(acc_full + a);        // accumulator complete -> epilogue
        }
        tc_commit(b_empty + s);           // both MMAs of this stage done -> producer may refill
      }
    }
  } else {
    // ===== epilogue: 8 warps, thread = one query row =====
    const int ew = warp - 2;                 // 0..7
    const int qh = ew >> 2;                  // query half
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int row = qh * 128 + quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(qh * 2 * TC_CT);
    // ---- sweep 1: m1 = largest approximate dot product of this row among the tiles seen, m2 = second largest
    // CHUNK maximum (chunks of 32 columns).  m2 <= the true second largest value, so thr = m2 - 2 eps is a valid
    // (slightly more inclusive) listing threshold, and a chunk costs 16 three-input max + 3 ops instead
    // of 96 (FMNMX runs at half rate).  Padding columns hold 0: never above a real score, because sets with
    // padding rows are checked to be non-negative (k_pack_f16).
    // TMEM reads are software-pipelined: the tcgen05.ld of the next 64 columns is in flight while the
    // current 64 are reduced (tcgen05.wait::ld waits for every outstanding load, so it sits after the
    // arithmetic), and the accumulator is handed back to the MMA warp as soon as its last columns are in
    // registers, before they are reduced.
    float m1 = -1.0f, m2 = -1.0f;
    auto top2 = [&](float c) {
      const float lo = fminf(m1, c);
      m1 = fmaxf(m1, c);
      m2 = fmaxf(m2, lo);
    };
    auto reduce64 = [&](const uint32_t (&r)[32], const uint32_t (&q)[32]) {
#ifdef TC_NOALU
      m1 = fmaxf(m1, __uint_as_float(r[0] ^ q[31] ^ r[17])); m2 = 0.2f;
      return;
#endif
      float g[8];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        g[k] = fmaxf(fmaxf(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1])), __uint_as_float(r[8 * k + 2]));
        g[k] = fmaxf(fmaxf(g[k], __uint_as_float(r[8 * k + 3])), __uint_as_float(r[8 * k + 4]));
        g[k] = fmaxf(fmaxf(g[k], __uint_as_float(r[8 * k + 5])), __uint_as_float(r[8 * k + 6]));
        g[k] = fmaxf(g[k], __uint_as_float(r[8 * k + 7]));
        g[4 + k] = fmaxf(fmaxf(__uint_as_float(q[8 * k]), __uint_as_float(q[8 * k + 1])), __uint_as_float(q[8 * k + 2]));
        g[4 + k] = fmaxf(fmaxf(g[4 + k], __uint_as_float(q[8 * k + 3])), __uint_as_float(q[8 * k + 4]));
        g[4 + k] = fmaxf(fmaxf(g[4 + k], __uint_as_float(q[8 * k + 5])), __uint_as_float(q[8 * k + 6]));
        g[4 + k] = fmaxf(g[4 + k], __uint_as_float(q[8 * k + 7]));
      }
      top2(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])));
      top2(fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
    };
    static_assert((TC_RAW & (TC_RAW - 1)) == 0, "raw list slots wrap with a mask");
    static_assert(TC_CT == 128, "both sweeps are unrolled for four 32-column chunks per tile");
    int it = 0;
    for (; it < n1; it++) {
      const int a = it & 1;
      const uint32_t taddr = tlane + (uint32_t)(a * TC_CT);
      mbar_wait(acc_full + qh * 2 + a, (it >> 1) & 1);
      tc_fence_after();
      uint32_t ra[32], qa[32], rb[32], qb[32];
      tc_ld32(taddr + 0, ra);
      tc_ld32(taddr + 32, qa);
      tc_ld_wait();
      tc_ld32(taddr + 64, rb);
      tc_ld32(taddr + 96, qb);
      reduce64(ra, qa);
      tc_ld_wait();
      tc_fence_before();
      mbar_arrive(acc_empty + qh * 2 + a);   // 128 arrivals free the accumulator
      reduce64(rb, qb);
    }
    // ---- sweep 2: list every candidate with approximate dot >= m2 - 2 eps ----
    // eps of this pair of sets from the rounding-error norms measured at pack time (header comment)
    const float eps2 = 2.0f * (1.002f * (sqrtf(__int_as_float(q_info[1])) + sqrtf(__int_as_float(c_info[1]))) + 2.0e-5f);
    float thr = m2 - eps2;
    // Hits go to a per-row raw list in shared memory ([entry][row]: the lanes of a warp write consecutive words),
    // value and index, because the threshold is still rising: once it is final the row's entries are filtered
    // again and only the survivors reach the global short list.  The append path runs for the whole warp whenever
    // ANY lane has a hit in an 8-column group, so it is kept short and branch-free inside.
    int cnt = 0;
    // one 32-column chunk: the four 8-column groups are tested against the threshold as it stands, then (tiles
    // that sweep 1 has not seen only: a value must not enter the top-2 twice) the chunk maximum raises it
    auto list32 = [&](const uint32_t (&r)[32], int cbase, bool unseen) {
#ifdef TC_NOALU
      if (__uint_as_float(r[3] ^ r[30]) == 123.25f) cnt++;
      return;
#endif
      float g[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        g[k] = fmaxf(fmaxf(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1])), __uint_as_float(r[8 * k + 2]));
        g[k] = fmaxf(fmaxf(g[k], __uint_as_float(r[8 * k + 3])), __uint_as_float(r[8 * k + 4]));
        g[k] = fmaxf(fmaxf(g[k], __uint_as_float(r[8 * k + 5])), __uint_as_float(r[8 * k + 6]));
        g[k] = fmaxf(g[k], __uint_as_float(r[8 * k + 7]));
        if (g[k] >= thr) {
          // Rare per lane, but the whole warp waits for the lanes in here, so the usual case (one hit in the group)
          // is straight-line code: a mask of the group's hits (8 independent compares), the first hit goes to slot
          // cnt mod TC_RAW with the GROUP maximum as its value (exact for a single hit, an upper bound otherwise: the
          // final filter then keeps a superset, never less).  A count above TC_RAW means the list wrapped and the
          // row is handed to the exact kernel.  Padding columns (score 0) are removed by the final filter.
          unsigned int m = 0;
#pragma unroll
          for (int j = 0; j < 8; j++) m |= (__uint_as_float(r[8 * k + j]) >= thr) ? (1u << j) : 0u;
          s_raw[(cnt & (TC_RAW - 1)) * TC_QT + row] = make_float2(g[k], __int_as_float(cbase + 8 * k + (__ffs(m) - 1)));
          cnt++;
          m &= m - 1;
          while (m) {                                // further hits in the same 8 columns: near-duplicates
            s_raw[(cnt & (TC_RAW - 1)) * TC_QT + row] = make_float2(g[k], __int_as_float(cbase + 8 * k + (__ffs(m) - 1)));
            cnt++;
            m &= m - 1;
          }
        }
      }
      if (unseen) {
        top2(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])));
        thr = m2 - eps2;
      }
    };
    // 64 columns (8 KB per warp) are in flight while the previous 64 are examined, and the first 64 of tile it + 1 are
    // requested before the last 64 of tile it are examined (the MMA warp is up to two accumulators ahead): with
    // 64 KB outstanding per SM the TMEM read port - 64 B/clk - always has a queue.  (32-column double buffering left
    // it idle whenever most warps were in their arithmetic; tcgen05.wait::ld waits for ALL outstanding loads, so
    // deeper pipelines have to come as larger batches.)
    uint32_t ra[32], qa[32], rb[32], qb[32];
    if (it < n_iter) {
      mbar_wait(acc_full + qh * 2 + (it & 1), (it >> 1) & 1);
      tc_fence_after();
      tc_ld32(tlane + (uint32_t)((it & 1) * TC_CT), ra);
      tc_ld32(tlane + (uint32_t)((it & 1) * TC_CT + 32), qa);
    }
    for (; it < n_iter; it++) {
      const int a = it & 1;
      const uint32_t taddr = tlane + (uint32_t)(a * TC_CT);
      const bool unseen = it < n_iter - n1;
      const int col0 = tile_of(it) * TC_CT;
      tc_ld_wait();
      tc_ld32(taddr + 64, rb);
      tc_ld32(taddr + 96, qb);
      list32(ra, col0, unseen);
      list32(qa, col0 + 32, unseen);
      tc_ld_wait();
      tc_fence_before();
      mbar_arrive(acc_empty + qh * 2 + a);
      if (it + 1 < n_iter) {
        mbar_wait(acc_full + qh * 2 + (a ^ 1), ((it + 1) >> 1) & 1);
        tc_fence_after();
        tc_ld32(tlane + (uint32_t)((a ^ 1) * TC_CT), ra);
        tc_ld32(tlane + (uint32_t)((a ^ 1) * TC_CT + 32), qa);
      }
      list32(rb, col0 + 64, unseen);
      list32(qb, col0 + 96, unseen);
    }
    // the threshold is final: keep the entries that reach it (each thread reads back its own writes only).
    // Short list of this (row, split): index -1 = unused slot, -2 in slot 0 = the list proves nothing (overflow);
    // value = approximate dot product (an upper bound for hits that shared an 8-column group); plus the
    // split's (m1, m2), from which the rescoring kernel derives the threshold over ALL splits.
    const size_t slot = (size_t)(qtile * TC_QT + row) * n_splits + split;
    int *const out_list = out_idx + slot * TC_TOPK;
    float *const out_lval = out_val + slot * TC_TOPK;
    // all slots are first cleared with 128-bit stores (a list is 64-byte aligned), then the survivors - a few - are
    // written over them (same thread: ordered)
    static_assert(TC_TOPK % 4 == 0, "vector clears");
#pragma unroll
    for (int k = 0; k < TC_TOPK / 4; k++) {
      reinterpret_cast<int4 *>(out_list)[k] = make_int4(-1, -1, -1, -1);
      reinterpret_cast<float4 *>(out_lval)[k] = make_float4(0.f, 0.f, 0.f, 0.f);   // the rescoring kernel loads values and indices together
    }
    int kept = 0;
    for (int e = 0; e < min(cnt, TC_RAW); e++) {
      const float2 en = s_raw[e * TC_QT + row];
      const int cidx = __float_as_int(en.y);
      if (en.x >= thr && cidx < nc) {
        if (kept < TC_TOPK) {
          out_list[kept] = cidx;
          out_lval[kept] = en.x;
        }
        kept++;
      }
    }
    if (cnt > TC_RAW || kept > TC_TOPK) out_list[0] = -2;   // wrapped raw list or more survivors than slots
    reinterpret_cast<float2 *>(out_val + (size_t)nq_pad * 4 * TC_TOPK)[slot] = make_float2(m1, m2);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---- 3. exact rescoring -----------------------------------------------------------------
__device__ __forceinline__ int bitrev4(int x) { return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3); }

// One warp rescoring RS_QPW queries at a time.  (The first version gave every query its own warp: about ten listed
// candidates = ten busy lanes, each walking a 128-step chain with two shared-memory reads per step, so the kernel
// was bound by the shared-memory pipe at a third of its lane capacity, and every split's private top-2 was rescored
// although only the top-2 over all splits matter.)
//   1. filter: lane = RS_H list entries of one query (n_splits x TC_TOPK <= 32 RS_H).  The splits' (m1, m2) pairs give a lower
//      bound of the second largest approximate dot product over ALL candidates (their chunks are disjoint); entries
//      whose value (an upper bound) is below it minus 2 eps cannot be best or second best and are dropped.
//      About 2.3 entries per query survive instead of 10.
//   2. the survivors of the warp's queries are packed into consecutive lanes (work items); their descriptors and
//      the queries' are staged in shared memory with asynchronous copies (all rows in flight at once);
//   3. lane = work item: exact score in the reference's rotated k order (matching.cu:84-89);
//   4. lane = query: FindMinCorr/FindMaxCorr's best / second-best rule over the query's few items.
constexpr int RS_QPW = 8;           // queries per warp and round (the (m1, m2) merge maps lane = 4 query + split)
constexpr int RS_WARPS = 4;
constexpr int RS_H = (4 * TC_TOPK + 31) / 32;   // list entries per lane (up to four splits x TC_TOPK entries per query)
constexpr int RS_QS = 129, RS_CS = 129;   // row strides (words): rows spread over the banks
constexpr size_t RS_SMEM_WARP = sizeof(float) * (RS_QPW * RS_QS + 32 * RS_CS) + sizeof(int) * (32 + 32);

template <bool kL2>
__global__ void __launch_bounds__(RS_WARPS * 32) k_rescore(csb_sift_point *__restrict__ s1, int n1,
                                                           const csb_sift_point *__restrict__ s2, int n2,
                                                           const float *__restrict__ sl_val, const int *__restrict__ sl_idx,
                                                           int n_splits, int nq_pad, const int *__restrict__ q_info,
                                                           const int *__restrict__ c_info, int *__restrict__ redo_flags,
                                                           int *__restrict__ redo_list, int *__restrict__ redo_count) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *sq = reinterpret_cast<float *>(rs_smem + wib * RS_SMEM_WARP);   // [RS_QPW][RS_QS]
  float *sc = sq + RS_QPW * RS_QS;                                       // [32][RS_CS]
  int *s_ci = reinterpret_cast<int *>(sc + 32 * RS_CS);                  // candidate index of work item
  int *s_qi = s_ci + 32;                                                 // local query of work item
  const int n_list = n_splits * TC_TOPK;          // <= 32 RS_H
  const float eps2 = 2.0f * (1.002f * (sqrtf(__int_as_float(q_info[1])) + sqrtf(__int_as_float(c_info[1]))) + 2.0e-5f);
  const float2 *sl_top = reinterpret_cast<const float2 *>(sl_val + (size_t)nq_pad * 4 * TC_TOPK);
  const float NONE = kL2 ? 999.0f : -1.0f;
  const unsigned FULLM = 0xffffffffu;

  for (int q0 = (blockIdx.x * RS_WARPS + wib) * RS_QPW; q0 < n1; q0 += gridDim.x * RS_WARPS * RS_QPW) {
    // ---- 1. filter; survivors stay in registers: lane l of query j -> keepm[j] bit l.  All loads of the warp's
    // queries are issued before the first is used (one L2 round trip, not eight).
    int my_ci[RS_QPW][RS_H];
    float my_v[RS_QPW][RS_H];
#pragma unroll
    for (int j = 0; j < RS_QPW; j++) {
#pragma unroll
      for (int h = 0; h < RS_H; h++) {
        my_ci[j][h] = -1;
        my_v[j][h] = -3.0e38f;
        if (q0 + j < n1 && lane + 32 * h < n_list) {
          my_ci[j][h] = sl_idx[(size_t)(q0 + j) * n_list + lane + 32 * h];
          my_v[j][h] = sl_val[(size_t)(q0 + j) * n_list + lane + 32 * h];
        }
      }
    }
    // second largest approximate value over all splits, from below: merge the splits' (m1, m2); lane = 4 j + split
    float a1 = -3.0e38f, a2 = -3.0e38f;
    {
      const int j = lane >> 2, sp = lane & 3;
      if (q0 + j < n1 && sp < n_splits) {
        const float2 t = sl_top[(size_t)(q0 + j) * n_splits + sp];
        a1 = t.x, a2 = t.y;
      }
#pragma unroll
      for (int o = 1; o < 4; o <<= 1) {
        const float b1 = __shfl_xor_sync(FULLM, a1, o), b2 = __shfl_xor_sync(FULLM, a2, o);
        const float lo = fminf(a1, b1);
        a1 = fmaxf(a1, b1);
        a2 = fmaxf(fmaxf(a2, b2), lo);
      }
    }
    unsigned int keepm[RS_QPW][RS_H];
    int kcnt[RS_QPW];
    unsigned int overflow_q = 0;                  // bit j: the lists of query j prove nothing
#pragma unroll
    for (int j = 0; j < RS_QPW; j++) {
      const float thr = __shfl_sync(FULLM, a2, 4 * j) - eps2;
      bool ovf = false;
      kcnt[j] = 0;
#pragma unroll
      for (int h = 0; h < RS_H; h++) {
        ovf = ovf || my_ci[j][h] == -2;
        const bool keep = my_ci[j][h] >= 0 && my_v[j][h] >= thr;
        if (!keep) my_ci[j][h] = -1;
        keepm[j][h] = __ballot_sync(FULLM, keep);
        kcnt[j] += __popc(keepm[j][h]);
      }
      if (__any_sync(FULLM, ovf) || kcnt[j] > 32) {   // (more than 32 survivors of one query do not fit a round either)
        overflow_q |= 1u << j;
        kcnt[j] = 0;
#pragma unroll
        for (int h = 0; h < RS_H; h++) {
          my_ci[j][h] = -1;
          keepm[j][h] = 0;
        }
      }
    }
    // ---- rounds: consecutive queries whose survivors fit the 32 lanes (a single query fits: kcnt <= 32)
    int jb = 0;
    while (jb < RS_QPW) {
      int je = jb, n_items = 0;
      int start[RS_QPW + 1];
#pragma unroll
      for (int j = 0; j < RS_QPW; j++) {
        start[j] = n_items;
        if (j >= jb && j == je && n_items + kcnt[j] <= 32) {
          n_items += kcnt[j];
          je = j + 1;
        }
      }
      start[RS_QPW] = n_items;
      // start[j] for j in [jb, je) is the first item of query j; queries outside the round have empty ranges
      __syncwarp();
      // ---- 2. work list + staging
#pragma unroll
      for (int j = 0; j < RS_QPW; j++) {
#pragma unroll
        for (int h = 0; h < RS_H; h++) {
          if (j >= jb && j < je && my_ci[j][h] >= 0) {
            int w = start[j] + __popc(keepm[j][h] & ((1u << lane) - 1u));
            for (int g = 0; g < h; g++) w += __popc(keepm[j][g]);
            s_ci[w] = my_ci[j][h];
            s_qi[w] = j;
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < RS_QPW; j++) {
        if (j >= jb && j < je && q0 + j < n1 && kcnt[j]) {
          const float *pq = s1[q0 + j].data;
#pragma unroll
          for (int t = 0; t < 4; t++) __pipeline_memcpy_async(&sq[j * RS_QS + lane + 32 * t], &pq[lane + 32 * t], 4);
        }
      }
      // candidate rows are stored ROTATED by tx = index % 16 (element k at (k - tx) mod 128), so that the chain below
      // reads its row front to back: lane w, step i -> bank (w + i) mod 32, conflict-free whatever the indices are
      for (int w = 0; w < n_items; w++) {
        const int cw = s_ci[w];
        const float *pb = s2[cw].data;            // 4-byte copies: data[] is only 4-byte aligned
        const int tx = cw & 15;
#pragma unroll
        for (int t = 0; t < 4; t++)
          __pipeline_memcpy_async(&sc[w * RS_CS + ((lane + 32 * t - tx) & 127)], &pb[lane + 32 * t], 4);
      }
      __pipeline_commit();
      // lane = work item from here on; its candidate's coordinates are fetched now and used in step 4
      int ci = -1, qi = 0;
      float cx = 0.0f, cy = 0.0f;
      if (lane < n_items) {
        ci = s_ci[lane];
        qi = s_qi[lane];
        cx = s2[ci].coords2D[0];
        cy = s2[ci].coords2D[1];
      }
      __pipeline_wait_prior(0);
      __syncwarp();
      // ---- 3. exact scores: sum over i of q[(i + tx) & 127] * c[(i + tx) & 127], i = 0..127 (matching.cu:84-89)
      float score = NONE;
      if (lane < n_items) {
        const float *pq = sq + qi * RS_QS, *pb = sc + lane * RS_CS;
        const int tx = ci & 15;
        float sum = 0.0f;
        // the first 112 steps never wrap on the query side (tx <= 15); only the last 16 need the mask
        const float *qa = pq + tx;
#pragma unroll
        for (int i = 0; i < 112; i++) sum = __fmaf_rn(qa[i], pb[i], sum);
#pragma unroll
        for (int i = 112; i < 128; i++) sum = __fmaf_rn(pq[(i + tx) & 127], pb[i], sum);
        score = kL2 ? __fsub_rn(2.0f, __fadd_rn(sum, sum)) : sum;
      }
      __syncwarp();
      // scores and indices of the items go through shared memory (the staging tile's first row is done with)
      float *s_sc = sc, *s_xy = sc + 32;          // 32 + 64 floats
      if (lane < n_items) {
        s_sc[lane] = score;
        s_xy[2 * lane] = cx;
        s_xy[2 * lane + 1] = cy;
      }
      __syncwarp();
      // ---- 4. lane j = query j of the round: best = lowest (score, bitrev4(col % 16), col / 16) as FindMinCorr's
      // winner, second = best score of the rest (an equal duplicate lands in `second`, matching.cu:229-235)
      if (lane >= jb && lane < je && q0 + lane < n1) {
        const int q = q0 + lane;
        int a = 0, b = 0;
#pragma unroll
        for (int j = 0; j < RS_QPW; j++)
          if (j == lane) a = start[j], b = start[j] + kcnt[j];
        if (overflow_q & (1u << lane)) {
          // the exact kernel redoes this block of 16 queries; the first query to flag a block lists it
          if (atomicExch(redo_flags + (q >> 4), 1) == 0) redo_list[atomicAdd(redo_count, 1)] = q >> 4;
        } else {
          auto better = [](float sa, int ia, float sb, int ib) {   // is a strictly preferred to b ?
            if (ib < 0) return ia >= 0;
            if (ia < 0) return false;
            if (sa != sb) return kL2 ? (sa < sb) : (sa > sb);
            const int ra = bitrev4(ia & 15), rb = bitrev4(ib & 15);
            if (ra != rb) return ra < rb;
            return ia < ib;
          };
          float bs = NONE;
          int bi = -1, bw = 0;
          for (int w = a; w < b; w++) {
            const float os = s_sc[w];
            const int oi = s_ci[w];
            if (better(os, oi, bs, bi)) bs = os, bi = oi, bw = w;
          }
          float ss = NONE;
          for (int w = a; w < b; w++)
            if (s_ci[w] != bi) ss = kL2 ? fminf(ss, s_sc[w]) : fmaxf(ss, s_sc[w]);
          csb_sift_point *o = s1 + q;
          o->score = bs;
          if (kL2) o->ambiguity = (float)((double)bs / ((double)ss + 1e-6));
          else o->ambiguity = (float)((double)__fsub_rn(1.0f, bs) / ((double)__fsub_rn(1.0f, ss) + 1e-6));
          o->match = bi;
          if (bi >= 0) {
            o->match_xpos = s_xy[2 * bw];
            o->match_ypos = s_xy[2 * bw + 1];
          }
        }
      }
      __syncwarp();
      jb = je;
    }
  }
}

}  // namespace

size_t tc_packed_bytes(int n) { return (size_t)((n + TC_QT - 1) / TC_QT * TC_QT) * 256; }
int tc_pad(int n) { return (n + TC_QT - 1) / TC_QT * TC_QT; }
int tc_splits(int n1, int n2, int sm_count) {
  const int qtiles = tc_pad(n1) / TC_QT, ctiles = tc_pad(n2) / TC_CT;
  int s = sm_count / (qtiles > 0 ? qtiles : 1);
  if (s < 1) s = 1;
  if (s > 4) s = 4;                       // the rescoring warp handles at most 4 x 8 listed candidates
  if (s > ctiles) s = ctiles;
  return s;
}

void launch_pack_f16(const csb_sift_point *pts, int n, void *packed, int *info, cudaStream_t st) {
  const int n_pad = tc_pad(n);
  const int threads = n_pad * 16;
  k_pack_f16<<<(threads + 255) / 256, 256, 0, st>>>(pts, n, n_pad, reinterpret_cast<__half *>(packed), info);
}

int launch_match_tc(const void *q_packed, int n1, const void *c_packed, int n2, int n_splits, float *sl_val, int *sl_idx,
                    const int *q_info, const int *c_info, int *redo_flags, int *redo_count, cudaStream_t st) {
  const int nq_pad = tc_pad(n1), nc_pad = tc_pad(n2);
  const int ctiles = nc_pad / TC_CT;
  const int tiles_per_split = (ctiles + n_splits - 1) / n_splits;
  {   // > 48 KB of dynamic shared memory needs the opt-in, once per device
    static bool opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
      cudaFuncSetAttribute(k_match_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TC);
      if (dev >= 0 && dev < 64) opted[dev] = true;
    }
  }
  dim3 grd(nq_pad / TC_QT, n_splits);
  k_match_tc<<<grd, TC_THREADS, SMEM_TC, st>>>(reinterpret_cast<const __half *>(q_packed), nq_pad,
                                               reinterpret_cast<const __half *>(c_packed), n2, nc_pad, tiles_per_split,
                                               sl_val, sl_idx, n_splits, q_info, c_info, redo_flags, (n1 + 15) / 16, redo_count);
  return (int)cudaPeekAtLastError();   // a failed launch must not be overwritten by the launches that follow
}

size_t tc_shortlist_ints(int n) { return (size_t)tc_pad(n) * 4 * TC_TOPK; }
size_t tc_shortlist_floats(int n) { return (size_t)tc_pad(n) * 4 * (TC_TOPK + 2); }   // entry values + (m1, m2) per split

void launch_rescore(csb_sift_point *s1, int n1, const csb_sift_point *s2, int n2, const float *sl_val, const int *sl_idx,
                    int n_splits, int distance, const int *q_info, const int *c_info, int *redo_flags, int *redo_list,
                    int *redo_count, cudaStream_t st) {
  // redo_flags / redo_count were reset by k_match_tc (launch_match_tc), which must precede this call on the stream
  constexpr size_t smem = RS_SMEM_WARP * RS_WARPS;
  {   // > 48 KB of dynamic shared memory needs the opt-in, once per device
    static bool opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !opted[dev]) {
      cudaFuncSetAttribute(k_rescore<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_rescore<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (dev >= 0 && dev < 64) opted[dev] = true;
    }
  }
  const int blocks = (n1 + RS_WARPS * RS_QPW - 1) / (RS_WARPS * RS_QPW);
  if (distance == 1)
    k_rescore<true><<<blocks, RS_WARPS * 32, smem, st>>>(s1, n1, s2, n2, sl_val, sl_idx, n_splits, tc_pad(n1), q_info, c_info, redo_flags, redo_list, redo_count);
  else
    k_rescore<false><<<blocks, RS_WARPS * 32, smem, st>>>(s1, n1, s2, n2, sl_val, sl_idx, n_splits, tc_pad(n1), q_info, c_info, redo_flags, redo_list, redo_count);
}
