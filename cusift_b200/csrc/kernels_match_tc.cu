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
//                               query tiles (once) and of the 256-candidate tiles (2 stages),
//                               completion on mbarriers (expect_tx);
//                      warp 1   one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                               M128 x N256 x K16, 8 per accumulator, operands straight from
//                               shared memory (smem descriptors, SWIZZLE_128B), fp32
//                               accumulators in TMEM: two 128x256 accumulators (512 columns),
//                               one per 128-query half, so every candidate tile is used twice;
//                               tcgen05.commit releases the smem stage / publishes the accumulator;
//                      warps 2-9  epilogue: thread = one query row (one TMEM lane), tcgen05.ld 32
//                               columns at a time.  The candidate slice is swept TWICE (the
//                               tensor pipe has time to spare, the epilogue does not):
//                               sweep 1 keeps the two largest approximate dot products
//                               m1 >= m2 per row, branch-free (3 FMNMX per value);
//                               sweep 2 lists every candidate with approximate dot >= m2 - 2 eps
//                               (chunk-max filter, so the divergent append path is rare).
//                               Any candidate left out has exact dot < m2 - eps <= the exact dots
//                               of at least two listed ones, so the exact best and second best
//                               are always in the list.
//  3. k_rescore      warp per query: exact fp32 scores of the listed candidates in the
//                    reference's rotated k order (bit-identical to ComputeDistance), then the
//                    reference's best / second-best rule incl. its tie-breaking
//                    (FindMinCorr/FindMaxCorr).  A query whose list overflowed (> 8 entries per
//                    split: massive near-ties) is flagged and redone by the exact fp32 kernel
//                    (kernels_match.cu).
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
constexpr int TC_CT = 256;        // candidates per tile (UMMA N)
constexpr int TC_STAGES = 2;
constexpr int TC_TOPK = 8;
constexpr int KHALF_BYTES_PER_ROW = 128;          // 64 fp16
constexpr uint32_t A_HALF_BYTES = 128 * KHALF_BYTES_PER_ROW;      // one 128-row x 64-K operand block: 16 KB
constexpr uint32_t B_HALF_BYTES = TC_CT * KHALF_BYTES_PER_ROW;    // 32 KB
constexpr uint32_t SMEM_A = 2 /*query halves*/ * 2 /*K halves*/ * A_HALF_BYTES;   // 64 KB
constexpr uint32_t SMEM_B_STAGE = 2 * B_HALF_BYTES;                                 // 64 KB
constexpr uint32_t SMEM_TC = SMEM_A + TC_STAGES * SMEM_B_STAGE + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 32 * (2 + 8);
constexpr int RS_ROWS = 16;        // staged candidate rows per rescoring round (per warp)
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
// Instruction descriptor, kind::f16: fp16 x fp16 -> fp32, both operands K-major, M=128, N=256.
constexpr uint32_t IDESC_F16_M128_N256 = (1u << 4)                 // D format = F32
                                         | (0u << 7) | (0u << 10)   // A, B format = F16
                                         | (0u << 15) | (0u << 16)  // A, B K-major
                                         | ((uint32_t)(TC_CT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// ---- 1. pack ----------------------------------------------------------------------
// Packed layout (per set): two K-halves; half kb is a [n_pad][64] fp16 matrix (128-byte rows) in
// which the 16-byte chunk c of row r is stored at chunk position c ^ (r & 7) (128B swizzle).
// Also validates the domain of the error bound and measures its ingredients (see the header): info[0] is set when a
// row has a non-finite / fp16-overflowing element or a squared norm above TC_MAX_NORM2; info[1] receives (as float
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
                                                            const int *__restrict__ q_info, const int *__restrict__ c_info) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the swizzle atoms
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char *sA = base;                          // [qhalf][khalf][128 rows][128 B]
  unsigned char *sB = base + SMEM_A;                 // [stage][khalf][256 rows][128 B]
  uint64_t *bars = reinterpret_cast<uint64_t *>(base + SMEM_A + TC_STAGES * SMEM_B_STAGE);
  uint64_t *a_full = bars + 0;
  uint64_t *b_full = bars + 1;                       // [TC_STAGES]
  uint64_t *b_empty = bars + 1 + TC_STAGES;          // [TC_STAGES]
  uint64_t *acc_full = bars + 1 + 2 * TC_STAGES;     // [2]
  uint64_t *acc_empty = bars + 3 + 2 * TC_STAGES;    // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 5 + 2 * TC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qtile = blockIdx.x, split = blockIdx.y;
  const int n_tiles_total = nc_pad / TC_CT;
  const int t0 = split * tiles_per_split;
  const int t1 = min(t0 + tiles_per_split, n_tiles_total);
  const int n_tiles = max(t1 - t0, 0);

  if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < TC_STAGES; s++) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(acc_full + a, 1);
      mbar_init(acc_empty + a, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM: all 512 columns (two 128 x 256 fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

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
      for (int i = 0; i < 2 * n_tiles; i++) {      // two sweeps over the candidate slice
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        const int tile = t0 + (i < n_tiles ? i : i - n_tiles);
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
      for (int i = 0; i < 2 * n_tiles; i++) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(b_full + s, ph);
        tc_fence_after();
        for (int qh = 0; qh < 2; qh++) {
          // accumulator qh is reused every tile: wait until the epilogue drained the previous one
          mbar_wait(acc_empty + qh, (i & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(qh * TC_CT);
#pragma unroll
          for (int kb = 0; kb < 2; kb++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint64_t ad = smem_desc_k128(smem_u32(sA + (qh * 2 + kb) * A_HALF_BYTES) + k * 32);
              const uint64_t bd = smem_desc_k128(smem_u32(sB + s * SMEM_B_STAGE + kb * B_HALF_BYTES) + k * 32);
              tc_mma_f16(d_tmem, ad, bd, IDESC_F16_M128_N256, (kb | k) ? 1u : 0u);
            }
          }
          tc_commit(acc_full + qh);       // accumulator qh complete -> epilogue
        }
        tc_commit(b_empty + s);           // both MMAs of this stage done -> producer may refill
      }
    }
  } else {
    // ===== epilogue: 8 warps, thread = one query row =====
    const int ew = warp - 2;                 // 0..7
    const int qh = ew >> 2;                  // query half (accumulator)
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int row = qh * 128 + quad * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(qh * TC_CT);
    // ---- sweep 1: m1 = largest approximate dot product of this row, m2 = second largest CHUNK maximum
    // (chunks of 32 columns).  m2 <= the true second largest value, so thr = m2 - 2 eps is still a valid
    // (slightly more inclusive) listing threshold, and a chunk costs 16 three-input max + 3 ops instead
    // of 96 (FMNMX runs at half rate: this epilogue is ALU-pipe bound).  Padding columns hold 0.
    // TMEM reads are software-pipelined: the tcgen05.ld of the next 64 columns is in flight while the
    // current 64 are reduced (tcgen05.wait::ld waits for every outstanding load, so it sits after the
    // arithmetic), and the accumulator is handed back to the MMA warp as soon as its last columns are in
    // registers, before they are reduced.
    float m1 = -1.0f, m2 = -1.0f;
    auto reduce64 = [&](const uint32_t (&r)[32], const uint32_t (&q)[32]) {
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
      const float ca = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
      const float cb = fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7]));
      float lo = fminf(m1, ca);
      m1 = fmaxf(m1, ca);
      m2 = fmaxf(m2, lo);
      lo = fminf(m1, cb);
      m1 = fmaxf(m1, cb);
      m2 = fmaxf(m2, lo);
    };
    static_assert(TC_CT / 64 == 4, "sweep 1 is unrolled for four 64-column chunks per tile");
    for (int i = 0; i < n_tiles; i++) {
      mbar_wait(acc_full + qh, i & 1);
      tc_fence_after();
      uint32_t ra[32], qa[32], rb[32], qb[32];
      tc_ld32(taddr + 0, ra);
      tc_ld32(taddr + 32, qa);
      tc_ld_wait();
      tc_ld32(taddr + 64, rb);
      tc_ld32(taddr + 96, qb);
      reduce64(ra, qa);
      tc_ld_wait();
      tc_ld32(taddr + 128, ra);
      tc_ld32(taddr + 160, qa);
      reduce64(rb, qb);
      tc_ld_wait();
      tc_ld32(taddr + 192, rb);
      tc_ld32(taddr + 224, qb);
      reduce64(ra, qa);
      tc_ld_wait();
      tc_fence_before();
      mbar_arrive(acc_empty + qh);           // 128 arrivals free the accumulator
      reduce64(rb, qb);
    }
    // ---- sweep 2: list every candidate with approximate dot >= m2 - 2 eps ----
    // eps of this pair of sets from the rounding-error norms measured at pack time (header comment)
    const float eps = 1.002f * (sqrtf(__int_as_float(q_info[1])) + sqrtf(__int_as_float(c_info[1]))) + 2.0e-5f;
    const float thr = m2 - 2.0f * eps;
    // Hits (about 2 per query and split) are appended straight to the global short list.  The append path
    // runs for the whole warp whenever ANY lane has a hit in an 8-column group (about a quarter of the
    // groups), so it must be short: a bit mask of the group's hits, then one iteration per set bit.
    int *const out_list = out_idx + ((size_t)(qtile * TC_QT + row) * n_splits + split) * TC_TOPK;
    int cnt = 0;
    auto list32 = [&](const uint32_t (&r)[32], int cbase) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float g = fmaxf(fmaxf(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1])), __uint_as_float(r[8 * k + 2]));
        g = fmaxf(fmaxf(g, __uint_as_float(r[8 * k + 3])), __uint_as_float(r[8 * k + 4]));
        g = fmaxf(fmaxf(g, __uint_as_float(r[8 * k + 5])), __uint_as_float(r[8 * k + 6]));
        g = fmaxf(g, __uint_as_float(r[8 * k + 7]));
        if (g >= thr) {
          unsigned int m = 0;
#pragma unroll
          for (int j = 0; j < 8; j++) m |= (__uint_as_float(r[8 * k + j]) >= thr) ? (1u << j) : 0u;
          while (m) {
            const int cidx = cbase + 8 * k + (__ffs(m) - 1);
            m &= m - 1;
            if (cidx < nc) {
              if (cnt < TC_TOPK) out_list[cnt] = cidx;
              cnt++;
            }
          }
        }
      }
    };
    for (int i = 0; i < n_tiles; i++) {
      mbar_wait(acc_full + qh, (n_tiles + i) & 1);
      tc_fence_after();
      const int col0 = (t0 + i) * TC_CT;
      uint32_t ra[32], rb[32];
      tc_ld32(taddr, ra);
      tc_ld_wait();
#pragma unroll 1
      for (int ch = 0; ch < TC_CT / 32; ch += 2) {        // same double buffering as sweep 1
        tc_ld32(taddr + (uint32_t)((ch + 1) * 32), rb);
        list32(ra, col0 + ch * 32);
        tc_ld_wait();
        if (ch + 2 < TC_CT / 32) {
          tc_ld32(taddr + (uint32_t)((ch + 2) * 32), ra);
        } else {
          tc_fence_before();
          mbar_arrive(acc_empty + qh);
        }
        list32(rb, col0 + (ch + 1) * 32);
        if (ch + 2 < TC_CT / 32) tc_ld_wait();
      }
    }
    for (int k = min(cnt, TC_TOPK); k < TC_TOPK; k++) out_list[k] = -1;
    out_val[(size_t)(qtile * TC_QT + row) * n_splits + split] = (float)cnt;   // entries found (may exceed TC_TOPK)
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

template <bool kL2>
__global__ void __launch_bounds__(128) k_rescore(csb_sift_point *__restrict__ s1, int n1,
                                                 const csb_sift_point *__restrict__ s2, int n2,
                                                 const float *__restrict__ sl_val, const int *__restrict__ sl_idx,
                                                 int n_splits, int *__restrict__ redo_flags) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n1) return;
  const int q = warp;
  const int n_list = n_splits * TC_TOPK;          // <= 32
  int ci = -1;
  if (lane < n_list) ci = sl_idx[(size_t)q * n_list + lane];
  // a split that found more candidates above its threshold than fit in its list: redo exactly
  bool overflow = false;
  if (lane < n_splits) overflow = sl_val[(size_t)q * n_splits + lane] > (float)TC_TOPK;
  const bool proven = !__any_sync(0xffffffffu, overflow);

  // exact score in the reference's rotated k order (matching.cu:84-89), lane = one listed candidate.
  // The listed descriptors are first staged in shared memory by the whole warp (coalesced 512-byte row
  // reads; 32 lanes each walking their own global row cost ~10 sectors per load instruction and made
  // this kernel as slow as the tensor-core scan), then every lane runs its 128-step FFMA chain from
  // shared memory: row stride 129 words spreads the lanes' rows over the banks.
  __shared__ float s_q[4][128];
  __shared__ float s_c[4][RS_ROWS * 129];
  const int wib = threadIdx.x >> 5;
  float *sq = s_q[wib], *sc = s_c[wib];
#pragma unroll
  for (int j = 0; j < 4; j++) sq[lane + 32 * j] = s1[q].data[lane + 32 * j];
  float score = kL2 ? 999.0f : -1.0f;
  // listed candidates are compacted to rows 0.. of the staging tile, RS_ROWS at a time (usually one round)
  const unsigned int have = __ballot_sync(0xffffffffu, ci >= 0);
  const int my_row = __popc(have & ((1u << lane) - 1u));
  const int n_have = __popc(have);
  for (int r0 = 0; r0 < n_have; r0 += RS_ROWS) {
    __syncwarp();
    unsigned int m = have;
    for (int r = 0; m; r++) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      if (r < r0 || r >= r0 + RS_ROWS) continue;
      const int c = __shfl_sync(0xffffffffu, ci, src);
      const float *pb = s2[c].data;
      // asynchronous global -> shared copies (LDGSTS): the rows of ALL listed candidates are in flight at once;
      // with ordinary loads every row waited for the previous row's data (load -> store dependency in a loop of
      // unknown length), i.e. one L2 round trip per candidate.  4-byte copies: data[] is only 4-byte aligned.
#pragma unroll
      for (int j = 0; j < 4; j++) __pipeline_memcpy_async(&sc[(r - r0) * 129 + lane + 32 * j], &pb[lane + 32 * j], 4);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncwarp();
    if (ci >= 0 && my_row >= r0 && my_row < r0 + RS_ROWS) {
      const float *pb = sc + (my_row - r0) * 129;
      const int tx = ci & 15;
      float sum = 0.0f;
      // k = (i + tx) & 127 for i = 0..127: the first 112 steps never wrap (tx <= 15), so they run off two
      // base pointers with immediate offsets; only the last 16 need the mask.  Same FFMA order.
      const float *qa = sq + tx, *pa = pb + tx;
#pragma unroll
      for (int i = 0; i < 112; i++) sum = __fmaf_rn(qa[i], pa[i], sum);
#pragma unroll
      for (int i = 112; i < 128; i++) {
        const int k = (i + tx) & 127;
        sum = __fmaf_rn(sq[k], pb[k], sum);
      }
      score = kL2 ? __fsub_rn(2.0f, __fadd_rn(sum, sum)) : sum;
    }
  }
  // best: FindMinCorr's winner = lowest (score, bitrev4(col % 16), col / 16); second = best of the rest
  // (an equal duplicate lands in `second`, matching.cu:229-235)
  auto better = [](float sa, int ia, float sb, int ib) {   // is a strictly preferred to b ?
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    if (sa != sb) return kL2 ? (sa < sb) : (sa > sb);
    const int ra = bitrev4(ia & 15), rb = bitrev4(ib & 15);
    if (ra != rb) return ra < rb;
    return ia < ib;
  };
  float bs = score;
  int bi = ci;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(os, oi, bs, bi)) {
      bs = os;
      bi = oi;
    }
  }
  float ss = (ci >= 0 && ci != bi) ? score : (kL2 ? 999.0f : -1.0f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, ss, o);
    ss = kL2 ? fminf(ss, os) : fmaxf(ss, os);
  }
  if (lane == 0) {
    if (!proven) {
      redo_flags[q >> 4] = 1;              // the exact kernel redoes this block of 16 queries
    } else {
      csb_sift_point *o = s1 + q;
      o->score = bs;
      if (kL2) o->ambiguity = (float)((double)bs / ((double)ss + 1e-6));
      else o->ambiguity = (float)((double)__fsub_rn(1.0f, bs) / ((double)__fsub_rn(1.0f, ss) + 1e-6));
      o->match = bi;
      if (bi >= 0) {
        o->match_xpos = s2[bi].coords2D[0];
        o->match_ypos = s2[bi].coords2D[1];
      }
    }
  }
}

// Compacts the flagged 16-query blocks into a list for the exact kernel.
__global__ void k_collect_redo(const int *__restrict__ flags, int n_blocks, int *__restrict__ list, int *__restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_blocks && flags[i]) list[atomicAdd(count, 1)] = i;
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

void launch_match_tc(const void *q_packed, int n1, const void *c_packed, int n2, int n_splits, float *sl_val, int *sl_idx,
                     const int *q_info, const int *c_info, cudaStream_t st) {
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
                                               sl_val, sl_idx, n_splits, q_info, c_info);
}

void launch_rescore(csb_sift_point *s1, int n1, const csb_sift_point *s2, int n2, const float *sl_val, const int *sl_idx,
                    int n_splits, int distance, int *redo_flags, int *redo_list, int *redo_count, cudaStream_t st) {
  const int n_blocks16 = (n1 + 15) / 16;
  cudaMemsetAsync(redo_flags, 0, sizeof(int) * n_blocks16, st);
  cudaMemsetAsync(redo_count, 0, sizeof(int), st);
  const int blocks = (n1 * 32 + 127) / 128;
  if (distance == 1) k_rescore<true><<<blocks, 128, 0, st>>>(s1, n1, s2, n2, sl_val, sl_idx, n_splits, redo_flags);
  else k_rescore<false><<<blocks, 128, 0, st>>>(s1, n1, s2, n2, sl_val, sl_idx, n_splits, redo_flags);
  k_collect_redo<<<(n_blocks16 + 255) / 256, 256, 0, st>>>(redo_flags, n_blocks16, redo_list, redo_count);
}
