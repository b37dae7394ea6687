// C ABI of cusift_b200 (include/cusift_b200.h): per-GPU context, frame slots,
// host-side orchestration of the extraction / matching / homography kernels.
//
// Host-side behaviour restated from the reference (danielsuo/cuSIFT):
//   SiftData::Extract / ExtractSiftLoop / ExtractSiftOctave  cuSIFT.cu:61-120,175-270
//   ScaleDown weights   cuSIFT.cu:320-338      LaplaceMulti weights  cuSIFT.cu:399-412
//   FindPointsMulti constants  cuSIFT.cu:424-444
// but with none of its per-frame cudaMalloc/cudaFree, cudaMemcpyToSymbol,
// cudaMemcpyFromSymbol round trips or texture-object churn: a slot owns a
// persistent pyramid workspace + texture objects, every parameter travels as a
// kernel argument, and a frame costs exactly one stream synchronisation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "csb_internal.h"

namespace {

inline int iAlignUp(int a, int b) { return (a % b != 0) ? (a - a % b + b) : a; }

struct Octave {
  int w = 0, h = 0, pitch = 0;
  float *base = nullptr;   // octave base image (octave 0: the caller's frame or slot->img0)
  float *dog = nullptr;    // 7 planes, layout CSB_DOG_PS / CSB_DOG_RS (csb_internal.h)
  cudaTextureObject_t tex = 0;
  CUtensorMap dog_map;     // 2-D TMA descriptor over dog (k_find_points)
  CUtensorMap src_map;     // 2-D TMA descriptor over base (k_pyramid); octave 0: valid for slot->img0 only
  bool src_map_ok = false;
};

struct TexCacheEntry {
  const float *ptr;
  int w, h, pitch;
  cudaTextureObject_t tex;
  CUtensorMap src_map;     // k_pyramid's descriptor over the same caller-owned frame
  bool src_map_ok;
};

struct ProfRec {
  int name_id;
  cudaEvent_t a, b;
};

struct Slot {
  cudaStream_t stream = nullptr;
  int w = 0, h = 0, n_oct = 0;
  float *arena = nullptr;
  size_t arena_bytes = 0;
  float *img0 = nullptr;   // upload target for host frames
  unsigned char *u8 = nullptr;            // device staging for 8-bit frames (csb_extract_batch_u8)
  size_t u8_cap = 0;
  Octave oct[CSB_MAX_OCTAVES];
  std::vector<TexCacheEntry> tex_cache;   // textures over caller-owned octave-0 frames
  unsigned int *d_counter = nullptr;
  KpStage *d_stage = nullptr;             // per-octave keypoint lists between k_find_points and k_orient_desc
  size_t stage_pts = 0;                   // capacity: octaves * max_pts entries
  int *h_count = nullptr;                 // pinned: keypoints found by the frame in flight
  csb_sift_point *h_stage = nullptr;      // pinned + mapped staging for pageable destinations
  size_t stage_cap = 0;
  csb_compact_point *d_compact = nullptr; // compact result records of the frame in flight (csb_extract_batch_compact)
  size_t compact_cap = 0;
  bool compact = false;
  // frame in flight: COMPUTING (kernels + count readback queued) -> COPYING (exact-size D2H queued) -> IDLE
  bool busy = false;
  bool copying = false;
  cudaEvent_t ev_count = nullptr;         // count has landed in h_count
  const csb_sift_point *cur_d_sift = nullptr;
  int cur_max_pts = 0;
  int cur_n = 0;
  void *user_h = nullptr;
  bool staged = false;
  int *user_num = nullptr;
  std::vector<ProfRec> prof_pending;
  std::vector<ProfRec> prof_free;
};

struct ProfEntry {
  std::string name;
  double total_ms = 0.0;
  long long launches = 0;
};

}  // namespace

struct csb_ctx {
  int device = 0;
  int sm_count = 148;
  int n_slots = 0;
  Slot *slots = nullptr;
  std::string err;
  bool profile = false;
  bool no_fuse = false;
  std::vector<ProfEntry> prof;
  long long launches = 0;
  // CSB_TRACE=1: host time spent queueing work vs waiting for the device, printed by csb_ctx_destroy
  bool trace = false;
  double host_enqueue_ms = 0.0, host_wait_ms = 0.0;
  long long frames = 0;
  // tensor-core matcher scratch
  bool match_exact = false;          // CSB_MATCH_EXACT=1: always use the fp32 CUDA-core kernel
  void *tc_pack[2] = {nullptr, nullptr};
  size_t tc_pack_cap[2] = {0, 0};
  float *tc_val = nullptr;
  int *tc_idx = nullptr, *tc_flags = nullptr, *tc_list = nullptr, *tc_count = nullptr;
  void *tc_part = nullptr;
  size_t tc_sl_cap = 0, tc_blk_cap = 0;
  int *h_tc_count = nullptr;
  long long tc_redo_blocks = 0;      // 16-query blocks redone exactly (diagnostics)
  long long tc_domain_fallbacks = 0; // calls / pairs routed to the exact kernel because a set violated the fp16 precondition
  int *ap_flags = nullptr, *h_ap_flags = nullptr;   // per-set out-of-domain flags of the all-pairs path
  size_t ap_flags_cap = 0;
  // homography scratch
  float *d_coord = nullptr, *d_homo = nullptr;
  int *d_rand = nullptr, *d_counts = nullptr;
  size_t coord_cap = 0, loops_cap = 0;
  int *h_counts = nullptr;
  // rigid-transform scratch (grow-only): coords | indices | Rt | counts | mask
  char *rt_dev = nullptr;
  size_t rt_cap = 0;
  // ImproveHomography scratch: device {H_in[9], H_out[9], numfit, job}, pinned host mirror
  char *ih_dev = nullptr, *ih_host = nullptr;
  char *ap_jobs = nullptr;           // all-pairs: one ImproveJob per pair
  size_t ap_jobs_cap = 0;
  // all-pairs scratch (grow-only, csb_allpairs_match_ransac)
  static constexpr int AP_STREAMS = 8;
  cudaStream_t ap_stream[AP_STREAMS] = {};
  cudaEvent_t ap_done[AP_STREAMS] = {};
  cudaEvent_t ap_packed = nullptr;
  char *ap_scratch = nullptr;        // AP_STREAMS carve-ups of ap_scratch_stride bytes
  size_t ap_scratch_cap = 0;
  char *ap_pack = nullptr;           // fp16 operand tiles of all sets
  size_t ap_pack_cap = 0;
  char *ap_result = nullptr;         // per pair: H[9], inliers, n_valid
  size_t ap_result_cap = 0;
};

namespace {

#define CSB_CHECK(ctx, call)                                                                         \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      char buf_[512];                                                                                \
      snprintf(buf_, sizeof(buf_), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
               cudaGetErrorString(e_));                                                              \
      (ctx)->err = buf_;                                                                             \
      return (int)e_;                                                                                \
    }                                                                                                \
  } while (0)

inline double host_now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int fail(csb_ctx *ctx, int code, const char *msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int prof_id(csb_ctx *ctx, const char *name) {
  for (size_t i = 0; i < ctx->prof.size(); i++)
    if (ctx->prof[i].name == name) return (int)i;
  ProfEntry e;
  e.name = name;
  ctx->prof.push_back(e);
  return (int)ctx->prof.size() - 1;
}

// Brackets one kernel launch with events on the slot's stream when profiling.
struct LaunchScope {
  csb_ctx *ctx;
  Slot *s;
  ProfRec rec;
  bool on;
  LaunchScope(csb_ctx *c, Slot *sl, const char *name) : ctx(c), s(sl), on(c->profile) {
    ctx->launches++;
    if (!on) return;
    if (!s->prof_free.empty()) {
      rec = s->prof_free.back();
      s->prof_free.pop_back();
    } else {
      cudaEventCreate(&rec.a);
      cudaEventCreate(&rec.b);
    }
    rec.name_id = prof_id(ctx, name);
    cudaEventRecord(rec.a, s->stream);
  }
  ~LaunchScope() {
    if (!on) return;
    cudaEventRecord(rec.b, s->stream);
    s->prof_pending.push_back(rec);
  }
};

void prof_collect(csb_ctx *ctx, Slot *s) {
  for (ProfRec &r : s->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ctx->prof[r.name_id].total_ms += ms;
      ctx->prof[r.name_id].launches++;
    }
    s->prof_free.push_back(r);
  }
  s->prof_pending.clear();
}

int make_texture(csb_ctx *ctx, const float *ptr, int w, int h, int pitch, cudaTextureObject_t *out) {
  // same descriptor as the reference, cuSIFT.cu:218-236
  cudaResourceDesc res;
  memset(&res, 0, sizeof(res));
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = const_cast<float *>(ptr);
  res.res.pitch2D.width = w;
  res.res.pitch2D.height = h;
  res.res.pitch2D.pitchInBytes = (size_t)pitch * sizeof(float);
  res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  cudaTextureDesc td;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = cudaAddressModeClamp;
  td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  CSB_CHECK(ctx, cudaCreateTextureObject(out, &res, &td, nullptr));
  return 0;
}

void slot_release_workspace(Slot *s) {
  for (int o = 0; o < CSB_MAX_OCTAVES; o++) {
    if (s->oct[o].tex) cudaDestroyTextureObject(s->oct[o].tex);
    s->oct[o] = Octave();
  }
  for (TexCacheEntry &e : s->tex_cache) cudaDestroyTextureObject(e.tex);
  s->tex_cache.clear();
  if (s->arena) cudaFree(s->arena);
  s->arena = nullptr;
  s->arena_bytes = 0;
  s->img0 = nullptr;
  s->w = s->h = s->n_oct = 0;
}

// (Re)builds the pyramid workspace of a slot for frames of w x h with n_oct octaves.
int slot_prepare(csb_ctx *ctx, Slot *s, int w, int h, int n_oct) {
  if (s->arena && s->w == w && s->h == h && s->n_oct == n_oct) return 0;
  slot_release_workspace(s);
  auto align512 = [](size_t b) { return (b + 511) & ~(size_t)511; };
  size_t total = 0;
  int ww = w, hh = h;
  size_t off_img0 = 0, off_base[CSB_MAX_OCTAVES], off_dog[CSB_MAX_OCTAVES];
  for (int o = 0; o < n_oct; o++) {
    const int p = iAlignUp(ww, 128);
    const size_t img = align512((size_t)p * (hh + 1) * sizeof(float));
    if (o == 0) { off_img0 = total; total += img; off_base[0] = off_img0; }
    else { off_base[o] = total; total += img; }
    off_dog[o] = total;
    total += align512((size_t)p * hh * CSB_NUM_DOG * sizeof(float));
    s->oct[o].w = ww; s->oct[o].h = hh; s->oct[o].pitch = p;
    ww /= 2; hh /= 2;
  }
  CSB_CHECK(ctx, cudaMalloc((void **)&s->arena, total));
  CSB_CHECK(ctx, cudaMemsetAsync(s->arena, 0, total, s->stream));
  s->arena_bytes = total;
  char *base = reinterpret_cast<char *>(s->arena);
  s->img0 = reinterpret_cast<float *>(base + off_img0);
  for (int o = 0; o < n_oct; o++) {
    s->oct[o].base = reinterpret_cast<float *>(base + off_base[o]);
    s->oct[o].dog = reinterpret_cast<float *>(base + off_dog[o]);
    if (make_dog_tensor_map(&s->oct[o].dog_map, s->oct[o].dog, s->oct[o].h, s->oct[o].pitch))
      return fail(ctx, CSB_E_INVALID, "cuTensorMapEncodeTiled failed for a DoG buffer");
    if (o > 0) {
      int rc = make_texture(ctx, s->oct[o].base, s->oct[o].w, s->oct[o].h, s->oct[o].pitch, &s->oct[o].tex);
      if (rc) return rc;
    }
    // octave 0's entry describes slot->img0 (the upload target); caller-owned frames go through tex_cache
    if (pyramid_source_map(&s->oct[o].src_map, o == 0 ? s->img0 : s->oct[o].base, s->oct[o].w, s->oct[o].h, s->oct[o].pitch))
      return fail(ctx, CSB_E_INVALID, "cuTensorMapEncodeTiled failed for an octave base");
    s->oct[o].src_map_ok = true;
  }
  s->w = w; s->h = h; s->n_oct = n_oct;
  return 0;
}

// Texture object + TMA descriptor over a caller-owned octave-0 frame (small per-slot cache keyed by geometry).
int slot_tex0(csb_ctx *ctx, Slot *s, const float *ptr, int w, int h, int pitch, cudaTextureObject_t *out,
              const CUtensorMap **map_out) {
  for (TexCacheEntry &e : s->tex_cache)
    if (e.ptr == ptr && e.w == w && e.h == h && e.pitch == pitch) {
      *out = e.tex;
      *map_out = e.src_map_ok ? &e.src_map : nullptr;
      return 0;
    }
  if (s->tex_cache.size() >= 256) {   // bounded: drop the oldest
    cudaDestroyTextureObject(s->tex_cache.front().tex);
    s->tex_cache.erase(s->tex_cache.begin());
  }
  TexCacheEntry e;
  memset(&e, 0, sizeof(e));
  e.ptr = ptr; e.w = w; e.h = h; e.pitch = pitch;
  int rc = make_texture(ctx, ptr, w, h, pitch, &e.tex);
  if (rc) return rc;
  e.src_map_ok = pyramid_tma_ok(ptr, pitch) && pyramid_source_map(&e.src_map, ptr, w, h, pitch) == 0;
  s->tex_cache.push_back(e);
  *out = e.tex;
  *map_out = s->tex_cache.back().src_map_ok ? &s->tex_cache.back().src_map : nullptr;
  return 0;
}

// cv::getGaussianKernel(3, 0.5, CV_32F): exp in double, float taps, float taps scaled by 1/sum (double)
void preblur_kernel(float *k0, float *k1) {
  const double sigma = 0.5, scale2x = -0.5 / (sigma * sigma);
  const float c = (float)exp(scale2x * 0.0), sd = (float)exp(scale2x * 1.0);
  const double inv = 1.0 / ((double)sd + (double)c + (double)sd);
  *k0 = (float)((double)c * inv);
  *k1 = (float)((double)sd * inv);
}

// ---- host-side constants, restated verbatim from the reference -----------------
void scale_down_kernel(float k3[3], float variance = 0.5f) {     // cuSIFT.cu:320-338; ExtractSiftLoop passes 0.5f (cuSIFT.cu:185)
  float h_Kernel[5], kernelSum = 0.0f;
  for (int j = 0; j < 5; j++) {
    h_Kernel[j] = (float)expf(-(double)(j - 2) * (j - 2) / 2.0 / variance);
    kernelSum += h_Kernel[j];
  }
  for (int j = 0; j < 5; j++) h_Kernel[j] /= kernelSum;
  k3[0] = h_Kernel[0]; k3[1] = h_Kernel[1]; k3[2] = h_Kernel[2];
}

void laplace_weights(float initBlur, DogWeights *W) {   // cuSIFT.cu:239-240,399-412
  const float baseBlur = pow(2.0f, -1.0f / CSB_NUM_SCALES);
  const float diffScale = pow(2.0f, 1.0f / CSB_NUM_SCALES);
  float kernel[12 * 16];
  float scale = baseBlur;
  for (int i = 0; i < CSB_NUM_LEVELS; i++) {
    float kernelSum = 0.0f;
    float var = scale * scale - initBlur * initBlur;
    for (int j = -4; j <= 4; j++) {
      kernel[16 * i + j + 4] = (float)expf(-(double)j * j / 2.0 / var);
      kernelSum += kernel[16 * i + j + 4];
    }
    for (int j = -4; j <= 4; j++) kernel[16 * i + j + 4] /= kernelSum;
    scale *= diffScale;
  }
  for (int i = 0; i < CSB_NUM_LEVELS; i++)
    for (int j = 0; j < 5; j++) W->k[i][j] = kernel[16 * i + j];
}

void extrema_params(const csb_params *p, ExtremaParams *E) {   // cuSIFT.cu:239-247,424-444 (the same for every octave)
  const float baseBlur = pow(2.0f, -1.0f / CSB_NUM_SCALES);
  const float diffScaleL = pow(2.0f, 1.0f / CSB_NUM_SCALES);
  const double sigma = baseBlur * diffScaleL;
  const float factor = 1.0f / CSB_NUM_SCALES;
  float scale = (float)sigma;
  const float diffScale = pow(2.0f, factor);
  for (int i = 0; i < CSB_NUM_SCALES; i++) {
    E->scales[i] = scale;
    scale *= diffScale;
  }
  E->thresh = p->peak_thresh;
  E->edge_limit = p->edge_thresh;
  E->factor = factor;
  E->n_oct = 0;
}

// Enqueues one frame on a slot.  d_img0/pitch0: octave-0 image on the device.
int enqueue_frame(csb_ctx *ctx, Slot *s, const float *d_img0, int w, int h, int pitch0, const csb_params *p,
                  csb_sift_point *d_sift, int max_pts, void *h_sift, int *num_pts, bool compact = false) {
  const int n_oct = p->num_octaves;
  s->compact = compact && h_sift != nullptr;
  if (s->compact && s->compact_cap < (size_t)max_pts) {
    if (s->d_compact) cudaFree(s->d_compact);
    s->d_compact = nullptr; s->compact_cap = 0;
    CSB_CHECK(ctx, cudaMalloc((void **)&s->d_compact, sizeof(csb_compact_point) * (size_t)max_pts));
    s->compact_cap = max_pts;
  }
  cudaStream_t st = s->stream;
  const double t_enq = ctx->trace ? host_now_ms() : 0.0;

  if (s->stage_pts < (size_t)n_oct * max_pts) {
    if (s->d_stage) cudaFree(s->d_stage);
    s->d_stage = nullptr; s->stage_pts = 0;
    CSB_CHECK(ctx, cudaMalloc((void **)&s->d_stage, sizeof(KpStage) * (size_t)n_oct * max_pts));
    s->stage_pts = (size_t)n_oct * max_pts;
  }

  // result destination: directly into the caller's buffer when it is page-locked, else via pinned staging
  s->staged = false;
  if (h_sift) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, h_sift);
    if (!(e == cudaSuccess && attr.type == cudaMemoryTypeHost)) {
      cudaGetLastError();
      const size_t need = (size_t)max_pts * sizeof(csb_sift_point);   // (also covers the smaller compact records)
      if (s->stage_cap < need) {
        if (s->h_stage) cudaFreeHost(s->h_stage);
        s->h_stage = nullptr;
        CSB_CHECK(ctx, cudaHostAlloc((void **)&s->h_stage, need, cudaHostAllocDefault));
        s->stage_cap = need;
      }
      s->staged = true;
    }
  }
  s->user_h = h_sift;
  s->user_num = num_pts;

  // profiling: hold the stream while the frame's launches and their event pairs are being queued, so that the events
  // measure device execution back to back rather than the host's enqueue pace
  if (ctx->profile) launch_delay(150000ull, st);
  CSB_CHECK(ctx, cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned int) * (1 + CSB_MAX_OCTAVES), st));   // count + run ends

  // octave geometry, blur schedule (cuSIFT.cu:188) and per-octave constants.  The DoG stack always uses the
  // slot's pitch (iAlignUp(w,128)); the caller's pitch only describes the octave-0 source image.
  Octave oct[CSB_MAX_OCTAVES];
  for (int o = 0; o < n_oct; o++) oct[o] = s->oct[o];
  oct[0].base = const_cast<float *>(d_img0);
  const int src_pitch0 = pitch0;
  const CUtensorMap *src_map0 = nullptr;
  if (d_img0 == s->img0 && pitch0 == s->oct[0].pitch) {
    if (!s->oct[0].tex) {
      int rc = make_texture(ctx, s->img0, w, h, pitch0, &s->oct[0].tex);
      if (rc) return rc;
    }
    oct[0].tex = s->oct[0].tex;
    src_map0 = &s->oct[0].src_map;
  } else {
    int rc = slot_tex0(ctx, s, d_img0, w, h, pitch0, &oct[0].tex, &src_map0);
    if (rc) return rc;
  }

  double initBlur[CSB_MAX_OCTAVES];
  float subs[CSB_MAX_OCTAVES];
  bool active[CSB_MAX_OCTAVES];
  initBlur[0] = p->init_blur;
  subs[0] = p->subsampling;
  for (int o = 0; o < n_oct; o++) {
    if (o > 0) {
      initBlur[o] = (float)sqrt(initBlur[o - 1] * initBlur[o - 1] + 0.5f * 0.5f) / 2.0f;   // cuSIFT.cu:188
      subs[o] = subs[o - 1] * 2.0f;
    }
    active[o] = p->lowest_scale < subs[o] * 2.0f;                                          // cuSIFT.cu:194
  }
  float k3[3];
  scale_down_kernel(k3);

  if (src_map0 && !ctx->no_fuse) {
    // TMA pyramid: (A) octave 0 fused with the downsample that seeds octave 1, (B) the remaining octave bases in one
    // chain launch, (C) ALL coarser octaves in one launch.  Three launches instead of one per octave.
    if (active[0]) {
      PyramidParams A;
      PyramidMaps MA;
      A.n_oct = 1;
      A.dk = DownK{k3[0], k3[1], k3[2]};
      A.oct[0].dog = oct[0].dog; A.oct[0].w = oct[0].w; A.oct[0].h = oct[0].h; A.oct[0].dpitch = oct[0].pitch;
      A.oct[0].next = n_oct > 1 ? oct[1].base : nullptr;
      A.oct[0].npitch = n_oct > 1 ? oct[1].pitch : 0;
      DogWeights W;
      laplace_weights((float)initBlur[0], &W);
      pyramid_set_weights(&A, 0, W);
      MA.m[0] = *src_map0;
      const int n_ctas = plan_pyramid(&A, ctx->sm_count);
      LaunchScope ls(ctx, s, "pyramid_o0");
      launch_pyramid(A, MA, n_ctas, n_oct > 1, st);
    } else if (n_oct > 1) {
      LaunchScope ls(ctx, s, "scale_down");
      launch_scale_down(oct[0].base, oct[0].w, oct[0].h, src_pitch0, oct[1].base, oct[1].pitch, k3, st);
    }
    if (n_oct > 2) {
      float *dst[CSB_MAX_OCTAVES];
      int dw[CSB_MAX_OCTAVES], dh[CSB_MAX_OCTAVES], dp[CSB_MAX_OCTAVES];
      for (int o = 2; o < n_oct; o++) { dst[o - 2] = oct[o].base; dw[o - 2] = oct[o].w; dh[o - 2] = oct[o].h; dp[o - 2] = oct[o].pitch; }
      LaunchScope ls(ctx, s, "down_chain");
      ctx->launches += (n_oct - 2 + 2) / 3 - 1;
      launch_down_chain(oct[1].base, oct[1].w, oct[1].h, oct[1].pitch, dst, dw, dh, dp, n_oct - 2, k3, st);
    }
    PyramidParams B;
    PyramidMaps MB;
    B.n_oct = 0;
    B.dk = DownK{k3[0], k3[1], k3[2]};
    for (int o = 1; o < n_oct; o++) {
      if (!active[o]) continue;
      PyramidOctave &X = B.oct[B.n_oct];
      X.dog = oct[o].dog; X.w = oct[o].w; X.h = oct[o].h; X.dpitch = oct[o].pitch; X.next = nullptr; X.npitch = 0;
      DogWeights W;
      laplace_weights((float)initBlur[o], &W);
      pyramid_set_weights(&B, B.n_oct, W);
      MB.m[B.n_oct] = oct[o].src_map;
      B.n_oct++;
    }
    if (B.n_oct > 0) {
      const int n_ctas = plan_pyramid(&B, ctx->sm_count);
      LaunchScope ls(ctx, s, "pyramid_rest");
      launch_pyramid(B, MB, n_ctas, false, st);
    }
  } else {
    // scalar fallback (source not addressable by the TMA unit, or CSB_NO_FUSE=1): fine -> coarse, one launch per octave
    static const char *kNameFused[CSB_MAX_OCTAVES] = {"blur_dog_down_o0", "blur_dog_down_o1", "blur_dog_down_o2",
                                                      "blur_dog_down_o3", "blur_dog_down_o4", "blur_dog_down_o5",
                                                      "blur_dog_down_o6", "blur_dog_down_o7"};
    static const char *kNameBlur[CSB_MAX_OCTAVES] = {"blur_dog_o0", "blur_dog_o1", "blur_dog_o2", "blur_dog_o3",
                                                     "blur_dog_o4", "blur_dog_o5", "blur_dog_o6", "blur_dog_o7"};
    for (int o = 0; o < n_oct; o++) {
      const bool need_down = (o + 1 < n_oct);
      const int sp = o == 0 ? src_pitch0 : oct[o].pitch;
      DogWeights W;
      if (active[o]) laplace_weights((float)initBlur[o], &W);
      if (active[o] && need_down && !ctx->no_fuse) {
        LaunchScope ls(ctx, s, kNameFused[o]);
        launch_blur_dog_down(oct[o].base, oct[o].w, oct[o].h, sp, oct[o].dog, oct[o].pitch, W, oct[o + 1].base,
                             oct[o + 1].pitch, k3, st);
      } else {
        if (need_down) {
          LaunchScope ls(ctx, s, "scale_down");
          launch_scale_down(oct[o].base, oct[o].w, oct[o].h, sp, oct[o + 1].base, oct[o + 1].pitch, k3, st);
        }
        if (active[o]) {
          LaunchScope ls(ctx, s, kNameBlur[o]);
          launch_blur_dog(oct[o].base, oct[o].w, oct[o].h, sp, oct[o].dog, oct[o].pitch, W, st);
        }
      }
    }
  }
  // extrema of all active octaves in ONE launch (octave 0 first: it owns most CTAs); every octave fills
  // its own list, k_orient_desc lays them out coarse -> fine, the reference's order (cuSIFT.cu:181-196)
  {
    ExtremaParams E;
    ExtremaMaps M;
    extrema_params(p, &E);
    for (int o = 0; o < n_oct; o++) {
      if (!active[o]) continue;
      M.m[E.n_oct] = oct[o].dog_map;
      ExtremaOctave &X = E.oct[E.n_oct++];
      X.dog = oct[o].dog; X.w = oct[o].w; X.h = oct[o].h; X.pitch = oct[o].pitch; X.octave = o;
    }
    const int n_ctas = plan_find_points(&E, ctx->sm_count);
    LaunchScope ls(ctx, s, "find_points");
    launch_find_points(E, M, n_ctas, s->d_stage, s->d_counter, max_pts, st);
  }
  {
    OctaveTexSet T;
    for (int o = 0; o < CSB_MAX_OCTAVES; o++) T.tex[o] = (o < n_oct) ? oct[o].tex : 0;
    LaunchScope ls(ctx, s, "orient_desc");
    launch_orient_desc(T, n_oct, subs, s->d_stage, d_sift, s->d_counter, max_pts, p->rootsift, ctx->sm_count, st);
  }
  if (s->compact) {
    LaunchScope ls(ctx, s, "compact");
    launch_compact(d_sift, s->d_counter, max_pts, s->d_compact, st);
  }
  // The count comes back first; the SiftPoint array follows as ONE exact-size copy-engine transfer
  // once the host knows it (start_copy).  A kernel storing into mapped host memory would save that
  // round trip, but its PCIe-bound CTAs slow every kernel sharing their SMs: measured 5.4 k -> 8 k
  // frames/s at 4 slots when the copy moved to the DMA engine.
  CSB_CHECK(ctx, cudaMemcpyAsync(s->h_count, s->d_counter, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CSB_CHECK(ctx, cudaEventRecord(s->ev_count, st));
  CSB_CHECK(ctx, cudaGetLastError());
  s->cur_d_sift = d_sift;
  s->cur_max_pts = max_pts;
  s->busy = true;
  s->copying = false;
  if (ctx->trace) {
    ctx->host_enqueue_ms += host_now_ms() - t_enq;
    ctx->frames++;
  }
  return 0;
}

// COMPUTING -> COPYING: waits for the frame's kernels, then queues the download of exactly n keypoints.
int start_copy(csb_ctx *ctx, Slot *s) {
  if (!s->busy || s->copying) return 0;
  const double t_w = ctx->trace ? host_now_ms() : 0.0;
  CSB_CHECK(ctx, cudaEventSynchronize(s->ev_count));
  if (ctx->trace) ctx->host_wait_ms += host_now_ms() - t_w;
  const unsigned int found = (unsigned int)s->h_count[0];
  s->cur_n = (int)(found < (unsigned int)s->cur_max_pts ? found : (unsigned int)s->cur_max_pts);
  if (s->user_h && s->cur_n > 0) {
    LaunchScope ls(ctx, s, "copy_out");
    ctx->launches--;                       // a copy-engine transfer, not a kernel
    const void *from = s->compact ? (const void *)s->d_compact : (const void *)s->cur_d_sift;
    const size_t rec = s->compact ? sizeof(csb_compact_point) : sizeof(csb_sift_point);
    CSB_CHECK(ctx, cudaMemcpyAsync(s->staged ? (void *)s->h_stage : s->user_h, from, (size_t)s->cur_n * rec,
                                   cudaMemcpyDeviceToHost, s->stream));
  }
  s->copying = true;
  return 0;
}

int finalize_frame(csb_ctx *ctx, Slot *s) {
  if (!s->busy) return 0;
  int rc = start_copy(ctx, s);
  if (rc) return rc;
  const double t_w = ctx->trace ? host_now_ms() : 0.0;
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  if (ctx->trace) ctx->host_wait_ms += host_now_ms() - t_w;
  s->busy = false;
  s->copying = false;
  const int n = s->cur_n;
  if (s->staged && s->user_h && n > 0)
    memcpy(s->user_h, s->h_stage, (size_t)n * (s->compact ? sizeof(csb_compact_point) : sizeof(csb_sift_point)));
  if (s->user_num) *s->user_num = n;
  if (ctx->profile) prof_collect(ctx, s);
  return 0;
}

int check_params(csb_ctx *ctx, int w, int h, const csb_params *p, int max_pts) {
  if (!ctx) return CSB_E_INVALID;
  if (!p || w < 16 || h < 16 || max_pts < 1) return fail(ctx, CSB_E_INVALID, "csb_extract: bad frame size / params");
  if (p->num_octaves < 1 || p->num_octaves > CSB_MAX_OCTAVES) return fail(ctx, CSB_E_TOOMANY, "num_octaves out of range");
  if ((w >> (p->num_octaves - 1)) < 2 || (h >> (p->num_octaves - 1)) < 2)
    return fail(ctx, CSB_E_INVALID, "frame too small for num_octaves");
  return 0;
}

}  // namespace

// =============================================================================
extern "C" {

int csb_version(void) { return CSB_VERSION; }
int csb_sizeof_sift_point(void) { return (int)sizeof(csb_sift_point); }

int csb_ctx_create(int device, int num_slots, csb_ctx **out) {
  if (!out) return CSB_E_INVALID;
  *out = nullptr;
  if (num_slots <= 0) num_slots = 4;
  if (num_slots > 64) return CSB_E_TOOMANY;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return (int)e;
  if (ndev == 0) return (int)cudaErrorNoDevice;
  if (device < 0) device = 0;
  if (device > ndev - 1) device = ndev - 1;   // InitCuda clamps the same way, cutils.h:71-80
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  csb_ctx *ctx = new csb_ctx();
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  const char *nf = getenv("CSB_NO_FUSE");
  ctx->no_fuse = nf && nf[0] == '1';
  const char *me = getenv("CSB_MATCH_EXACT");
  ctx->match_exact = me && me[0] == '1';
  ctx->trace = getenv("CSB_TRACE") != nullptr;
  ctx->n_slots = num_slots;
  ctx->slots = new Slot[num_slots];
  for (int i = 0; i < num_slots; i++) {
    Slot *s = &ctx->slots[i];
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess) goto bad;
    if ((e = cudaMalloc((void **)&s->d_counter, 256)) != cudaSuccess) goto bad;
    if ((e = cudaHostAlloc((void **)&s->h_count, 256, cudaHostAllocDefault)) != cudaSuccess) goto bad;
    if ((e = cudaEventCreateWithFlags(&s->ev_count, cudaEventDisableTiming)) != cudaSuccess) goto bad;
    s->h_count[0] = s->h_count[1] = 0;
  }
  *out = ctx;
  return 0;
bad:
  csb_ctx_destroy(ctx);
  return (int)e;
}

void csb_ctx_destroy(csb_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  csb_dist_release(ctx);
  if (ctx->trace && ctx->frames)
    fprintf(stderr, "[csb] %lld frames on %d slots: host queueing %.1f us/frame, host waiting %.1f us/frame\n", ctx->frames,
            ctx->n_slots, ctx->host_enqueue_ms * 1e3 / ctx->frames, ctx->host_wait_ms * 1e3 / ctx->frames);
  for (int i = 0; i < ctx->n_slots; i++) {
    Slot *s = &ctx->slots[i];
    if (s->stream) cudaStreamSynchronize(s->stream);
    slot_release_workspace(s);
    for (ProfRec &r : s->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (ProfRec &r : s->prof_free) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    if (s->d_counter) cudaFree(s->d_counter);
    if (s->d_stage) cudaFree(s->d_stage);
    if (s->d_compact) cudaFree(s->d_compact);
    if (s->u8) cudaFree(s->u8);
    if (s->h_count) cudaFreeHost(s->h_count);
    if (s->ev_count) cudaEventDestroy(s->ev_count);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    if (s->stream) cudaStreamDestroy(s->stream);
  }
  delete[] ctx->slots;
  for (int i = 0; i < 2; i++)
    if (ctx->tc_pack[i]) cudaFree(ctx->tc_pack[i]);
  if (ctx->tc_val) cudaFree(ctx->tc_val);
  if (ctx->tc_idx) cudaFree(ctx->tc_idx);
  if (ctx->tc_flags) cudaFree(ctx->tc_flags);
  if (ctx->tc_list) cudaFree(ctx->tc_list);
  if (ctx->tc_part) cudaFree(ctx->tc_part);
  for (int t = 0; t < csb_ctx::AP_STREAMS; t++) {
    if (ctx->ap_stream[t]) cudaStreamDestroy(ctx->ap_stream[t]);
    if (ctx->ap_done[t]) cudaEventDestroy(ctx->ap_done[t]);
  }
  if (ctx->ap_packed) cudaEventDestroy(ctx->ap_packed);
  if (ctx->ap_scratch) cudaFree(ctx->ap_scratch);
  if (ctx->ap_pack) cudaFree(ctx->ap_pack);
  if (ctx->ap_result) cudaFree(ctx->ap_result);
  if (ctx->rt_dev) cudaFree(ctx->rt_dev);
  if (ctx->ih_dev) cudaFree(ctx->ih_dev);
  if (ctx->ih_host) cudaFreeHost(ctx->ih_host);
  if (ctx->ap_jobs) cudaFree(ctx->ap_jobs);
  if (ctx->ap_flags) cudaFree(ctx->ap_flags);
  if (ctx->h_ap_flags) cudaFreeHost(ctx->h_ap_flags);
  if (ctx->tc_count) cudaFree(ctx->tc_count);
  if (ctx->h_tc_count) cudaFreeHost(ctx->h_tc_count);
  if (ctx->d_coord) cudaFree(ctx->d_coord);
  if (ctx->d_homo) cudaFree(ctx->d_homo);
  if (ctx->d_rand) cudaFree(ctx->d_rand);
  if (ctx->d_counts) cudaFree(ctx->d_counts);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  delete ctx;
}

int csb_ctx_device(const csb_ctx *ctx) { return ctx ? ctx->device : -1; }
int csb_ctx_num_slots(const csb_ctx *ctx) { return ctx ? ctx->n_slots : 0; }
const char *csb_last_error(const csb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int csb_host_alloc(void **ptr, unsigned long long bytes) {
  if (!ptr) return CSB_E_INVALID;
  return (int)cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocMapped | cudaHostAllocPortable);
}
int csb_host_free(void *ptr) { return ptr ? (int)cudaFreeHost(ptr) : 0; }

int csb_device_alloc(csb_ctx *ctx, void **d_ptr, unsigned long long bytes) {
  if (!ctx || !d_ptr) return CSB_E_INVALID;
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  CSB_CHECK(ctx, cudaMalloc(d_ptr, (size_t)bytes));
  return 0;
}
int csb_forget_image(csb_ctx *ctx, const void *d_ptr) {
  if (!ctx) return CSB_E_INVALID;
  if (!d_ptr) return 0;
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  for (int i = 0; i < ctx->n_slots; i++) {   // drop cached textures / TMA descriptors over this buffer
    Slot *s = &ctx->slots[i];
    for (size_t k = 0; k < s->tex_cache.size();) {
      if (s->tex_cache[k].ptr == d_ptr) {
        cudaStreamSynchronize(s->stream);
        cudaDestroyTextureObject(s->tex_cache[k].tex);
        s->tex_cache.erase(s->tex_cache.begin() + k);
      } else {
        k++;
      }
    }
  }
  return 0;
}
int csb_device_free(csb_ctx *ctx, void *d_ptr) {
  if (!ctx) return CSB_E_INVALID;
  if (!d_ptr) return 0;
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  for (int i = 0; i < ctx->n_slots; i++) {   // drop cached textures over this buffer
    Slot *s = &ctx->slots[i];
    for (size_t k = 0; k < s->tex_cache.size();) {
      if (s->tex_cache[k].ptr == d_ptr) {
        cudaStreamSynchronize(s->stream);
        cudaDestroyTextureObject(s->tex_cache[k].tex);
        s->tex_cache.erase(s->tex_cache.begin() + k);
      } else {
        k++;
      }
    }
  }
  CSB_CHECK(ctx, cudaFree(d_ptr));
  return 0;
}
int csb_memcpy_h2d(csb_ctx *ctx, void *d_dst, const void *h_src, unsigned long long bytes) {
  if (!ctx) return CSB_E_INVALID;
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  CSB_CHECK(ctx, cudaMemcpy(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice));
  return 0;
}
int csb_memcpy_d2h(csb_ctx *ctx, void *h_dst, const void *d_src, unsigned long long bytes) {
  if (!ctx) return CSB_E_INVALID;
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  CSB_CHECK(ctx, cudaMemcpy(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
  return 0;
}
int csb_upload_image(csb_ctx *ctx, float *d_img, int pitch_floats, const float *h_img, int w, int h) {
  if (!ctx || !d_img || !h_img || pitch_floats < w) return fail(ctx, CSB_E_INVALID, "csb_upload_image: bad argument");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  CSB_CHECK(ctx, cudaMemcpy2D(d_img, sizeof(float) * pitch_floats, h_img, sizeof(float) * w, sizeof(float) * w, h,
                              cudaMemcpyHostToDevice));
  return 0;
}
int csb_download_image(csb_ctx *ctx, float *h_img, const float *d_img, int pitch_floats, int w, int h) {
  if (!ctx || !d_img || !h_img || pitch_floats < w) return fail(ctx, CSB_E_INVALID, "csb_download_image: bad argument");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  CSB_CHECK(ctx, cudaMemcpy2D(h_img, sizeof(float) * w, d_img, sizeof(float) * pitch_floats, sizeof(float) * w, h,
                              cudaMemcpyDeviceToHost));
  return 0;
}

int csb_extract(csb_ctx *ctx, const float *d_img, int w, int h, int pitch_floats, const csb_params *p, void *d_sift,
                int max_pts, void *h_sift, int *num_pts) {
  int rc = check_params(ctx, w, h, p, max_pts);
  if (rc) return rc;
  if (!d_img || !d_sift || pitch_floats < w) return fail(ctx, CSB_E_INVALID, "csb_extract: null image/output or pitch < width");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  if ((rc = finalize_frame(ctx, s))) return rc;
  if ((rc = slot_prepare(ctx, s, w, h, p->num_octaves))) return rc;
  if ((rc = enqueue_frame(ctx, s, d_img, w, h, pitch_floats, p, (csb_sift_point *)d_sift, max_pts, h_sift, num_pts)))
    return rc;
  return finalize_frame(ctx, s);
}

int csb_extract_host(csb_ctx *ctx, const float *h_img, int w, int h, const csb_params *p, void *d_sift, int max_pts,
                     void *h_sift, int *num_pts) {
  int rc = check_params(ctx, w, h, p, max_pts);
  if (rc) return rc;
  if (!h_img || !d_sift) return fail(ctx, CSB_E_INVALID, "csb_extract_host: null image/output");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  if ((rc = finalize_frame(ctx, s))) return rc;
  if ((rc = slot_prepare(ctx, s, w, h, p->num_octaves))) return rc;
  const int pitch = s->oct[0].pitch;
  CSB_CHECK(ctx, cudaMemcpy2DAsync(s->img0, sizeof(float) * pitch, h_img, sizeof(float) * w, sizeof(float) * w, h,
                                   cudaMemcpyHostToDevice, s->stream));
  if ((rc = enqueue_frame(ctx, s, s->img0, w, h, pitch, p, (csb_sift_point *)d_sift, max_pts, h_sift, num_pts)))
    return rc;
  return finalize_frame(ctx, s);
}

// A frame whose output buffers are still owned by a frame in flight on ANOTHER slot must wait for it
// (a caller may reuse one d_sift / h_sift for every frame of a batch).
static int wait_for_aliases(csb_ctx *ctx, const Slot *self, const void *d_sift, const void *h_sift) {
  for (int i = 0; i < ctx->n_slots; i++) {
    Slot *t = &ctx->slots[i];
    if (t == self || !t->busy) continue;
    if (t->cur_d_sift == d_sift || (h_sift != nullptr && t->user_h == h_sift)) {
      int rc = finalize_frame(ctx, t);
      if (rc) return rc;
    }
  }
  return 0;
}

int csb_extract_batch(csb_ctx *ctx, int n_frames, const float *const *imgs, int imgs_on_host, int w, int h,
                      int pitch_floats, const csb_params *p, void *const *d_sifts, void *const *h_sifts, int max_pts,
                      int *num_pts) {
  int rc = check_params(ctx, w, h, p, max_pts);
  if (rc) return rc;
  if (n_frames < 0 || !imgs || !d_sifts || !num_pts) return fail(ctx, CSB_E_INVALID, "csb_extract_batch: null argument");
  if (!imgs_on_host && pitch_floats < w) return fail(ctx, CSB_E_INVALID, "csb_extract_batch: pitch < width");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  for (int f = 0; f < n_frames; f++) {
    Slot *s = &ctx->slots[f % ctx->n_slots];
    if ((rc = finalize_frame(ctx, s))) return rc;
    if ((rc = slot_prepare(ctx, s, w, h, p->num_octaves))) return rc;
    const float *d_img = imgs[f];
    int pitch = pitch_floats;
    if (imgs_on_host) {
      pitch = s->oct[0].pitch;
      CSB_CHECK(ctx, cudaMemcpy2DAsync(s->img0, sizeof(float) * pitch, imgs[f], sizeof(float) * w, sizeof(float) * w,
                                       h, cudaMemcpyHostToDevice, s->stream));
      d_img = s->img0;
    }
    void *hs = h_sifts ? h_sifts[f] : nullptr;
    if ((rc = wait_for_aliases(ctx, s, d_sifts[f], hs))) return rc;
    if ((rc = enqueue_frame(ctx, s, d_img, w, h, pitch, p, (csb_sift_point *)d_sifts[f], max_pts, hs, &num_pts[f])))
      return rc;
    // two-stage pipeline: the frame queued n_slots/2 iterations ago moves on to its download while the
    // younger frames keep the SMs busy
    const int lag = ctx->n_slots / 2;
    if (f >= lag && ctx->slots[(f - lag) % ctx->n_slots].user_h && (rc = start_copy(ctx, &ctx->slots[(f - lag) % ctx->n_slots])))
      return rc;
  }
  for (int i = 0; i < ctx->n_slots; i++)
    if ((rc = finalize_frame(ctx, &ctx->slots[i]))) return rc;
  return 0;
}

int csb_extract_batch_compact(csb_ctx *ctx, int n_frames, const void *const *imgs, int imgs_on_host, int w, int h,
                              int pitch_floats, const csb_params *p, void *const *d_sifts, void *const *h_compact,
                              int max_pts, int *num_pts) {
  int rc = check_params(ctx, w, h, p, max_pts);
  if (rc) return rc;
  if (n_frames < 0 || !imgs || !d_sifts || !num_pts || imgs_on_host < 0 || imgs_on_host > 2)
    return fail(ctx, CSB_E_INVALID, "csb_extract_batch_compact: bad argument");
  if (imgs_on_host != 1 && pitch_floats < w) return fail(ctx, CSB_E_INVALID, "csb_extract_batch_compact: pitch < width");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  for (int f = 0; f < n_frames; f++) {
    Slot *s = &ctx->slots[f % ctx->n_slots];
    if ((rc = finalize_frame(ctx, s))) return rc;
    if ((rc = slot_prepare(ctx, s, w, h, p->num_octaves))) return rc;
    const float *d_img = (const float *)imgs[f];
    int pitch = pitch_floats;
    if (imgs_on_host == 1) {
      pitch = s->oct[0].pitch;
      CSB_CHECK(ctx, cudaMemcpy2DAsync(s->img0, sizeof(float) * pitch, imgs[f], sizeof(float) * w, sizeof(float) * w, h,
                                       cudaMemcpyHostToDevice, s->stream));
      d_img = s->img0;
    } else if (imgs_on_host == 2) {
      const size_t need = (size_t)w * h;
      if (s->u8_cap < need) {
        if (s->u8) cudaFree(s->u8);
        s->u8 = nullptr; s->u8_cap = 0;
        CSB_CHECK(ctx, cudaMalloc((void **)&s->u8, need));
        s->u8_cap = need;
      }
      pitch = s->oct[0].pitch;
      CSB_CHECK(ctx, cudaMemcpy2DAsync(s->u8, w, imgs[f], pitch_floats, w, h, cudaMemcpyHostToDevice, s->stream));
      {
        LaunchScope ls(ctx, s, "ingest_u8");
        launch_ingest_u8(s->u8, w, w, h, s->img0, pitch, 0, 0.f, 0.f, s->stream);
      }
      d_img = s->img0;
    }
    void *hs = h_compact ? h_compact[f] : nullptr;
    if ((rc = wait_for_aliases(ctx, s, d_sifts[f], hs))) return rc;
    if ((rc = enqueue_frame(ctx, s, d_img, w, h, pitch, p, (csb_sift_point *)d_sifts[f], max_pts, hs, &num_pts[f], true))) return rc;
    const int lag = ctx->n_slots / 2;
    if (f >= lag && ctx->slots[(f - lag) % ctx->n_slots].user_h && (rc = start_copy(ctx, &ctx->slots[(f - lag) % ctx->n_slots])))
      return rc;
  }
  for (int i = 0; i < ctx->n_slots; i++)
    if ((rc = finalize_frame(ctx, &ctx->slots[i]))) return rc;
  return 0;
}

int csb_rigid_transform(csb_ctx *ctx, const float *h_coord, int num_pts, int type, const int *h_indices, int num_loops,
                        float thresh2, unsigned int seed, float *Rt12, int *num_inliers, char *h_inliers) {
  if (!ctx || !h_coord || !Rt12 || !num_inliers || num_loops <= 0 || (type != 0 && type != 1))
    return fail(ctx, CSB_E_INVALID, "csb_rigid_transform: bad argument");
  *num_inliers = 0;
  for (int i = 0; i < 12; i++) Rt12[i] = (i % 5 == 0) ? 1.0f : 0.0f;   // identity [I | 0]
  if (num_pts < 3) return 0;
  if (h_indices)
    for (int i = 0; i < 3 * num_loops; i++)
      if (h_indices[i] < 0 || h_indices[i] >= num_pts) return fail(ctx, CSB_E_INVALID, "csb_rigid_transform: index out of range");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  cudaStream_t st = s->stream;
  int rc = finalize_frame(ctx, s);
  if (rc) return rc;
  // one scratch allocation: coords | indices | Rt | counts | mask
  const size_t b_coord = ((size_t)num_pts * 6 * 4 + 255) & ~(size_t)255, b_idx = ((size_t)num_loops * 12 + 255) & ~(size_t)255,
               b_rt = ((size_t)num_loops * 48 + 255) & ~(size_t)255, b_cnt = ((size_t)num_loops * 4 + 255) & ~(size_t)255,
               b_mask = ((size_t)num_pts + 255) & ~(size_t)255;
  // the scratch lives in the context and only grows (no cudaMalloc / cudaFree per call)
  const size_t need = b_coord + b_idx + b_rt + b_cnt + b_mask;
  if (ctx->rt_cap < need) {
    if (ctx->rt_dev) cudaFree(ctx->rt_dev);
    ctx->rt_dev = nullptr; ctx->rt_cap = 0;
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->rt_dev, need));
    ctx->rt_cap = need;
  }
  char *d = ctx->rt_dev;
  float *d_coord = (float *)d;
  int *d_idx = (int *)(d + b_coord);
  float *d_rt = (float *)(d + b_coord + b_idx);
  int *d_cnt = (int *)(d + b_coord + b_idx + b_rt);
  char *d_mask = d + b_coord + b_idx + b_rt + b_cnt;
  std::vector<int> counts(num_loops);
  std::vector<char> mask(num_pts);
  cudaError_t e = cudaMemcpyAsync(d_coord, h_coord, (size_t)num_pts * 24, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && h_indices) e = cudaMemcpyAsync(d_idx, h_indices, (size_t)num_loops * 12, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    ctx->launches += 2;
    launch_rigid_hypotheses(d_coord, num_pts, d_idx, h_indices ? 0 : 1, seed, type, num_loops, thresh2, d_rt, d_cnt, st);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(counts.data(), d_cnt, (size_t)num_loops * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  int best = -1, best_cnt = -1;
  if (e == cudaSuccess) {
    for (int i = 0; i < num_loops; i++)          // rigidTransform.cu:451-457: `>=`, the last maximum wins
      if (counts[i] >= best_cnt) { best_cnt = counts[i]; best = i; }
    ctx->launches += 1;
    launch_rigid_mask(d_coord, num_pts, d_rt, best, thresh2, d_mask, st);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(mask.data(), d_mask, (size_t)num_pts, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(Rt12, d_rt + 12 * (size_t)best, 48, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    ctx->err = std::string("csb_rigid_transform: ") + cudaGetErrorString(e);
    return (int)e;
  }
  *num_inliers = best_cnt;
  if (h_inliers) memcpy(h_inliers, mask.data(), (size_t)num_pts);
  if (type == 1 && best_cnt >= 3) {              // rigidTransform.cu:476-484: refit on every inlier of the winner
    std::vector<int> idx;
    for (int i = 0; i < num_pts; i++)
      if (mask[i] == 1) idx.push_back(i);
    rigid_refit_host(h_coord, idx.data(), (int)idx.size(), Rt12);
  }
  return 0;
}

unsigned int csb_rigid_sample_hash(unsigned int seed, unsigned int loop, unsigned int k, unsigned int attempt) {
  return rigid_hash_host(seed, loop, k, attempt);
}

int csb_ingest_u8(csb_ctx *ctx, const unsigned char *src, int src_on_host, int w, int h, int stride, int preblur,
                  float *d_dst, int dst_pitch_floats) {
  if (!ctx || !src || !d_dst || w < 1 || h < 1 || stride < w || dst_pitch_floats < w)
    return fail(ctx, CSB_E_INVALID, "csb_ingest_u8: bad argument");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  int rc = finalize_frame(ctx, s);
  if (rc) return rc;
  const unsigned char *d_src = src;
  int d_stride = stride;
  if (src_on_host) {
    const size_t need = (size_t)w * h;
    if (s->u8_cap < need) {
      if (s->u8) cudaFree(s->u8);
      s->u8 = nullptr; s->u8_cap = 0;
      CSB_CHECK(ctx, cudaMalloc((void **)&s->u8, need));
      s->u8_cap = need;
    }
    CSB_CHECK(ctx, cudaMemcpy2DAsync(s->u8, w, src, stride, w, h, cudaMemcpyHostToDevice, s->stream));
    d_src = s->u8;
    d_stride = w;
  }
  float k0, k1;
  preblur_kernel(&k0, &k1);
  {
    LaunchScope ls(ctx, s, "ingest_u8");
    launch_ingest_u8(d_src, d_stride, w, h, d_dst, dst_pitch_floats, preblur, k0, k1, s->stream);
  }
  CSB_CHECK(ctx, cudaGetLastError());
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  if (ctx->profile) prof_collect(ctx, s);
  return 0;
}

int csb_extract_batch_u8(csb_ctx *ctx, int n_frames, const unsigned char *const *h_imgs, int w, int h, int stride,
                         int preblur, const csb_params *p, void *const *d_sifts, void *const *h_sifts, int max_pts,
                         int *num_pts) {
  int rc = check_params(ctx, w, h, p, max_pts);
  if (rc) return rc;
  if (n_frames < 0 || !h_imgs || !d_sifts || !num_pts || stride < w)
    return fail(ctx, CSB_E_INVALID, "csb_extract_batch_u8: bad argument");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  float k0, k1;
  preblur_kernel(&k0, &k1);
  for (int f = 0; f < n_frames; f++) {
    Slot *s = &ctx->slots[f % ctx->n_slots];
    if ((rc = finalize_frame(ctx, s))) return rc;
    if ((rc = slot_prepare(ctx, s, w, h, p->num_octaves))) return rc;
    const size_t need = (size_t)w * h;
    if (s->u8_cap < need) {
      if (s->u8) cudaFree(s->u8);
      s->u8 = nullptr; s->u8_cap = 0;
      CSB_CHECK(ctx, cudaMalloc((void **)&s->u8, need));
      s->u8_cap = need;
    }
    const int pitch = s->oct[0].pitch;
    CSB_CHECK(ctx, cudaMemcpy2DAsync(s->u8, w, h_imgs[f], stride, w, h, cudaMemcpyHostToDevice, s->stream));
    {
      LaunchScope ls(ctx, s, "ingest_u8");
      launch_ingest_u8(s->u8, w, w, h, s->img0, pitch, preblur, k0, k1, s->stream);
    }
    void *hs = h_sifts ? h_sifts[f] : nullptr;
    if ((rc = wait_for_aliases(ctx, s, d_sifts[f], hs))) return rc;
    if ((rc = enqueue_frame(ctx, s, s->img0, w, h, pitch, p, (csb_sift_point *)d_sifts[f], max_pts, hs, &num_pts[f])))
      return rc;
    const int lag = ctx->n_slots / 2;
    if (f >= lag && ctx->slots[(f - lag) % ctx->n_slots].user_h && (rc = start_copy(ctx, &ctx->slots[(f - lag) % ctx->n_slots])))
      return rc;
  }
  for (int i = 0; i < ctx->n_slots; i++)
    if ((rc = finalize_frame(ctx, &ctx->slots[i]))) return rc;
  return 0;
}

int csb_scale_down(csb_ctx *ctx, const float *d_src, int w, int h, int src_pitch, float *d_dst, int dst_pitch) {
  return csb_scale_down_var(ctx, d_src, w, h, src_pitch, d_dst, dst_pitch, 0.5f);
}

int csb_scale_down_var(csb_ctx *ctx, const float *d_src, int w, int h, int src_pitch, float *d_dst, int dst_pitch,
                       float variance) {
  if (!ctx || !d_src || !d_dst || w < 2 || h < 2 || !(variance > 0.0f))
    return fail(ctx, CSB_E_INVALID, "csb_scale_down: bad argument");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  float k3[3];
  scale_down_kernel(k3, variance);
  {
    LaunchScope ls(ctx, s, "scale_down");
    launch_scale_down(d_src, w, h, src_pitch, d_dst, dst_pitch, k3, s->stream);
  }
  CSB_CHECK(ctx, cudaGetLastError());
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  if (ctx->profile) prof_collect(ctx, s);
  return 0;
}

int csb_rootsift(csb_ctx *ctx, void *d_sift, int n) {
  if (!ctx || !d_sift || n < 0) return fail(ctx, CSB_E_INVALID, "csb_rootsift: bad argument");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  {
    LaunchScope ls(ctx, s, "rootsift");
    launch_rootsift((csb_sift_point *)d_sift, n, s->stream);
  }
  CSB_CHECK(ctx, cudaGetLastError());
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  if (ctx->profile) prof_collect(ctx, s);
  return 0;
}

int csb_match(csb_ctx *ctx, void *d_sift1, int n1, const void *d_sift2, int n2, int distance, void *h_sift1) {
  if (!ctx || n1 < 0 || n2 < 0 || (distance != 0 && distance != 1)) return fail(ctx, CSB_E_INVALID, "csb_match: bad argument");
  if (n1 == 0 || n2 == 0) return 0;   // matching.cu:282-283: nothing to do
  if (!d_sift1 || !d_sift2) return fail(ctx, CSB_E_INVALID, "csb_match: null device data");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  if (ctx->profile) launch_delay(150000ull, s->stream);   // see enqueue_frame
  const bool use_tc = !ctx->match_exact && n1 >= 256 && n2 >= 256;
  if (!use_tc) {
    LaunchScope ls(ctx, s, "match");
    launch_match((csb_sift_point *)d_sift1, n1, (const csb_sift_point *)d_sift2, n2, distance, s->stream);
  } else {
    // tensor-core path: pack -> tcgen05 scan (short list per query and split) -> exact rescoring
    const int ns[2] = {n1, n2};
    const void *sets[2] = {d_sift1, d_sift2};
    for (int i = 0; i < 2; i++) {
      const size_t need = tc_packed_bytes(ns[i]);
      if (ctx->tc_pack_cap[i] < need) {
        if (ctx->tc_pack[i]) cudaFree(ctx->tc_pack[i]);
        ctx->tc_pack[i] = nullptr;
        CSB_CHECK(ctx, cudaMalloc(&ctx->tc_pack[i], need));
        ctx->tc_pack_cap[i] = need;
      }
    }
    const int splits = tc_splits(n1, n2, ctx->sm_count);
    const size_t sl_need = tc_shortlist_ints(n1);           // up to 4 splits x TC_TOPK entries per query
    if (ctx->tc_sl_cap < sl_need) {
      if (ctx->tc_val) cudaFree(ctx->tc_val);
      if (ctx->tc_idx) cudaFree(ctx->tc_idx);
      ctx->tc_val = nullptr; ctx->tc_idx = nullptr;
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->tc_val, sizeof(float) * tc_shortlist_floats(n1)));
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->tc_idx, sizeof(int) * sl_need));
      ctx->tc_sl_cap = sl_need;
    }
    const size_t blk_need = (size_t)(n1 + 15) / 16;
    if (ctx->tc_blk_cap < blk_need) {
      if (ctx->tc_flags) cudaFree(ctx->tc_flags);
      if (ctx->tc_list) cudaFree(ctx->tc_list);
      if (ctx->tc_part) cudaFree(ctx->tc_part);
      ctx->tc_flags = nullptr; ctx->tc_list = nullptr; ctx->tc_part = nullptr;
      CSB_CHECK(ctx, cudaMalloc(&ctx->tc_part, match_redo_scratch_bytes((int)blk_need)));
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->tc_flags, sizeof(int) * blk_need));
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->tc_list, sizeof(int) * blk_need));
      ctx->tc_blk_cap = blk_need;
    }
    if (!ctx->tc_count) {
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->tc_count, 256));
      CSB_CHECK(ctx, cudaMemset(ctx->tc_count, 0, 256));
      CSB_CHECK(ctx, cudaHostAlloc((void **)&ctx->h_tc_count, 256, cudaHostAllocDefault));
    }
    CSB_CHECK(ctx, cudaMemsetAsync(ctx->tc_count + 4, 0, 4 * sizeof(int), s->stream));   // {out-of-domain, err^2} x 2 sets
    {
      LaunchScope ls(ctx, s, "match_pack");
      ctx->launches += 1;
      launch_pack_f16((const csb_sift_point *)sets[0], ns[0], ctx->tc_pack[0], ctx->tc_count + 4, s->stream);
      launch_pack_f16((const csb_sift_point *)sets[1], ns[1], ctx->tc_pack[1], ctx->tc_count + 6, s->stream);
    }
    {
      LaunchScope ls(ctx, s, "match_tc");
      CSB_CHECK(ctx, (cudaError_t)launch_match_tc(ctx->tc_pack[0], n1, ctx->tc_pack[1], n2, splits, ctx->tc_val, ctx->tc_idx,
                                                  ctx->tc_count + 4, ctx->tc_count + 6, ctx->tc_flags, ctx->tc_count, s->stream));
    }
    {
      LaunchScope ls(ctx, s, "match_rescore");
      ctx->launches += 1;
      launch_rescore((csb_sift_point *)d_sift1, n1, (const csb_sift_point *)d_sift2, n2, ctx->tc_val, ctx->tc_idx, splits,
                     distance, ctx->tc_count + 4, ctx->tc_count + 6, ctx->tc_flags, ctx->tc_list, ctx->tc_count, s->stream);
    }
    {
      LaunchScope ls(ctx, s, "match_redo");
      ctx->launches += 1;
      launch_match_blocks((csb_sift_point *)d_sift1, n1, (const csb_sift_point *)d_sift2, n2, distance, ctx->tc_list,
                          ctx->tc_count, (int)blk_need, ctx->tc_part, s->stream);
    }
    CSB_CHECK(ctx, cudaMemcpyAsync(ctx->h_tc_count, ctx->tc_count, 8 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  }
  CSB_CHECK(ctx, cudaGetLastError());
  if (use_tc) {
    CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
    ctx->tc_redo_blocks += ctx->h_tc_count[0];
    if (ctx->h_tc_count[4] || ctx->h_tc_count[6]) {
      // a descriptor set outside the domain in which the fp16 prefilter's error bound holds (not finite in fp16, or
      // squared norm > 1.002): the short lists prove nothing, so the exact fp32 kernel recomputes every query
      ctx->tc_domain_fallbacks++;
      LaunchScope ls(ctx, s, "match");
      launch_match((csb_sift_point *)d_sift1, n1, (const csb_sift_point *)d_sift2, n2, distance, s->stream);
      CSB_CHECK(ctx, cudaGetLastError());
    }
  }
  if (h_sift1) {   // the five match fields, strided (matching.cu:352-356)
    const csb_sift_point *d = (const csb_sift_point *)d_sift1;
    csb_sift_point *hp = (csb_sift_point *)h_sift1;
    CSB_CHECK(ctx, cudaMemcpy2DAsync(&hp[0].score, sizeof(csb_sift_point), &d[0].score, sizeof(csb_sift_point),
                                     5 * sizeof(float), n1, cudaMemcpyDeviceToHost, s->stream));
  }
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  if (ctx->profile) prof_collect(ctx, s);
  return 0;
}

long long csb_match_redo_blocks(const csb_ctx *ctx) { return ctx ? ctx->tc_redo_blocks : 0; }
long long csb_match_domain_fallbacks(const csb_ctx *ctx) { return ctx ? ctx->tc_domain_fallbacks : 0; }

int csb_find_homography(csb_ctx *ctx, const void *d_sift, int n, const int *h_rand_pts, int num_loops, float thresh,
                        float *H9, int *num_inliers) {
  if (!ctx || !H9 || !num_inliers) return fail(ctx, CSB_E_INVALID, "csb_find_homography: null output");
  *num_inliers = 0;
  H9[0] = H9[4] = H9[8] = 1.0f;
  H9[1] = H9[2] = H9[3] = H9[5] = H9[6] = H9[7] = 0.0f;
  if (!d_sift || !h_rand_pts || num_loops <= 0 || (num_loops % 16) != 0)
    return fail(ctx, CSB_E_INVALID, "csb_find_homography: bad argument (num_loops must be a positive multiple of 16)");
  if (n < 8) return 0;   // homography.cu:207-208
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  const int n_up = ((n + 15) / 16) * 16;
  if (ctx->coord_cap < (size_t)n_up) {
    if (ctx->d_coord) cudaFree(ctx->d_coord);
    ctx->d_coord = nullptr;
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->d_coord, sizeof(float) * 4 * (size_t)n_up));
    ctx->coord_cap = n_up;
  }
  if (ctx->loops_cap < (size_t)num_loops) {
    if (ctx->d_homo) cudaFree(ctx->d_homo);
    if (ctx->d_rand) cudaFree(ctx->d_rand);
    if (ctx->d_counts) cudaFree(ctx->d_counts);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    ctx->d_homo = nullptr; ctx->d_rand = nullptr; ctx->d_counts = nullptr; ctx->h_counts = nullptr;
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->d_homo, sizeof(float) * 8 * (size_t)num_loops));
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->d_rand, sizeof(int) * 4 * (size_t)num_loops));
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->d_counts, sizeof(int) * (size_t)num_loops));
    CSB_CHECK(ctx, cudaHostAlloc((void **)&ctx->h_counts, sizeof(int) * (size_t)num_loops, cudaHostAllocDefault));
    ctx->loops_cap = num_loops;
  }
  for (int i = 0; i < 4 * num_loops; i++)
    if (h_rand_pts[i] < 0 || h_rand_pts[i] >= n) return fail(ctx, CSB_E_INVALID, "csb_find_homography: sample index out of range");
  CSB_CHECK(ctx, cudaMemcpyAsync(ctx->d_rand, h_rand_pts, sizeof(int) * 4 * (size_t)num_loops, cudaMemcpyHostToDevice,
                                 s->stream));
  {
    LaunchScope ls(ctx, s, "homography");
    ctx->launches += 2;
    launch_homography((const csb_sift_point *)d_sift, n, n_up, ctx->d_coord, ctx->d_rand, ctx->d_homo, ctx->d_counts,
                      num_loops, thresh * thresh, s->stream);
  }
  CSB_CHECK(ctx, cudaGetLastError());
  CSB_CHECK(ctx, cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(int) * (size_t)num_loops, cudaMemcpyDeviceToHost,
                                 s->stream));
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  int maxIndex = -1, maxCount = -1;   // first maximum wins, homography.cu:259-264
  for (int i = 0; i < num_loops; i++)
    if (ctx->h_counts[i] > maxCount) { maxCount = ctx->h_counts[i]; maxIndex = i; }
  *num_inliers = maxCount;
  CSB_CHECK(ctx, cudaMemcpy2D(H9, sizeof(float), &ctx->d_homo[maxIndex], sizeof(float) * num_loops, sizeof(float), 8,
                              cudaMemcpyDeviceToHost));
  H9[8] = 1.0f;
  if (ctx->profile) prof_collect(ctx, s);
  return 0;
}

int csb_improve_homography(csb_ctx *ctx, void *d_sift, int n, float *H9, int num_loops, float min_score, float max_ambiguity,
                           float thresh, int *num_fit, void *h_sift) {
  if (!ctx || !H9 || !num_fit || n < 0 || num_loops < 0) return fail(ctx, CSB_E_INVALID, "csb_improve_homography: bad argument");
  *num_fit = 0;
  if (n == 0) return 0;
  if (!d_sift) return fail(ctx, CSB_E_INVALID, "csb_improve_homography: null device data");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  if (!ctx->ih_dev) {
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->ih_dev, 512));
    CSB_CHECK(ctx, cudaMemset(ctx->ih_dev, 0, 512));
    CSB_CHECK(ctx, cudaHostAlloc((void **)&ctx->ih_host, 512, cudaHostAllocDefault));
    memset(ctx->ih_host, 0, 512);
  }
  // layout (device and pinned mirror): [0,36) H_in, [64,100) H_out, [128,132) numfit, [256,..) job record
  float *dH_in = (float *)ctx->ih_dev, *dH_out = (float *)(ctx->ih_dev + 64);
  int *d_nf = (int *)(ctx->ih_dev + 128);
  memcpy(ctx->ih_host, H9, 9 * sizeof(float));
  improve_job_fill(ctx->ih_host + 256, d_sift, n, dH_in, dH_out, d_nf);
  CSB_CHECK(ctx, cudaMemcpyAsync(ctx->ih_dev, ctx->ih_host, 36, cudaMemcpyHostToDevice, s->stream));
  CSB_CHECK(ctx, cudaMemcpyAsync(ctx->ih_dev + 256, ctx->ih_host + 256, improve_job_bytes(), cudaMemcpyHostToDevice, s->stream));
  {
    LaunchScope ls(ctx, s, "improve_homography");
    launch_improve_homography(ctx->ih_dev + 256, 1, num_loops, min_score, max_ambiguity, thresh * thresh, s->stream);
  }
  CSB_CHECK(ctx, cudaGetLastError());
  CSB_CHECK(ctx, cudaMemcpyAsync(ctx->ih_host + 64, ctx->ih_dev + 64, 68, cudaMemcpyDeviceToHost, s->stream));
  if (h_sift) {   // match_error is the one field the reference's ImproveHomography writes (homography.cu:339)
    const csb_sift_point *d = (const csb_sift_point *)d_sift;
    csb_sift_point *hp = (csb_sift_point *)h_sift;
    CSB_CHECK(ctx, cudaMemcpy2DAsync(&hp[0].match_error, sizeof(csb_sift_point), &d[0].match_error, sizeof(csb_sift_point),
                                     sizeof(float), n, cudaMemcpyDeviceToHost, s->stream));
  }
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  memcpy(H9, ctx->ih_host + 64, 9 * sizeof(float));
  *num_fit = *(int *)(ctx->ih_host + 128);
  if (ctx->profile) prof_collect(ctx, s);
  return 0;
}

unsigned int csb_sample_hash(unsigned int seed, unsigned int pair, unsigned int loop, unsigned int k,
                             unsigned int attempt) {
  return csb_sample_hash_host(seed, pair, loop, k, attempt);
}

int csb_allpairs_match_ransac(csb_ctx *ctx, int n_sets, void *const *d_sifts, const int *counts, int n_pairs,
                              const int *pair_i, const int *pair_j, const unsigned int *pair_ids, int distance,
                              int num_loops, float min_score, float max_ambiguity, float thresh, unsigned int seed,
                              float *H_out, int *inliers_out, int *nvalid_out) {
  return csb_allpairs_match_ransac_improve(ctx, n_sets, d_sifts, counts, n_pairs, pair_i, pair_j, pair_ids, distance, num_loops,
                                           min_score, max_ambiguity, thresh, seed, 0, 0.0f, H_out, inliers_out, nvalid_out,
                                           nullptr, nullptr);
}

int csb_allpairs_match_ransac_improve(csb_ctx *ctx, int n_sets, void *const *d_sifts, const int *counts, int n_pairs,
                                      const int *pair_i, const int *pair_j, const unsigned int *pair_ids, int distance,
                                      int num_loops, float min_score, float max_ambiguity, float thresh, unsigned int seed,
                                      int improve_loops, float improve_thresh, float *H_out, int *inliers_out,
                                      int *nvalid_out, float *H_improved_out, int *numfit_out) {
  const bool improve = improve_loops > 0;
  if (improve && (!H_improved_out || !numfit_out)) return fail(ctx, CSB_E_INVALID, "csb_allpairs: improve outputs missing");
  if (!ctx || n_sets <= 0 || !d_sifts || !counts || n_pairs < 0 || !pair_i || !pair_j || !H_out || !inliers_out ||
      !nvalid_out || num_loops <= 0 || (num_loops % 16) != 0 || (distance != 0 && distance != 1))
    return fail(ctx, CSB_E_INVALID, "csb_allpairs_match_ransac: bad argument");
  if (n_pairs == 0) return 0;
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  Slot *s = &ctx->slots[0];
  cudaStream_t st = s->stream;
  int max_n = 0;
  for (int i = 0; i < n_sets; i++) {
    if (counts[i] < 0 || (counts[i] > 0 && !d_sifts[i])) return fail(ctx, CSB_E_INVALID, "csb_allpairs: bad set");
    if (counts[i] > max_n) max_n = counts[i];
  }
  for (int k = 0; k < n_pairs; k++)
    if (pair_i[k] < 0 || pair_i[k] >= n_sets || pair_j[k] < 0 || pair_j[k] >= n_sets)
      return fail(ctx, CSB_E_INVALID, "csb_allpairs: pair index out of range");
  // Pairs run round-robin on AP_STREAMS streams so that the small RANSAC kernels of one pair overlap
  // the tensor-core scan of the next.  Pairs that share the query set i write the same match fields,
  // so stream = i % n_streams keeps them in issue order.  Scratch lives in the context and only grows
  // (cudaMalloc/cudaFree cost more than a whole pair).
  constexpr int AP_STREAMS = csb_ctx::AP_STREAMS;
  const bool trace = getenv("CSB_AP_TRACE") != nullptr;
  auto now_ms = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now_ms();
  struct Scratch {
    float *sl_val, *d_coord, *d_homo;
    int *sl_idx, *flags, *list, *cnt, *d_valid, *d_nvalid, *d_rand, *d_counts;
    void *part;
    cudaStream_t st;
  } sc[AP_STREAMS];
  const bool use_tc = !ctx->match_exact;
  const int n_up = ((max_n + 15) / 16) * 16;
  int n_streams = AP_STREAMS;
  if (const char *e = getenv("CSB_AP_STREAMS")) n_streams = atoi(e) < 1 ? 1 : (atoi(e) > AP_STREAMS ? AP_STREAMS : atoi(e));
  if (n_pairs < n_streams) n_streams = n_pairs;
  if (!ctx->ap_packed) {
    for (int t = 0; t < AP_STREAMS; t++) {
      CSB_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->ap_stream[t], cudaStreamNonBlocking));
      CSB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ap_done[t], cudaEventDisableTiming));
    }
    CSB_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ap_packed, cudaEventDisableTiming));
  }
  {
    const size_t sl_need = tc_shortlist_ints(max_n > 0 ? max_n : 1);
    const size_t nblk = (size_t)max_n / 16 + 2;
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_slv = carve(4 * tc_shortlist_floats(max_n > 0 ? max_n : 1)), o_sli = carve(4 * sl_need), o_flags = carve(4 * nblk), o_list = carve(4 * nblk),
                 o_cnt = carve(256), o_part = carve(match_redo_scratch_bytes((int)nblk)), o_valid = carve(4 * (size_t)(max_n + 1)),
                 o_nvalid = carve(256), o_coord = carve(16 * (size_t)(n_up + 16)), o_rand = carve(16 * (size_t)num_loops),
                 o_homo = carve(32 * (size_t)num_loops), o_counts = carve(4 * (size_t)num_loops);
    if (ctx->ap_scratch_cap < off * AP_STREAMS) {
      if (ctx->ap_scratch) cudaFree(ctx->ap_scratch);
      ctx->ap_scratch = nullptr; ctx->ap_scratch_cap = 0;
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->ap_scratch, off * AP_STREAMS));
      ctx->ap_scratch_cap = off * AP_STREAMS;
    }
    for (int t = 0; t < AP_STREAMS; t++) {
      Scratch &c = sc[t];
      char *base = ctx->ap_scratch + off * t;
      c.sl_val = (float *)(base + o_slv); c.sl_idx = (int *)(base + o_sli); c.flags = (int *)(base + o_flags);
      c.list = (int *)(base + o_list); c.cnt = (int *)(base + o_cnt); c.part = base + o_part;
      c.d_valid = (int *)(base + o_valid); c.d_nvalid = (int *)(base + o_nvalid); c.d_coord = (float *)(base + o_coord);
      c.d_rand = (int *)(base + o_rand); c.d_homo = (float *)(base + o_homo); c.d_counts = (int *)(base + o_counts);
      c.st = ctx->ap_stream[t];
    }
  }
  const size_t res_need = (size_t)n_pairs * 21 * 4;   // H[9], inliers, n_valid, H_improved[9], numfit
  if (ctx->ap_result_cap < res_need) {
    if (ctx->ap_result) cudaFree(ctx->ap_result);
    ctx->ap_result = nullptr; ctx->ap_result_cap = 0;
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->ap_result, res_need));
    ctx->ap_result_cap = res_need;
  }
  float *d_H = (float *)ctx->ap_result;
  int *d_inl = (int *)(ctx->ap_result + (size_t)n_pairs * 36), *d_nv = d_inl + n_pairs;
  float *d_Himp = (float *)(d_nv + n_pairs);
  int *d_nfit = (int *)(d_Himp + 9 * (size_t)n_pairs);
  const size_t job_b = improve_job_bytes();
  if (improve) {
    if (ctx->ap_jobs_cap < job_b * n_pairs) {
      if (ctx->ap_jobs) cudaFree(ctx->ap_jobs);
      ctx->ap_jobs = nullptr; ctx->ap_jobs_cap = 0;
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->ap_jobs, job_b * n_pairs));
      ctx->ap_jobs_cap = job_b * n_pairs;
    }
    std::vector<char> jobs(job_b * n_pairs);
    for (int k = 0; k < n_pairs; k++)
      improve_job_fill(jobs.data() + job_b * k, d_sifts[pair_i[k]], counts[pair_i[k]], d_H + 9 * (size_t)k, d_Himp + 9 * (size_t)k,
                       d_nfit + k);
    CSB_CHECK(ctx, cudaMemcpyAsync(ctx->ap_jobs, jobs.data(), jobs.size(), cudaMemcpyHostToDevice, st));
    CSB_CHECK(ctx, cudaStreamSynchronize(st));      // `jobs` is pageable and goes out of scope
  }
  std::vector<void *> packed(n_sets, nullptr);
  if (ctx->ap_flags_cap < (size_t)n_sets) {
    if (ctx->ap_flags) cudaFree(ctx->ap_flags);
    if (ctx->h_ap_flags) cudaFreeHost(ctx->h_ap_flags);
    ctx->ap_flags = nullptr; ctx->h_ap_flags = nullptr; ctx->ap_flags_cap = 0;
    CSB_CHECK(ctx, cudaMalloc((void **)&ctx->ap_flags, 2 * sizeof(int) * (size_t)n_sets));
    CSB_CHECK(ctx, cudaHostAlloc((void **)&ctx->h_ap_flags, 2 * sizeof(int) * (size_t)n_sets, cudaHostAllocDefault));
    ctx->ap_flags_cap = n_sets;
  }
  for (int i = 0; i < 2 * n_sets; i++) ctx->h_ap_flags[i] = 0;
  if (use_tc) {
    CSB_CHECK(ctx, cudaMemsetAsync(ctx->ap_flags, 0, 2 * sizeof(int) * (size_t)n_sets, st));
    size_t need = 0;
    for (int i = 0; i < n_sets; i++) if (counts[i] >= 256) need += (tc_packed_bytes(counts[i]) + 1023) & ~(size_t)1023;
    if (ctx->ap_pack_cap < need) {
      if (ctx->ap_pack) cudaFree(ctx->ap_pack);
      ctx->ap_pack = nullptr; ctx->ap_pack_cap = 0;
      CSB_CHECK(ctx, cudaMalloc((void **)&ctx->ap_pack, need));
      ctx->ap_pack_cap = need;
    }
    size_t off = 0;
    for (int i = 0; i < n_sets; i++) {
      if (counts[i] < 256) continue;
      packed[i] = ctx->ap_pack + off;
      off += (tc_packed_bytes(counts[i]) + 1023) & ~(size_t)1023;
      launch_pack_f16((const csb_sift_point *)d_sifts[i], counts[i], packed[i], ctx->ap_flags + 2 * i, st);
      ctx->launches++;
    }
    // which sets satisfy the fp16 prefilter's precondition?  (one small read-back per batch of pairs)
    CSB_CHECK(ctx, cudaMemcpyAsync(ctx->h_ap_flags, ctx->ap_flags, 2 * sizeof(int) * (size_t)n_sets, cudaMemcpyDeviceToHost, st));
    CSB_CHECK(ctx, cudaStreamSynchronize(st));
  }
  CSB_CHECK(ctx, cudaEventRecord(ctx->ap_packed, st));
  for (int t = 0; t < n_streams; t++) CSB_CHECK(ctx, cudaStreamWaitEvent(sc[t].st, ctx->ap_packed, 0));
  const double t_alloc = now_ms();
  for (int k = 0; k < n_pairs; k++) {
    const int i = pair_i[k], j = pair_j[k];
    const int n1 = counts[i], n2 = counts[j];
    Scratch &c = sc[i % n_streams];
    csb_sift_point *s1 = (csb_sift_point *)d_sifts[i];
    const csb_sift_point *s2 = (const csb_sift_point *)d_sifts[j];
    if (n1 > 0 && n2 > 0) {
      const bool in_domain = !ctx->h_ap_flags[2 * i] && !ctx->h_ap_flags[2 * j];
      if (use_tc && n1 >= 256 && n2 >= 256 && !in_domain) ctx->tc_domain_fallbacks++;
      if (use_tc && n1 >= 256 && n2 >= 256 && in_domain) {
        const int splits = tc_splits(n1, n2, ctx->sm_count);
        CSB_CHECK(ctx, (cudaError_t)launch_match_tc(packed[i], n1, packed[j], n2, splits, c.sl_val, c.sl_idx, ctx->ap_flags + 2 * i,
                                                    ctx->ap_flags + 2 * j, c.flags, c.cnt, c.st));
        launch_rescore(s1, n1, s2, n2, c.sl_val, c.sl_idx, splits, distance, ctx->ap_flags + 2 * i, ctx->ap_flags + 2 * j, c.flags,
                       c.list, c.cnt, c.st);
        launch_match_blocks(s1, n1, s2, n2, distance, c.list, c.cnt, (n1 + 15) / 16, c.part, c.st);
        ctx->launches += 5;
      } else {
        launch_match(s1, n1, s2, n2, distance, c.st);
        ctx->launches += 1;
      }
    }
    const int nu = ((n1 + 15) / 16) * 16;
    launch_pair_ransac(s1, n1, nu > 0 ? nu : 16, min_score, max_ambiguity, c.d_valid, c.d_nvalid, c.d_coord, c.d_rand,
                       c.d_homo, c.d_counts, num_loops, thresh * thresh, seed, pair_ids ? pair_ids[k] : (unsigned int)k,
                       d_H + 9 * (size_t)k, d_inl + k, d_nv + k, c.st);
    ctx->launches += 6;
    if (improve && n1 > 0) {   // main.cpp:335: ImproveHomography right after FindHomography
      launch_improve_homography(ctx->ap_jobs + job_b * k, 1, improve_loops, min_score, max_ambiguity,
                                improve_thresh * improve_thresh, c.st);
      ctx->launches += 1;
    }
  }
  CSB_CHECK(ctx, cudaGetLastError());
  for (int t = 0; t < n_streams; t++) {
    CSB_CHECK(ctx, cudaEventRecord(ctx->ap_done[t], sc[t].st));
    CSB_CHECK(ctx, cudaStreamWaitEvent(st, ctx->ap_done[t], 0));
  }
  CSB_CHECK(ctx, cudaMemcpyAsync(H_out, d_H, sizeof(float) * 9 * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
  CSB_CHECK(ctx, cudaMemcpyAsync(inliers_out, d_inl, sizeof(int) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
  CSB_CHECK(ctx, cudaMemcpyAsync(nvalid_out, d_nv, sizeof(int) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
  if (improve) {
    CSB_CHECK(ctx, cudaMemcpyAsync(H_improved_out, d_Himp, sizeof(float) * 9 * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
    CSB_CHECK(ctx, cudaMemcpyAsync(numfit_out, d_nfit, sizeof(int) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
  }
  const double t_issue = now_ms();
  CSB_CHECK(ctx, cudaStreamSynchronize(st));
  if (trace)
    fprintf(stderr, "[csb allpairs] %d pairs on %d streams: prepare+pack %.2f ms, issue %.2f ms, drain %.2f ms\n", n_pairs,
            n_streams, t_alloc - t_begin, t_issue - t_alloc, now_ms() - t_issue);
  return 0;
}

int csb_debug_octave(csb_ctx *ctx, int oct, float *h_base, float *h_dog, int *w, int *h) {
  if (!ctx) return CSB_E_INVALID;
  Slot *s = &ctx->slots[0];
  if (oct < 0 || oct >= s->n_oct) return fail(ctx, CSB_E_INVALID, "csb_debug_octave: no such octave in slot 0");
  CSB_CHECK(ctx, cudaSetDevice(ctx->device));
  CSB_CHECK(ctx, cudaStreamSynchronize(s->stream));
  const Octave &o = s->oct[oct];
  if (w) *w = o.w;
  if (h) *h = o.h;
  if (h_base) {
    if (oct == 0) return fail(ctx, CSB_E_INVALID, "csb_debug_octave: octave 0 base is the caller's frame");
    CSB_CHECK(ctx, cudaMemcpy2D(h_base, sizeof(float) * o.w, o.base, sizeof(float) * o.pitch, sizeof(float) * o.w, o.h,
                                cudaMemcpyDeviceToHost));
  }
  if (h_dog) {
    for (int i = 0; i < CSB_NUM_DOG; i++)
      CSB_CHECK(ctx, cudaMemcpy2D(h_dog + (size_t)i * o.w * o.h, sizeof(float) * o.w, o.dog + (size_t)i * CSB_DOG_PS(o.pitch, o.h),
                                  sizeof(float) * CSB_DOG_RS(o.pitch), sizeof(float) * o.w, o.h, cudaMemcpyDeviceToHost));
  }
  return 0;
}

int csb_profile_enable(csb_ctx *ctx, int on) {
  if (!ctx) return CSB_E_INVALID;
  ctx->profile = on != 0;
  return 0;
}
int csb_profile_reset(csb_ctx *ctx) {
  if (!ctx) return CSB_E_INVALID;
  for (ProfEntry &e : ctx->prof) { e.total_ms = 0.0; e.launches = 0; }
  return 0;
}
int csb_profile_count(const csb_ctx *ctx) { return ctx ? (int)ctx->prof.size() : 0; }
int csb_profile_get(csb_ctx *ctx, int idx, const char **name, double *total_ms, long long *launches) {
  if (!ctx || idx < 0 || idx >= (int)ctx->prof.size()) return CSB_E_INVALID;
  if (name) *name = ctx->prof[idx].name.c_str();
  if (total_ms) *total_ms = ctx->prof[idx].total_ms;
  if (launches) *launches = ctx->prof[idx].launches;
  return 0;
}
long long csb_launch_count(const csb_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
