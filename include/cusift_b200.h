/*
 * cusift_b200 — C ABI of the B200-native SIFT hot path.
 *
 * This is the drop-in boundary underneath the reference-compatible C++ headers in
 * include/cusift/ (cuImage.h, cuSIFT.h, extras/matching.h, extras/homography.h).
 * Plain pointers and sizes only; every entry point returns 0 on success or a
 * non-zero status (a cudaError_t value, or CSB_E_* below) and never exits the
 * process.  Each declaration cites the reference interface it replaces
 * (file:line under danielsuo/cuSIFT).
 *
 * SiftPoint records crossing this boundary use the reference layout
 * (cuSIFT.h:10-30): 588 bytes, see csb_sift_point below.
 */
#ifndef CUSIFT_B200_H
#define CUSIFT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CSB_VERSION 100
#define CSB_MAX_OCTAVES 8
#define CSB_NUM_SCALES 5            /* cuSIFT_D.h:8  NUM_SCALES */

#define CSB_E_INVALID   10001       /* bad argument */
#define CSB_E_NOMEM     10002
#define CSB_E_TOOMANY   10003       /* more octaves / slots than supported */

/* cuSIFT.h:10-30 (class SiftPoint), byte-for-byte. */
typedef struct csb_sift_point {
  float coords2D[2];
  float scale;
  float sharpness;
  float edgeness;
  float orientation;     /* degrees, [0,360) */
  float score;
  float ambiguity;
  int   match;
  float match_xpos;
  float match_ypos;
  float match_error;
  float subsampling;
  float empty[3];
  float data[128];
  float coords3D[3];
} csb_sift_point;

/* Opt-in compact result record (csb_extract_batch_compact): what a consumer of keypoints + descriptors needs, with the
 * descriptor rounded to IEEE fp16 (SIFT / RootSIFT descriptor entries lie in [0, 1]: relative error <= 2^-11).
 * 288 bytes instead of 588: halves the device-to-host traffic of a frame.  NOT the reference layout - the default
 * entry points keep delivering SiftPoint records. */
typedef struct csb_compact_point {
  float x, y;            /* coords2D */
  float scale;
  float orientation;     /* degrees */
  float sharpness;
  float edgeness;
  float subsampling;
  float reserved;        /* 0 */
  unsigned short data[128];   /* descriptor, fp16 bits */
} csb_compact_point;

/* Extraction parameters: the public fields of SiftData (cuSIFT.h:45-51) plus the
 * subsampling argument of SiftData::Extract (cuSIFT.h:63). */
typedef struct csb_params {
  int    num_octaves;    /* SiftData::numOctaves                                  */
  double init_blur;      /* SiftData::initBlur                                    */
  float  peak_thresh;    /* SiftData::peakThresh  (DoG contrast threshold)        */
  float  edge_thresh;    /* SiftData::edgeThresh  (tra^2 < edge*det)              */
  float  lowest_scale;   /* SiftData::lowestScale (octave skipped unless < 2*sub) */
  float  subsampling;    /* Extract(..., subsampling = 1.0f)                      */
  int    rootsift;       /* !=0: ConvertSiftToRootSift after extraction           */
                         /* (legacy ExtractRootSift, cuSIFT.cu:122-134)           */
} csb_params;

typedef struct csb_ctx csb_ctx;

/* ---- context -------------------------------------------------------------
 * One context per GPU (replaces the reference's file-scope __constant__/__device__
 * state, cuSIFT_D.cu:13-20, and InitCuda, cutils.h:71-92).  `num_slots` frames may
 * be in flight at once (each slot owns a stream, a pyramid workspace, per-octave
 * texture objects and pinned staging); 0 picks the default (4). */
int  csb_ctx_create(int device, int num_slots, csb_ctx **out);
void csb_ctx_destroy(csb_ctx *ctx);
int  csb_ctx_device(const csb_ctx *ctx);
int  csb_ctx_num_slots(const csb_ctx *ctx);
const char *csb_last_error(const csb_ctx *ctx);
int  csb_version(void);
int  csb_sizeof_sift_point(void);

/* ---- pinned host buffers ---------------------------------------------------
 * Host SiftPoint arrays handed to csb_extract* may be any host memory; arrays
 * obtained here are page-locked + device-mapped, which lets the result be written
 * by the GPU directly (no staging copy).  Replaces malloc/free of
 * SiftData::h_data (cuSIFT.cu:22-25,43-46). */
int  csb_host_alloc(void **ptr, unsigned long long bytes);
int  csb_host_free(void *ptr);

/* ---- device buffers (for FFI callers without a CUDA runtime binding) ---------
 * Replace cudaMalloc in the SiftData constructor (cuSIFT.cu:27-30) and
 * cuImage::Allocate / HostToDevice / DeviceToHost (cuImage.cu:16-45,83-117).
 * csb_upload_image / csb_download_image copy a dense host frame (row stride w)
 * to / from a pitched device image (pitch in floats). */
int  csb_device_alloc(csb_ctx *ctx, void **d_ptr, unsigned long long bytes);
int  csb_device_free(csb_ctx *ctx, void *d_ptr);
/* A caller that frees a frame it allocated itself (cudaFree) tells the context, which caches a texture object and a
 * TMA descriptor per caller-owned frame (keyed by pointer and geometry); csb_device_free does this implicitly. */
int  csb_forget_image(csb_ctx *ctx, const void *d_ptr);
int  csb_memcpy_h2d(csb_ctx *ctx, void *d_dst, const void *h_src, unsigned long long bytes);
int  csb_memcpy_d2h(csb_ctx *ctx, void *h_dst, const void *d_src, unsigned long long bytes);
int  csb_upload_image(csb_ctx *ctx, float *d_img, int pitch_floats, const float *h_img, int w, int h);
int  csb_download_image(csb_ctx *ctx, float *h_img, const float *d_img, int pitch_floats, int w, int h);

/* ---- extraction ------------------------------------------------------------
 * csb_extract: legacy ExtractSift / ExtractRootSift contract (main.cpp:102,327;
 *   cuSIFT.cu:122-134,175-270): the frame is already on the device as a pitched
 *   fp32 image (cuImage, cuImage.h:8-26; pitch in floats).  On return d_sift
 *   holds *num_pts SiftPoints (at most max_pts), h_sift (optional) a copy.
 * csb_extract_host: HEAD SiftData::Extract(float*, w, h, subsampling)
 *   (cuSIFT.cu:61-120): dense host frame (row stride = w), upload included.
 * Both block until the result is complete.  Octaves are processed coarsest
 * first like ExtractSiftLoop (cuSIFT.cu:175-202); the order of points inside an
 * octave is unspecified (it is in the reference too: atomicInc, cuSIFT_D.cu:513).
 * Points beyond max_pts are dropped (the reference overwrites slot max_pts-1).  */
int csb_extract(csb_ctx *ctx, const float *d_img, int w, int h, int pitch_floats, const csb_params *p,
                void *d_sift, int max_pts, void *h_sift, int *num_pts);
int csb_extract_host(csb_ctx *ctx, const float *h_img, int w, int h, const csb_params *p,
                     void *d_sift, int max_pts, void *h_sift, int *num_pts);

/* Pipelined batch of independent frames of one shape (the frame-sharded
 * throughput path; the reference has no equivalent — a caller would loop over
 * SiftData::Extract).  imgs[i] is a device pitched image (imgs_on_host == 0) or a
 * dense host frame (imgs_on_host != 0).  d_sifts[i] / h_sifts[i] receive frame i
 * (h_sifts or any h_sifts[i] may be NULL); num_pts[i] its count.  Frames are in flight on up to
 * num_slots slots at once; a frame whose d_sifts[i] / h_sifts[i] is still in use by an earlier frame
 * waits for it, so reusing one buffer for every frame is legal but serialises the pipeline. */
int csb_extract_batch(csb_ctx *ctx, int n_frames, const float *const *imgs, int imgs_on_host, int w, int h,
                      int pitch_floats, const csb_params *p, void *const *d_sifts, void *const *h_sifts,
                      int max_pts, int *num_pts);

/* csb_extract_batch with compact host results: h_compact[i] receives num_pts[i] csb_compact_point records (288 B each)
 * instead of SiftPoint records; d_sifts[i] still receives the full SiftPoint array on the device (matching etc. work on
 * it).  Frames may be device images (imgs_on_host = 0), dense fp32 host frames (1) or dense 8-bit host frames (2: the
 * ingest path of csb_extract_batch_u8 without pre-blur; `pitch_floats` is then the row stride in bytes). */
int csb_extract_batch_compact(csb_ctx *ctx, int n_frames, const void *const *imgs, int imgs_on_host, int w, int h,
                              int pitch_floats, const csb_params *p, void *const *d_sifts, void *const *h_compact,
                              int max_pts, int *num_pts);

/* Rigid-transform RANSAC (EstimateRigidTransformH, extras/rigidTransform.cu:387-520).
 * h_coord: num_pts x 6 floats (reference-frame xyz, moving-frame xyz); type 0 = planar two-point fit
 * (x, z; y ignored), 1 = 3-D three-point quaternion fit.  h_indices: num_loops x 3 point indices, or
 * NULL to draw them on the device (counter-based hash of `seed`; the reference seeds cuRAND with
 * time(0)).  Returns the hypothesis with the most inliers (the LAST one on ties, like the reference's
 * `>=` scan), its inlier count and mask (h_inliers: num_pts bytes, may be NULL); for type 1 the
 * transform is re-estimated from all inliers on the host, as the reference does. */
int csb_rigid_transform(csb_ctx *ctx, const float *h_coord, int num_pts, int type, const int *h_indices, int num_loops,
                        float thresh2, unsigned int seed, float *Rt12, int *num_inliers, char *h_inliers);

/* the device-side sample generator of csb_rigid_transform (h_indices == NULL), for tests */
unsigned int csb_rigid_sample_hash(unsigned int seed, unsigned int loop, unsigned int k, unsigned int attempt);

/* Frame ingest (reference: the host-side preparation in main.cpp:301-318).
 * csb_ingest_u8: 8-bit grey frame (host or device memory, `stride` bytes per row) -> pitched fp32
 * device image; preblur != 0 applies cv::GaussianBlur(Size(3,3), 0.5) (main.cpp:308-309), bit-identical
 * to OpenCV 4's CV_32F result where OpenCV runs its SIMD loop body (within 1 ulp on its scalar row tails).  Synchronous. */
int csb_ingest_u8(csb_ctx *ctx, const unsigned char *src, int src_on_host, int w, int h, int stride, int preblur,
                  float *d_dst, int dst_pitch_floats);
/* csb_extract_batch over 8-bit HOST frames: per frame one w*h-byte upload (4x less PCIe traffic than
 * the fp32 frame), conversion (+ optional pre-blur) on the device into the slot's octave-0 image, then
 * the same pipeline and result delivery as csb_extract_batch. */
int csb_extract_batch_u8(csb_ctx *ctx, int n_frames, const unsigned char *const *h_imgs, int w, int h, int stride,
                         int preblur, const csb_params *p, void *const *d_sifts, void *const *h_sifts, int max_pts,
                         int *num_pts);

/* ScaleDown(cuImage &res, cuImage &src, 0.5f) (cuSIFT.h:76, cuSIFT.cu:313-353):
 * dst is (w/2) x (h/2).  Stores are guarded (the reference's are not). */
int csb_scale_down(csb_ctx *ctx, const float *d_src, int w, int h, int src_pitch, float *d_dst, int dst_pitch);
/* the same with the caller's variance (cuSIFT.cu:313-338 builds the 5 taps from it; csb_scale_down uses 0.5f) */
int csb_scale_down_var(csb_ctx *ctx, const float *d_src, int w, int h, int src_pitch, float *d_dst, int dst_pitch,
                       float variance);

/* SiftData::ConvertSiftToRootSift (cuSIFT.cu:383-395) on n device points. */
int csb_rootsift(csb_ctx *ctx, void *d_sift, int n);

/* ---- matching --------------------------------------------------------------
 * Device part of MatchSiftData (extras/matching.cu:272-357): for every point of
 * d_sift1 finds the best and second-best point of d_sift2 and writes score,
 * ambiguity, match, match_xpos, match_ypos into d_sift1 (and, when h_sift1 is
 * given, into the same five fields of the host copy, matching.cu:352-356).
 * distance: 0 = MatchSiftDistanceDotProduct, 1 = MatchSiftDistanceL2
 * (extras/matching.h:10-13). */
int csb_match(csb_ctx *ctx, void *d_sift1, int n1, const void *d_sift2, int n2, int distance, void *h_sift1);
/* Diagnostics of the tensor-core matcher (used for sets of >= 256 points; set CSB_MATCH_EXACT=1
 * to force the fp32 CUDA-core kernel): number of 16-query blocks whose short list could not be
 * proven complete and were redone exactly, accumulated since the context was created. */
long long csb_match_redo_blocks(const csb_ctx *ctx);
/* The tensor-core path prefilters with fp16 dot products whose error bound holds for descriptors that are
 * finite in fp16 with squared norm <= 1.002 (SIFT / RootSIFT descriptors are unit vectors) and, when the set's size
 * is not a multiple of 256, non-negative (the scan pads the set with zero rows).  The packing kernel
 * checks this (and measures the actual rounding-error norms, from which the per-pair bound ~6e-4 is derived); calls whose
 * sets fall outside are computed by the exact fp32 kernel instead, so csb_match equals the
 * reference for ANY SiftPoint.data.  Number of calls / pairs that took that route since the context was created: */
long long csb_match_domain_fallbacks(const csb_ctx *ctx);

/* ---- homography -------------------------------------------------------------
 * FindHomography (extras/homography.h:8, homography.cu:191-278).  Valid points
 * are those with score > min_score && ambiguity < max_ambiguity.  The 4-point
 * samples are drawn by the caller — h_rand_pts is int[4][num_loops] holding
 * indices into the point array (the reference draws them with libc rand(),
 * homography.cu:232-244; the C++ shim does exactly that) — so results are
 * reproducible.  num_loops must be a multiple of 16.  H9 gets the best
 * hypothesis (H9[8] = 1), num_inliers its inlier count over all n points. */
int csb_find_homography(csb_ctx *ctx, const void *d_sift, int n, const int *h_rand_pts, int num_loops,
                        float thresh, float *H9, int *num_inliers);

/* ---- all-pairs matching + RANSAC (BASELINE config 5) -----------------------------------------
 * For every listed pair (i, j): MatchSiftData(set_i, set_j, distance) — which, as in the reference,
 * overwrites set_i's five match fields in place — followed by FindHomography(set_i, num_loops,
 * min_score, max_ambiguity, thresh).  Everything stays on the device (descriptors are packed for the
 * tensor cores once per set, samples come from the counter-based generator csb_sample_hash, the
 * winning hypothesis is picked on the device); one synchronisation at the end.  Outputs, per pair:
 * H[9], the winner's inlier count, and the number of valid points.  The reference has no batched
 * entry point: a caller would loop over MatchSiftData + FindHomography (main.cpp:331-334).
 * pair_ids (optional) are the indices fed to the sample generator (default: position in the list),
 * so that a pair gets the same samples no matter which rank processes it. */
int csb_allpairs_match_ransac(csb_ctx *ctx, int n_sets, void *const *d_sifts, const int *counts, int n_pairs,
                              const int *pair_i, const int *pair_j, const unsigned int *pair_ids, int distance,
                              int num_loops, float min_score, float max_ambiguity, float thresh, unsigned int seed,
                              float *H_out, int *inliers_out, int *nvalid_out);
unsigned int csb_sample_hash(unsigned int seed, unsigned int pair, unsigned int loop, unsigned int k,
                             unsigned int attempt);
/* The same with ImproveHomography(set_i, H, improve_loops, min_score, max_ambiguity, improve_thresh) appended to every
 * pair (main.cpp:334-335), on the device: H_improved_out gets the refined homography, numfit_out its inlier count. */
int csb_allpairs_match_ransac_improve(csb_ctx *ctx, int n_sets, void *const *d_sifts, const int *counts, int n_pairs,
                                      const int *pair_i, const int *pair_j, const unsigned int *pair_ids, int distance,
                                      int num_loops, float min_score, float max_ambiguity, float thresh, unsigned int seed,
                                      int improve_loops, float improve_thresh, float *H_out, int *inliers_out,
                                      int *nvalid_out, float *H_improved_out, int *numfit_out);

/* ---- multi-GPU all-pairs (BASELINE config 5; the reference is single-GPU) ------------------------------------
 * One process per GPU.  Every rank owns `sets_per_rank` keypoint sets (device SiftPoint arrays of at most `cap`
 * points); global set g = rank * sets_per_rank + local index.  csb_allpairs_distributed all-gathers the sets with NCCL
 * (one ncclAllGather per local set index, on its own stream, underneath the pairs whose two sets are local),
 * partitions the n(n-1)/2 unordered pairs (i < j, query = i) cyclically by flattened pair index, runs
 * csb_allpairs_match_ransac_improve on the rank's share and all-gathers the per-pair results: on return EVERY rank
 * holds H / inliers / n_valid (/ H_improved / numfit) of ALL pairs in flattened (i-major) order; the sample generator
 * is fed the flattened pair index, so results do not depend on the number of ranks.
 * `comm` is an ncclComm_t over the same ranks (NCCL is bound at run time with dlopen("libnccl.so.2"), the library has
 * no link-time dependency on it).  Callers without an NCCL binding of their own create it here: rank 0 fills a
 * 128-byte id with csb_nccl_unique_id and distributes it by any means, every rank calls csb_nccl_comm_create.
 * world == 1 works without NCCL (comm may be NULL).  timings_ms (optional, 4 doubles): local pairs + queueing the
 * exchange, waiting for the exchange, remaining pairs, result exchange. */
int csb_nccl_unique_id(void *id128);
int csb_nccl_comm_create(csb_ctx *ctx, int rank, int world, const void *id128, void **comm_out);
int csb_nccl_comm_destroy(void *comm);
int csb_allpairs_distributed(csb_ctx *ctx, void *comm, int rank, int world, int sets_per_rank, void *const *d_local_sifts,
                             const int *local_counts, int cap, int distance, int num_loops, float min_score, float max_ambiguity,
                             float thresh, unsigned int seed, int improve_loops, float improve_thresh, float *H_out,
                             int *inliers_out, int *nvalid_out, float *H_improved_out, int *numfit_out, double *timings_ms);

/* ImproveHomography (extras/homography.cu:271-337; declared in main.cpp:19) on the DEVICE copy of the points:
 * `num_loops` rounds of re-weighted least squares (weights thresh^2 / (err + thresh^2), points with
 * score < min_score || ambiguity > max_ambiguity skipped, 8x8 normal equations in fp64, Cholesky), then the number of
 * points with err < thresh^2 (over ALL points, like the reference) and match_error = sqrt(err) of every point, written
 * to d_sift and, when h_sift is given, to the host records.  H9 is read (divided by H9[8]) and overwritten.
 * The reference does this on the host with OpenCV's cv::solve; the C++ shim keeps a host version for callers that
 * only have host data. */
int csb_improve_homography(csb_ctx *ctx, void *d_sift, int n, float *H9, int num_loops, float min_score, float max_ambiguity,
                           float thresh, int *num_fit, void *h_sift);

/* ---- debugging / measurement --------------------------------------------------
 * csb_debug_octave: after a csb_extract* call on slot 0, copies octave `oct`'s
 *   base image (dense w x h) and/or its 7 DoG planes (dense [7][h][w]) to host.
 * csb_profile_*: when enabled every kernel launch of the extraction path is
 *   bracketed by CUDA events on its own stream; csb_profile_get returns the
 *   accumulated device time (ms) and launch count per kernel name. */
int csb_debug_octave(csb_ctx *ctx, int oct, float *h_base, float *h_dog, int *w, int *h);
int csb_profile_enable(csb_ctx *ctx, int on);
int csb_profile_reset(csb_ctx *ctx);
int csb_profile_count(const csb_ctx *ctx);
int csb_profile_get(csb_ctx *ctx, int idx, const char **name, double *total_ms, long long *launches);
long long csb_launch_count(const csb_ctx *ctx);   /* kernels launched since creation */

#ifdef __cplusplus
}
#endif
#endif /* CUSIFT_B200_H */
