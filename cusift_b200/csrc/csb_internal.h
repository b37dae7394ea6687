// Internal declarations shared by the kernels and the C-ABI layer of cusift_b200.
#ifndef CSB_INTERNAL_H
#define CSB_INTERNAL_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cusift_b200.h"

#define CSB_NUM_LEVELS (CSB_NUM_SCALES + 3)   // blur levels per octave (LAPLACE_S, cuSIFT_D.h:23)
#define CSB_NUM_DOG (CSB_NUM_SCALES + 2)      // DoG planes per octave

// Layout of the 7 DoG planes of an octave (internal: only k_blur_dog*, k_find_points and csb_debug_octave
// see it).  Element (plane p, row y, column x) lives at  dog[p * csb_dog_ps + y * csb_dog_rs + x].
//   planar      (the reference's): ps = pitch * h, rs = pitch
//   interleaved (default here)   : ps = pitch,     rs = 7 * pitch  — the 7 planes of one image row are
//     adjacent, so a CTA that streams a band of rows through all planes walks ONE contiguous region
//     instead of seven regions megabytes apart.
#ifndef CSB_DOG_PLANAR
#define CSB_DOG_PS(pitch, h) ((size_t)(pitch))
#define CSB_DOG_RS(pitch) ((size_t)(pitch) * CSB_NUM_DOG)
#else
#define CSB_DOG_PS(pitch, h) ((size_t)(pitch) * (size_t)(h))
#define CSB_DOG_RS(pitch) ((size_t)(pitch))
#endif

// 9-tap symmetric Gaussian weights of the 8 blur levels of one octave, k[s][0..4]
// = taps at distance 4,3,2,1,0 (same indexing as d_Kernel2, cuSIFT.cu:405-410).
struct DogWeights {
  float k[CSB_NUM_LEVELS][5];
};

// Keypoint as k_find_points leaves it: octave-pixel coordinates, before orientation / descriptor.
// One list per octave (max_pts entries each) in slot-owned memory; k_orient_desc turns it into the
// caller's SiftPoint array, coarse octaves first (the reference's order, cuSIFT.cu:181-196).
struct KpStage {
  float x, y, scale, sharp, edge;
};

// Parameters of the extrema kernel (d_Threshold/d_EdgeLimit/d_Scales/d_Factor in the reference,
// cuSIFT_D.cu:13-15, pushed per launch at cuSIFT.cu:441-444).  ALL octaves of a frame go through
// ONE launch: a CTA finds its octave from cta_begin.
struct ExtremaOctave {
  const float *dog;                // 7 DoG planes, plane stride pitch*h
  int w, h, pitch;
  int tiles_x;                     // CTAs per row band
  int cta_begin;                   // first linear CTA index of this octave (octave 0 first: longest work first)
  int octave;
};
struct ExtremaParams {
  float thresh;                    // peakThresh
  float edge_limit;                // edgeThresh
  float scales[CSB_NUM_SCALES];    // sigma * 2^(i/5)
  float factor;                    // 1/NUM_SCALES
  int rows;                        // output rows per CTA (multiple of 3)
  int n_oct;                       // entries used in oct[]
  ExtremaOctave oct[CSB_MAX_OCTAVES];
};

// One 2-D TMA descriptor per octave over its row-interleaved DoG buffer viewed as a (pitch x 7h) matrix:
// a box of 128 columns x 21 matrix rows is "3 image rows x 7 planes", i.e. one pipeline stage of
// k_find_points in ONE cp.async.bulk.tensor (index = ExtremaParams::oct[] index).
struct alignas(64) ExtremaMaps {
  CUtensorMap m[CSB_MAX_OCTAVES];
};
// encodes the descriptor of one octave (host; returns 0 on success)
int make_dog_tensor_map(CUtensorMap *out, const float *dog, int h, int pitch);

// Per-octave view handed to the orientation/descriptor kernel.
struct OctaveTexSet {
  cudaTextureObject_t tex[CSB_MAX_OCTAVES];
};

// ---- k_pyramid (kernels_pyramid.cu): blur + DoG (+ downsample) of any set of octaves in ONE launch ----
struct DownK {
  float k0, k1, k2;                // ScaleDown taps (cuSIFT.cu:320-338): k0 = outer, k1 = inner, k2 = centre
};
struct PyramidOctave {
  float *dog;                      // 7 DoG planes, layout CSB_DOG_PS / CSB_DOG_RS of dpitch
  float *next;                     // base of the next octave (written when the launch downsamples), else unused
  int w, h, dpitch, npitch;
  int rows;                        // output rows per CTA (multiple of 4)
  int tiles_x;                     // CTAs per row band (one CTA = two 120-column strips)
  int cta_begin;                   // first linear CTA index of this octave
};
struct PyramidParams {
  int n_oct;
  DownK dk;
  PyramidOctave oct[CSB_MAX_OCTAVES];
  DogWeights W[CSB_MAX_OCTAVES];   // index = oct[] index (the weights depend on the octave's initBlur)
};
// one 2-D TMA descriptor per octave over its BASE image (w x h floats, row pitch in bytes): box = 248 x 4
struct alignas(64) PyramidMaps {
  CUtensorMap m[CSB_MAX_OCTAVES];
};
int pyramid_source_map(CUtensorMap *out, const float *base, int w, int h, int pitch);
bool pyramid_tma_ok(const float *base, int pitch);     // 16-byte aligned base, pitch multiple of 4 floats
void pyramid_set_weights(PyramidParams *pp, int idx, const DogWeights &wts);
// fills rows / tiles_x / cta_begin from the geometry already in pp->oct[]; returns the grid size
int plan_pyramid(PyramidParams *pp, int sm_count);
void launch_pyramid(const PyramidParams &pp, const PyramidMaps &maps, int n_ctas, bool down, cudaStream_t st);
// octave bases dst[0..n-1] (each half the size of the one before) from `src` in ceil(n/3) launches
void launch_down_chain(const float *src, int sw, int sh, int spitch, float *const *dst, const int *dw, const int *dh,
                       const int *dpitch, int n, const float k[3], cudaStream_t st);

// ---- kernel launchers (defined in the .cu files) ---------------------------
void launch_scale_down(const float *src, int w, int h, int spitch, float *dst, int dpitch, const float k[3],
                       cudaStream_t st);
// scalar fallback of k_pyramid (sources the TMA unit cannot address; CSB_NO_FUSE=1)
void launch_blur_dog(const float *base, int w, int h, int spitch, float *dog, int dpitch, const DogWeights &wts,
                     cudaStream_t st);
void launch_blur_dog_down(const float *base, int w, int h, int spitch, float *dog, int dpitch, const DogWeights &wts,
                          float *next, int npitch, const float k[3], cudaStream_t st);
// fills ep.rows / tiles_x / cta_begin from the octave geometry already in ep.oct[]; returns the grid size
int plan_find_points(ExtremaParams *ep, int sm_count);
void launch_find_points(const ExtremaParams &ep, const ExtremaMaps &maps, int n_ctas, KpStage *d_stage,
                        unsigned int *d_counter, int max_pts, cudaStream_t st);
// subs[o] = subsampling of octave o (multiplies coords2D / scale in the final record)
void launch_orient_desc(const OctaveTexSet &texs, int n_oct, const float *subs, const KpStage *d_stage, csb_sift_point *d_sift,
                        unsigned int *d_counter, int max_pts, int rootsift, int sm_count, cudaStream_t st);
void launch_ingest_u8(const unsigned char *d_src, int stride, int w, int h, float *d_dst, int pitch, int preblur, float k0,
                      float k1, cudaStream_t st);
void launch_delay(unsigned long long ns, cudaStream_t st);   // measurement aid (profiling mode only)
void launch_rootsift(csb_sift_point *d_sift, int n, cudaStream_t st);
// SiftPoint -> csb_compact_point for the first min(*d_count, max_pts) points
void launch_compact(const csb_sift_point *d_sift, const unsigned int *d_count, int max_pts, csb_compact_point *d_out,
                    cudaStream_t st);
void launch_match(csb_sift_point *d_sift1, int n1, const csb_sift_point *d_sift2, int n2, int distance,
                  cudaStream_t st);
#define CSB_REDO_SLICES 32
size_t match_redo_scratch_bytes(int max_blocks);
void launch_match_blocks(csb_sift_point *d_sift1, int n1, const csb_sift_point *d_sift2, int n2, int distance,
                         const int *block_list, const int *block_count, int max_blocks, void *part, cudaStream_t st);
// tensor-core matcher (kernels_match_tc.cu)
size_t tc_packed_bytes(int n);
int tc_pad(int n);
int tc_splits(int n1, int n2, int sm_count);
// info = 2 device ints zeroed by the caller: [0] set when the set violates the fp16 error bound's precondition,
// [1] = float bits of the largest squared fp16 rounding-error norm of a row (the pair's eps is derived from it)
void launch_pack_f16(const csb_sift_point *pts, int n, void *packed, int *info, cudaStream_t st);
// resets redo_flags[(n1 + 15) / 16] and *redo_count for the launch_rescore that must follow; returns the launch's cudaError_t
int launch_match_tc(const void *q_packed, int n1, const void *c_packed, int n2, int n_splits, float *sl_val, int *sl_idx,
                    const int *q_info, const int *c_info, int *redo_flags, int *redo_count, cudaStream_t st);
void launch_rescore(csb_sift_point *s1, int n1, const csb_sift_point *s2, int n2, const float *sl_val, const int *sl_idx,
                    int n_splits, int distance, const int *q_info, const int *c_info, int *redo_flags, int *redo_list,
                    int *redo_count, cudaStream_t st);
size_t tc_shortlist_floats(int n);   // floats of short-list scratch (sl_val) a query set of n points needs
size_t tc_shortlist_ints(int n);     // ints of short-list scratch (sl_idx)
void launch_homography(const csb_sift_point *d_sift, int n, int n_up, float *d_coord, const int *d_rand, float *d_homo,
                       int *d_counts, int num_loops, float thresh2, cudaStream_t st);

void launch_pair_ransac(const csb_sift_point *d_sift, int n, int n_up, float min_score, float max_amb, int *d_valid,
                        int *d_nvalid, float *d_coord, int *d_rand, float *d_homo, int *d_counts, int num_loops,
                        float thresh2, unsigned int seed, unsigned int pair, float *H_out, int *inl_out, int *nvalid_out,
                        cudaStream_t st);
// device ImproveHomography: one 8-CTA cluster per job (kernels_homography.cu); jobs = array of improve_job_bytes() records
void launch_improve_homography(const void *d_jobs, int n_jobs, int num_loops, float min_score, float max_amb, float limit,
                               cudaStream_t st);
size_t improve_job_bytes();
void improve_job_fill(void *h_job, void *d_pts, int n, const float *d_H_in, float *d_H_out, int *d_numfit);
// frees the multi-GPU exchange buffers of a context (csb_dist.cu); called by csb_ctx_destroy
void csb_dist_release(csb_ctx *ctx);
// rigid-transform RANSAC (kernels_rigid.cu)
void launch_rigid_hypotheses(const float *d_coord, int num_pts, int *d_indices, int draw, unsigned int seed, int type3d,
                             int num_loops, float thresh2, float *d_Rt, int *d_counts, cudaStream_t st);
void launch_rigid_mask(const float *d_coord, int num_pts, const float *d_Rt, int loop, float thresh2, char *d_mask,
                       cudaStream_t st);
void rigid_refit_host(const float *h_coord, const int *idx, int n, float *Rt);
unsigned int rigid_hash_host(unsigned int seed, unsigned int loop, unsigned int k, unsigned int attempt);
unsigned int csb_sample_hash_host(unsigned int seed, unsigned int pair, unsigned int loop, unsigned int k,
                                  unsigned int attempt);

#endif
