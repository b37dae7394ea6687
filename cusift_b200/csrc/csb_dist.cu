// Multi-GPU all-pairs matching (BASELINE config 5): one process per GPU, the keypoint sets are
// exchanged with NCCL all-gathers issued from here, the unordered pairs are partitioned cyclically, every rank runs
// the batched match + RANSAC (+ ImproveHomography) pipeline on its share and the per-pair results are all-gathered.
// The reference is single-GPU and has no equivalent; a caller would loop MatchSiftData + FindHomography
// (main.cpp:331-335) over the pairs.
//
// What is exchanged: the SiftPoint records themselves (588 B: coordinates + the fp32 descriptor).  The tensor-core
// prefilter only needs fp16 descriptors, but the exact fp32 rescoring that makes score / ambiguity / match
// bit-identical to the reference reads the fp32 descriptors of both sets, so they must travel; the header fields
// are 13 % of the record.  The exchange is cut into one all-gather per local set index ("chunk") on its own stream
// and runs underneath the pairs whose two sets are local to the rank.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 - the copy PyTorch has already loaded when the caller is the
// Python harness), so libcusift_b200.so itself has no link-time dependency on it.
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "csb_internal.h"

namespace {

// the subset of nccl.h used here (ABI-stable since NCCL 2.x)
typedef int ncclResult_t;
typedef void *ncclComm_t;
struct ncclUniqueId_ { char internal[128]; };
enum { kNcclInt8 = 0, kNcclInt32 = 2, kNcclFloat32 = 7 };

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi *nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) api.err = "libnccl lacks a required symbol";
  });
  return &api;
}

struct DistState {
  char *gathered = nullptr;      // [sets_per_rank][world][cap] SiftPoint records
  size_t gathered_cap = 0;
  int *d_counts = nullptr;       // [world][sets_per_rank] (+ send area behind it)
  size_t counts_cap = 0;
  float *d_res = nullptr;        // result exchange: [world][max_local][21] (+ send area)
  size_t res_cap = 0;
  cudaStream_t comm_stream = nullptr;
};
std::map<csb_ctx *, DistState> g_state;
std::mutex g_mu;

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

void csb_dist_release(csb_ctx *ctx) {
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_state.find(ctx);
  if (it == g_state.end()) return;
  DistState &s = it->second;
  if (s.gathered) cudaFree(s.gathered);
  if (s.d_counts) cudaFree(s.d_counts);
  if (s.d_res) cudaFree(s.d_res);
  if (s.comm_stream) cudaStreamDestroy(s.comm_stream);
  g_state.erase(it);
}

extern "C" {

int csb_nccl_unique_id(void *id128) {
  NcclApi *n = nccl();
  if (!id128 || !n->err.empty()) return CSB_E_INVALID;
  ncclUniqueId_ id;
  const ncclResult_t r = n->GetUniqueId(&id);
  if (r != 0) return 20000 + r;
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int csb_nccl_comm_create(csb_ctx *ctx, int rank, int world, const void *id128, void **comm_out) {
  NcclApi *n = nccl();
  if (!ctx || !id128 || !comm_out || rank < 0 || rank >= world || !n->err.empty()) return CSB_E_INVALID;
  cudaSetDevice(csb_ctx_device(ctx));
  ncclUniqueId_ id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  const ncclResult_t r = n->CommInitRank(&comm, world, id, rank);
  if (r != 0) return 20000 + r;
  *comm_out = comm;
  return 0;
}

int csb_nccl_comm_destroy(void *comm) {
  NcclApi *n = nccl();
  if (!comm || !n->err.empty()) return CSB_E_INVALID;
  const ncclResult_t r = n->CommDestroy((ncclComm_t)comm);
  return r == 0 ? 0 : 20000 + r;
}

int csb_allpairs_distributed(csb_ctx *ctx, void *comm, int rank, int world, int sets_per_rank, void *const *d_local_sifts,
                             const int *local_counts, int cap, int distance, int num_loops, float min_score, float max_ambiguity,
                             float thresh, unsigned int seed, int improve_loops, float improve_thresh, float *H_out,
                             int *inliers_out, int *nvalid_out, float *H_improved_out, int *numfit_out, double *timings_ms) {
  NcclApi *n = nccl();
  if (!ctx || world < 1 || rank < 0 || rank >= world || sets_per_rank < 1 || !d_local_sifts || !local_counts || cap < 1 ||
      !H_out || !inliers_out || !nvalid_out || (world > 1 && (!comm || !n->err.empty())))
    return CSB_E_INVALID;
  const bool improve = improve_loops > 0;
  if (improve && (!H_improved_out || !numfit_out)) return CSB_E_INVALID;
  for (int c = 0; c < sets_per_rank; c++)
    if (local_counts[c] < 0 || local_counts[c] > cap || !d_local_sifts[c]) return CSB_E_INVALID;
  cudaSetDevice(csb_ctx_device(ctx));
#define DCHECK(call)                      \
  do {                                    \
    cudaError_t e_ = (call);              \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)
#define NCHECK(call)                        \
  do {                                      \
    ncclResult_t r_ = (call);               \
    if (r_ != 0) return 20000 + (int)r_;    \
  } while (0)
  DistState *S;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    S = &g_state[ctx];
  }
  const int spr = sets_per_rank, n_sets = world * spr;
  const size_t set_bytes = (size_t)cap * sizeof(csb_sift_point);
  if (!S->comm_stream) DCHECK(cudaStreamCreateWithFlags(&S->comm_stream, cudaStreamNonBlocking));
  if (S->gathered_cap < set_bytes * n_sets) {
    if (S->gathered) cudaFree(S->gathered);
    S->gathered = nullptr; S->gathered_cap = 0;
    DCHECK(cudaMalloc((void **)&S->gathered, set_bytes * n_sets));
    S->gathered_cap = set_bytes * n_sets;
  }
  if (S->counts_cap < (size_t)n_sets + spr) {
    if (S->d_counts) cudaFree(S->d_counts);
    S->d_counts = nullptr; S->counts_cap = 0;
    DCHECK(cudaMalloc((void **)&S->d_counts, sizeof(int) * ((size_t)n_sets + spr)));
    S->counts_cap = (size_t)n_sets + spr;
  }
  cudaStream_t cs = S->comm_stream;
  const double t0 = now_ms();

  // ---- the exchange, queued on the communication stream: counts, then one all-gather per local set index ----------
  // global set g = r * spr + c lives at gathered[c][r]
  auto slot = [&](int g) { return S->gathered + ((size_t)(g % spr) * world + (g / spr)) * set_bytes; };
  std::vector<int> counts(n_sets);
  DCHECK(cudaMemcpyAsync(S->d_counts + n_sets, local_counts, sizeof(int) * spr, cudaMemcpyHostToDevice, cs));
  if (world > 1) NCHECK(n->AllGather(S->d_counts + n_sets, S->d_counts, (size_t)spr, kNcclInt32, (ncclComm_t)comm, cs));
  else DCHECK(cudaMemcpyAsync(S->d_counts, S->d_counts + n_sets, sizeof(int) * spr, cudaMemcpyDeviceToDevice, cs));
  DCHECK(cudaMemcpyAsync(counts.data(), S->d_counts, sizeof(int) * n_sets, cudaMemcpyDeviceToHost, cs));
  for (int c = 0; c < spr; c++) {
    char *mine = slot(rank * spr + c);
    DCHECK(cudaMemcpyAsync(mine, d_local_sifts[c], (size_t)local_counts[c] * sizeof(csb_sift_point), cudaMemcpyDeviceToDevice, cs));
    if (world > 1)   // in place: this rank's block of the receive buffer is its send buffer
      NCHECK(n->AllGather(mine, S->gathered + (size_t)c * world * set_bytes, set_bytes, kNcclInt8, (ncclComm_t)comm, cs));
  }

  // ---- this rank's pairs: flattened index k over (i < j), owner k % world ------------------------------------------
  std::vector<int> loc_i, loc_j, rem_i, rem_j;
  std::vector<unsigned int> loc_id, rem_id;
  long long k = 0;
  for (int i = 0; i < n_sets; i++)
    for (int j = i + 1; j < n_sets; j++, k++) {
      if (k % world != rank) continue;
      const bool both_local = i / spr == rank && j / spr == rank;
      (both_local ? loc_i : rem_i).push_back(i);
      (both_local ? loc_j : rem_j).push_back(j);
      (both_local ? loc_id : rem_id).push_back((unsigned int)k);
    }
  const long long n_pairs_total = k;
  const int n_loc = (int)loc_i.size(), n_rem = (int)rem_i.size(), n_mine = n_loc + n_rem;
  std::vector<float> H((size_t)9 * (n_mine + 1)), H2((size_t)9 * (n_mine + 1));
  std::vector<int> inl(n_mine + 1), nv(n_mine + 1), nf(n_mine + 1);

  // phase 1 (underneath the exchange): pairs whose two sets are local, on the caller's own arrays
  int rc = 0;
  if (n_loc > 0) {
    std::vector<void *> ptrs(n_sets, nullptr);
    std::vector<int> cnts(n_sets, 0);
    for (int c = 0; c < spr; c++) { ptrs[rank * spr + c] = d_local_sifts[c]; cnts[rank * spr + c] = local_counts[c]; }
    rc = csb_allpairs_match_ransac_improve(ctx, n_sets, ptrs.data(), cnts.data(), n_loc, loc_i.data(), loc_j.data(), loc_id.data(),
                                           distance, num_loops, min_score, max_ambiguity, thresh, seed, improve_loops,
                                           improve_thresh, H.data(), inl.data(), nv.data(), improve ? H2.data() : nullptr,
                                           improve ? nf.data() : nullptr);
    if (rc) return rc;
  }
  const double t1 = now_ms();
  DCHECK(cudaStreamSynchronize(cs));                 // the gathered sets (and their counts) are complete
  const double t2 = now_ms();
  // phase 2: everything else, on the gathered copies
  if (n_rem > 0) {
    std::vector<void *> ptrs(n_sets);
    for (int g = 0; g < n_sets; g++) ptrs[g] = slot(g);
    rc = csb_allpairs_match_ransac_improve(ctx, n_sets, ptrs.data(), counts.data(), n_rem, rem_i.data(), rem_j.data(), rem_id.data(),
                                           distance, num_loops, min_score, max_ambiguity, thresh, seed, improve_loops,
                                           improve_thresh, H.data() + 9 * (size_t)n_loc, inl.data() + n_loc, nv.data() + n_loc,
                                           improve ? H2.data() + 9 * (size_t)n_loc : nullptr, improve ? nf.data() + n_loc : nullptr);
    if (rc) return rc;
  }
  const double t3 = now_ms();

  // ---- result exchange: fixed-size records {pair id, H[9], inliers, n_valid, H_improved[9], numfit} ------------------
  const int REC = 22;
  const int max_local = (int)((n_pairs_total + world - 1) / world);
  std::vector<float> send((size_t)REC * max_local, 0.0f), recv((size_t)REC * max_local * world);
  auto put = [&](int at, unsigned int id, int src) {
    float *r = send.data() + (size_t)REC * at;
    memcpy(r, &id, 4);
    memcpy(r + 1, H.data() + 9 * (size_t)src, 36);
    memcpy(r + 10, &inl[src], 4);
    memcpy(r + 11, &nv[src], 4);
    memcpy(r + 12, H2.data() + 9 * (size_t)src, 36);
    memcpy(r + 21, &nf[src], 4);
  };
  for (int a = 0; a < max_local; a++) { const unsigned int none = 0xffffffffu; memcpy(send.data() + (size_t)REC * a, &none, 4); }
  for (int a = 0; a < n_loc; a++) put(a, loc_id[a], a);
  for (int a = 0; a < n_rem; a++) put(n_loc + a, rem_id[a], n_loc + a);
  const size_t rec_bytes = sizeof(float) * REC * max_local;
  if (S->res_cap < rec_bytes * (world + 1)) {
    if (S->d_res) cudaFree(S->d_res);
    S->d_res = nullptr; S->res_cap = 0;
    DCHECK(cudaMalloc((void **)&S->d_res, rec_bytes * (world + 1)));
    S->res_cap = rec_bytes * (world + 1);
  }
  float *d_send = S->d_res + (size_t)REC * max_local * world;
  DCHECK(cudaMemcpyAsync(d_send, send.data(), rec_bytes, cudaMemcpyHostToDevice, cs));
  if (world > 1) NCHECK(n->AllGather(d_send, S->d_res, (size_t)REC * max_local, kNcclFloat32, (ncclComm_t)comm, cs));
  else DCHECK(cudaMemcpyAsync(S->d_res, d_send, rec_bytes, cudaMemcpyDeviceToDevice, cs));
  DCHECK(cudaMemcpyAsync(recv.data(), S->d_res, rec_bytes * world, cudaMemcpyDeviceToHost, cs));
  DCHECK(cudaStreamSynchronize(cs));
  for (size_t a = 0; a < (size_t)max_local * world; a++) {
    const float *r = recv.data() + (size_t)REC * a;
    unsigned int id;
    memcpy(&id, r, 4);
    if (id == 0xffffffffu || id >= (unsigned long long)n_pairs_total) continue;
    memcpy(H_out + 9 * (size_t)id, r + 1, 36);
    memcpy(inliers_out + id, r + 10, 4);
    memcpy(nvalid_out + id, r + 11, 4);
    if (improve) {
      memcpy(H_improved_out + 9 * (size_t)id, r + 12, 36);
      memcpy(numfit_out + id, r + 21, 4);
    }
  }
  if (timings_ms) {
    timings_ms[0] = t1 - t0;      // queueing the exchange + the pairs local to the rank (they run underneath it)
    timings_ms[1] = t2 - t1;      // waiting for the rest of the exchange
    timings_ms[2] = t3 - t2;      // remaining pairs
    timings_ms[3] = now_ms() - t3;   // result exchange
  }
#undef DCHECK
#undef NCHECK
  return 0;
}

}  // extern "C"
