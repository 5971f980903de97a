// C ABI (include/block_b200.h) of the B200 DMRG hot path: context, device arenas, orchestration.
// Compiled by nvcc as host C++ (no kernels here: those are in kernels.cu / gemm_grouped.cuh).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <set>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/block_b200.h"
#include "kernels.h"
#include "plan.hpp"
#include "opbuild.hpp"
#include "guess.hpp"
#include "eig_block_jacobi.cuh"

using namespace b2d;

namespace {

std::string g_create_error;

// cudaMalloc that, when the device is full, first lets the live contexts on this device give memory back (empty arena slabs, spare
// buffers, then cached blocks that are not children of the current block iteration move to pinned host memory) and tries again
bool release_device_memory(size_t bytes);
inline cudaError_t device_malloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    if (release_device_memory(bytes)) e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) cudaGetLastError();
  }
  return e;
}

struct DevBuf {   // growable raw device buffer
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = device_malloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// device copy of a host Schedule: one allocation, typed views per chunk
struct DevSchedule {
  DevBuf buf;
  struct DC { DevBatch s1, s2; };
  std::vector<DC> chunks;
  void clear() { chunks.clear(); }
};

// NCCL through dlopen: the library ships inside the torch wheel (nvidia/nccl/lib/libnccl.so.2), no link-time dependency
struct Nccl {
  struct Id128 { char b[128]; };   // ncclUniqueId (passed by value)
  void* h = nullptr;
  void* comm = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (h) return true;
    const char* env = getenv("B2D_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { err = std::string("cannot dlopen NCCL (set B2D_NCCL_LIB): ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "NCCL symbols missing"; return false; }
    return true;
  }
};

// cuSOLVER through dlopen (ships with the CUDA toolkit of the image): Dsyevd for the LARGE density-matrix sectors.
// The eigen-decomposition is O(sum d^3) once per block iteration (< 0.1 % of one sigma's flops): a plain library call,
// like the reference's dsyev_ (MatrixBLAS.C:396-401); small sectors use the hand-written Jacobi kernel.
struct Cusolver {
  void* h = nullptr;
  void* handle = nullptr;
  int (*Create)(void**) = nullptr;
  int (*Destroy)(void*) = nullptr;
  int (*SetStream)(void*, cudaStream_t) = nullptr;
  int (*Dsyevd_bufferSize)(void*, int, int, int, const double*, int, const double*, int*) = nullptr;
  int (*Dsyevd)(void*, int, int, int, double*, int, double*, double*, int, int*) = nullptr;
  bool load(std::string& err) {
    if (handle) return true;
    if (!h) {
      const char* env = getenv("B2D_CUSOLVER_LIB");
      const char* names[] = {env, "libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so"};
      for (const char* n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
      }
      if (!h) { err = std::string("cannot dlopen cuSOLVER (set B2D_CUSOLVER_LIB): ") + dlerror(); return false; }
      Create = (decltype(Create))dlsym(h, "cusolverDnCreate");
      Destroy = (decltype(Destroy))dlsym(h, "cusolverDnDestroy");
      SetStream = (decltype(SetStream))dlsym(h, "cusolverDnSetStream");
      Dsyevd_bufferSize = (decltype(Dsyevd_bufferSize))dlsym(h, "cusolverDnDsyevd_bufferSize");
      Dsyevd = (decltype(Dsyevd))dlsym(h, "cusolverDnDsyevd");
      if (!Create || !Destroy || !SetStream || !Dsyevd_bufferSize || !Dsyevd) { err = "cuSOLVER symbols missing"; return false; }
    }
    if (Create(&handle) != 0) { err = "cusolverDnCreate failed"; handle = nullptr; return false; }
    return true;
  }
};

}  // namespace

struct b2d_ctx {
  int device = -1;
  bool has_device = false;
  cudaStream_t stream = nullptr;
  cudaStream_t side_streams[B2D_NUM_TILE_CLASSES - 1] = {};   // one per tile class beyond the first (run_schedule)
  cudaEvent_t fork_ev = nullptr, join_ev[B2D_NUM_TILE_CLASSES - 1] = {};
  bool multi_stream = true;
  std::string err;
  AngMom am;

  Side side[2];
  PsiLayout psi;
  bool planned = false;
  double core_energy = 0.0;
  bool hubbard = false;
  int norbs = 0, rank = 0, nranks = 1;
  std::vector<Term> terms_all, terms_mine;
  Schedule sched;
  DevSchedule dsched;
  double flops_all = 0.0;

  // options
  double workspace_mb = 2048.0;
  int max_davidson_iter = 2000;
  int forced_class = -1;
  bool sync_debug = false;

  // operator arena: slabs with bump allocation
  struct Slab { char* p; size_t cap, used; };
  std::vector<Slab> slabs;
  int64_t arena_doubles = 0;

  // operator uploads are batched: host images and block descriptors of the operators added since the last flush
  // (one H2D copy + one pack launch per flush instead of three synchronisations per operator)
  double* pend_pinned = nullptr;          // pinned host staging (cudaMallocHost): the H2D copy of a flush is a true DMA
  struct PinnedBuf {                      // grow-only pinned host buffer for large descriptor arrays (scatter tasks and their tiles)
    void* p = nullptr; size_t cap = 0;
    bool reserve(size_t bytes) {
      if (bytes <= cap) return true;
      if (p) cudaFreeHost(p);
      p = nullptr; cap = 0;
      size_t want = std::max(bytes + bytes / 2, (size_t)1 << 20);
      if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; return false; }
      cap = want;
      return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  } kron_pinned_tasks, kron_pinned_tiles;
  size_t pend_cap = 0, pend_used = 0;     // doubles
  std::vector<BlockDesc> pend_desc;

  // scratch
  DevBuf staging, desc_scratch, work, flat_in, flat_out;
  DevBuf trace_buf;        // B2D_TRACE diagnostic
  DevBuf parts;            // split-K partial copies of the destination wavefunction (run_sigma_schedule)
  int slice_iters = 256;   // pipeline iterations per split-K slice (0: no split)
  bool partition_renorm = false;   // option "partition_renormalisation": several ranks that ALL hold the whole block divide the eigen-decomposition (sectors) and the operator rotation (operators) and all-reduce the results
  int slice_iters_narrow = 0;   // option "slice_iters_narrow": finer slices for the narrow tiles of a sigma block (0: the same slices as its 128 x 128 tiles)
  DevBuf psi_blocks;       // BlockDesc per psi block
  DevBuf diag_tasks, diag_begin, diag_gather, diag_pool, diag_regions;
  DevBuf partials, scalars;   // level-1 partial sums; G / theta / alpha / misc scalars
  double* h_pinned = nullptr; // 64 doubles

  // wavefunction slots
  DevBuf user_pool, dav_pool;
  int nuser = 0, ndav = 0;

  // density / rotation
  DevBuf rho, eig_g, eig_vt, eig_vals, eig_sweeps, sector_desc, rot, gather_desc, gather_rows;
  std::vector<int64_t> rho_off;          // per left sector offset in rho (padded)
  int64_t rho_padded = 0;
  std::vector<std::vector<double>> evals; // per sector ascending, clamped
  std::vector<std::vector<int>> eval_row; // per sector: Jacobi row index of the i-th ascending eigenvalue
  bool have_rho = false;                  // b2d_make_density / b2d_density_upload ran for the current block
  bool have_eig = false;
  std::vector<int> kept;                  // kept columns per sector
  std::vector<int64_t> rot_off;           // per sector offset in rot (padded ld = pad_ld(kept))
  bool have_rot = false;
  // rotated operators
  Side rotated;                           // sectors + ops of the renormalised left block
  std::vector<int> rotated_old;           // old sector index per new sector
  DevBuf rotated_arena;
  bool have_rotated = false;

  // timing / accounting
  int64_t launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool timing_valid = false;
  std::vector<cudaEvent_t> phase_events;
  bool phase_timing = false;
  double last_step_ms[2] = {0, 0};
  double class_ms[2][B2D_NUM_TILE_CLASSES] = {};
  int class_launches[2][B2D_NUM_TILE_CLASSES] = {};

  // enlarged-block operator construction: product StateInfo of side[0] (x) side[1] and the operators built on it
  struct Product {
    double* identity = nullptr;             // factorised mode: identity block (max child sector size) in the arena, the `A` of TensorTrace factors
    int identity_ld = 0;
    std::vector<std::vector<std::pair<size_t, SubBlock>>> pend_subs;   // factorised mode: per operator, (stored block index, factor) in product order
    std::map<int, std::vector<int>> pair_hits;   // deferred scatter: per operator, contributions received so far by each (row piece, column piece)
    // Scatter tasks of one product as a TEMPLATE: every product of an operator whose left factor has the same quantum numbers, orientation
    // and allowed pattern and whose right factor is the same operator has the same tasks up to the left operator's base address and the
    // scale (a complementary operator is a sum of hundreds of such products)
    struct FactorTemplate { size_t blk; SubBlock sb; int64_t a_off; bool identity; double m2, h; };   // alpha = ((scale * sb.alpha) * m2) * h: the reference's order of operations   // factorised operators: the same idea for the factor lists
    struct KronTemplate { std::vector<KronTask> tasks; std::vector<double> m2; std::vector<int> pair; int pair_count = 0; double bytes = 0.0; int left_owner = -1; std::vector<FactorTemplate> factors; };   // task coefficient = (scale * tasks[i].coef) * m2[i]
    std::map<std::array<int64_t, 8>, KronTemplate> kron_templates;
    struct Combo { std::vector<std::pair<const double*, bool>> parts; std::vector<double> ratios; const double* block; };
    std::map<std::array<int64_t, 3>, std::vector<Combo>> combos;       // (first part address, m, n) -> pre-summed blocks, for sharing
    Side side;                              // collected sectors of the enlarged block + the operators built so far
    std::vector<int> lmap, rmap, unc_dims;  // leftUnMapQuanta / rightUnMapQuanta / unCollectedStateInfo->quantaStates
    std::vector<std::vector<int>> old_to_new;
    bool set = false;
  } product;
  DevBuf kron_tasks, kron_tiles;
  bool factorised = false;                  // option "factorised": operators of an enlarged block (renormalised block x one-site dot) are kept as lists of
                                            // scaled sub-blocks of the renormalised operators instead of being materialised (16x less memory, no structural zeros)
  bool presum_identity = true;              // option "presum_identity": a piece pair that receives {child block, identity} is pre-summed into one block (one K segment
                                            // instead of two: ~7 % fewer executed flops at M = 4000 for ~5 GB more memory); off: both stay direct factors
  int64_t combo_doubles = 0;                // pre-summed factor blocks allocated since the last reset
  int64_t nsubs_direct = 0, nsubs_combo = 0, ncombos = 0;
  bool balance_terms = false;               // option "balance_terms": cost-weighted term ownership under a term partition (default: the reference's rule)
  bool opbuild_batch = true;                // b2d_build_enlarged_op defers its scatter tasks: one launch per ROUND for a whole child block
  std::vector<KronTask> pend_kron;          // deferred tasks ...
  std::vector<int> pend_kron_round;         // ... and the round of each: how many earlier tasks hit the same destination piece
  double kron_bytes = 0.0;                  // algorithmic bytes of the scatter tasks planned since b2d_set_product_stateinfo (8 x (|A| + |B| + 2 |dst piece|))
  int64_t kron_ntasks = 0, kron_nproducts = 0;
  int kron_last_rounds = 0;                 // rounds (launches) of the last batched flush
  Side stash[2];                            // children of the big block parked by b2d_stash_product / b2d_stash_side
  bool stash_set[2] = {false, false};
  Integrals integrals;                      // one- / two-electron integrals for the complementary operators (b2d_set_integrals)
  GuessPlan guess;                          // guess-wavefunction transform of the next block iteration (b2d_guess_plan)
  DevBuf guess_image, guess_trial, materialised;

  // Device-side shadow of the reference's scratch files (SpinBlock::store / restore, save_load_block.C:23-108): renormalised
  // blocks stay resident between block iterations, keyed by a token the binding hides in the host copy.  An entry owns one buffer: device
  // memory while the cache is under its budget, pinned host memory beyond (copied back into the arena when the block is used).
  struct CachedBlock {
    Side side;                 // sectors + operators; op.dev holds the OFFSET (in doubles) from the buffer base while cached
    DevBuf dev;                // device copy (cap == 0: spilled)
    double* pinned = nullptr;  // pinned host copy of a spilled entry
    int64_t doubles = 0;
  };
  std::map<uint64_t, CachedBlock> cache;
  uint64_t cache_next_token = 1;
  std::set<uint64_t> cache_in_use;    // entries that are children of the current block iteration (until b2d_reset): never evicted
  int64_t cache_evictions = 0;
  double cache_device_mb = 0.0;       // option "cache_device_mb": device memory the cache may hold before it spills to pinned host memory (<= 0: automatic)
  int64_t cache_device_doubles = 0, cache_hits = 0, cache_puts = 0;
  std::vector<DevBuf> spare_bufs;     // buffers of dropped entries, reused by the next b2d_transform_operators (cudaMalloc of hundreds of MB costs ~10 ms)
  std::map<std::vector<int>, PsiLayout> layouts;   // wavefunction layouts for other target quanta (noise: O.psi sectors)
  DevBuf dm_noise;
  Nccl nccl;
  Cusolver cusolver;
  DevBuf eig_work, eig_info, eig_pairs, eig_tmp;
  bool eig_cusolver = false;   // diagnostic option: cusolverDnDsyevd for the large sectors instead of the block Jacobi kernel
  int eig_block_sweeps = 0;    // sweeps of the last block-Jacobi solve
  bool persistent = false;   // option "persistent": 128 x 128 class as a persistent kernel with a cross-tile pipeline (measured: no gain, see profiles/README.md)
  DevBuf tile_counter;
  int eig_jacobi_max = 64;   // sectors up to this size use the hand-written Jacobi kernel, larger ones cusolverDnDsyevd
};

namespace {

int fail(b2d_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) return fail(ctx, B2D_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)
#define NEED_DEVICE()                                                                                     \
  do {                                                                                                    \
    if (!ctx) return B2D_ERR_ARG;                                                                         \
    if (!ctx->has_device) return fail(ctx, B2D_ERR_NO_DEVICE, "no CUDA device bound to this context (planning-only); there is no CPU fallback"); \
  } while (0)
#define NEED_PLAN()                                                                     \
  do {                                                                                  \
    if (!ctx->planned) return fail(ctx, B2D_ERR_ARG, "call b2d_plan first");            \
  } while (0)

cudaError_t arena_alloc(b2d_ctx* ctx, size_t bytes, double** out) {
  bytes = (bytes + 255) / 256 * 256;
  for (auto& s : ctx->slabs)
    if (s.cap - s.used >= bytes) { *out = (double*)(s.p + s.used); s.used += bytes; return cudaSuccess; }
  size_t cap = std::max(bytes, (size_t)256 << 20);
  char* p = nullptr;
  cudaError_t e = device_malloc((void**)&p, cap);
  if (e != cudaSuccess) {   // fall back to an exact-size slab
    cap = bytes;
    e = device_malloc((void**)&p, cap);
    if (e != cudaSuccess) return e;
  }
  ctx->slabs.push_back({p, cap, bytes});
  *out = (double*)p;
  return cudaSuccess;
}

// arena allocation that also works on a planning-only context (CPU tests of the factorised planner): addresses from a private
// counter that nothing dereferences
cudaError_t arena_alloc_any(b2d_ctx* ctx, size_t bytes, double** out) {
  if (ctx->has_device) return arena_alloc(ctx, bytes, out);
  static uintptr_t fake = (uintptr_t)1 << 44;
  *out = (double*)fake;
  fake += (bytes + 255) / 256 * 256;
  return cudaSuccess;
}

double* user_vec(b2d_ctx* ctx, int slot) { return (double*)ctx->user_pool.p + (int64_t)slot * ctx->psi.Wp; }
double* dav_vec(b2d_ctx* ctx, int slot) { return (double*)ctx->dav_pool.p + (int64_t)slot * ctx->psi.Wp; }

// BlockDesc table of one operator (host)
std::vector<BlockDesc> op_blocks(const Side& s, const OpRec& op) {
  std::vector<BlockDesc> v;
  int64_t ref = 0;
  for (int i = 0; i < s.nq; ++i)
    for (int j = 0; j < s.nq; ++j)
      if (op.allowed[(size_t)i * s.nq + j]) {
        BlockDesc d;
        d.ref_off = ref; d.dev_off = op.off[(size_t)i * s.nq + j];
        d.rows = s.dims[i]; d.cols = s.dims[j]; d.ld = pad_ld(s.dims[j]); d.pad = 0;
        ref += (int64_t)d.rows * d.cols;
        v.push_back(d);
      }
  return v;
}

int upload_desc(b2d_ctx* ctx, DevBuf& buf, const void* host, size_t bytes) {
  if (bytes == 0) return B2D_OK;
  CU(buf.reserve(bytes));
  CU(cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));   // host vectors are temporaries
  return B2D_OK;
}

// Scatter tasks round by round: tasks[round_begin[r] .. round_begin[r + 1]) never overlap in their destinations; the rounds run in stream
// order.  One upload of the tasks, one of the band list (every task cut into bands of KRON_BAND destination rows for load balance).
void begin_timing(b2d_ctx* ctx);
void end_timing(b2d_ctx* ctx);
// tasks_pinned: the array lives in pinned host memory (ctx->kron_pinned_tasks): its upload is one DMA without the driver's staging copy
int run_kron_rounds(b2d_ctx* ctx, const KronTask* tasks, size_t ntasks, const std::vector<int>& round_begin, bool timed = false, bool tasks_pinned = false) {
  if (ntasks == 0) return B2D_OK;
  int rc = B2D_OK;
  if (tasks_pinned) {
    CU(ctx->kron_tasks.reserve(ntasks * sizeof(KronTask)));
    CU(cudaMemcpyAsync(ctx->kron_tasks.p, tasks, ntasks * sizeof(KronTask), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    rc = upload_desc(ctx, ctx->kron_tasks, tasks, ntasks * sizeof(KronTask));
    if (rc) return rc;
  }
  size_t ntiles = 0;
  for (size_t t = 0; t < ntasks; ++t) ntiles += (size_t)(tasks[t].a_rows * tasks[t].b_rows + KRON_BAND - 1) / KRON_BAND;
  std::vector<KronTile> tiles_pageable;
  KronTile* tiles = nullptr;
  const bool tiles_pinned = ctx->kron_pinned_tiles.reserve(std::max<size_t>(ntiles, 1) * sizeof(KronTile));
  if (tiles_pinned) tiles = (KronTile*)ctx->kron_pinned_tiles.p;
  else { tiles_pageable.resize(std::max<size_t>(ntiles, 1)); tiles = tiles_pageable.data(); }
  std::vector<int> tile_begin(round_begin.size(), 0);
  size_t nt = 0;
  for (size_t r = 0; r + 1 < round_begin.size(); ++r) {
    tile_begin[r] = (int)nt;
    for (int t = round_begin[r]; t < round_begin[r + 1]; ++t) {
      const int rows = tasks[t].a_rows * tasks[t].b_rows;
      for (int b = 0; b * KRON_BAND < rows; ++b) tiles[nt++] = KronTile{t, b};
    }
  }
  tile_begin[round_begin.size() - 1] = (int)nt;
  if (tiles_pinned) {
    CU(ctx->kron_tiles.reserve(std::max<size_t>(nt, 1) * sizeof(KronTile)));
    CU(cudaMemcpyAsync(ctx->kron_tiles.p, tiles, nt * sizeof(KronTile), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // the pinned buffers are reused by the next flush
  } else {
    rc = upload_desc(ctx, ctx->kron_tiles, tiles, nt * sizeof(KronTile));
    if (rc) return rc;
  }
  if (timed) begin_timing(ctx);   // b2d_last_timing: device time of the scatter launches alone (descriptors are resident)
  for (size_t r = 0; r + 1 < round_begin.size(); ++r)
    CU(launch_kron_scatter((const KronTask*)ctx->kron_tasks.p, (const KronTile*)ctx->kron_tiles.p + tile_begin[r], tile_begin[r + 1] - tile_begin[r], ctx->stream,
                           &ctx->launches));
  if (timed) end_timing(ctx);
  return B2D_OK;
}

// pack every pending operator image into its padded device blocks (dev_off is absolute: base pointer 0)
int run_kron_rounds(b2d_ctx* ctx, const std::vector<KronTask>& tasks, const std::vector<int>& round_begin, bool timed = false) {
  return run_kron_rounds(ctx, tasks.data(), tasks.size(), round_begin, timed, false);
}

int flush_pending_ops(b2d_ctx* ctx) {
  if (ctx->pend_desc.empty()) { ctx->pend_used = 0; return B2D_OK; }
  CU(cudaSetDevice(ctx->device));
  CU(ctx->staging.reserve(ctx->pend_used * 8));
  CU(cudaMemcpyAsync(ctx->staging.p, ctx->pend_pinned, ctx->pend_used * 8, cudaMemcpyHostToDevice, ctx->stream));
  int rc = upload_desc(ctx, ctx->desc_scratch, ctx->pend_desc.data(), ctx->pend_desc.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_pack((const BlockDesc*)ctx->desc_scratch.p, (int)ctx->pend_desc.size(), (const double*)ctx->staging.p, (double*)nullptr, ctx->stream, &ctx->launches));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->pend_desc.clear();
  ctx->pend_used = 0;
  return B2D_OK;
}

// room for n more doubles in the pinned staging buffer (flushing what is pending first if necessary); nullptr on failure
double* pending_room(b2d_ctx* ctx, size_t n, int* rc) {
  *rc = B2D_OK;
  const size_t chunk = ((size_t)64 << 20) / 8;
  if (ctx->pend_used + n > ctx->pend_cap) {
    *rc = flush_pending_ops(ctx);
    if (*rc) return nullptr;
    if (n > ctx->pend_cap) {
      if (ctx->pend_pinned) cudaFreeHost(ctx->pend_pinned);
      ctx->pend_pinned = nullptr; ctx->pend_cap = 0;
      size_t want = std::max(n, chunk);
      if (cudaMallocHost(&ctx->pend_pinned, want * 8) != cudaSuccess) { *rc = fail(ctx, B2D_ERR_CUDA, "cudaMallocHost failed for the operator staging buffer"); return nullptr; }
      ctx->pend_cap = want;
    }
  }
  double* p = ctx->pend_pinned + ctx->pend_used;
  ctx->pend_used += n;
  return p;
}

int upload_schedule(b2d_ctx* ctx, const Schedule& S, DevSchedule& D) {
  D.clear();
  size_t bytes = 0;
  auto add = [&](size_t n) { size_t o = bytes; bytes += (n + 255) / 256 * 256; return o; };
  struct Off { size_t segs, groups, tiles[B2D_NUM_TILE_CLASSES]; };
  std::vector<Off> o1(S.chunks.size()), o2(S.chunks.size());
  auto plan = [&](const GemmBatch& b, Off& o) {
    o.segs = add(b.segs.size() * sizeof(GSeg));
    o.groups = add(b.groups.size() * sizeof(GGroup));
    for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) o.tiles[c] = add(b.tiles[c].size() * sizeof(GTile));
  };
  for (size_t i = 0; i < S.chunks.size(); ++i) { plan(S.chunks[i].step1, o1[i]); plan(S.chunks[i].step2, o2[i]); }
  if (bytes == 0) return B2D_OK;
  std::vector<char> host(bytes);
  auto fill = [&](const GemmBatch& b, const Off& o) {
    if (!b.segs.empty()) memcpy(&host[o.segs], b.segs.data(), b.segs.size() * sizeof(GSeg));
    if (!b.groups.empty()) memcpy(&host[o.groups], b.groups.data(), b.groups.size() * sizeof(GGroup));
    for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c)
      if (!b.tiles[c].empty()) memcpy(&host[o.tiles[c]], b.tiles[c].data(), b.tiles[c].size() * sizeof(GTile));
  };
  for (size_t i = 0; i < S.chunks.size(); ++i) { fill(S.chunks[i].step1, o1[i]); fill(S.chunks[i].step2, o2[i]); }
  int rc = upload_desc(ctx, D.buf, host.data(), bytes);
  if (rc) return rc;
  char* base = (char*)D.buf.p;
  auto view = [&](const GemmBatch& b, const Off& o) {
    DevBatch d;
    d.segs = (const GSeg*)(base + o.segs);
    d.groups = (const GGroup*)(base + o.groups);
    for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) { d.tiles[c] = (const GTile*)(base + o.tiles[c]); d.ntiles[c] = (int)b.tiles[c].size(); }
    d.unit_alpha = b.unit_alpha;
    return d;
  };
  D.chunks.resize(S.chunks.size());
  for (size_t i = 0; i < S.chunks.size(); ++i) { D.chunks[i].s1 = view(S.chunks[i].step1, o1[i]); D.chunks[i].s2 = view(S.chunks[i].step2, o2[i]); }
  return B2D_OK;
}

// run a two-step schedule: bases for SRC / DST / AUX supplied by the caller, WORK = ctx->work.
// With phase_timing every launch is bracketed by events and the times are summed per (step, tile class).
int run_schedule(b2d_ctx* ctx, const Schedule& S, const DevSchedule& D, double* src, double* dst, double* aux) {
  CU(ctx->work.reserve((size_t)std::max<int64_t>(S.work_max, 16) * 8));
  double* bases[B2D_NUM_BASES] = {nullptr, src, (double*)ctx->work.p, dst, aux};
  if (!ctx->phase_timing) {
    // The tile classes of one step are independent: the first non-empty class stays on the main stream, the others
    // are forked onto side streams so that their CTAs fill the SMs the main launch leaves idle in its tail.
    const char* trace_path = getenv("B2D_TRACE");   // diagnostic: per-launch (stream, class, start, end) timeline
    struct TraceRec { int chunk, step, cls; };
    std::vector<TraceRec> trace;
    int cur_chunk = 0, cur_step = 0;
    const size_t trace_cap = D.chunks.size() * 2 * B2D_NUM_TILE_CLASSES;
    if (trace_path) {   // one {min start, max end} slot per launch, filled by the kernels from %globaltimer
      CU(ctx->trace_buf.reserve(trace_cap * 16));
      std::vector<unsigned long long> init(trace_cap * 2);
      for (size_t i = 0; i < trace_cap; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
      CU(cudaMemcpyAsync(ctx->trace_buf.p, init.data(), trace_cap * 16, cudaMemcpyHostToDevice, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
    auto traced_launch = [&](const DevBatch& b, int c, cudaStream_t st) -> int {
      unsigned long long* slot = nullptr;
      if (trace_path) {
        slot = (unsigned long long*)ctx->trace_buf.p + 2 * trace.size();
        trace.push_back(TraceRec{cur_chunk, cur_step, c});
      }
      int* counter = (ctx->persistent && c == 0 && st == ctx->stream) ? (int*)ctx->tile_counter.p : nullptr;
      CU(launch_gemm_class(b, c, bases, st, &ctx->launches, slot, counter));
      return B2D_OK;
    };
    auto run_batch = [&](const DevBatch& b) -> int {
      int first = -1;
      for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) if (b.ntiles[c] > 0) { first = c; break; }
      if (first < 0) return B2D_OK;
      // the big class goes first and on the high-priority main stream: it takes every SM, the narrower classes (queued
      // behind it on low-priority side streams) are dispatched as SMs drain in its tail
      bool more = false;
      for (int c = first + 1; c < B2D_NUM_TILE_CLASSES; ++c) more = more || b.ntiles[c] > 0;
      if (ctx->multi_stream && more) CU(cudaEventRecord(ctx->fork_ev, ctx->stream));
      // persistent 128 x 128 class: the narrower classes are queued FIRST and the persistent CTAs (which claim tiles dynamically,
      // so a CTA that starts late simply takes fewer tiles) fill the SMs as they free up - nothing is left for a serial tail
      const bool big_last = ctx->persistent && ctx->multi_stream && first == 0 && more;
      if (!big_last) { int rc = traced_launch(b, first, ctx->stream); if (rc) return rc; }
      if (ctx->multi_stream) {
        for (int c = first + 1; c < B2D_NUM_TILE_CLASSES; ++c) {
          if (b.ntiles[c] <= 0) continue;
          cudaStream_t side = ctx->side_streams[c - 1];
          CU(cudaStreamWaitEvent(side, ctx->fork_ev, 0));
          { int rc = traced_launch(b, c, side); if (rc) return rc; }
          CU(cudaEventRecord(ctx->join_ev[c - 1], side));
        }
      }
      if (big_last) { int rc = traced_launch(b, first, ctx->stream); if (rc) return rc; }
      for (int c = first + 1; c < B2D_NUM_TILE_CLASSES; ++c) {
        if (b.ntiles[c] <= 0) continue;
        if (ctx->multi_stream) CU(cudaStreamWaitEvent(ctx->stream, ctx->join_ev[c - 1], 0));
        else CU(launch_gemm_class(b, c, bases, ctx->stream, &ctx->launches));
      }
      return B2D_OK;
    };
    for (size_t i = 0; i < D.chunks.size(); ++i) {
      cur_chunk = (int)i; cur_step = 0;
      if (S.chunks[i].zero_work) CU(cudaMemsetAsync(ctx->work.p, 0, (size_t)S.chunks[i].work * 8, ctx->stream));
      int rc = run_batch(D.chunks[i].s1);
      if (rc) return rc;
      cur_step = 1;
      rc = run_batch(D.chunks[i].s2);
      if (rc) return rc;
      if (ctx->sync_debug) CU(cudaStreamSynchronize(ctx->stream));
    }
    if (trace_path && !trace.empty()) {
      CU(cudaStreamSynchronize(ctx->stream));
      std::vector<unsigned long long> t(trace.size() * 2);
      CU(cudaMemcpy(t.data(), ctx->trace_buf.p, t.size() * 8, cudaMemcpyDeviceToHost));
      FILE* f = fopen(trace_path, "w");
      if (f) {
        fprintf(f, "chunk,step,class,start_ms,end_ms\n");
        for (size_t i = 0; i < trace.size(); ++i)
          fprintf(f, "%d,%d,%d,%.4f,%.4f\n", trace[i].chunk, trace[i].step, trace[i].cls, (double)(t[2 * i] - t[0]) * 1e-6, (double)(t[2 * i + 1] - t[0]) * 1e-6);
        fclose(f);
      }
    }
    return B2D_OK;
  }
  struct Mark { int step, cls; };
  std::vector<Mark> marks;
  size_t nev = 0;
  auto event = [&](size_t i) -> cudaEvent_t {
    while (ctx->phase_events.size() <= i) {
      cudaEvent_t e = nullptr;
      cudaEventCreate(&e);
      ctx->phase_events.push_back(e);
    }
    return ctx->phase_events[i];
  };
  for (size_t i = 0; i < D.chunks.size(); ++i)
    for (int step = 0; step < 2; ++step) {
      if (step == 0 && S.chunks[i].zero_work) CU(cudaMemsetAsync(ctx->work.p, 0, (size_t)S.chunks[i].work * 8, ctx->stream));
      const DevBatch& b = step == 0 ? D.chunks[i].s1 : D.chunks[i].s2;
      for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) {
        if (b.ntiles[c] <= 0) continue;
        CU(cudaEventRecord(event(nev++), ctx->stream));
        CU(launch_gemm_class(b, c, bases, ctx->stream, &ctx->launches, nullptr, (ctx->persistent && c == 0) ? (int*)ctx->tile_counter.p : nullptr));
        CU(cudaEventRecord(event(nev++), ctx->stream));
        marks.push_back(Mark{step, c});
      }
    }
  CU(cudaStreamSynchronize(ctx->stream));
  for (int st = 0; st < 2; ++st)
    for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) { ctx->class_ms[st][c] = 0.0; ctx->class_launches[st][c] = 0; }
  ctx->last_step_ms[0] = ctx->last_step_ms[1] = 0.0;
  for (size_t k = 0; k < marks.size(); ++k) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->phase_events[2 * k], ctx->phase_events[2 * k + 1]);
    ctx->class_ms[marks[k].step][marks[k].cls] += ms;
    ctx->class_launches[marks[k].step][marks[k].cls] += 1;
    ctx->last_step_ms[marks[k].step] += ms;
  }
  return B2D_OK;
}

int allreduce(b2d_ctx* ctx, double* p, int64_t n) {
  if (ctx->nranks <= 1 || !ctx->nccl.comm) return B2D_OK;
  int r = ctx->nccl.AllReduce(p, p, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->nccl.comm, ctx->stream);
  if (r != 0) return fail(ctx, B2D_ERR_NCCL, std::string("ncclAllReduce: ") + (ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(r) : "error"));
  return B2D_OK;
}

// run a build_schedule() schedule: dst += (terms) src, with the split-K partial copies zeroed before and summed after
int run_sigma_schedule(b2d_ctx* ctx, const Schedule& S, const DevSchedule& D, double* src, double* dst) {
  const int nparts = S.nslices - 1;
  const int64_t Wp = ctx->psi.Wp;
  if (nparts > 0) {
    CU(ctx->parts.reserve((size_t)nparts * Wp * 8));
    CU(cudaMemsetAsync(ctx->parts.p, 0, (size_t)nparts * Wp * 8, ctx->stream));
  }
  int rc = run_schedule(ctx, S, D, src, dst, (double*)ctx->parts.p);
  if (rc) return rc;
  if (nparts > 0) CU(launch_sum_parts(dst, (const double*)ctx->parts.p, nparts, Wp, Wp, ctx->stream, &ctx->launches));
  return B2D_OK;
}

int sigma_dev(b2d_ctx* ctx, double* src, double* dst, bool accumulate, bool reduce) {
  if (!accumulate) CU(cudaMemsetAsync(dst, 0, (size_t)ctx->psi.Wp * 8, ctx->stream));
  int rc = run_sigma_schedule(ctx, ctx->sched, ctx->dsched, src, dst);
  if (rc) return rc;
  if (reduce) return allreduce(ctx, dst, ctx->psi.Wp);
  return B2D_OK;
}

void begin_timing(b2d_ctx* ctx) { cudaEventRecord(ctx->ev[0], ctx->stream); ctx->timing_valid = true; }
void end_timing(b2d_ctx* ctx) { cudaEventRecord(ctx->ev[1], ctx->stream); }

}  // namespace

extern "C" {

int b2d_abi_version(void) { return 4; }   // 4: factorised operators, b2d_alloc_ops, block cache (b2d_cache_*); 3: stash / assemble, guess transform

static std::vector<b2d_ctx*> g_live_ctx;   // contexts the memory-pressure handler may ask for memory

extern "C++" {
namespace {
bool release_device_memory(size_t bytes) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  auto enough = [&]() {
    size_t free_b = 0, total_b = 0;
    return cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b >= bytes + ((size_t)256 << 20);
  };
  bool any = false;
  for (b2d_ctx* c : g_live_ctx) {
    if (!c->has_device || c->device != dev) continue;
    cudaStreamSynchronize(c->stream);
    for (size_t i = 0; i < c->slabs.size();)
      if (c->slabs[i].used == 0) { cudaFree(c->slabs[i].p); c->slabs.erase(c->slabs.begin() + i); any = true; }
      else ++i;
    for (DevBuf& b : c->spare_bufs) { if (b.p) any = true; b.release(); }
    c->spare_bufs.clear();
  }
  if (any && enough()) return true;
  for (b2d_ctx* c : g_live_ctx) {
    if (!c->has_device || c->device != dev) continue;
    for (auto& kv : c->cache) {               // oldest first: in a sweep the oldest blocks of the other direction are needed last
      b2d_ctx::CachedBlock& cb = kv.second;
      if (!cb.dev.p || cb.doubles <= 0 || c->cache_in_use.count(kv.first)) continue;
      double* host = nullptr;
      if (cudaMallocHost(&host, (size_t)cb.doubles * 8) != cudaSuccess) { cudaGetLastError(); return any && enough(); }
      if (cudaMemcpy(host, cb.dev.p, (size_t)cb.doubles * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFreeHost(host); cudaGetLastError(); return any && enough(); }
      cb.pinned = host;
      cb.dev.release();
      c->cache_device_doubles -= cb.doubles;
      ++c->cache_evictions;
      any = true;
      if (enough()) return true;
    }
  }
  return any && enough();
}
}   // namespace
}   // extern "C++"

int b2d_create(int device, b2d_ctx** out) {
  if (!out) return B2D_ERR_ARG;
  *out = nullptr;
  std::unique_ptr<b2d_ctx> c(new b2d_ctx());
  b2d_ctx* ctx = nullptr;   // errors before the context exists go to the global slot
  if (device >= 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device >= n)
      return fail(nullptr, B2D_ERR_NO_DEVICE, std::string("CUDA device ") + std::to_string(device) + " not available: " + (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range"));
    CU(cudaSetDevice(device));
    c->device = device;
    c->has_device = true;
    int prio_least = 0, prio_greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
    for (auto& ev : c->ev) CU(cudaEventCreate(&ev));
    for (auto& st : c->side_streams) CU(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_least));
    CU(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    for (auto& ev : c->join_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(gemm_init());
    CU(cudaMallocHost(&c->h_pinned, 64 * sizeof(double)));
    CU(c->partials.reserve((size_t)L1_MAX_BLOCKS * L1_MAX_VECS * 8));
    CU(c->scalars.reserve(8192 * 8));
    CU(c->tile_counter.reserve(256));
  }
  *out = c.release();
  g_live_ctx.push_back(*out);
  return B2D_OK;
}

void b2d_destroy(b2d_ctx* ctx) {
  if (!ctx) return;
  g_live_ctx.erase(std::remove(g_live_ctx.begin(), g_live_ctx.end(), ctx), g_live_ctx.end());
  if (ctx->has_device) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl.comm) ctx->nccl.CommDestroy(ctx->nccl.comm);
    if (ctx->cusolver.handle) ctx->cusolver.Destroy(ctx->cusolver.handle);
    for (auto& s : ctx->slabs) cudaFree(s.p);
    DevBuf* bufs[] = {&ctx->staging, &ctx->desc_scratch, &ctx->work, &ctx->parts, &ctx->trace_buf, &ctx->eig_work, &ctx->eig_info, &ctx->dm_noise, &ctx->flat_in, &ctx->flat_out, &ctx->psi_blocks, &ctx->diag_tasks,
                      &ctx->diag_begin, &ctx->partials, &ctx->scalars, &ctx->user_pool, &ctx->dav_pool, &ctx->rho, &ctx->eig_g, &ctx->eig_vt,
                      &ctx->eig_vals, &ctx->eig_sweeps, &ctx->sector_desc, &ctx->rot, &ctx->gather_desc, &ctx->gather_rows,
                      &ctx->rotated_arena, &ctx->dsched.buf, &ctx->tile_counter, &ctx->diag_gather, &ctx->diag_pool, &ctx->kron_tasks,
                      &ctx->guess_image, &ctx->guess_trial, &ctx->eig_pairs, &ctx->diag_regions, &ctx->materialised, &ctx->eig_tmp, &ctx->kron_tiles};
    for (DevBuf* b : bufs) b->release();
    for (auto& kv : ctx->cache) { kv.second.dev.release(); if (kv.second.pinned) cudaFreeHost(kv.second.pinned); }
    for (DevBuf& b : ctx->spare_bufs) b.release();
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->pend_pinned) cudaFreeHost(ctx->pend_pinned);
    ctx->kron_pinned_tasks.release(); ctx->kron_pinned_tiles.release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->phase_events) cudaEventDestroy(ev);
    for (auto& st : ctx->side_streams) if (st) cudaStreamDestroy(st);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    for (auto& ev : ctx->join_ev) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
  }
  delete ctx;
}

// Forget the block description (both children, their operators, the plan, wavefunction slots, density / rotation state) but
// keep everything that is expensive to create: streams, events, pinned memory, the cuSOLVER / NCCL handles, the operator arena
// slabs and every scratch buffer.  A sweep creates ONE context and resets it between block iterations.
int b2d_reset(b2d_ctx* ctx) {
  if (!ctx) return B2D_ERR_ARG;
  if (ctx->has_device) {
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->dsched.buf.release();
  }
  ctx->side[0] = Side(); ctx->side[1] = Side();
  ctx->psi = PsiLayout();
  ctx->planned = false;
  ctx->terms_all.clear(); ctx->terms_mine.clear();
  ctx->sched = Schedule();
  ctx->dsched = DevSchedule();
  ctx->flops_all = 0.0;
  for (auto& sl : ctx->slabs) sl.used = 0;
  ctx->cache_in_use.clear();
  ctx->arena_doubles = 0;
  ctx->pend_used = 0; ctx->pend_desc.clear();
  ctx->nuser = 0; ctx->ndav = 0;
  ctx->rho_off.clear(); ctx->rho_padded = 0;
  ctx->evals.clear(); ctx->eval_row.clear(); ctx->kept.clear(); ctx->rot_off.clear();
  ctx->have_eig = ctx->have_rot = ctx->have_rotated = false;
  ctx->have_rho = false;
  ctx->rotated = Side(); ctx->rotated_old.clear();
  ctx->layouts.clear();
  ctx->product = b2d_ctx::Product();
  ctx->stash[0] = Side(); ctx->stash[1] = Side(); ctx->stash_set[0] = ctx->stash_set[1] = false;
  ctx->guess = GuessPlan();
  ctx->pend_kron.clear(); ctx->pend_kron_round.clear(); ctx->product.pair_hits.clear();
  ctx->combo_doubles = 0; ctx->nsubs_direct = ctx->nsubs_combo = ctx->ncombos = 0;
  ctx->timing_valid = false;   // (the integrals belong to the whole calculation: b2d_reset keeps them)
  ctx->err.clear();
  return B2D_OK;
}

const char* b2d_last_error(const b2d_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int b2d_set_option(b2d_ctx* ctx, const char* key, double value) {
  if (!ctx || !key) return B2D_ERR_ARG;
  std::string k(key);
  if (k == "workspace_mb") ctx->workspace_mb = value;
  else if (k == "max_davidson_iter") ctx->max_davidson_iter = (int)value;
  else if (k == "tile_class") ctx->forced_class = (int)value;   // -1 auto, 0/1/2: square 128/64/32 tiles everywhere
  else if (k == "sync_debug") ctx->sync_debug = value != 0;
  else if (k == "multi_stream") ctx->multi_stream = value != 0;
  else if (k == "slice_iters") ctx->slice_iters = (int)value;
  else if (k == "slice_iters_narrow") ctx->slice_iters_narrow = (int)value;
  else if (k == "partition_renormalisation") ctx->partition_renorm = value != 0.0;
  else if (k == "eig_jacobi_max") ctx->eig_jacobi_max = (int)value;
  else if (k == "eig_cusolver") ctx->eig_cusolver = value != 0;
  else if (k == "persistent") ctx->persistent = value != 0;
  else if (k == "phase_timing") ctx->phase_timing = value != 0;
  else if (k == "opbuild_batch") ctx->opbuild_batch = value != 0;
  else if (k == "factorised") ctx->factorised = value != 0;
  else if (k == "presum_identity") ctx->presum_identity = value != 0;
  else if (k == "balance_terms") ctx->balance_terms = value != 0;
  else if (k == "cache_device_mb") ctx->cache_device_mb = value;
  else return fail(ctx, B2D_ERR_ARG, "unknown option " + k);
  return B2D_OK;
}

int b2d_set_block(b2d_ctx* ctx, int side, int nq, const int32_t* q, const int32_t* dims, int is_loop, int nsites, const int32_t* sites) {
  if (!ctx || side < 0 || side > 1 || nq <= 0 || !q || !dims) return fail(ctx, B2D_ERR_ARG, "b2d_set_block: bad arguments");
  Side& s = ctx->side[side];
  s = Side();
  s.nq = nq;
  s.q.assign(q, q + 3 * nq);
  s.dims.assign(dims, dims + nq);
  for (int d : s.dims) if (d <= 0) return fail(ctx, B2D_ERR_ARG, "b2d_set_block: sector with no states");
  s.loop = is_loop != 0;
  if (sites) s.sites.assign(sites, sites + nsites);
  ctx->planned = false;
  return B2D_OK;
}

static int add_op_impl(b2d_ctx* ctx, int side, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq, int fermion,
                       const uint8_t* allowed, const double* data, const double* const* blocks, int* op_id);

int b2d_add_op(b2d_ctx* ctx, int side, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq, int fermion,
               const uint8_t* allowed, const double* data, int* op_id) {
  return add_op_impl(ctx, side, optype, norb, orbs, comp, dq, fermion, allowed, data, nullptr, op_id);
}

int b2d_add_op_blocks(b2d_ctx* ctx, int side, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq, int fermion,
                      const uint8_t* allowed, const double* const* blocks, int* op_id) {
  if (!blocks) return fail(ctx, B2D_ERR_ARG, "b2d_add_op_blocks: null block table");
  return add_op_impl(ctx, side, optype, norb, orbs, comp, dq, fermion, allowed, nullptr, blocks, op_id);
}

static int add_op_impl(b2d_ctx* ctx, int side, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq, int fermion,
                       const uint8_t* allowed, const double* data, const double* const* blocks, int* op_id) {
  if (!ctx || side < 0 || side > 1 || norb < 0 || norb > 2 || !dq || !allowed) return fail(ctx, B2D_ERR_ARG, "b2d_add_op: bad arguments");
  Side& s = ctx->side[side];
  if (s.nq == 0) return fail(ctx, B2D_ERR_ARG, "b2d_add_op: call b2d_set_block first");
  OpRec op;
  op.optype = optype; op.norb = norb; op.comp = comp;
  for (int k = 0; k < norb; ++k) op.orbs[k] = orbs[k];
  memcpy(op.dq, dq, sizeof(op.dq));
  op.fermion = fermion != 0;
  op.allowed.assign(allowed, allowed + (size_t)s.nq * s.nq);
  layout_op(s, op);
  if ((data || blocks) && op.packed_size > 0 && op.packed_size <= 256) {   // the one-site dot: its 1 x 1 elements become alphas of factorised operators
    op.host.resize((size_t)op.packed_size);
    if (data) memcpy(op.host.data(), data, (size_t)op.packed_size * 8);
    else {
      int64_t off = 0, k = 0;
      for (int i = 0; i < s.nq; ++i)
        for (int j = 0; j < s.nq; ++j)
          if (op.allowed[(size_t)i * s.nq + j]) { const int64_t n = (int64_t)s.dims[i] * s.dims[j]; memcpy(op.host.data() + off, blocks[k++], (size_t)n * 8); off += n; }
    }
  }
  // data == NULL: the blocks are materialised (zero-filled) by b2d_plan, and only if one of this rank's terms uses
  // the operator - under a term partition a rank never holds the other ranks' operators
  if (ctx->has_device && (data || blocks)) {
    CU(cudaSetDevice(ctx->device));
    if (op.dev_size > 0) {
      CU(arena_alloc(ctx, (size_t)op.dev_size * 8, &op.dev));
      ctx->arena_doubles += op.dev_size;
      CU(cudaMemsetAsync(op.dev, 0, (size_t)op.dev_size * 8, ctx->stream));
      {
        std::vector<BlockDesc> bd = op_blocks(s, op);
        int rc = B2D_OK;
        double* room = pending_room(ctx, (size_t)op.packed_size, &rc);
        if (!room) return rc;
        const int64_t base = room - ctx->pend_pinned;
        const int64_t dev_base = (int64_t)((uintptr_t)op.dev / sizeof(double));
        for (BlockDesc& d : bd) { d.ref_off += base; d.dev_off += dev_base; }
        ctx->pend_desc.insert(ctx->pend_desc.end(), bd.begin(), bd.end());
        if (data) {
          memcpy(room, data, (size_t)op.packed_size * 8);
        } else {   // scatter-gather form: one pointer per allowed block, copied straight from the caller's matrices
          int64_t off = 0, k = 0;
          for (int i = 0; i < s.nq; ++i)
            for (int j = 0; j < s.nq; ++j)
              if (op.allowed[(size_t)i * s.nq + j]) {
                const int64_t n = (int64_t)s.dims[i] * s.dims[j];
                memcpy(room + off, blocks[k++], (size_t)n * 8);
                off += n;
              }
        }
      }
    }
  }
  s.ops.push_back(std::move(op));
  if (op_id) *op_id = (int)s.ops.size() - 1;
  ctx->planned = false;
  return B2D_OK;
}

// Allocate (zero-filled) every operator of a side that was added without data: children of a product block never see b2d_plan, which
// is where data-less operators are otherwise materialised (synthetic benchmark: b2d_fill_op_random fills them afterwards).
int b2d_alloc_ops(b2d_ctx* ctx, int side) {
  NEED_DEVICE();
  if (side < 0 || side > 1) return fail(ctx, B2D_ERR_ARG, "b2d_alloc_ops: bad side");
  CU(cudaSetDevice(ctx->device));
  for (OpRec& op : ctx->side[side].ops) {
    if (op.dev || op.dev_size == 0 || op.factorised) continue;
    CU(arena_alloc(ctx, (size_t)op.dev_size * 8, &op.dev));
    ctx->arena_doubles += op.dev_size;
    CU(cudaMemsetAsync(op.dev, 0, (size_t)op.dev_size * 8, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int64_t b2d_op_size(const b2d_ctx* ctx, int side, int op_id) {
  if (!ctx || side < 0 || side > 1 || op_id < 0 || op_id >= (int)ctx->side[side].ops.size()) return -1;
  return ctx->side[side].ops[op_id].packed_size;
}

static int materialise_factorised(b2d_ctx* ctx, const Side& S, const OpRec& op, DevBuf& buf);
int b2d_download_op(b2d_ctx* ctx, int side, int op_id, double* data) {
  NEED_DEVICE();
  { int frc = flush_pending_ops(ctx); if (frc) return frc; }
  if (side < 0 || side > 1 || op_id < 0 || op_id >= (int)ctx->side[side].ops.size() || !data) return fail(ctx, B2D_ERR_ARG, "b2d_download_op: bad arguments");
  const Side& s = ctx->side[side];
  const OpRec& op = s.ops[op_id];
  if (op.packed_size == 0) return B2D_OK;
  const double* dev = op.dev;
  if (op.factorised) {
    int mrc = materialise_factorised(ctx, s, op, ctx->materialised);
    if (mrc) return mrc;
    dev = (const double*)ctx->materialised.p;
  }
  if (!dev) return fail(ctx, B2D_ERR_ARG, "b2d_download_op: operator is not resident on this rank (no term of this rank uses it)");
  std::vector<BlockDesc> bd = op_blocks(s, op);
  CU(ctx->staging.reserve((size_t)op.packed_size * 8));
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), dev, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(data, ctx->staging.p, (size_t)op.packed_size * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int b2d_fill_op_random(b2d_ctx* ctx, int side, int op_id, uint64_t seed, double amplitude, int symmetric) {
  NEED_DEVICE();
  { int frc = flush_pending_ops(ctx); if (frc) return frc; }
  if (side < 0 || side > 1 || op_id < 0 || op_id >= (int)ctx->side[side].ops.size()) return fail(ctx, B2D_ERR_ARG, "b2d_fill_op_random: bad arguments");
  const Side& s = ctx->side[side];
  const OpRec& op = s.ops[op_id];
  if (op.factorised) return fail(ctx, B2D_ERR_ARG, "b2d_fill_op_random: a factorised operator has no storage of its own");
  if (op.packed_size == 0 || !op.dev) return B2D_OK;   // not resident on this rank: nothing to fill
  std::vector<BlockDesc> bd = op_blocks(s, op);
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_fill_random(op.dev, (const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), seed, amplitude, ctx->stream, &ctx->launches));
  if (symmetric) {
    // self-adjoint in the reduced sense: A[(j,i)] = A[(i,j)]^T / scaling_t(i,j)   (Transposeview::get_scaling, BaseOperator.C:56-91)
    if (op.dq[0] != 0 || op.dq[2] != 0) return fail(ctx, B2D_ERR_ARG, "symmetric fill needs a particle-number conserving, totally symmetric operator");
    std::vector<SymPair> pairs;
    try {
      for (int i = 0; i < s.nq; ++i)
        for (int j = i; j < s.nq; ++j) {
          if (!op.allowed[(size_t)i * s.nq + j]) continue;
          if (!op.allowed[(size_t)j * s.nq + i]) return fail(ctx, B2D_ERR_ARG, "symmetric fill: allowed mask is not symmetric");
          SymPair p;
          p.off_a = op.off[(size_t)i * s.nq + j]; p.off_b = op.off[(size_t)j * s.nq + i];
          p.rows = s.dims[i]; p.cols = s.dims[j]; p.ld_a = pad_ld(s.dims[j]); p.ld_b = pad_ld(s.dims[i]);
          p.f = 1.0 / ctx->am.transpose_scaling(op.dq[1], s.quantum(i)[1], s.quantum(j)[1]);
          pairs.push_back(p);
        }
    } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, e.what()); }
    rc = upload_desc(ctx, ctx->desc_scratch, pairs.data(), pairs.size() * sizeof(SymPair));
    if (rc) return rc;
    CU(launch_symmetrise(op.dev, (const SymPair*)ctx->desc_scratch.p, (int)pairs.size(), ctx->stream, &ctx->launches));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int b2d_plan(b2d_ctx* ctx, const int32_t* psi_dq, double core_energy, int hubbard, int norbs, int rank, int nranks) {
  if (!ctx || !psi_dq || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, B2D_ERR_ARG, "b2d_plan: bad arguments");
  if (ctx->side[0].nq == 0 || ctx->side[1].nq == 0) return fail(ctx, B2D_ERR_ARG, "b2d_plan: set both blocks first");
  if (ctx->has_device) { int frc = flush_pending_ops(ctx); if (frc) return frc; }
  try {
    ctx->core_energy = core_energy; ctx->hubbard = hubbard != 0; ctx->norbs = norbs; ctx->rank = rank; ctx->nranks = nranks;
    int dq[3] = {psi_dq[0], psi_dq[1], psi_dq[2]};
    ctx->psi.build(ctx->side[0], ctx->side[1], dq);
    ctx->terms_all = enumerate_terms(ctx->side[0], ctx->side[1], core_energy, ctx->hubbard, norbs, nranks, ctx->am);
    const int64_t budget0 = (int64_t)(ctx->workspace_mb * 1024.0 * 1024.0 / 8.0);
    ctx->flops_all = -1.0;
    if (nranks > 1 && ctx->balance_terms) {
      // option "balance_terms": replace the reference's static rule (i % n, trimap_2d % n: max / mean flops 1.03 at 8 ranks) by a
      // cost-weighted assignment; the SUM over ranks is the same multiplyH
      Schedule all = build_schedule(ctx->side[0], ctx->side[1], ctx->psi, ctx->terms_all, 0, budget0, ctx->forced_class, ctx->am);
      ctx->flops_all = all.flops_alg;
      balance_owners(ctx->terms_all, all.term_flops, nranks);
    }
    ctx->terms_mine.clear();
    for (const Term& t : ctx->terms_all) if (t.owner == rank) ctx->terms_mine.push_back(t);
    if (ctx->has_device) {   // materialise the device-allocated operators this rank's terms use
      CU(cudaSetDevice(ctx->device));
      std::vector<OpRec*> need;
      size_t need_bytes = 0;
      for (const Term& t : ctx->terms_mine) {
        OpRec* used[2] = {&ctx->side[0].ops[t.lop], &ctx->side[1].ops[t.rop]};
        for (OpRec* op : used) {
          if (op->dev || op->dev_size == 0 || op->pending || op->factorised) continue;
          // room left in a slab that b2d_reset emptied (one context per sweep): reuse it, so that reset-and-plan cycles do not grow
          // the arena; only what does not fit goes into a new exact-size slab below
          const size_t bytes = ((size_t)op->dev_size * 8 + 255) / 256 * 256;
          bool placed = false;
          for (auto& sl : ctx->slabs)
            if (sl.cap - sl.used >= bytes) {
              op->dev = (double*)(sl.p + sl.used);
              sl.used += bytes;
              CU(cudaMemsetAsync(op->dev, 0, bytes, ctx->stream));
              ctx->arena_doubles += op->dev_size;
              placed = true;
              break;
            }
          if (placed) continue;
          op->pending = true;
          need.push_back(op);
          need_bytes += bytes;
        }
      }
      if (need_bytes > 0) {   // one exact-size slab: at benchmark scale the arena is most of the GPU's memory
        char* slab = nullptr;
        CU(device_malloc((void**)&slab, need_bytes));
        ctx->slabs.push_back({slab, need_bytes, need_bytes});
        CU(cudaMemsetAsync(slab, 0, need_bytes, ctx->stream));
        size_t off = 0;
        for (OpRec* op : need) {
          op->dev = (double*)(slab + off);
          op->pending = false;
          off += ((size_t)op->dev_size * 8 + 255) / 256 * 256;
          ctx->arena_doubles += op->dev_size;
        }
      }
    }
    int64_t budget = (int64_t)(ctx->workspace_mb * 1024.0 * 1024.0 / 8.0);
    ctx->sched = build_schedule(ctx->side[0], ctx->side[1], ctx->psi, ctx->terms_mine, 0, budget, ctx->forced_class, ctx->am, ctx->slice_iters, ctx->slice_iters_narrow);
    if (nranks > 1 && ctx->flops_all < 0.0) {
      // algorithmic flops of the whole sigma (all ranks) without keeping the other ranks' schedules
      Schedule all = build_schedule(ctx->side[0], ctx->side[1], ctx->psi, ctx->terms_all, 0, budget, ctx->forced_class, ctx->am);
      ctx->flops_all = all.flops_alg;
    } else if (nranks <= 1) ctx->flops_all = ctx->sched.flops_alg;
  } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, std::string("b2d_plan: ") + e.what()); }
  ctx->planned = true;
  if (getenv("B2D_PLAN_DEBUG")) {   // per step and tile class: tiles, pipeline iterations in total and of the longest tile (its serial length)
    for (int st = 0; st < 2; ++st)
      for (int k = 0; k < B2D_NUM_TILE_CLASSES; ++k) {
        int64_t ntiles = 0, iters = 0, longest = 0, launches = 0, sum_longest = 0;
        for (const Chunk& c : ctx->sched.chunks) {
          const GemmBatch& b = st ? c.step2 : c.step1;
          if (b.tiles[k].empty()) continue;
          ++launches;
          int64_t lmax = 0;
          for (const GTile& t : b.tiles[k]) { const int64_t it = b.groups[t.group].kiters; iters += it; lmax = std::max(lmax, it); }
          ntiles += (int64_t)b.tiles[k].size(); longest = std::max(longest, lmax); sum_longest += lmax;
        }
        if (launches) fprintf(stderr, "B2D_PLAN step %d class %d: launches %lld tiles %lld iterations %lld longest tile %lld sum over launches of the longest %lld\n", st + 1, k,
                              (long long)launches, (long long)ntiles, (long long)iters, (long long)longest, (long long)sum_longest);
      }
  }
  ctx->layouts.clear();
  ctx->have_eig = ctx->have_rot = ctx->have_rotated = false;
  if (ctx->has_device) {
    CU(cudaSetDevice(ctx->device));
    int rc = upload_schedule(ctx, ctx->sched, ctx->dsched);
    if (rc) return rc;
    std::vector<BlockDesc> bd(ctx->psi.nblocks());
    for (int p = 0; p < ctx->psi.nblocks(); ++p) {
      bd[p].ref_off = ctx->psi.ref_off[p]; bd[p].dev_off = ctx->psi.dev_off[p];
      bd[p].rows = ctx->psi.rows[p]; bd[p].cols = ctx->psi.cols[p]; bd[p].ld = ctx->psi.ld[p]; bd[p].pad = 0;
    }
    rc = upload_desc(ctx, ctx->psi_blocks, bd.data(), bd.size() * sizeof(BlockDesc));
    if (rc) return rc;
    CU(ctx->flat_in.reserve((size_t)std::max<int64_t>(ctx->psi.W, 1) * 8));
    CU(ctx->flat_out.reserve((size_t)std::max<int64_t>(ctx->psi.W, 1) * 8));
    // a re-plan changes Wp: drop the slots
    ctx->user_pool.release(); ctx->dav_pool.release(); ctx->nuser = ctx->ndav = 0;
  }
  return B2D_OK;
}

int64_t b2d_psi_size(const b2d_ctx* ctx) { return ctx && ctx->planned ? ctx->psi.W : -1; }
int64_t b2d_psi_padded_size(const b2d_ctx* ctx) { return ctx && ctx->planned ? ctx->psi.Wp : -1; }
int b2d_psi_num_blocks(const b2d_ctx* ctx) { return ctx && ctx->planned ? ctx->psi.nblocks() : -1; }
int b2d_psi_blocks(const b2d_ctx* ctx, int32_t* lq, int32_t* rq, int64_t* offset) {
  if (!ctx || !ctx->planned) return B2D_ERR_ARG;
  for (int p = 0; p < ctx->psi.nblocks(); ++p) {
    if (lq) lq[p] = ctx->psi.bl[p];
    if (rq) rq[p] = ctx->psi.br[p];
    if (offset) offset[p] = ctx->psi.ref_off[p];
  }
  return B2D_OK;
}

int b2d_num_terms(const b2d_ctx* ctx, int all_ranks) {
  if (!ctx || !ctx->planned) return -1;
  return (int)(all_ranks ? ctx->terms_all.size() : ctx->terms_mine.size());
}
int b2d_terms(const b2d_ctx* ctx, int all_ranks, int32_t* left_op, int32_t* right_op, int32_t* flags, double* scale, int32_t* owner) {
  if (!ctx || !ctx->planned) return B2D_ERR_ARG;
  const std::vector<Term>& T = all_ranks ? ctx->terms_all : ctx->terms_mine;
  for (size_t i = 0; i < T.size(); ++i) {
    if (left_op) left_op[i] = T[i].lop;
    if (right_op) right_op[i] = T[i].rop;
    if (flags) flags[i] = (T[i].lt ? 1 : 0) | (T[i].rt ? 2 : 0);
    if (scale) scale[i] = T[i].scale;
    if (owner) owner[i] = T[i].owner;
  }
  return B2D_OK;
}
double b2d_sigma_flops(const b2d_ctx* ctx, int all_ranks) {
  if (!ctx || !ctx->planned) return -1.0;
  return all_ranks ? ctx->flops_all : ctx->sched.flops_alg;
}
int b2d_plan_stats(const b2d_ctx* ctx, double* out, int n) {
  if (!ctx || !ctx->planned || !out) return B2D_ERR_ARG;
  int launches = 0;
  for (const Chunk& c : ctx->sched.chunks)
    for (int k = 0; k < B2D_NUM_TILE_CLASSES; ++k) launches += (c.step1.tiles[k].empty() ? 0 : 1) + (c.step2.tiles[k].empty() ? 0 : 1);
  double useful = 0.0, issued = 0.0;
  for (const Chunk& c : ctx->sched.chunks)
    for (int k = 0; k < B2D_NUM_TILE_CLASSES; ++k) {
      useful += c.step1.class_flops[k] + c.step2.class_flops[k];
      issued += c.step1.class_padded[k] + c.step2.class_padded[k];
    }
  double v[14] = {(double)ctx->sched.chunks.size(), (double)ctx->sched.n_step1, (double)ctx->sched.n_step2, (double)ctx->sched.n_tiles,
                  (double)ctx->sched.work_max, (double)ctx->arena_doubles, (double)launches, ctx->sched.flops_exec, useful, issued,
                  (double)ctx->combo_doubles, (double)ctx->nsubs_direct, (double)ctx->nsubs_combo, (double)ctx->ncombos};
  for (int i = 0; i < n && i < 14; ++i) out[i] = v[i];
  return B2D_OK;
}

// ---- slots --------------------------------------------------------------------------------------------------
int b2d_vec_reserve(b2d_ctx* ctx, int nslots) {
  NEED_DEVICE(); NEED_PLAN();
  if (nslots <= ctx->nuser) return B2D_OK;
  CU(cudaSetDevice(ctx->device));
  size_t bytes = (size_t)nslots * ctx->psi.Wp * 8;
  if (bytes <= ctx->user_pool.cap) {   // capacity left over (b2d_reset keeps the pool): only the new slots need zeroing
    size_t have = (size_t)ctx->nuser * ctx->psi.Wp * 8;
    CU(cudaMemsetAsync((char*)ctx->user_pool.p + have, 0, bytes - have, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->nuser = nslots;
    return B2D_OK;
  }
  DevBuf nb;
  CU(nb.reserve(std::max<size_t>(bytes, 256)));
  CU(cudaMemsetAsync(nb.p, 0, std::max<size_t>(bytes, 256), ctx->stream));
  if (ctx->nuser > 0) CU(cudaMemcpyAsync(nb.p, ctx->user_pool.p, (size_t)ctx->nuser * ctx->psi.Wp * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->user_pool.release();
  ctx->user_pool = nb;
  ctx->nuser = nslots;
  return B2D_OK;
}
#define CHECK_SLOT(s)                                                                                     \
  do {                                                                                                    \
    if ((s) < 0 || (s) >= ctx->nuser) return fail(ctx, B2D_ERR_ARG, "wavefunction slot out of range (b2d_vec_reserve)"); \
  } while (0)

int b2d_vec_upload(b2d_ctx* ctx, int slot, const double* flat) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(slot);
  if (!flat) return fail(ctx, B2D_ERR_ARG, "null buffer");
  CU(cudaMemcpyAsync(ctx->flat_in.p, flat, (size_t)ctx->psi.W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_pack((const BlockDesc*)ctx->psi_blocks.p, ctx->psi.nblocks(), (const double*)ctx->flat_in.p, user_vec(ctx, slot), ctx->stream, &ctx->launches));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}
int b2d_vec_download(b2d_ctx* ctx, int slot, double* flat) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(slot);
  if (!flat) return fail(ctx, B2D_ERR_ARG, "null buffer");
  CU(launch_unpack((const BlockDesc*)ctx->psi_blocks.p, ctx->psi.nblocks(), user_vec(ctx, slot), (double*)ctx->flat_out.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(flat, ctx->flat_out.p, (size_t)ctx->psi.W * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}
int b2d_vec_dot(b2d_ctx* ctx, int a, int b, double* out) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(a); CHECK_SLOT(b);
  VecList x; x.p[0] = user_vec(ctx, a);
  double* sc = (double*)ctx->scalars.p;
  CU(launch_multi_dot(1, x, user_vec(ctx, b), ctx->psi.Wp, (double*)ctx->partials.p, sc, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(ctx->h_pinned, sc, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *out = ctx->h_pinned[0];
  return B2D_OK;
}
int b2d_vec_axpy(b2d_ctx* ctx, double alpha, int x, int y) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(x); CHECK_SLOT(y);
  CU(launch_axpy(user_vec(ctx, y), user_vec(ctx, x), nullptr, alpha, ctx->psi.Wp, ctx->stream, &ctx->launches));
  return B2D_OK;
}
int b2d_vec_scale(b2d_ctx* ctx, double alpha, int x) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(x);
  CU(launch_scale(user_vec(ctx, x), alpha, ctx->psi.Wp, ctx->stream, &ctx->launches));
  return B2D_OK;
}
int b2d_vec_copy(b2d_ctx* ctx, int src, int dst) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(src); CHECK_SLOT(dst);
  if (src != dst) CU(cudaMemcpyAsync(user_vec(ctx, dst), user_vec(ctx, src), (size_t)ctx->psi.Wp * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  return B2D_OK;
}
int b2d_vec_clear(b2d_ctx* ctx, int slot) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(slot);
  CU(cudaMemsetAsync(user_vec(ctx, slot), 0, (size_t)ctx->psi.Wp * 8, ctx->stream));
  return B2D_OK;
}

// ---- sigma ----------------------------------------------------------------------------------------------------
int b2d_sigma(b2d_ctx* ctx, int src_slot, int dst_slot, int accumulate) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(src_slot); CHECK_SLOT(dst_slot);
  if (src_slot == dst_slot) return fail(ctx, B2D_ERR_ARG, "b2d_sigma: src and dst must differ");
  CU(cudaSetDevice(ctx->device));
  begin_timing(ctx);
  int rc;
  if (accumulate && ctx->nranks > 1 && ctx->nccl.comm) {
    // partial sums must be reduced before they are added to v: go through a scratch slot of the Davidson pool
    if (ctx->ndav < 1) { CU(ctx->dav_pool.reserve((size_t)ctx->psi.Wp * 8)); ctx->ndav = 1; }
    rc = sigma_dev(ctx, user_vec(ctx, src_slot), dav_vec(ctx, 0), false, true);
    if (rc) return rc;
    CU(launch_axpy(user_vec(ctx, dst_slot), dav_vec(ctx, 0), nullptr, 1.0, ctx->psi.Wp, ctx->stream, &ctx->launches));
  } else {
    rc = sigma_dev(ctx, user_vec(ctx, src_slot), user_vec(ctx, dst_slot), accumulate != 0, true);
    if (rc) return rc;
  }
  end_timing(ctx);
  return B2D_OK;
}

int b2d_multiplyH_host(b2d_ctx* ctx, const double* c_flat, double* v_flat, int accumulate) {
  NEED_DEVICE(); NEED_PLAN();
  if (!c_flat || !v_flat) return fail(ctx, B2D_ERR_ARG, "null buffer");
  CU(cudaSetDevice(ctx->device));
  if (ctx->ndav < 2) {
    CU(ctx->dav_pool.reserve((size_t)2 * ctx->psi.Wp * 8));
    CU(cudaMemsetAsync(ctx->dav_pool.p, 0, (size_t)2 * ctx->psi.Wp * 8, ctx->stream));
    ctx->ndav = std::max(ctx->ndav, 2);
  }
  double* c = dav_vec(ctx, 0);
  double* v = dav_vec(ctx, 1);
  const BlockDesc* bd = (const BlockDesc*)ctx->psi_blocks.p;
  begin_timing(ctx);
  CU(cudaMemcpyAsync(ctx->flat_in.p, c_flat, (size_t)ctx->psi.W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_pack(bd, ctx->psi.nblocks(), (const double*)ctx->flat_in.p, c, ctx->stream, &ctx->launches));
  int rc = sigma_dev(ctx, c, v, false, true);
  if (rc) return rc;
  CU(launch_unpack(bd, ctx->psi.nblocks(), v, (double*)ctx->flat_out.p, ctx->stream, &ctx->launches));
  if (accumulate) {
    CU(cudaMemcpyAsync(ctx->flat_in.p, v_flat, (size_t)ctx->psi.W * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_axpy((double*)ctx->flat_out.p, (const double*)ctx->flat_in.p, nullptr, 1.0, ctx->psi.W, ctx->stream, &ctx->launches));
  }
  CU(cudaMemcpyAsync(v_flat, ctx->flat_out.p, (size_t)ctx->psi.W * 8, cudaMemcpyDeviceToHost, ctx->stream));
  end_timing(ctx);
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int b2d_tensor_multiply(b2d_ctx* ctx, int left_op, int right_op, int flags, int opq_spin, double scale, int src_slot, int dst_slot) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(src_slot); CHECK_SLOT(dst_slot);
  if (left_op < 0 || right_op < 0 || left_op >= (int)ctx->side[0].ops.size() || right_op >= (int)ctx->side[1].ops.size())
    return fail(ctx, B2D_ERR_ARG, "b2d_tensor_multiply: operator id out of range (the one-operator form is not available yet)");
  if (src_slot == dst_slot) return fail(ctx, B2D_ERR_ARG, "b2d_tensor_multiply: src and dst must differ");
  std::vector<Term> one(1, Term{left_op, right_op, (flags & 1) != 0, (flags & 2) != 0, scale, ctx->rank, TERM_PAIR});
  Schedule S;
  try {
    S = build_schedule(ctx->side[0], ctx->side[1], ctx->psi, one, opq_spin, (int64_t)1 << 60, ctx->forced_class, ctx->am, ctx->slice_iters);
  } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, e.what()); }
  DevSchedule D;
  int rc = upload_schedule(ctx, S, D);
  if (rc) return rc;
  rc = run_sigma_schedule(ctx, S, D, user_vec(ctx, src_slot), user_vec(ctx, dst_slot));
  CU(cudaStreamSynchronize(ctx->stream));
  D.buf.release();
  return rc;
}

int b2d_diagonal(b2d_ctx* ctx, int dst_slot) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(dst_slot);
  CU(cudaSetDevice(ctx->device));
  std::vector<DiagTask> tasks;
  std::vector<int> begin;
  std::vector<BlockDesc> regions;   // the psi blocks, or their (row piece, column piece) sub-blocks when a child is a product of factorised operators
  try {
    build_diag_tasks(ctx->side[0], ctx->side[1], ctx->psi, ctx->terms_mine, ctx->core_energy, ctx->hubbard, ctx->am, tasks, begin, regions);
  } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, e.what()); }
  // every (operator, sector) diagonal is read by ~d_other x #tasks threads with stride ld + 1: gather each distinct one ONCE into
  // a compact pool and point the tasks at it (stride 1: coalesced along j, broadcast along i)
  std::vector<DiagGather> gather;
  {
    std::map<std::array<int64_t, 3>, int64_t> seen;   // (address, stride, length) -> pool offset; the length is part of the key: the identity
    int64_t pool = 0;                                  // block of a product is shared by pieces of every size
    auto remap = [&](int64_t& addr, int32_t& stride, int n) {
      if (!addr) return;
      const std::array<int64_t, 3> key{addr, (int64_t)stride, (int64_t)n};
      auto it = seen.find(key);
      if (it == seen.end()) {
        it = seen.emplace(key, pool).first;
        gather.push_back(DiagGather{addr, pool, n, stride});
        pool += (n + 1) & ~1;
      }
      addr = it->second;   // pool offset for now; turned into an address below
      stride = 1;
    };
    for (size_t p = 0; p < regions.size(); ++p)
      for (int t = begin[p]; t < begin[p + 1]; ++t) {
        remap(tasks[t].a, tasks[t].sa, regions[p].rows);
        remap(tasks[t].b, tasks[t].sb, regions[p].cols);
      }
    CU(ctx->diag_pool.reserve((size_t)std::max<int64_t>(pool, 2) * 8));
    const int64_t base = (int64_t)(intptr_t)ctx->diag_pool.p;
    for (size_t p = 0; p < regions.size(); ++p)
      for (int t = begin[p]; t < begin[p + 1]; ++t) {
        if (tasks[t].sa == 1) tasks[t].a = base + 8 * tasks[t].a;
        if (tasks[t].sb == 1) tasks[t].b = base + 8 * tasks[t].b;
      }
  }
  int rc = upload_desc(ctx, ctx->diag_gather, gather.data(), gather.size() * sizeof(DiagGather));
  if (rc) return rc;
  rc = upload_desc(ctx, ctx->diag_tasks, tasks.data(), tasks.size() * sizeof(DiagTask));
  if (rc) return rc;
  rc = upload_desc(ctx, ctx->diag_begin, begin.data(), begin.size() * sizeof(int));
  if (rc) return rc;
  rc = upload_desc(ctx, ctx->diag_regions, regions.data(), regions.size() * sizeof(BlockDesc));
  if (rc) return rc;
  begin_timing(ctx);
  CU(cudaMemsetAsync(user_vec(ctx, dst_slot), 0, (size_t)ctx->psi.Wp * 8, ctx->stream));
  CU(launch_gather_diag((const DiagGather*)ctx->diag_gather.p, (int)gather.size(), (double*)ctx->diag_pool.p, ctx->stream, &ctx->launches));
  CU(launch_diag((const BlockDesc*)ctx->diag_regions.p, (int)regions.size(), (const DiagTask*)ctx->diag_tasks.p, (const int*)ctx->diag_begin.p,
                 user_vec(ctx, dst_slot), ctx->stream, &ctx->launches));
  rc = allreduce(ctx, user_vec(ctx, dst_slot), ctx->psi.Wp);   // every rank added the terms it owns
  if (rc) return rc;
  end_timing(ctx);
  return B2D_OK;
}

// ---- Davidson ---------------------------------------------------------------------------------------------------
int b2d_davidson(b2d_ctx* ctx, int nroots, int guess_slot0, int diag_slot, double normtol, int deflation_min, int deflation_max,
                 double* evals, int* n_multiply, double* residual) {
  return b2d_davidson_lower(ctx, nroots, guess_slot0, diag_slot, normtol, deflation_min, deflation_max, 0, 0, evals, n_multiply, residual);
}

int b2d_davidson_lower(b2d_ctx* ctx, int nroots, int guess_slot0, int diag_slot, double normtol, int deflation_min, int deflation_max,
                       int n_lower, int lower_slot0, double* evals, int* n_multiply, double* residual) {
  NEED_DEVICE(); NEED_PLAN();
  if (n_lower < 0 || n_lower > 32) return fail(ctx, B2D_ERR_ARG, "b2d_davidson_lower: 0 <= n_lower <= 32");
  for (int i = 0; i < n_lower; ++i) CHECK_SLOT(lower_slot0 + i);
  if (nroots < 1 || deflation_max > 30 || deflation_min < nroots || deflation_max <= deflation_min)
    return fail(ctx, B2D_ERR_ARG, "b2d_davidson: need nroots >= 1 and nroots <= deflation_min < deflation_max <= 30");
  for (int i = 0; i < nroots; ++i) CHECK_SLOT(guess_slot0 + i);
  CHECK_SLOT(diag_slot);
  CU(cudaSetDevice(ctx->device));
  const int64_t n = ctx->psi.Wp;
  const int maxsub = deflation_max + 1;
  const int need = 2 * maxsub + 1;
  if (ctx->ndav < need) {
    CU(ctx->dav_pool.reserve((size_t)need * n * 8));
    ctx->ndav = need;
  }
  // slot pointers: B[i], S[i] (i < maxsub) and the residual scratch R
  std::vector<double*> B(maxsub), Sg(maxsub);
  for (int i = 0; i < maxsub; ++i) { B[i] = dav_vec(ctx, i); Sg[i] = dav_vec(ctx, maxsub + i); }
  double* R = dav_vec(ctx, 2 * maxsub);
  double* diag = user_vec(ctx, diag_slot);
  double* partials = (double*)ctx->partials.p;
  // scalar area: Gt[32*32] | theta[32] | alpha[32*32] | misc[64]
  double* sc = (double*)ctx->scalars.p;
  double* Gt = sc; double* theta = sc + 1024; double* alpha = sc + 1056; double* misc = sc + 2080;
  const int LDG = 32;
  cudaStream_t st = ctx->stream;
  int64_t* L = &ctx->launches;
  begin_timing(ctx);
  // lower states of a state-specific solve (linear.C:201-208, 311-317, 369-375): r <- r - <r|l>/<l|l> l, in the order given
  std::vector<double> inv_ll(n_lower, 0.0);
  if (n_lower > 0) {
    for (int i = 0; i < n_lower; ++i) {
      VecList x; x.p[0] = user_vec(ctx, lower_slot0 + i);
      CU(launch_multi_dot(1, x, user_vec(ctx, lower_slot0 + i), n, partials, misc + 16 + i, st, L));
    }
    CU(cudaMemcpyAsync(ctx->h_pinned, misc + 16, (size_t)n_lower * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < n_lower; ++i) {
      if (!(ctx->h_pinned[i] > 0.0)) return fail(ctx, B2D_ERR_ARG, "b2d_davidson_lower: a lower state has zero norm");
      inv_ll[i] = 1.0 / ctx->h_pinned[i];
    }
  }
  auto project_lower = [&](double* r) -> cudaError_t {
    for (int i = 0; i < n_lower; ++i) {
      VecList x; x.p[0] = user_vec(ctx, lower_slot0 + i);
      cudaError_t e = launch_multi_dot(1, x, r, n, partials, misc + 2, st, L);
      if (e != cudaSuccess) return e;
      e = launch_axpy(r, user_vec(ctx, lower_slot0 + i), misc + 2, -inv_ll[i], n, st, L);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  for (int i = 0; i < nroots; ++i) CU(cudaMemcpyAsync(B[i], user_vec(ctx, guess_slot0 + i), (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  // Gram-Schmidt of the guesses (linear.C:190-198)
  for (int i = 0; i < nroots; ++i) {
    for (int j = 0; j < i; ++j) {
      VecList x; x.p[0] = B[j];
      CU(launch_multi_dot(1, x, B[i], n, partials, misc, st, L));
      CU(launch_axpy(B[i], B[j], misc, -1.0, n, st, L));
    }
    CU(launch_normalise(B[i], n, partials, misc, st, L));
  }
  if (n_lower > 0) {                                                   // linear.C:201-208: only b[0]
    CU(project_lower(B[0]));
    CU(launch_normalise(B[0], n, partials, misc, st, L));
  }
  int nb = nroots, nsig = 0, converged = 0, nmult = 0;
  double rnorm = 0.0;
  int rc = B2D_OK;
  for (int iter = 0;; ++iter) {
    if (iter >= ctx->max_davidson_iter) { rc = fail(ctx, B2D_ERR_NOCONV, "b2d_davidson: iteration cap reached"); break; }
    for (int i = nsig; i < nb; ++i) {                                  // linear.C:234-257
      rc = sigma_dev(ctx, B[i], Sg[i], false, true);
      if (rc) return rc;
      ++nmult;
    }
    nsig = nb;
    // subspace matrix (linear.C:266-270): Gt[j][i] = <b_i|sigma_j>, i >= j
    for (int j = 0; j < nb; ++j) {
      VecList x;
      for (int i = j; i < nb; ++i) x.p[i - j] = B[i];
      CU(launch_multi_dot(nb - j, x, Sg[j], n, partials, Gt + j * LDG + j, st, L));
    }
    CU(launch_subspace_eig(Gt, nb, LDG, theta, alpha, st, L));         // :273
    {                                                                  // Ritz rotation of b and sigma (:279-294)
      VecList xb, xs;
      for (int i = 0; i < nb; ++i) { xb.p[i] = B[i]; xs.p[i] = Sg[i]; }
      CU(launch_rotate(nb, nb, xb, alpha, LDG, n, st, L));
      CU(launch_rotate(nb, nb, xs, alpha, LDG, n, st, L));
    }
    // roots that were converged must still be (:298-307)
    if (converged > 0) {
      for (int i = 0; i < converged; ++i) CU(launch_residual(Sg[i], B[i], theta + i, R, n, partials, misc + 8 + i, st, L));
      CU(cudaMemcpyAsync(ctx->h_pinned, misc + 8, (size_t)converged * 8, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      for (int i = 0; i < converged; ++i)
        if (ctx->h_pinned[i] > normtol) { converged = i; break; }
    }
    CU(launch_residual(Sg[converged], B[converged], theta + converged, R, n, partials, misc + 4, st, L));   // :308-309, :323
    if (n_lower > 0) {                                                 // :311-317, then rnorm = <r|r> of the projected residual
      CU(project_lower(R));
      VecList x; x.p[0] = R;
      CU(launch_multi_dot(1, x, R, n, partials, misc + 4, st, L));
    }
    CU(cudaMemcpyAsync(ctx->h_pinned, misc + 4, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));                                     // the one host read per iteration
    rnorm = ctx->h_pinned[0];
    if (rnorm < normtol) {                                             // :335-348
      ++converged;
      if (converged == nroots) break;
      continue;
    }
    CU(launch_olsen(R, B[converged], diag, theta + converged, n, partials, misc, st, L));                   // :331
    if (nb >= deflation_max) { nb = deflation_min; nsig = deflation_min; }                                  // :352-357
    for (int j = 0; j < nb; ++j) CU(launch_mgs_step(R, B[j], n, partials, misc, st, L));                    // :358-366
    if (n_lower > 0) CU(project_lower(R));                                                                  // :369-375
    CU(launch_normalise(R, n, partials, misc, st, L));
    std::swap(R, B[nb]);
    ++nb;
  }
  {   // also when the iteration cap stopped the solve: the caller gets the current Ritz pairs and B2D_ERR_NOCONV
    CU(cudaMemcpyAsync(ctx->h_pinned, theta, (size_t)nroots * 8, cudaMemcpyDeviceToHost, st));
    for (int i = 0; i < nroots; ++i) CU(cudaMemcpyAsync(user_vec(ctx, guess_slot0 + i), B[i], (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    end_timing(ctx);
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < nroots; ++i) evals[i] = ctx->h_pinned[i];
  }
  if (n_multiply) *n_multiply = nmult;
  if (residual) *residual = rnorm;
  return rc;
}

// ---- renormalisation ----------------------------------------------------------------------------------------------
static void density_layout(b2d_ctx* ctx) {
  const Side& L = ctx->side[0];
  ctx->rho_off.assign(L.nq, 0);
  int64_t off = 0;
  for (int q = 0; q < L.nq; ++q) { ctx->rho_off[q] = off; off += align_up((int64_t)L.dims[q] * pad_ld(L.dims[q]), BLK_ALIGN); }
  ctx->rho_padded = off;
}

int b2d_make_density(b2d_ctx* ctx, int nroots, int slot0, const double* weights) {
  NEED_DEVICE(); NEED_PLAN();
  if (nroots < 1 || !weights) return fail(ctx, B2D_ERR_ARG, "b2d_make_density: bad arguments");
  for (int i = 0; i < nroots; ++i) CHECK_SLOT(slot0 + i);
  CU(cudaSetDevice(ctx->device));
  const Side& L = ctx->side[0];
  const PsiLayout& P = ctx->psi;
  density_layout(ctx);
  CU(ctx->rho.reserve((size_t)ctx->rho_padded * 8));
  // one group per left sector, one segment per (root, rQ): rho[q] += w psi[q,r] psi[q,r]^T   (operatorfunctions.C:640-649)
  Schedule S;
  Chunk ch;
  std::vector<std::vector<GSeg>> per(L.nq);
  for (int i = 0; i < nroots; ++i) {
    if (std::fabs(weights[i]) < 1e-20) continue;    // density.C:86 skips negligible weights
    add_density_segments(P, B2D_BASE_SRC, (int64_t)(slot0 + i) * P.Wp, weights[i], per);
  }
  for (int q = 0; q < L.nq; ++q) {
    GGroup G;
    memset(&G, 0, sizeof(G));
    G.c = ctx->rho_off[q]; G.c_base = B2D_BASE_AUX; G.ldc = pad_ld(L.dims[q]); G.m = G.n = L.dims[q]; G.accumulate = 0;
    G.seg_begin = (int)ch.step2.segs.size();
    ch.step2.segs.insert(ch.step2.segs.end(), per[q].begin(), per[q].end());
    G.seg_end = (int)ch.step2.segs.size();
    ch.step2.groups.push_back(G);
  }
  make_tiles(ch.step2, ctx->forced_class);
  ch.nterms = 1;
  S.chunks.push_back(std::move(ch));
  DevSchedule D;
  int rc = upload_schedule(ctx, S, D);
  if (rc) return rc;
  begin_timing(ctx);
  CU(cudaMemsetAsync(ctx->rho.p, 0, (size_t)ctx->rho_padded * 8, ctx->stream));
  rc = run_schedule(ctx, S, D, (double*)ctx->user_pool.p, nullptr, (double*)ctx->rho.p);
  end_timing(ctx);
  CU(cudaStreamSynchronize(ctx->stream));
  D.buf.release();
  ctx->have_eig = ctx->have_rot = ctx->have_rotated = false;
  ctx->have_rho = rc == B2D_OK;
  return rc;
}

int64_t b2d_density_size(const b2d_ctx* ctx) {
  if (!ctx || ctx->side[0].nq == 0) return -1;
  int64_t n = 0;
  for (int d : ctx->side[0].dims) n += (int64_t)d * d;
  return n;
}

static std::vector<BlockDesc> density_blocks(b2d_ctx* ctx) {
  const Side& L = ctx->side[0];
  std::vector<BlockDesc> bd(L.nq);
  int64_t ref = 0;
  for (int q = 0; q < L.nq; ++q) {
    bd[q].ref_off = ref; bd[q].dev_off = ctx->rho_off[q]; bd[q].rows = bd[q].cols = L.dims[q]; bd[q].ld = pad_ld(L.dims[q]); bd[q].pad = 0;
    ref += (int64_t)L.dims[q] * L.dims[q];
  }
  return bd;
}

int b2d_density_download(b2d_ctx* ctx, double* rho) {
  NEED_DEVICE(); NEED_PLAN();
  if (!ctx->rho.p || !rho) return fail(ctx, B2D_ERR_ARG, "no density matrix yet");
  std::vector<BlockDesc> bd = density_blocks(ctx);
  int64_t n = b2d_density_size(ctx);
  CU(ctx->staging.reserve((size_t)n * 8));
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), (const double*)ctx->rho.p, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(rho, ctx->staging.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int b2d_density_upload(b2d_ctx* ctx, const double* rho) {
  NEED_DEVICE(); NEED_PLAN();
  if (!rho) return fail(ctx, B2D_ERR_ARG, "null buffer");
  density_layout(ctx);
  CU(ctx->rho.reserve((size_t)ctx->rho_padded * 8));
  CU(cudaMemsetAsync(ctx->rho.p, 0, (size_t)ctx->rho_padded * 8, ctx->stream));
  std::vector<BlockDesc> bd = density_blocks(ctx);
  int64_t n = b2d_density_size(ctx);
  CU(ctx->staging.reserve((size_t)n * 8));
  CU(cudaMemcpyAsync(ctx->staging.p, rho, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_pack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), (const double*)ctx->staging.p, (double*)ctx->rho.p, ctx->stream, &ctx->launches));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->have_eig = ctx->have_rot = ctx->have_rotated = false;
  ctx->have_rho = true;
  return B2D_OK;
}

int b2d_diagonalise_dm(b2d_ctx* ctx, double* evals_out) {
  NEED_DEVICE(); NEED_PLAN();
  if (!ctx->rho.p || !ctx->have_rho) return fail(ctx, B2D_ERR_ARG, "no density matrix yet");
  CU(cudaSetDevice(ctx->device));
  const Side& L = ctx->side[0];
  int64_t nev = 0;
  std::vector<BlockDesc> sd(L.nq);
  for (int q = 0; q < L.nq; ++q) {
    sd[q].rows = sd[q].cols = L.dims[q]; sd[q].ld = pad_ld(L.dims[q]); sd[q].dev_off = ctx->rho_off[q]; sd[q].ref_off = nev; sd[q].pad = 0;
    nev += L.dims[q];
  }
  CU(ctx->eig_g.reserve((size_t)ctx->rho_padded * 8));
  CU(ctx->eig_vt.reserve((size_t)ctx->rho_padded * 8));
  CU(ctx->eig_vals.reserve((size_t)nev * 8));
  CU(ctx->eig_sweeps.reserve((size_t)L.nq * 4));
  // small sectors: one CTA each in the Jacobi kernel; large sectors: cusolverDnDsyevd in place on the eigenvector buffer
  std::vector<BlockDesc> small;
  std::vector<int> large;
  for (int q = 0; q < L.nq; ++q) {
    if (L.dims[q] <= ctx->eig_jacobi_max) small.push_back(sd[q]); else large.push_back(q);
  }
  // Several ranks holding the same density matrix (option partition_renormalisation; SURVEY 8e "rho eigensolve: sectors round-robin"): the
  // LARGE sectors are divided by cost (d^3, longest first), rank 0 also takes the small ones, everybody starts from zero-filled results
  // and one all-reduce at the end gives every rank every sector - each value is computed by exactly one rank, so all ranks hold the
  // same bits.
  const bool part = ctx->partition_renorm && ctx->nranks > 1 && ctx->nccl.comm && !ctx->eig_cusolver;
  if (part) {
    std::vector<int> order = large;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return L.dims[a] > L.dims[b]; });
    std::vector<double> load(ctx->nranks, 0.0);
    std::vector<int> mine;
    for (int q : order) {
      int best = 0;
      for (int r = 1; r < ctx->nranks; ++r) if (load[r] < load[best]) best = r;
      load[best] += (double)L.dims[q] * L.dims[q] * L.dims[q];
      if (best == ctx->rank) mine.push_back(q);
    }
    std::sort(mine.begin(), mine.end());
    large.swap(mine);
    if (ctx->rank != 0) small.clear();
    CU(cudaMemsetAsync(ctx->eig_vt.p, 0, (size_t)ctx->rho_padded * 8, ctx->stream));
    CU(cudaMemsetAsync(ctx->eig_vals.p, 0, (size_t)nev * 8, ctx->stream));
  }
  int rc = upload_desc(ctx, ctx->sector_desc, small.data(), small.size() * sizeof(BlockDesc));
  if (rc) return rc;
  begin_timing(ctx);
  CU(cudaMemcpyAsync(ctx->eig_g.p, ctx->rho.p, (size_t)ctx->rho_padded * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  if (!large.empty() && ctx->eig_cusolver)   // Dsyevd works in place: rho_q -> eigenvectors (the Jacobi kernels initialise their own sectors of eig_vt)
    CU(cudaMemcpyAsync(ctx->eig_vt.p, ctx->rho.p, (size_t)ctx->rho_padded * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(launch_sector_eig((const BlockDesc*)ctx->sector_desc.p, (int)small.size(), (double*)ctx->eig_g.p, (double*)ctx->eig_vt.p, (double*)ctx->eig_vals.p,
                       (int*)ctx->eig_sweeps.p, ctx->stream, &ctx->launches));
  if (!large.empty() && !ctx->eig_cusolver) {
    // LARGE sectors: block one-sided Jacobi (eig_block_jacobi.cuh).  One launch = one step of the round-robin tournament over the 32-row
    // blocks of every sector (one CTA per block pair); after each sweep (every block has met every other block of its sector once) the
    // host reads one flag per sector: a sector in which a whole sweep applied no rotation has converged and its CTAs exit at once.
    const int nl = (int)large.size();
    std::vector<BJPair> sect(nl);
    std::vector<int> nbe(nl);   // blocks per sector, rounded up to even (a dummy block sits out)
    int steps = 1;
    for (int k = 0; k < nl; ++k) {
      const int q = large[k], d = L.dims[q];
      BJPair s{};
      s.off = ctx->rho_off[q]; s.d = d; s.ld = pad_ld(d); s.sector = k;
      sect[k] = s;
      const int nb = (d + BJ_B - 1) / BJ_B;
      nbe[k] = nb + (nb & 1);
      steps = std::max(steps, std::max(nbe[k] - 1, 1));
    }
    std::vector<BJPair> pairs;
    std::vector<int> step_begin(steps + 1, 0);
    for (int st = 0; st < steps; ++st) {
      step_begin[st] = (int)pairs.size();
      for (int k = 0; k < nl; ++k) {
        const int d = sect[k].d, nb = (d + BJ_B - 1) / BJ_B, n = nbe[k];
        auto rows = [&](int b) { return std::min(BJ_B, d - b * BJ_B); };
        if (n <= 2) {                                   // one or two blocks: a single pair, every step
          BJPair p = sect[k];
          p.i0 = 0; p.ni = rows(0); p.j0 = BJ_B; p.nj = nb > 1 ? rows(1) : 0;
          pairs.push_back(p);
          continue;
        }
        const int r = st % (n - 1);                     // a sector with fewer blocks than the largest one simply starts its next sweep
        for (int t = 0; t < n / 2; ++t) {
          int a = (r + t) % (n - 1);
          int b = t == 0 ? n - 1 : (r - t + (n - 1)) % (n - 1);
          if (a > b) std::swap(a, b);
          if (a >= nb) continue;
          BJPair p = sect[k];
          p.i0 = a * BJ_B; p.ni = rows(a);
          if (b < nb) { p.j0 = b * BJ_B; p.nj = rows(b); } else { p.j0 = 0; p.nj = 0; }   // partner is the dummy block
          pairs.push_back(p);
        }
      }
    }
    step_begin[steps] = (int)pairs.size();
    rc = upload_desc(ctx, ctx->eig_pairs, pairs.data(), pairs.size() * sizeof(BJPair));
    if (rc) return rc;
    rc = upload_desc(ctx, ctx->eig_work, sect.data(), sect.size() * sizeof(BJPair));
    if (rc) return rc;
    CU(ctx->eig_info.reserve((size_t)2 * nl * sizeof(int)));
    int* d_active = (int*)ctx->eig_info.p;
    int* d_rotated = d_active + nl;
    std::vector<int> flags(2 * nl, 0);
    for (int k = 0; k < nl; ++k) flags[k] = 1;
    CU(cudaMemcpyAsync(d_active, flags.data(), (size_t)2 * nl * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_block_jacobi_init((const BJPair*)ctx->eig_work.p, nl, (double*)ctx->eig_vt.p, ctx->stream, &ctx->launches));
    ctx->eig_block_sweeps = 0;
    for (int sweep = 0; sweep < 40; ++sweep) {
      for (int st = 0; st < steps; ++st)
        CU(launch_block_jacobi_step((const BJPair*)ctx->eig_pairs.p + step_begin[st], step_begin[st + 1] - step_begin[st], (double*)ctx->eig_g.p,
                                    (double*)ctx->eig_vt.p, d_active, d_rotated, 1e-15, ctx->stream, &ctx->launches));
      CU(cudaMemcpyAsync(flags.data(), d_rotated, (size_t)nl * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      ++ctx->eig_block_sweeps;
      bool any = false;
      for (int k = 0; k < nl; ++k) any = any || flags[k] != 0;
      if (!any) break;
      for (int k = 0; k < nl; ++k) flags[nl + k] = 0;   // active <- rotated, rotated <- 0
      CU(cudaMemcpyAsync(d_active, flags.data(), (size_t)2 * nl * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  if (!large.empty() && !ctx->eig_cusolver) {
    // The accumulated rotations leave V orthonormal to ~(rotations per row) x eps ~ 1e-12 for sectors of several hundred states; the
    // reference's dsyev_ eigenvectors are orthonormal to ~d x eps.  One Newton-Schulz step  V <- (3/2 I - 1/2 V V^T) V  squares the defect
    // (1e-12 -> rounding level) with two dense products per sector in the grouped contraction kernel; rows stay eigenvectors to O(defect).
    CU(ctx->eig_tmp.reserve((size_t)ctx->rho_padded * 8));
    CU(cudaMemcpyAsync(ctx->eig_tmp.p, ctx->eig_vt.p, (size_t)ctx->rho_padded * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(launch_scale((double*)ctx->eig_tmp.p, 1.5, ctx->rho_padded, ctx->stream, &ctx->launches));
    Schedule S;
    Chunk ch;
    for (int q : large) {
      const int d = L.dims[q];
      GSeg sg;
      memset(&sg, 0, sizeof(sg));
      sg.a = sg.b = ctx->rho_off[q]; sg.a_base = sg.b_base = B2D_BASE_SRC;
      sg.a_trans = 0; sg.b_kmajor = 1;                      // S = V V^T
      sg.lda = sg.ldb = pad_ld(d); sg.k = d; sg.alpha = 1.0;
      GGroup G;
      memset(&G, 0, sizeof(G));
      G.c = ctx->rho_off[q]; G.c_base = B2D_BASE_AUX; G.ldc = pad_ld(d); G.m = G.n = d; G.accumulate = 0;
      G.seg_begin = (int)ch.step1.segs.size();
      ch.step1.segs.push_back(sg);
      G.seg_end = (int)ch.step1.segs.size();
      ch.step1.groups.push_back(G);
      memset(&sg, 0, sizeof(sg));
      sg.a = ctx->rho_off[q]; sg.a_base = B2D_BASE_AUX; sg.a_trans = 0;   // S (symmetric)
      sg.b = ctx->rho_off[q]; sg.b_base = B2D_BASE_SRC; sg.b_kmajor = 0;   // V
      sg.lda = sg.ldb = pad_ld(d); sg.k = d; sg.alpha = -0.5;
      G.c_base = B2D_BASE_DST; G.accumulate = 1;            // 3/2 V - 1/2 S V
      G.seg_begin = (int)ch.step2.segs.size();
      ch.step2.segs.push_back(sg);
      G.seg_end = (int)ch.step2.segs.size();
      ch.step2.groups.push_back(G);
    }
    make_tiles(ch.step1, ctx->forced_class);
    make_tiles(ch.step2, ctx->forced_class);
    ch.nterms = 1;
    S.chunks.push_back(std::move(ch));
    DevSchedule D;
    rc = upload_schedule(ctx, S, D);
    if (rc) return rc;
    rc = run_schedule(ctx, S, D, (double*)ctx->eig_vt.p, (double*)ctx->eig_tmp.p, (double*)ctx->eig_g.p);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    D.buf.release();
    std::swap(ctx->eig_vt, ctx->eig_tmp);                   // small sectors: 1.5 x their rows sit in the old buffer, so copy them over unscaled
    for (int q = 0; q < L.nq; ++q)
      if (L.dims[q] <= ctx->eig_jacobi_max)   // (partitioned: on ranks other than 0 these are zeros in both buffers)
        CU(cudaMemcpyAsync((double*)ctx->eig_vt.p + ctx->rho_off[q], (const double*)ctx->eig_tmp.p + ctx->rho_off[q],
                           (size_t)L.dims[q] * pad_ld(L.dims[q]) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (getenv("B2D_EIG_DEBUG")) fprintf(stderr, "b2d eig: %d large sectors (max %d states), %d block-Jacobi sweeps\n", (int)large.size(),
                                         L.dims[*std::max_element(large.begin(), large.end(), [&](int a, int b) { return L.dims[a] < L.dims[b]; })], ctx->eig_block_sweeps);
  }
  if (!large.empty() && ctx->eig_cusolver) {   // diagnostic option "eig_cusolver": the library call round 1 used (never the default)
    std::string err;
    if (!ctx->cusolver.load(err)) return fail(ctx, B2D_ERR_CUDA, err);
    if (ctx->cusolver.SetStream(ctx->cusolver.handle, ctx->stream) != 0) return fail(ctx, B2D_ERR_CUDA, "cusolverDnSetStream failed");
    int lwork_max = 0;
    for (int q : large) {
      int lw = 0;
      const int d = L.dims[q];
      if (ctx->cusolver.Dsyevd_bufferSize(ctx->cusolver.handle, 1, 0, d, (const double*)ctx->eig_vt.p + ctx->rho_off[q], pad_ld(d),
                                          (const double*)ctx->eig_vals.p + sd[q].ref_off, &lw) != 0)
        return fail(ctx, B2D_ERR_CUDA, "cusolverDnDsyevd_bufferSize failed");
      lwork_max = std::max(lwork_max, lw);
    }
    CU(ctx->eig_work.reserve((size_t)std::max(lwork_max, 1) * 8));
    CU(ctx->eig_info.reserve(large.size() * sizeof(int)));
    for (size_t k = 0; k < large.size(); ++k) {
      const int q = large[k], d = L.dims[q];
      // rho_q is symmetric, so its row-major block IS a column-major matrix with lda = ld; on exit column k (= our row k) is
      // eigenvector k and W is ascending: exactly the layout the Jacobi kernel leaves in eig_vt / eig_vals
      int st = ctx->cusolver.Dsyevd(ctx->cusolver.handle, /*CUSOLVER_EIG_MODE_VECTOR*/ 1, /*CUBLAS_FILL_MODE_LOWER*/ 0, d,
                                    (double*)ctx->eig_vt.p + ctx->rho_off[q], pad_ld(d), (double*)ctx->eig_vals.p + sd[q].ref_off,
                                    (double*)ctx->eig_work.p, lwork_max, (int*)ctx->eig_info.p + k);
      if (st != 0) return fail(ctx, B2D_ERR_CUDA, "cusolverDnDsyevd failed with status " + std::to_string(st));
      ++ctx->launches;
    }
    std::vector<int> info(large.size());
    CU(cudaMemcpyAsync(info.data(), ctx->eig_info.p, large.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int v : info) if (v != 0) return fail(ctx, B2D_ERR_CUDA, "cusolverDnDsyevd did not converge (info " + std::to_string(v) + ")");
  }
  if (!large.empty()) {
    // Eigenvalues as Rayleigh quotients v_i.(rho v_i) of the final eigenvectors with the ORIGINAL rho: G_q = Vt_q rho_q by the grouped
    // contraction kernel, then one dot product per row - accurate to the rounding of one FP64 product (~1e-17 absolute), which is what the
    // reference's clamp (1e-14) and keep threshold (1e-13, rotationmat.C:161,274) need, and free of the drift G accumulates over the sweeps.
    Schedule S;
    Chunk ch;
    std::vector<BlockDesc> lsd;
    for (int q : large) {
      const int d = L.dims[q];
      GSeg sg;
      memset(&sg, 0, sizeof(sg));
      sg.a = (int64_t)(uintptr_t)((double*)ctx->eig_vt.p + ctx->rho_off[q]);
      sg.b = (int64_t)(uintptr_t)((double*)ctx->rho.p + ctx->rho_off[q]);
      sg.a_base = sg.b_base = B2D_BASE_ABS;
      sg.a_trans = 0; sg.b_kmajor = 0;
      sg.lda = sg.ldb = pad_ld(d);
      sg.k = d;
      sg.alpha = 1.0;
      GGroup G;
      memset(&G, 0, sizeof(G));
      G.c = ctx->rho_off[q]; G.c_base = B2D_BASE_AUX; G.ldc = pad_ld(d); G.m = G.n = d; G.accumulate = 0;
      G.seg_begin = (int)ch.step2.segs.size();
      ch.step2.segs.push_back(sg);
      G.seg_end = (int)ch.step2.segs.size();
      ch.step2.groups.push_back(G);
      lsd.push_back(sd[q]);
    }
    make_tiles(ch.step2, ctx->forced_class);
    ch.nterms = 1;
    S.chunks.push_back(std::move(ch));
    DevSchedule D;
    rc = upload_schedule(ctx, S, D);
    if (rc) return rc;
    rc = run_schedule(ctx, S, D, nullptr, nullptr, (double*)ctx->eig_g.p);
    if (rc) return rc;
    rc = upload_desc(ctx, ctx->sector_desc, lsd.data(), lsd.size() * sizeof(BlockDesc));
    if (rc) return rc;
    CU(launch_rayleigh((const BlockDesc*)ctx->sector_desc.p, (int)lsd.size(), (const double*)ctx->eig_g.p, (const double*)ctx->eig_vt.p,
                       (double*)ctx->eig_vals.p, ctx->stream, &ctx->launches));
    CU(cudaStreamSynchronize(ctx->stream));
    D.buf.release();
  }
  if (part) {
    rc = allreduce(ctx, (double*)ctx->eig_vt.p, ctx->rho_padded);
    if (rc) return rc;
    rc = allreduce(ctx, (double*)ctx->eig_vals.p, nev);
    if (rc) return rc;
  }
  end_timing(ctx);
  std::vector<double> raw(nev);
  CU(cudaMemcpyAsync(raw.data(), ctx->eig_vals.p, (size_t)nev * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  // ascending order per sector (dsyev), clamp < 1e-14 to 0 (rotationmat.C:274-276)
  ctx->evals.assign(L.nq, std::vector<double>());
  ctx->eval_row.assign(L.nq, std::vector<int>());
  int64_t o = 0;
  for (int q = 0; q < L.nq; ++q) {
    int d = L.dims[q];
    std::vector<int> idx(d);
    for (int i = 0; i < d; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return raw[o + a] < raw[o + b]; });
    ctx->eval_row[q] = idx;
    ctx->evals[q].resize(d);
    for (int i = 0; i < d; ++i) { double w = raw[o + idx[i]]; ctx->evals[q][i] = w < 1e-14 ? 0.0 : w; }
    if (evals_out) for (int i = 0; i < d; ++i) evals_out[o + i] = ctx->evals[q][i];
    o += d;
  }
  ctx->have_eig = true;
  ctx->have_rot = ctx->have_rotated = false;
  return B2D_OK;
}

static void rotation_layout(b2d_ctx* ctx) {
  const Side& L = ctx->side[0];
  ctx->rot_off.assign(L.nq, 0);
  int64_t off = 0;
  for (int q = 0; q < L.nq; ++q) { ctx->rot_off[q] = off; off += align_up((int64_t)L.dims[q] * pad_ld(std::max(ctx->kept[q], 1)), BLK_ALIGN); }
  ctx->rot.reserve((size_t)std::max<int64_t>(off, 16) * 8);
}

int b2d_select_states(b2d_ctx* ctx, int keep_states, int32_t* kept_counts, double* discarded) {
  NEED_DEVICE(); NEED_PLAN();
  if (!ctx->have_eig) return fail(ctx, B2D_ERR_ARG, "call b2d_diagonalise_dm first");
  const Side& L = ctx->side[0];
  std::vector<std::vector<int>> kept;
  double err = select_states(ctx->evals, keep_states, kept);
  ctx->kept.assign(L.nq, 0);
  for (int q = 0; q < L.nq; ++q) ctx->kept[q] = (int)kept[q].size();
  rotation_layout(ctx);
  if (!ctx->rot.p) return fail(ctx, B2D_ERR_CUDA, "rotation buffer allocation failed");
  std::vector<GatherDesc> gd;
  std::vector<int> rows;
  for (int q = 0; q < L.nq; ++q) {
    if (kept[q].empty()) continue;
    GatherDesc g;
    g.vt_off = ctx->rho_off[q]; g.u_off = ctx->rot_off[q]; g.d = L.dims[q]; g.ld_vt = pad_ld(L.dims[q]);
    g.ncols = (int)kept[q].size(); g.ld_u = pad_ld(g.ncols); g.row_begin = (int)rows.size(); g.pad = 0;
    for (int s : kept[q]) rows.push_back(ctx->eval_row[q][s]);
    gd.push_back(g);
  }
  CU(cudaMemsetAsync(ctx->rot.p, 0, ctx->rot.cap, ctx->stream));
  int rc = upload_desc(ctx, ctx->gather_desc, gd.data(), gd.size() * sizeof(GatherDesc));
  if (rc) return rc;
  rc = upload_desc(ctx, ctx->gather_rows, rows.data(), rows.size() * sizeof(int));
  if (rc) return rc;
  CU(launch_gather_rotation((const GatherDesc*)ctx->gather_desc.p, (int)gd.size(), (const int*)ctx->gather_rows.p, (const double*)ctx->eig_vt.p,
                            (double*)ctx->rot.p, ctx->stream, &ctx->launches));
  CU(cudaStreamSynchronize(ctx->stream));
  if (kept_counts) for (int q = 0; q < L.nq; ++q) kept_counts[q] = ctx->kept[q];
  if (discarded) *discarded = err;
  ctx->have_rot = true;
  ctx->have_rotated = false;
  return B2D_OK;
}

int64_t b2d_rotation_size(const b2d_ctx* ctx) {
  if (!ctx || !ctx->have_rot) return -1;
  int64_t n = 0;
  for (int q = 0; q < ctx->side[0].nq; ++q) n += (int64_t)ctx->side[0].dims[q] * ctx->kept[q];
  return n;
}

static std::vector<BlockDesc> rotation_blocks(b2d_ctx* ctx) {
  const Side& L = ctx->side[0];
  std::vector<BlockDesc> bd;
  int64_t ref = 0;
  for (int q = 0; q < L.nq; ++q) {
    if (ctx->kept[q] == 0) continue;
    BlockDesc d;
    d.ref_off = ref; d.dev_off = ctx->rot_off[q]; d.rows = L.dims[q]; d.cols = ctx->kept[q]; d.ld = pad_ld(ctx->kept[q]); d.pad = 0;
    ref += (int64_t)d.rows * d.cols;
    bd.push_back(d);
  }
  return bd;
}

int b2d_rotation_download(b2d_ctx* ctx, double* rot) {
  NEED_DEVICE(); NEED_PLAN();
  if (!ctx->have_rot || !rot) return fail(ctx, B2D_ERR_ARG, "no rotation matrices yet");
  std::vector<BlockDesc> bd = rotation_blocks(ctx);
  int64_t n = b2d_rotation_size(ctx);
  if (n == 0) return B2D_OK;
  CU(ctx->staging.reserve((size_t)n * 8));
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), (const double*)ctx->rot.p, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(rot, ctx->staging.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int b2d_rotation_upload(b2d_ctx* ctx, const int32_t* kept_counts, const double* rot) {
  NEED_DEVICE(); NEED_PLAN();
  if (!kept_counts || !rot) return fail(ctx, B2D_ERR_ARG, "null buffer");
  const Side& L = ctx->side[0];
  ctx->kept.assign(kept_counts, kept_counts + L.nq);
  rotation_layout(ctx);
  if (!ctx->rot.p) return fail(ctx, B2D_ERR_CUDA, "rotation buffer allocation failed");
  ctx->have_rot = true;
  CU(cudaMemsetAsync(ctx->rot.p, 0, ctx->rot.cap, ctx->stream));
  std::vector<BlockDesc> bd = rotation_blocks(ctx);
  int64_t n = b2d_rotation_size(ctx);
  if (n > 0) {
    CU(ctx->staging.reserve((size_t)n * 8));
    CU(cudaMemcpyAsync(ctx->staging.p, rot, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
    if (rc) return rc;
    CU(launch_pack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), (const double*)ctx->staging.p, (double*)ctx->rot.p, ctx->stream, &ctx->launches));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->have_rotated = false;
  return B2D_OK;
}

int b2d_transform_operators(b2d_ctx* ctx) {
  NEED_DEVICE(); NEED_PLAN();
  if (!ctx->have_rot) return fail(ctx, B2D_ERR_ARG, "no rotation matrices (b2d_select_states / b2d_rotation_upload)");
  CU(cudaSetDevice(ctx->device));
  const Side& L = ctx->side[0];
  // new StateInfo: sectors that kept >= 1 state, in old order (save_load_block.C:270-283)
  Side& N = ctx->rotated;
  N = Side();
  ctx->rotated_old.clear();
  for (int q = 0; q < L.nq; ++q)
    if (ctx->kept[q] > 0) {
      ctx->rotated_old.push_back(q);
      N.q.insert(N.q.end(), L.q.begin() + 3 * q, L.q.begin() + 3 * q + 3);
      N.dims.push_back(ctx->kept[q]);
    }
  N.nq = (int)N.dims.size();
  N.loop = L.loop; N.sites = L.sites;
  int64_t total = 0;
  for (const OpRec& o : L.ops) {
    OpRec r;
    r.optype = o.optype; r.norb = o.norb; r.orbs[0] = o.orbs[0]; r.orbs[1] = o.orbs[1]; r.comp = o.comp;
    memcpy(r.dq, o.dq, sizeof(r.dq)); r.fermion = o.fermion;
    r.allowed.assign((size_t)N.nq * N.nq, 0);
    for (int a = 0; a < N.nq; ++a)
      for (int b = 0; b < N.nq; ++b)
        r.allowed[(size_t)a * N.nq + b] = o.allowed[(size_t)ctx->rotated_old[a] * L.nq + ctx->rotated_old[b]];
    layout_op(N, r);
    if (o.dev || o.factorised) total += align_up(r.dev_size, 32);   // operators another rank holds are rotated there
    N.ops.push_back(std::move(r));
  }
  {
    const size_t need_bytes = (size_t)std::max<int64_t>(total, 16) * 8;
    if (ctx->rotated_arena.cap < need_bytes) {   // best-fitting buffer of a dropped cache entry before a fresh cudaMalloc
      int best = -1;
      for (size_t i = 0; i < ctx->spare_bufs.size(); ++i)
        if (ctx->spare_bufs[i].cap >= need_bytes && (best < 0 || ctx->spare_bufs[i].cap < ctx->spare_bufs[best].cap)) best = (int)i;
      if (best >= 0) {
        ctx->rotated_arena.release();
        ctx->rotated_arena = ctx->spare_bufs[best];
        ctx->spare_bufs.erase(ctx->spare_bufs.begin() + best);
      }
    }
    CU(ctx->rotated_arena.reserve(need_bytes));
    CU(cudaMemsetAsync(ctx->rotated_arena.p, 0, need_bytes, ctx->stream));
  }
  {
    int64_t off = 0;
    for (size_t m = 0; m < N.ops.size(); ++m) {
      OpRec& r = N.ops[m];
      if (!(L.ops[m].dev || L.ops[m].factorised)) { r.dev = nullptr; continue; }
      r.dev = (double*)ctx->rotated_arena.p + off;
      off += align_up(r.dev_size, 32);
    }
  }
  // schedule: step 1  tmp = O[Q,Q'] U_Q'   step 2  O'[a,b] = U_Q^T tmp     (MatrixRotate, MatrixBLAS.C:553-572)
  Schedule S;
  const int64_t budget = (int64_t)(ctx->workspace_mb * 1024.0 * 1024.0 / 8.0);
  Chunk cur;
  auto close = [&]() {
    if (cur.nterms == 0) return;
    make_tiles(cur.step1, ctx->forced_class);
    make_tiles(cur.step2, ctx->forced_class);
    S.work_max = std::max(S.work_max, cur.work);
    S.chunks.push_back(std::move(cur));
    cur = Chunk();
  };
  // Several ranks that all hold the whole block (option partition_renormalisation; SURVEY 8e "rotation"): the operators are divided by
  // cost (longest first), every rank rotates its share into the zero-filled arena - the same layout everywhere - and one all-reduce
  // completes the block on every rank with identical bits.
  const bool part = ctx->partition_renorm && ctx->nranks > 1 && ctx->nccl.comm;
  std::vector<char> mine_op(L.ops.size(), 1);
  if (part) {
    std::vector<double> cost(L.ops.size(), 0.0);
    for (size_t m = 0; m < L.ops.size(); ++m) {
      const OpRec& r = N.ops[m];
      if (!(L.ops[m].dev || L.ops[m].factorised)) continue;
      for (int a = 0; a < N.nq; ++a)
        for (int b = 0; b < N.nq; ++b)
          if (r.allowed[(size_t)a * N.nq + b]) {
            const double dQ = L.dims[ctx->rotated_old[a]], dQp = L.dims[ctx->rotated_old[b]];
            cost[m] += dQ * (dQp + N.dims[a]) * N.dims[b];
          }
    }
    std::vector<int> order(L.ops.size());
    for (size_t m = 0; m < order.size(); ++m) order[m] = (int)m;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    std::vector<double> load(ctx->nranks, 0.0);
    for (int m : order) {
      int best = 0;
      for (int r = 1; r < ctx->nranks; ++r) if (load[r] < load[best]) best = r;
      load[best] += cost[m];
      mine_op[m] = best == ctx->rank;
    }
  }
  for (size_t m = 0; m < L.ops.size(); ++m) {
    const OpRec& o = L.ops[m];
    const OpRec& r = N.ops[m];
    if (!(o.dev || o.factorised) || !mine_op[m]) continue;
    View ov{&L, &o, false};
    int64_t need = 0;
    for (int a = 0; a < N.nq; ++a)
      for (int b = 0; b < N.nq; ++b)
        if (r.allowed[(size_t)a * N.nq + b]) need += align_up((int64_t)L.dims[ctx->rotated_old[a]] * pad_ld(N.dims[b]), BLK_ALIGN);
    if (cur.nterms > 0 && cur.work + need > budget) close();
    cur.nterms++;
    if (o.factorised) cur.zero_work = true;   // row pieces without a factor stay zero
    std::vector<SubBlock> subs;
    for (int a = 0; a < N.nq; ++a)
      for (int b = 0; b < N.nq; ++b) {
        if (!r.allowed[(size_t)a * N.nq + b]) continue;
        const int Q = ctx->rotated_old[a], Qp = ctx->rotated_old[b];
        const int dQ = L.dims[Q], mQ = N.dims[a], mQp = N.dims[b];
        const int64_t toff = cur.work;
        const int ldt = pad_ld(mQp);
        cur.work += align_up((int64_t)dQ * ldt, BLK_ALIGN);
        // tmp = O[Q,Q'] U_Q': one product for a materialised block; for a factorised one a group per row piece with one K segment per
        // (row piece, column piece) factor, each reading its rows of U_Q'
        subs.clear();
        ov.for_each_sub(Q, Qp, [&](const SubBlock& sb) { subs.push_back(sb); });
        std::stable_sort(subs.begin(), subs.end(), [](const SubBlock& x, const SubBlock& y) { return x.r0 < y.r0; });
        for (size_t k0 = 0; k0 < subs.size();) {
          size_t k1 = k0;
          while (k1 < subs.size() && subs[k1].r0 == subs[k0].r0) ++k1;
          GGroup g1;
          memset(&g1, 0, sizeof(g1));
          g1.c = toff + (int64_t)subs[k0].r0 * ldt; g1.c_base = B2D_BASE_WORK; g1.ldc = ldt; g1.m = subs[k0].m; g1.n = mQp; g1.accumulate = 0;
          g1.seg_begin = (int)cur.step1.segs.size();
          for (size_t k = k0; k < k1; ++k) {
            const SubBlock& sb = subs[k];
            GSeg s1;
            memset(&s1, 0, sizeof(s1));
            s1.a = (int64_t)(intptr_t)sb.a; s1.a_base = B2D_BASE_ABS; s1.a_trans = sb.t ? 1 : 0; s1.lda = sb.lda;
            s1.b = ctx->rot_off[Qp] + (int64_t)sb.c0 * pad_ld(mQp); s1.b_base = B2D_BASE_AUX; s1.b_kmajor = 0; s1.ldb = pad_ld(mQp);
            s1.k = sb.n; s1.alpha = sb.alpha;
            cur.step1.segs.push_back(s1);
          }
          g1.seg_end = (int)cur.step1.segs.size();
          cur.step1.groups.push_back(g1);
          k0 = k1;
        }
        GSeg s2;
        memset(&s2, 0, sizeof(s2));
        s2.a = ctx->rot_off[Q]; s2.a_base = B2D_BASE_AUX; s2.a_trans = 1; s2.lda = pad_ld(mQ);
        s2.b = toff; s2.b_base = B2D_BASE_WORK; s2.b_kmajor = 0; s2.ldb = ldt;
        s2.k = dQ; s2.alpha = 1.0;
        GGroup g2;
        memset(&g2, 0, sizeof(g2));
        g2.c = (r.dev - (double*)ctx->rotated_arena.p) + r.off[(size_t)a * N.nq + b]; g2.c_base = B2D_BASE_DST; g2.ldc = pad_ld(mQp); g2.m = mQ; g2.n = mQp;
        g2.accumulate = 0;
        g2.seg_begin = (int)cur.step2.segs.size(); g2.seg_end = g2.seg_begin + 1;
        cur.step2.segs.push_back(s2); cur.step2.groups.push_back(g2);
      }
  }
  close();
  DevSchedule D;
  int rc = upload_schedule(ctx, S, D);
  if (rc) return rc;
  begin_timing(ctx);
  rc = run_schedule(ctx, S, D, nullptr, (double*)ctx->rotated_arena.p, (double*)ctx->rot.p);
  if (!rc && part) rc = allreduce(ctx, (double*)ctx->rotated_arena.p, total);
  end_timing(ctx);
  CU(cudaStreamSynchronize(ctx->stream));
  D.buf.release();
  if (rc) return rc;
  ctx->have_rotated = true;
  return B2D_OK;
}

int b2d_rotated_num_sectors(const b2d_ctx* ctx) { return ctx && ctx->have_rotated ? ctx->rotated.nq : -1; }
int b2d_rotated_sectors(const b2d_ctx* ctx, int32_t* old_index, int32_t* dims) {
  if (!ctx || !ctx->have_rotated) return B2D_ERR_ARG;
  for (int a = 0; a < ctx->rotated.nq; ++a) {
    if (old_index) old_index[a] = ctx->rotated_old[a];
    if (dims) dims[a] = ctx->rotated.dims[a];
  }
  return B2D_OK;
}
int64_t b2d_rotated_op_size(const b2d_ctx* ctx, int op_id) {
  if (!ctx || !ctx->have_rotated || op_id < 0 || op_id >= (int)ctx->rotated.ops.size()) return -1;
  return ctx->rotated.ops[op_id].packed_size;
}
int b2d_rotated_op_download(b2d_ctx* ctx, int op_id, uint8_t* allowed, double* data) {
  NEED_DEVICE();
  if (!ctx->have_rotated || op_id < 0 || op_id >= (int)ctx->rotated.ops.size()) return fail(ctx, B2D_ERR_ARG, "no rotated operators");
  const Side& N = ctx->rotated;
  const OpRec& op = N.ops[op_id];
  if (allowed) memcpy(allowed, op.allowed.data(), op.allowed.size());
  if (!data || op.packed_size == 0) return B2D_OK;
  if (!op.dev) return fail(ctx, B2D_ERR_ARG, "b2d_rotated_op_download: operator is not resident on this rank");
  std::vector<BlockDesc> bd = op_blocks(N, op);
  CU(ctx->staging.reserve((size_t)op.packed_size * 8));
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), op.dev, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(data, ctx->staging.p, (size_t)op.packed_size * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int64_t b2d_rotated_total_size(const b2d_ctx* ctx) {
  if (!ctx || !ctx->have_rotated) return -1;
  int64_t n = 0;
  for (const OpRec& op : ctx->rotated.ops) n += op.packed_size;
  return n;
}

int b2d_rotated_download_all(b2d_ctx* ctx, double* data) {
  NEED_DEVICE();
  if (!ctx->have_rotated || !data) return fail(ctx, B2D_ERR_ARG, "no rotated operators");
  const Side& N = ctx->rotated;
  std::vector<BlockDesc> all;
  int64_t base = 0;
  for (const OpRec& op : N.ops) {
    if (op.packed_size == 0) continue;
    if (!op.dev) return fail(ctx, B2D_ERR_ARG, "b2d_rotated_download_all: an operator is not resident on this rank");
    std::vector<BlockDesc> bd = op_blocks(N, op);
    const int64_t dev_base = op.dev - (const double*)ctx->rotated_arena.p;
    for (BlockDesc& d : bd) { d.ref_off += base; d.dev_off += dev_base; }
    all.insert(all.end(), bd.begin(), bd.end());
    base += op.packed_size;
  }
  if (base == 0) return B2D_OK;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->staging.reserve((size_t)base * 8));
  int rc = upload_desc(ctx, ctx->desc_scratch, all.data(), all.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)all.size(), (const double*)ctx->rotated_arena.p, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(data, ctx->staging.p, (size_t)base * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

namespace {
const PsiLayout& layout_for(b2d_ctx* ctx, const int* dq) {
  std::vector<int> key(dq, dq + 3);
  auto it = ctx->layouts.find(key);
  if (it == ctx->layouts.end()) {
    PsiLayout P;
    P.build(ctx->side[0], ctx->side[1], dq);
    it = ctx->layouts.emplace(key, std::move(P)).first;
  }
  return it->second;
}

// the operator arrays add_onedot_noise loops over (density.C:360-379), this rank's share
std::vector<int> noise_operators(b2d_ctx* ctx) {
  const Side& L = ctx->side[0];
  bool has_cc = false, has_ddcomp = false;
  for (const OpRec& o : L.ops) { has_cc = has_cc || o.optype == OP_CRE_CRE; has_ddcomp = has_ddcomp || o.optype == OP_DES_DESCOMP; }
  std::vector<int> types = {OP_CRE};
  if (has_cc) { types.push_back(OP_CRE_CRE); types.push_back(OP_CRE_DES); }
  else if (has_ddcomp) { types.push_back(OP_DES_DESCOMP); types.push_back(OP_CRE_DESCOMP); }
  std::vector<int> out;
  for (int ty : types)
    for (size_t m = 0; m < L.ops.size(); ++m) {
      const OpRec& o = L.ops[m];
      if (o.optype != ty || !(o.dev || o.factorised)) continue;
      int owner = 0;
      if (ctx->nranks > 1) owner = o.norb == 1 ? o.orbs[0] % ctx->nranks : trimap_2d(o.orbs[0], o.orbs[1], ctx->norbs) % ctx->nranks;
      if (ctx->nranks > 1 && ctx->balance_terms) owner = (int)(m % (size_t)ctx->nranks);   // any fixed partition of the operators will do: rho_n is all-reduced
      if (owner == ctx->rank) out.push_back((int)m);
    }
  return out;
}
}  // namespace

int64_t b2d_wavefunction_size(b2d_ctx* ctx, const int32_t* dq) {
  if (!ctx || !ctx->planned || !dq) return -1;
  int q[3] = {dq[0], dq[1], dq[2]};
  return layout_for(ctx, q).W;
}

int b2d_tensor_multiply_one_host(b2d_ctx* ctx, int side, int op_id, int transposed, const int32_t* dst_dq, double scale, const double* c_flat,
                                 double* v_flat) {
  NEED_DEVICE(); NEED_PLAN();
  if (side < 0 || side > 1 || op_id < 0 || op_id >= (int)ctx->side[side].ops.size() || !dst_dq || !c_flat || !v_flat)
    return fail(ctx, B2D_ERR_ARG, "b2d_tensor_multiply_one_host: bad arguments");
  const OpRec& op = ctx->side[side].ops[op_id];
  if (!op.resident()) return fail(ctx, B2D_ERR_ARG, "b2d_tensor_multiply_one_host: operator is not resident on this rank");
  CU(cudaSetDevice(ctx->device));
  int q[3] = {dst_dq[0], dst_dq[1], dst_dq[2]};
  const PsiLayout& Pd = layout_for(ctx, q);
  const PsiLayout& Ps = ctx->psi;
  if (Pd.W == 0) return B2D_OK;
  // WORK: [ src (padded) | dst (padded) ], host images staged through flat_in / staging
  Schedule S;
  Chunk ch;
  try {
    add_one_op_groups(ctx->side[0], ctx->side[1], Ps, Pd, side, op, transposed != 0, scale, B2D_BASE_WORK, 0, B2D_BASE_WORK, Ps.Wp, ctx->am, ch.step1);
  } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, e.what()); }
  make_tiles(ch.step1, ctx->forced_class);
  ch.nterms = 1;
  ch.work = Ps.Wp + Pd.Wp;
  S.work_max = ch.work;
  S.chunks.push_back(std::move(ch));
  DevSchedule D;
  int rc = upload_schedule(ctx, S, D);
  if (rc) return rc;
  CU(ctx->work.reserve((size_t)S.work_max * 8));
  double* src = (double*)ctx->work.p;
  double* dst = src + Ps.Wp;
  std::vector<BlockDesc> bd(Pd.nblocks());
  for (int p = 0; p < Pd.nblocks(); ++p) { bd[p].ref_off = Pd.ref_off[p]; bd[p].dev_off = Pd.dev_off[p]; bd[p].rows = Pd.rows[p]; bd[p].cols = Pd.cols[p]; bd[p].ld = Pd.ld[p]; bd[p].pad = 0; }
  rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(ctx->staging.reserve((size_t)Pd.W * 8));
  CU(cudaMemsetAsync(ctx->work.p, 0, (size_t)S.work_max * 8, ctx->stream));
  CU(cudaMemcpyAsync(ctx->flat_in.p, c_flat, (size_t)Ps.W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_pack((const BlockDesc*)ctx->psi_blocks.p, Ps.nblocks(), (const double*)ctx->flat_in.p, src, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(ctx->staging.p, v_flat, (size_t)Pd.W * 8, cudaMemcpyHostToDevice, ctx->stream));      // v += ...
  CU(launch_pack((const BlockDesc*)ctx->desc_scratch.p, Pd.nblocks(), (const double*)ctx->staging.p, dst, ctx->stream, &ctx->launches));
  rc = run_schedule(ctx, S, D, nullptr, nullptr, nullptr);
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, Pd.nblocks(), dst, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(v_flat, ctx->staging.p, (size_t)Pd.W * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  D.buf.release();
  return B2D_OK;
}

// rho += weight * w w^T for a host wavefunction of ANY target quantum dq (MultiplyProduct(w, Transpose(w), dm, weight),
// operatorfunctions.C:630-650): the device half of DensityMatrix::add_twodot_noise (density.C:92-165), whose random
// wavefunctions are drawn on the host (glibc rand(), Wavefunction::Randomise) so that the stream stays the reference's.
int b2d_add_wavefunction_density(b2d_ctx* ctx, const int32_t* dq, const double* flat, double weight) {
  NEED_DEVICE(); NEED_PLAN();
  if (!dq || !flat) return fail(ctx, B2D_ERR_ARG, "b2d_add_wavefunction_density: bad arguments");
  if (!ctx->rho.p || !ctx->have_rho) return fail(ctx, B2D_ERR_ARG, "b2d_add_wavefunction_density: call b2d_make_density first");
  CU(cudaSetDevice(ctx->device));
  const Side& L = ctx->side[0];
  int q[3] = {dq[0], dq[1], dq[2]};
  const PsiLayout& Pd = layout_for(ctx, q);
  if (Pd.W == 0) return B2D_OK;
  Schedule S;
  Chunk ch;
  std::vector<std::vector<GSeg>> per(L.nq);
  try {
    add_density_segments(Pd, B2D_BASE_WORK, 0, weight, per);
  } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, e.what()); }
  for (int sq = 0; sq < L.nq; ++sq) {
    if (per[sq].empty()) continue;
    GGroup G;
    memset(&G, 0, sizeof(G));
    G.c = ctx->rho_off[sq]; G.c_base = B2D_BASE_AUX; G.ldc = pad_ld(L.dims[sq]); G.m = G.n = L.dims[sq]; G.accumulate = 1;
    G.seg_begin = (int)ch.step2.segs.size();
    ch.step2.segs.insert(ch.step2.segs.end(), per[sq].begin(), per[sq].end());
    G.seg_end = (int)ch.step2.segs.size();
    ch.step2.groups.push_back(G);
  }
  make_tiles(ch.step2, ctx->forced_class);
  ch.nterms = 1; ch.work = Pd.Wp;
  S.work_max = Pd.Wp;
  S.chunks.push_back(std::move(ch));
  DevSchedule D;
  int rc = upload_schedule(ctx, S, D);
  if (rc) return rc;
  CU(ctx->work.reserve((size_t)std::max<int64_t>(Pd.Wp, 16) * 8));
  CU(cudaMemsetAsync(ctx->work.p, 0, (size_t)Pd.Wp * 8, ctx->stream));
  std::vector<BlockDesc> bd(Pd.nblocks());
  for (int p = 0; p < Pd.nblocks(); ++p) { bd[p].ref_off = Pd.ref_off[p]; bd[p].dev_off = Pd.dev_off[p]; bd[p].rows = Pd.rows[p]; bd[p].cols = Pd.cols[p]; bd[p].ld = Pd.ld[p]; bd[p].pad = 0; }
  rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(ctx->staging.reserve((size_t)Pd.W * 8));
  CU(cudaMemcpyAsync(ctx->staging.p, flat, (size_t)Pd.W * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_pack((const BlockDesc*)ctx->desc_scratch.p, Pd.nblocks(), (const double*)ctx->staging.p, (double*)ctx->work.p, ctx->stream, &ctx->launches));
  double* bases[B2D_NUM_BASES] = {nullptr, (double*)ctx->user_pool.p, (double*)ctx->work.p, nullptr, (double*)ctx->rho.p};
  CU(launch_gemm_batch(D.chunks[0].s2, bases, ctx->stream, &ctx->launches));
  CU(cudaStreamSynchronize(ctx->stream));
  D.buf.release();
  ctx->have_eig = ctx->have_rot = ctx->have_rotated = false;
  return B2D_OK;
}

int b2d_add_onedot_noise(b2d_ctx* ctx, int nroots, int slot0, double noise) {
  NEED_DEVICE(); NEED_PLAN();
  if (!ctx->rho.p || !ctx->have_rho) return fail(ctx, B2D_ERR_ARG, "b2d_add_onedot_noise: call b2d_make_density first");
  for (int i = 0; i < nroots; ++i) CHECK_SLOT(slot0 + i);
  if (!(noise > 1e-15) || nroots < 1) return B2D_OK;                          // NUMERICAL_ZERO, density.C:40
  if (ctx->hubbard) return B2D_OK;                                             // reference quirk: rho_n is never added for HUBBARD (density.C:358-392)
  CU(cudaSetDevice(ctx->device));
  const Side& L = ctx->side[0];
  const PsiLayout& Ps = ctx->psi;
  const std::vector<int> ops = noise_operators(ctx);
  // work items: (operator, +dQ with the operator / -dQ with its transpose, coupled spin)      density.C:203-237
  struct Item { int op; bool t; int dq[3]; };
  std::vector<Item> items;
  for (int m : ops) {
    const OpRec& o = L.ops[m];
    const int irrep = Ps.dq[2] ^ o.dq[2];
    for (int spin = std::abs(Ps.dq[1] - o.dq[1]); spin <= Ps.dq[1] + o.dq[1]; spin += 2)
      for (int sign = +1; sign >= -1; sign -= 2) {
        Item it{m, sign < 0, {Ps.dq[0] + sign * o.dq[0], spin, irrep}};
        if (layout_for(ctx, it.dq).W > 0) items.push_back(it);
      }
  }
  std::vector<BlockDesc> sectors = density_blocks(ctx);
  int rc = upload_desc(ctx, ctx->sector_desc, sectors.data(), sectors.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(ctx->dm_noise.reserve((size_t)ctx->rho_padded * 8));
  double* partials = (double*)ctx->partials.p;
  double* misc = (double*)ctx->scalars.p + 2080;
  const int64_t budget = (int64_t)(ctx->workspace_mb * 1024.0 * 1024.0 / 8.0);
  begin_timing(ctx);
  for (int root = 0; root < nroots; ++root) {
    CU(cudaMemsetAsync(ctx->dm_noise.p, 0, (size_t)ctx->rho_padded * 8, ctx->stream));
    size_t next = 0;
    while (next < items.size()) {
      // chunk of items whose O.psi vectors fit the workspace
      Schedule S;
      Chunk ch;
      std::vector<std::pair<int64_t, int64_t>> vecs;   // (offset, padded length) of each O.psi in WORK
      std::vector<std::vector<GSeg>> per(L.nq);
      try {
        while (next < items.size()) {
          const Item& it = items[next];
          const PsiLayout& Pd = layout_for(ctx, it.dq);
          if (!vecs.empty() && ch.work + Pd.Wp > budget) break;
          add_one_op_groups(ctx->side[0], ctx->side[1], Ps, Pd, 0, L.ops[it.op], it.t, 1.0, B2D_BASE_SRC, (int64_t)(slot0 + root) * Ps.Wp,
                            B2D_BASE_WORK, ch.work, ctx->am, ch.step1);
          add_density_segments(Pd, B2D_BASE_WORK, ch.work, 1.0, per);        // MultiplyProduct(opxwave, Transpose(opxwave), dm, 1.0)
          vecs.emplace_back(ch.work, Pd.Wp);
          ch.work += Pd.Wp;
          ++next;
        }
      } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, e.what()); }
      for (int q = 0; q < L.nq; ++q) {
        if (per[q].empty()) continue;
        GGroup G;
        memset(&G, 0, sizeof(G));
        G.c = ctx->rho_off[q]; G.c_base = B2D_BASE_AUX; G.ldc = pad_ld(L.dims[q]); G.m = G.n = L.dims[q]; G.accumulate = 1;
        G.seg_begin = (int)ch.step2.segs.size();
        ch.step2.segs.insert(ch.step2.segs.end(), per[q].begin(), per[q].end());
        G.seg_end = (int)ch.step2.segs.size();
        ch.step2.groups.push_back(G);
      }
      make_tiles(ch.step1, ctx->forced_class);
      make_tiles(ch.step2, ctx->forced_class);
      ch.nterms = (int)vecs.size();
      S.work_max = ch.work;
      S.chunks.push_back(std::move(ch));
      DevSchedule D;
      rc = upload_schedule(ctx, S, D);
      if (rc) return rc;
      CU(ctx->work.reserve((size_t)std::max<int64_t>(S.work_max, 16) * 8));
      CU(cudaMemsetAsync(ctx->work.p, 0, (size_t)S.work_max * 8, ctx->stream));
      double* bases[B2D_NUM_BASES] = {nullptr, (double*)ctx->user_pool.p, (double*)ctx->work.p, nullptr, (double*)ctx->dm_noise.p};
      CU(launch_gemm_batch(D.chunks[0].s1, bases, ctx->stream, &ctx->launches));                     // every O.psi of the chunk
      for (const auto& v : vecs)                                                                    // normalise, vanishing ones -> 0 (:220-224)
        CU(launch_normalise_guarded((double*)ctx->work.p + v.first, v.second, 1e-15, partials, misc, ctx->stream, &ctx->launches));
      CU(launch_gemm_batch(D.chunks[0].s2, bases, ctx->stream, &ctx->launches));                     // rho_n += (O psi)(O psi)^T
      CU(cudaStreamSynchronize(ctx->stream));
      D.buf.release();
    }
    rc = allreduce(ctx, (double*)ctx->dm_noise.p, ctx->rho_padded);                                  // distributedaccumulate(dm[0]), density.C:255
    if (rc) return rc;
    CU(launch_trace((const BlockDesc*)ctx->sector_desc.p, L.nq, (const double*)ctx->dm_noise.p, misc, ctx->stream, &ctx->launches));
    CU(cudaMemcpyAsync(ctx->h_pinned, misc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const double norm = ctx->h_pinned[0];
    if (norm > 1.0)                                                                                 // density.C:388-389
      CU(launch_axpy((double*)ctx->rho.p, (const double*)ctx->dm_noise.p, nullptr, (noise / nroots) / norm, ctx->rho_padded, ctx->stream, &ctx->launches));
  }
  end_timing(ctx);
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->have_eig = ctx->have_rot = ctx->have_rotated = false;
  return B2D_OK;
}

int b2d_renormalise_from(b2d_ctx* ctx, int nroots, int guess_slot0, const double* weights, double normtol, int keep_states, int deflation_min,
                         int deflation_max, double noise, double* energies, int32_t* kept_counts, double* discarded, int* n_multiply) {
  NEED_DEVICE(); NEED_PLAN();
  // diag(H) goes to the slot after the guesses (solver.C:34)
  int diag_slot = guess_slot0 + nroots;
  if (ctx->nuser <= diag_slot) {
    int rc = b2d_vec_reserve(ctx, diag_slot + 1);
    if (rc) return rc;
  }
  int rc = b2d_diagonal(ctx, diag_slot);
  if (rc) return rc;
  rc = b2d_davidson(ctx, nroots, guess_slot0, diag_slot, normtol, deflation_min, deflation_max, energies, n_multiply, nullptr);   // solver.C:91
  if (rc) return rc;
  rc = b2d_make_density(ctx, nroots, guess_slot0, weights);                                                                        // renormalise.C:104
  if (rc) return rc;
  rc = b2d_add_onedot_noise(ctx, nroots, guess_slot0, noise);                                                                      // density.C:40-60
  if (rc) return rc;
  rc = b2d_diagonalise_dm(ctx, nullptr);                                                                                           // :113 -> rotationmat.C:258
  if (rc) return rc;
  return b2d_select_states(ctx, keep_states, kept_counts, discarded);
}

// ---- enlarged-block operator construction: TensorProduct / TensorTrace on the device --------------------
int b2d_set_product_stateinfo(b2d_ctx* ctx, int nq, const int32_t* q, const int32_t* dims, int nunc, const int32_t* lmap, const int32_t* rmap,
                              const int32_t* unc_dims, const int32_t* old_to_new_begin, const int32_t* old_to_new) {
  if (!ctx || nq <= 0 || !q || !dims || nunc <= 0 || !lmap || !rmap || !unc_dims || !old_to_new_begin || !old_to_new)
    return fail(ctx, B2D_ERR_ARG, "b2d_set_product_stateinfo: bad arguments");
  ctx->pend_kron.clear(); ctx->pend_kron_round.clear(); ctx->product.pair_hits.clear();   // deferred tasks of a product block nobody stashed
  ctx->kron_bytes = 0.0; ctx->kron_ntasks = ctx->kron_nproducts = 0; ctx->kron_last_rounds = 0;
  const Side& L = ctx->side[0];
  const Side& R = ctx->side[1];
  if (L.nq == 0 || R.nq == 0) return fail(ctx, B2D_ERR_ARG, "b2d_set_product_stateinfo: set both children first (b2d_set_block)");
  b2d_ctx::Product P;
  P.side.nq = nq;
  P.side.q.assign(q, q + 3 * nq);
  P.side.dims.assign(dims, dims + nq);
  P.lmap.assign(lmap, lmap + nunc); P.rmap.assign(rmap, rmap + nunc); P.unc_dims.assign(unc_dims, unc_dims + nunc);
  P.old_to_new.resize(nq);
  for (int c = 0; c < nq; ++c) {
    int sum = 0;
    for (int k = old_to_new_begin[c]; k < old_to_new_begin[c + 1]; ++k) {
      const int u = old_to_new[k];
      if (u < 0 || u >= nunc || lmap[u] < 0 || lmap[u] >= L.nq || rmap[u] < 0 || rmap[u] >= R.nq) return fail(ctx, B2D_ERR_ARG, "b2d_set_product_stateinfo: index out of range");
      if (unc_dims[u] != L.dims[lmap[u]] * R.dims[rmap[u]]) return fail(ctx, B2D_ERR_ARG, "b2d_set_product_stateinfo: uncollected sector size is not d_left * d_right");
      if (!qn_allow(&P.side.q[3 * c], L.quantum(lmap[u]), R.quantum(rmap[u]))) return fail(ctx, B2D_ERR_ARG, "b2d_set_product_stateinfo: piece does not couple to its sector");
      P.old_to_new[c].push_back(u);
      sum += unc_dims[u];
    }
    if (sum != dims[c]) return fail(ctx, B2D_ERR_ARG, "b2d_set_product_stateinfo: pieces do not add up to the sector size");
  }
  if (!ctx->has_device)   // planning-only context: distinct (never dereferenced) addresses per child operator, so that the factor bookkeeping of
    for (int sd = 0; sd < 2; ++sd)   // factorised operators (sharing of pre-summed blocks, memory accounting) is the device run's
      for (OpRec& op : ctx->side[sd].ops)
        if (!op.dev && op.dev_size > 0) { arena_alloc_any(ctx, (size_t)op.dev_size * 8, &op.dev); ctx->arena_doubles += op.dev_size; }
  // un-collected pieces of every collected sector: (offset, size) in concatenation order
  P.side.pieces.resize(nq);
  for (int c = 0; c < nq; ++c) {
    int off = 0;
    for (int u : P.old_to_new[c]) { P.side.pieces[c].emplace_back(off, unc_dims[u]); off += unc_dims[u]; }
  }
  P.set = true;
  ctx->product = std::move(P);
  return B2D_OK;
}

// factorised form is possible when the right child is a one-site dot: every sector holds ONE state, so a product a (x) b is the
// child operator's block times a number
static bool product_is_factorisable(const b2d_ctx* ctx) {
  if (!ctx->factorised || !ctx->product.set) return false;
  for (int d : ctx->side[1].dims) if (d != 1) return false;
  return true;
}

int b2d_product_op_create(b2d_ctx* ctx, const int32_t* dq, int fermion, int* prod_id) {
  if (!ctx) return B2D_ERR_ARG;
  if (!ctx->has_device && !product_is_factorisable(ctx)) NEED_DEVICE();   // a factorised operator is host bookkeeping: planning-only contexts can build it
  if (!ctx->product.set || !dq || !prod_id) return fail(ctx, B2D_ERR_ARG, "b2d_product_op_create: call b2d_set_product_stateinfo first");
  if (ctx->has_device) CU(cudaSetDevice(ctx->device));
  Side& S = ctx->product.side;
  OpRec op;
  memcpy(op.dq, dq, sizeof(op.dq));
  op.fermion = fermion != 0;
  op.allowed.assign((size_t)S.nq * S.nq, 0);
  for (int i = 0; i < S.nq; ++i)                       // SparseMatrix::allocate BaseOperator.C:123-145
    for (int j = 0; j < S.nq; ++j) op.allowed[(size_t)i * S.nq + j] = qn_allow(S.quantum(i), dq, S.quantum(j)) ? 1 : 0;
  layout_op(S, op);
  if (product_is_factorisable(ctx)) {
    op.factorised = true;   // no storage: b2d_product_op_accumulate records scaled sub-blocks of the left child's operators
    b2d_ctx::Product& P = ctx->product;
    if (!P.identity) {      // the `A` of TensorTrace products (identity on the renormalised block): one block of the largest sector size
      int maxd = 1;
      for (int d : ctx->side[0].dims) maxd = std::max(maxd, d);
      P.identity_ld = pad_ld(maxd);
      const size_t n = (size_t)maxd * P.identity_ld;
      CU(arena_alloc_any(ctx, n * 8, &P.identity));
      ctx->arena_doubles += (int64_t)n;
      if (ctx->has_device) {
        std::vector<double> eye(n, 0.0);
        for (int i = 0; i < maxd; ++i) eye[(size_t)i * P.identity_ld + i] = 1.0;
        CU(cudaMemcpyAsync(P.identity, eye.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
      }
    }
  } else if (op.dev_size > 0) {
    CU(arena_alloc(ctx, (size_t)op.dev_size * 8, &op.dev));
    ctx->arena_doubles += op.dev_size;
    CU(cudaMemsetAsync(op.dev, 0, (size_t)op.dev_size * 8, ctx->stream));
  }
  S.ops.push_back(std::move(op));
  ctx->product.pend_subs.resize(S.ops.size());
  *prod_id = (int)S.ops.size() - 1;
  return B2D_OK;
}

// Deferred scatter tasks (option opbuild_batch): executed round by round - tasks of one round hit pairwise different destination pieces,
// a piece receives its contributions in the order they were planned (same summation order as the immediate path: bit-identical).
static int flush_product_tasks(b2d_ctx* ctx) {
  if (ctx->pend_kron.empty()) return B2D_OK;
  { int frc = flush_pending_ops(ctx); if (frc) return frc; }
  CU(cudaSetDevice(ctx->device));
  int nrounds = 0;
  for (int r : ctx->pend_kron_round) nrounds = std::max(nrounds, r + 1);
  std::vector<int> count(nrounds + 1, 0);
  for (int r : ctx->pend_kron_round) count[r + 1]++;
  for (int r = 0; r < nrounds; ++r) count[r + 1] += count[r];
  const size_t ntasks = ctx->pend_kron.size();
  std::vector<KronTask> sorted_pageable;
  const bool pinned = ctx->kron_pinned_tasks.reserve(ntasks * sizeof(KronTask));   // counting sort straight into pinned memory
  KronTask* sorted = nullptr;
  if (pinned) sorted = (KronTask*)ctx->kron_pinned_tasks.p;
  else { sorted_pageable.resize(ntasks); sorted = sorted_pageable.data(); }
  {
    std::vector<int> at(count.begin(), count.end() - 1);
    for (size_t i = 0; i < ntasks; ++i) sorted[at[ctx->pend_kron_round[i]]++] = ctx->pend_kron[i];
  }
  // round 0 = the first contribution to a piece of freshly zero-filled storage: it is stored, not read-modified-written
  for (int i = 0; i < count[1]; ++i) sorted[i].pad |= 1;
  int rc = run_kron_rounds(ctx, sorted, ntasks, count, true, pinned);
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->kron_last_rounds = nrounds;
  ctx->pend_kron.clear(); ctx->pend_kron_round.clear(); ctx->product.pair_hits.clear();
  return B2D_OK;
}

static int product_op_accumulate_impl(b2d_ctx* ctx, int prod_id, int left_op, int left_transposed, int right_op, int right_transposed, double scale, bool defer);
int b2d_product_op_accumulate(b2d_ctx* ctx, int prod_id, int left_op, int left_transposed, int right_op, int right_transposed, double scale) {
  NEED_DEVICE();
  { int frc = flush_product_tasks(ctx); if (frc) return frc; }   // keep the order of contributions
  return product_op_accumulate_impl(ctx, prod_id, left_op, left_transposed, right_op, right_transposed, scale, false);
}
static int product_op_accumulate_impl(b2d_ctx* ctx, int prod_id, int left_op, int left_transposed, int right_op, int right_transposed, double scale, bool defer) {
  if (!ctx) return B2D_ERR_ARG;
  b2d_ctx::Product& P = ctx->product;
  const bool host_only = !ctx->has_device && P.set && prod_id >= 0 && prod_id < (int)P.side.ops.size() && P.side.ops[prod_id].factorised;
  if (!host_only) NEED_DEVICE();
  const Side& L = ctx->side[0];
  const Side& R = ctx->side[1];
  if (!P.set || prod_id < 0 || prod_id >= (int)P.side.ops.size() || left_op >= (int)L.ops.size() || right_op >= (int)R.ops.size() || (left_op < 0 && right_op < 0))
    return fail(ctx, B2D_ERR_ARG, "b2d_product_op_accumulate: bad arguments");
  if (std::fabs(scale) < 1e-20) return B2D_OK;        // TINY, operatorfunctions.C:148
  if (!host_only) {
    { int frc = flush_pending_ops(ctx); if (frc) return frc; }
    CU(cudaSetDevice(ctx->device));
  }
  const OpRec& c = P.side.ops[prod_id];
  const bool trace_l = left_op < 0, trace_r = right_op < 0;      // identity on that child: TensorTrace
  const OpRec* la = trace_l ? nullptr : &L.ops[left_op];
  const OpRec* rb = trace_r ? nullptr : &R.ops[right_op];
  if (!host_only && ((la && la->dev_size > 0 && !la->dev) || (rb && rb->dev_size > 0 && !rb->dev && !c.factorised)))
    return fail(ctx, B2D_ERR_ARG, "b2d_product_op_accumulate: child operator is not resident");
  View a{&L, la, left_transposed != 0}, b{&R, rb, right_transposed != 0};
  const int sa = la ? la->dq[1] : 0, sb = rb ? rb->dq[1] : 0, sc = c.dq[1];
  if (c.factorised) {
    // factorised operator: the same loop as below, but a product becomes one SCALED SUB-BLOCK per (row piece, column piece) - the child
    // operator's block (or the identity block) times  f x (1 x 1 element of the dot operator)  - instead of a scatter into storage
    if (rb && rb->host.empty() && rb->packed_size > 0) return fail(ctx, B2D_ERR_ARG, "factorised operator: the dot operator has no host copy (b2d_add_op with data)");
    if (la && la->factorised) return fail(ctx, B2D_ERR_ARG, "factorised operator: the left child's operators must be materialised");
    std::vector<int64_t> rb_ref;   // packed (host) offset of each allowed block of the dot operator
    if (rb) {
      rb_ref.assign((size_t)R.nq * R.nq, -1);
      int64_t off = 0;
      for (int i = 0; i < R.nq; ++i)
        for (int j = 0; j < R.nq; ++j)
          if (rb->allowed[(size_t)i * R.nq + j]) { rb_ref[(size_t)i * R.nq + j] = off; off += (int64_t)R.dims[i] * R.dims[j]; }
    }
    auto& out = P.pend_subs[prod_id];
    // the factors for scale = 1 and a left operator stored at address 0: a template shared by the products of this operator with the same
    // structure (see KronTemplate)
    const bool left_is_identity = trace_l || la->optype == OP_OVERLAP;
    const std::array<int64_t, 8> fkey{(int64_t)prod_id, trace_l ? -1 : (int64_t)la->dq[0], trace_l ? 0 : (int64_t)la->dq[1], trace_l ? 0 : (int64_t)la->dq[2],
                                      (int64_t)(left_transposed != 0), (int64_t)right_op, (int64_t)(right_transposed != 0), (int64_t)(left_is_identity ? 1 : 0) + 2};
    b2d_ctx::Product::KronTemplate* tp = &P.kron_templates[fkey];
    b2d_ctx::Product::KronTemplate fresh;
    if (tp->left_owner >= 0 && la && (L.ops[tp->left_owner].dev_size != la->dev_size || L.ops[tp->left_owner].packed_size != la->packed_size)) tp = &fresh;
    if (tp->pair_count == 0) {
      b2d_ctx::Product::KronTemplate& T = *tp;
      T.left_owner = la ? left_op : -1;
      try {
        for (int cq = 0; cq < P.side.nq; ++cq)
          for (int cqp = 0; cqp < P.side.nq; ++cqp) {
            if (!c.allowed[(size_t)cq * P.side.nq + cqp]) continue;
            int row = 0;
            for (int oi : P.old_to_new[cq]) {
              int col = 0;
              for (int oj : P.old_to_new[cqp]) {
                const int aq = P.lmap[oi], aqp = P.lmap[oj], bq = P.rmap[oi], bqp = P.rmap[oj];
                const bool a_ok = trace_l ? aq == aqp : a.allowed(aq, aqp);
                const bool b_ok = trace_r ? bq == bqp : b.allowed(bq, bqp);
                if (a_ok && b_ok) {
                  // f = ((scale * 9j) * scalings) * dot element, sign in between (exact): the three factors are kept apart so that a product
                  // evaluates them in the reference's order (operatorfunctions.C:205-218) - the template changes no bit
                  const double n9 = ctx->am.ninej(L.quantum(aqp)[1], R.quantum(bqp)[1], P.side.quantum(cqp)[1], sa, sb, sc, L.quantum(aq)[1], R.quantum(bq)[1],
                                                  P.side.quantum(cq)[1]);
                  double m2 = (!trace_l && !trace_r) ? a.scaling(ctx->am, aq, aqp) * b.scaling(ctx->am, bq, bqp) : 1.0;
                  if (rb && rb->fermion && (L.quantum(aqp)[0] & 1)) m2 = -m2;
                  const double h = rb ? rb->host[(size_t)(b.t ? rb_ref[(size_t)bqp * R.nq + bq] : rb_ref[(size_t)bq * R.nq + bqp])] : 1.0;   // 1 x 1 block: its own transpose
                  if (n9 * m2 * h != 0.0) {
                    b2d_ctx::Product::FactorTemplate ft;
                    ft.blk = (size_t)cq * P.side.nq + cqp;
                    ft.m2 = m2; ft.h = h;
                    ft.sb.r0 = row; ft.sb.c0 = col; ft.sb.m = L.dims[aq]; ft.sb.n = L.dims[aqp]; ft.sb.alpha = n9;
                    // the OVERLAP operator of a renormalised block is the identity (same bra and ket states; what the rotation leaves of it is
                    // I + O(1e-16)): it is contracted as THE identity block, so that "child block + identity" stays a pair of direct factors
                    ft.identity = left_is_identity;
                    if (ft.identity) { ft.sb.a = P.identity; ft.sb.lda = P.identity_ld; ft.sb.t = false; ft.a_off = 0; }
                    else { ft.sb.a = nullptr; ft.a_off = a.stored_off(aq, aqp); ft.sb.lda = a.stored_ld(aq, aqp); ft.sb.t = a.t; }
                    T.factors.push_back(ft);
                  }
                }
                col += P.unc_dims[oj];
              }
              row += P.unc_dims[oi];
            }
          }
      } catch (const std::exception& e) { T = b2d_ctx::Product::KronTemplate(); return fail(ctx, B2D_ERR_ARG, e.what()); }
      T.pair_count = 1;
    }
    for (const b2d_ctx::Product::FactorTemplate& ft : tp->factors) {
      SubBlock sbk = ft.sb;
      if (!ft.identity) sbk.a = la->dev + ft.a_off;
      sbk.alpha = ((scale * ft.sb.alpha) * ft.m2) * ft.h;
      if (sbk.alpha != 0.0) out.emplace_back(ft.blk, sbk);
    }
    return B2D_OK;
  }
  // the tasks for scale = 1 and a left operator stored at address 0 (template, shared by the products of this operator with the same
  // structure); the loop below runs once per template
  const std::array<int64_t, 8> tkey{(int64_t)prod_id, trace_l ? -1 : (int64_t)la->dq[0], trace_l ? 0 : (int64_t)la->dq[1], trace_l ? 0 : (int64_t)la->dq[2],
                                    (int64_t)(left_transposed != 0), (int64_t)right_op, (int64_t)(right_transposed != 0), trace_l ? 0 : (int64_t)la->allowed.size()};
  b2d_ctx::Product::KronTemplate* tp = &P.kron_templates[tkey];
  b2d_ctx::Product::KronTemplate fresh;
  if (tp->left_owner >= 0 && la && (L.ops[tp->left_owner].dev_size != la->dev_size || L.ops[tp->left_owner].packed_size != la->packed_size))
    tp = &fresh;   // same quantum numbers but another block pattern (never with the reference's allocate rule, BaseOperator.C:123-145): no sharing
  if (tp->pair_count == 0) {
    b2d_ctx::Product::KronTemplate& T = *tp;
    T.left_owner = la ? left_op : -1;
    int pair_index = 0;
    try {
      for (int cq = 0; cq < P.side.nq; ++cq)
        for (int cqp = 0; cqp < P.side.nq; ++cqp) {
          if (!c.allowed[(size_t)cq * P.side.nq + cqp]) continue;
          int row = 0;
          for (int oi : P.old_to_new[cq]) {
            int col = 0;
            for (int oj : P.old_to_new[cqp]) {
              const int aq = P.lmap[oi], aqp = P.lmap[oj], bq = P.rmap[oi], bqp = P.rmap[oj];
              const bool a_ok = trace_l ? aq == aqp : a.allowed(aq, aqp);
              const bool b_ok = trace_r ? bq == bqp : b.allowed(bq, bqp);
              ++pair_index;
              if (a_ok && b_ok) {
                T.pair.push_back(pair_index - 1);
                // operatorfunctions.C:205-218 (TensorProduct) / :83-107 (TensorTrace: no get_scaling there)
                // coefficient = (scale * 9j) * scalings with the sign (exact) folded into the second factor: a product evaluates it in the
                // reference's order, the template changes no bit
                const double n9 = ctx->am.ninej(L.quantum(aqp)[1], R.quantum(bqp)[1], P.side.quantum(cqp)[1], sa, sb, sc, L.quantum(aq)[1], R.quantum(bq)[1],
                                                P.side.quantum(cq)[1]);
                double m2 = (!trace_l && !trace_r) ? a.scaling(ctx->am, aq, aqp) * b.scaling(ctx->am, bq, bqp) : 1.0;
                if (rb && rb->fermion && (L.quantum(aqp)[0] & 1)) m2 = -m2;
                T.m2.push_back(m2);
                KronTask k;
                memset(&k, 0, sizeof(k));
                k.coef = n9;
                k.a_rows = L.dims[aq]; k.a_cols = L.dims[aqp]; k.b_rows = R.dims[bq]; k.b_cols = R.dims[bqp];
                if (!trace_l) { k.a = 8 * a.stored_off(aq, aqp); k.lda = a.stored_ld(aq, aqp); k.a_t = a.t ? 1 : 0; }   // + the operator's address, per product
                if (!trace_r) { k.b = (int64_t)(intptr_t)rb->dev + 8 * b.stored_off(bq, bqp); k.ldb = b.stored_ld(bq, bqp); k.b_t = b.t ? 1 : 0; }
                k.dst = (int64_t)(intptr_t)c.dev + 8 * c.off[(size_t)cq * P.side.nq + cqp];
                k.ldd = pad_ld(P.side.dims[cqp]);
                k.row0 = row; k.col0 = col;
                T.tasks.push_back(k);
                T.bytes += 8.0 * ((trace_l ? 0.0 : (double)k.a_rows * k.a_cols) + (trace_r ? 0.0 : (double)k.b_rows * k.b_cols) + 2.0 * (double)k.a_rows * k.b_rows * k.a_cols * k.b_cols);
              }
              col += P.unc_dims[oj];
            }
            row += P.unc_dims[oi];
          }
        }
    } catch (const std::exception& e) { T = b2d_ctx::Product::KronTemplate(); return fail(ctx, B2D_ERR_ARG, e.what()); }
    T.pair_count = std::max(pair_index, 1);
  }
  const b2d_ctx::Product::KronTemplate& T = *tp;
  const std::vector<int>& task_pair = T.pair;
  const int pair_index = T.pair_count;
  const int64_t a_base = trace_l ? 0 : (int64_t)(intptr_t)la->dev;
  std::vector<KronTask> tasks(T.tasks);
  for (size_t i = 0; i < tasks.size(); ++i) { tasks[i].a += a_base; tasks[i].coef = (scale * tasks[i].coef) * T.m2[i]; }
  ctx->kron_bytes += T.bytes;
  ctx->kron_ntasks += (int64_t)tasks.size();
  ctx->kron_nproducts += 1;
  if (defer) {
    // round of a task = how many earlier products of THIS operator hit the same piece pair (per-operator counters; no search structure)
    std::vector<int>& hits = P.pair_hits[prod_id];
    if ((int)hits.size() < pair_index) hits.resize(pair_index, 0);
    for (size_t i = 0; i < tasks.size(); ++i) {
      ctx->pend_kron.push_back(tasks[i]);
      ctx->pend_kron_round.push_back(hits[task_pair[i]]++);
    }
    return B2D_OK;
  }
  int rc = run_kron_rounds(ctx, tasks, std::vector<int>{0, (int)tasks.size()});
  if (rc) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

// Factorised operator complete (every product recorded): turn the product-ordered factor list into the per-block table the planners
// read.  Factors that hit the same (row piece, column piece) are PRE-SUMMED into one block of the child's size ("combo", built on the
// device by the scatter kernel in product order) - a complementary operator has one factor per site of the renormalised block there, and
// contracting them one by one would multiply the flops -, except the common pair {child block, identity}: the identity stays a factor of
// its own (no storage).  Combos with the same parts and proportional coefficients are shared: the spin-recoupling variants of a piece pair
// (S_dot = 1/2 couples to S +- 1/2) differ only by their 9j prefactor, which goes into alpha.
static int finalise_factorised_op(b2d_ctx* ctx, int prod_id) {
  b2d_ctx::Product& P = ctx->product;
  OpRec& op = P.side.ops[prod_id];
  if (!op.factorised || !op.sub_begin.empty()) return B2D_OK;
  auto& pend = P.pend_subs[prod_id];
  // by stored block, then by (row piece, column piece); stable: the contributions to one piece keep the order of the products
  std::stable_sort(pend.begin(), pend.end(), [](const std::pair<size_t, SubBlock>& x, const std::pair<size_t, SubBlock>& y) {
    if (x.first != y.first) return x.first < y.first;
    if (x.second.r0 != y.second.r0) return x.second.r0 < y.second.r0;
    return x.second.c0 < y.second.c0;
  });
  const size_t nb = (size_t)P.side.nq * P.side.nq;
  op.sub_begin.assign(nb + 1, 0);
  op.subs.clear();
  size_t k = 0;
  std::vector<const SubBlock*> list;
  std::unordered_map<uint64_t, size_t> part_index;
  std::vector<std::pair<const double*, bool>> parts;
  std::vector<double> alphas, ratios;
  std::vector<int> ldas;
  for (size_t b = 0; b < nb; ++b) {
    op.sub_begin[b] = (int32_t)op.subs.size();
    size_t k1 = k;
    while (k1 < pend.size() && pend[k1].first == b) ++k1;
    for (size_t i = k, i1; i < k1; i = i1) {
      list.clear();
      for (i1 = i; i1 < k1 && pend[i1].second.r0 == pend[i].second.r0 && pend[i1].second.c0 == pend[i].second.c0; ++i1) list.push_back(&pend[i1].second);
      // merge repeated parts (same child block, same orientation); parts keep the order of their first appearance
      parts.clear(); alphas.clear(); ldas.clear();
      if (list.size() == 1) { parts.emplace_back(list[0]->a, list[0]->t); alphas.push_back(list[0]->alpha); ldas.push_back(list[0]->lda); }
      else if (list.size() <= 8) {
        for (const SubBlock* sb : list) {
          bool merged = false;
          for (size_t q = 0; q < parts.size(); ++q)
            if (parts[q].first == sb->a && parts[q].second == sb->t) { alphas[q] += sb->alpha; merged = true; break; }
          if (!merged) { parts.emplace_back(sb->a, sb->t); alphas.push_back(sb->alpha); ldas.push_back(sb->lda); }
        }
      } else {
      part_index.clear();
      for (const SubBlock* sb : list) {
        const uint64_t pk = ((uint64_t)(intptr_t)sb->a << 1) | (sb->t ? 1u : 0u);   // operator blocks are 8-byte aligned
        auto ins = part_index.emplace(pk, parts.size());
        if (!ins.second) alphas[ins.first->second] += sb->alpha;
        else { parts.emplace_back(sb->a, sb->t); alphas.push_back(sb->alpha); ldas.push_back(sb->lda); }
      }
      }
      const SubBlock& first = *list[0];
      int nident = 0;
      for (const auto& pr : parts) nident += pr.first == P.identity;
      if (parts.size() == 1 || (parts.size() == 2 && nident == 1 && !ctx->presum_identity)) {
        for (size_t q = 0; q < parts.size(); ++q) {
          if (alphas[q] == 0.0) continue;
          SubBlock sb = first;
          sb.a = parts[q].first; sb.t = parts[q].second; sb.lda = ldas[q]; sb.alpha = alphas[q];
          op.subs.push_back(sb);
          ++ctx->nsubs_direct;
        }
        continue;
      }
      size_t lead = 0;
      while (lead < alphas.size() && alphas[lead] == 0.0) ++lead;
      if (lead == alphas.size()) continue;
      ratios.resize(alphas.size());
      for (size_t q = 0; q < alphas.size(); ++q) ratios[q] = alphas[q] / alphas[lead];
      // bucket key: the parts (addresses, orientation) and the coefficient ratios quantised to ~1e-8 of the largest one - candidates in a
      // bucket are then compared exactly (1e-12); two equal vectors straddling a quantisation boundary merely miss the sharing
      uint64_t hk = 1469598103934665603ull;
      auto mixin = [&](uint64_t v) { hk ^= v; hk *= 1099511628211ull; };
      double rmax = 0.0;
      for (double r : ratios) rmax = std::max(rmax, std::fabs(r));
      for (size_t q = 0; q < parts.size(); ++q) {
        mixin((uint64_t)(intptr_t)parts[q].first); mixin(parts[q].second ? 2 : 1);
        mixin((uint64_t)(int64_t)std::llround(ratios[q] / rmax * 67108864.0));
      }
      const std::array<int64_t, 3> key{(int64_t)hk, (int64_t)first.m, (int64_t)first.n};
      std::vector<b2d_ctx::Product::Combo>& cands = P.combos[key];
      const double* block = nullptr;
      for (const auto& c : cands) {
        if (c.parts != parts) continue;
        bool same = true;
        for (size_t q = 0; q < ratios.size() && same; ++q) same = std::fabs(c.ratios[q] - ratios[q]) <= 1e-12 * std::max(std::fabs(ratios[q]), 1e-300);
        if (same) { block = c.block; break; }
      }
      const int ldc = pad_ld(first.n);
      if (!block) {
        double* dst = nullptr;
        const size_t n = (size_t)align_up((int64_t)first.m * ldc, BLK_ALIGN);
        CU(arena_alloc_any(ctx, n * 8, &dst));
        if (ctx->has_device) CU(cudaMemsetAsync(dst, 0, n * 8, ctx->stream));
        ctx->arena_doubles += (int64_t)n; ctx->combo_doubles += (int64_t)n;
        if (getenv("B2D_FACT_DEBUG")) { static std::map<int, std::array<double, 3>> acc; auto& a = acc[op.optype]; a[0] += (double)n; a[1] += 1; a[2] += (double)parts.size(); fprintf(stderr, "FACT optype %d combos %.0f doubles %.3e avg parts %.1f | this: parts %d m %d n %d ident %d r0 %d c0 %d blk %zu\n", op.optype, a[1], a[0], a[2] / a[1], (int)parts.size(), first.m, first.n, nident, first.r0, first.c0, b); }
        int round = 0;   // the block is fresh: its parts simply take successive rounds (no lookup in the hit table)
        for (size_t q = 0; q < parts.size(); ++q) {
          if (ratios[q] == 0.0) continue;
          KronTask kt;
          memset(&kt, 0, sizeof(kt));
          kt.a = (int64_t)(intptr_t)parts[q].first; kt.b = 0; kt.dst = (int64_t)(intptr_t)dst; kt.coef = ratios[q];
          kt.a_rows = first.m; kt.a_cols = first.n; kt.lda = ldas[q]; kt.a_t = parts[q].second ? 1 : 0;
          kt.b_rows = kt.b_cols = 1; kt.ldb = 1; kt.b_t = 0; kt.row0 = kt.col0 = 0; kt.ldd = ldc;
          ctx->pend_kron.push_back(kt);
          ctx->pend_kron_round.push_back(round++);
          ctx->kron_bytes += 8.0 * 3.0 * (double)first.m * first.n;
          ++ctx->kron_ntasks;
        }
        cands.push_back(b2d_ctx::Product::Combo{parts, ratios, dst});
        block = dst;
        ++ctx->ncombos;
      }
      SubBlock sb = first;
      sb.a = block; sb.lda = ldc; sb.t = false; sb.alpha = alphas[lead];
      op.subs.push_back(sb);
      ++ctx->nsubs_combo;
    }
    k = k1;
  }
  op.sub_begin[nb] = (int32_t)op.subs.size();
  std::vector<std::pair<size_t, SubBlock>>().swap(pend);
  return B2D_OK;
}

int b2d_set_integrals(b2d_ctx* ctx, int norbs, const double* v1, const double* v2, const int32_t* orbital_irreps, double one_tol, double two_tol) {
  if (!ctx || norbs <= 0 || !v1 || !v2 || !orbital_irreps) return fail(ctx, B2D_ERR_ARG, "b2d_set_integrals: bad arguments");
  Integrals& I = ctx->integrals;
  I.n = norbs;
  I.h1.assign(v1, v1 + (size_t)norbs * norbs);
  I.h2.assign(v2, v2 + (size_t)norbs * norbs * norbs * norbs);
  I.irrep.assign(orbital_irreps, orbital_irreps + norbs);
  I.one_tol = one_tol; I.two_tol = two_tol;
  I.set = true;
  return B2D_OK;
}

static int plan_enlarged_op(b2d_ctx* ctx, int optype, int norb, const int32_t* orbs, const int32_t* dq, int hubbard, std::vector<ProductCall>& calls) {
  if (!ctx || !dq || norb < 0 || norb > 2 || (norb > 0 && !orbs)) return fail(ctx, B2D_ERR_ARG, "enlarged-block operator: bad arguments");
  if (ctx->side[0].nq == 0 || ctx->side[1].nq == 0) return fail(ctx, B2D_ERR_ARG, "enlarged-block operator: set both children first");
  int o[2] = {norb > 0 ? orbs[0] : -1, norb > 1 ? orbs[1] : -1};
  int q[3] = {dq[0], dq[1], dq[2]};
  try {
    OpBuildPlanner P(ctx->side[0], ctx->side[1], ctx->integrals, hubbard != 0, ctx->am);
    calls = P.plan(optype, o, q);
  } catch (const std::exception& e) { return fail(ctx, B2D_ERR_ARG, e.what()); }
  return B2D_OK;
}

int b2d_enlarged_op_products(b2d_ctx* ctx, int optype, int norb, const int32_t* orbs, const int32_t* dq, int hubbard, int max_products,
                             int32_t* left_op, int32_t* right_op, int32_t* flags, double* scale) {
  std::vector<ProductCall> calls;
  int rc = plan_enlarged_op(ctx, optype, norb, orbs, dq, hubbard, calls);
  if (rc) return -rc;
  for (int k = 0; k < (int)calls.size() && k < max_products; ++k) {
    if (left_op) left_op[k] = calls[k].lop;
    if (right_op) right_op[k] = calls[k].rop;
    if (flags) flags[k] = (calls[k].lt ? 1 : 0) | (calls[k].rt ? 2 : 0);
    if (scale) scale[k] = calls[k].scale;
  }
  return (int)calls.size();
}

int b2d_build_enlarged_op(b2d_ctx* ctx, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq, int fermion, int hubbard, int* prod_id) {
  if (!ctx) return B2D_ERR_ARG;
  if (!ctx->has_device && !product_is_factorisable(ctx)) NEED_DEVICE();
  if (!prod_id) return fail(ctx, B2D_ERR_ARG, "b2d_build_enlarged_op: bad arguments");
  std::vector<ProductCall> calls;
  int rc = plan_enlarged_op(ctx, optype, norb, orbs, dq, hubbard, calls);
  if (rc) return rc;
  rc = b2d_product_op_create(ctx, dq, fermion, prod_id);
  if (rc) return rc;
  {   // identity of the operator inside its array: the term planner (b2d_plan) finds operators by (type, orbitals, component)
    OpRec& op = ctx->product.side.ops[*prod_id];
    op.optype = optype; op.norb = norb; op.comp = comp;
    for (int k = 0; k < norb; ++k) op.orbs[k] = orbs[k];
  }
  for (const ProductCall& c : calls) {
    rc = (ctx->opbuild_batch || ctx->product.side.ops[*prod_id].factorised)
             ? product_op_accumulate_impl(ctx, *prod_id, c.lop, c.lt ? 1 : 0, c.rop, c.rt ? 1 : 0, c.scale, true)
             : b2d_product_op_accumulate(ctx, *prod_id, c.lop, c.lt ? 1 : 0, c.rop, c.rt ? 1 : 0, c.scale);
    if (rc) return rc;
  }
  return finalise_factorised_op(ctx, *prod_id);
}

// One child of the big block is ready (built on the device from ITS children, or uploaded as it is): park it, so that side 0 / side 1
// are free to describe the children of the other one.  b2d_assemble_big then makes the two parked blocks the children of the big block.
int b2d_stash_product(b2d_ctx* ctx, int slot, int is_loop, int nsites, const int32_t* sites) {
  if (!ctx || slot < 0 || slot > 1 || !ctx->product.set) return fail(ctx, B2D_ERR_ARG, "b2d_stash_product: no product block");
  for (size_t m = 0; m < ctx->product.side.ops.size(); ++m) { int frc = finalise_factorised_op(ctx, (int)m); if (frc) return frc; }
  if (ctx->has_device) { int frc = flush_product_tasks(ctx); if (frc) return frc; }
  Side s = std::move(ctx->product.side);
  s.loop = is_loop != 0;
  s.sites.clear();
  if (sites) s.sites.assign(sites, sites + nsites);
  ctx->stash[slot] = std::move(s);
  ctx->stash_set[slot] = true;
  ctx->product = b2d_ctx::Product();
  return B2D_OK;
}
int b2d_stash_side(b2d_ctx* ctx, int slot, int from_side) {
  if (!ctx || slot < 0 || slot > 1 || from_side < 0 || from_side > 1 || ctx->side[from_side].nq == 0) return fail(ctx, B2D_ERR_ARG, "b2d_stash_side: bad arguments");
  if (ctx->has_device) { int frc = flush_pending_ops(ctx); if (frc) return frc; }
  ctx->stash[slot] = std::move(ctx->side[from_side]);
  ctx->side[from_side] = Side();
  ctx->stash_set[slot] = true;
  return B2D_OK;
}
int b2d_assemble_big(b2d_ctx* ctx) {
  if (!ctx || !ctx->stash_set[0] || !ctx->stash_set[1]) return fail(ctx, B2D_ERR_ARG, "b2d_assemble_big: both children must be stashed first");
  ctx->side[0] = std::move(ctx->stash[0]);
  ctx->side[1] = std::move(ctx->stash[1]);
  ctx->stash[0] = Side(); ctx->stash[1] = Side();
  ctx->stash_set[0] = ctx->stash_set[1] = false;
  ctx->planned = false;
  return B2D_OK;
}

// ---- guess wavefunction of the next block iteration (GuessWave::transform_previous_wavefunction) ------------------------
int b2d_guess_plan(b2d_ctx* ctx, const b2d_guess_desc* desc, double* out, int n) {
  if (!ctx || !desc) return fail(ctx, B2D_ERR_ARG, "b2d_guess_plan: null argument");
  try {
    ctx->guess = plan_guess_transform(*desc, ctx->am, ctx->forced_class);
  } catch (const std::exception& e) {
    ctx->guess = GuessPlan();
    return fail(ctx, B2D_ERR_ARG, e.what());
  }
  const GuessPlan& P = ctx->guess;
  int64_t ntasks = 0;
  for (const auto& r : P.rounds) ntasks += (int64_t)r.size();
  const double v[8] = {(double)P.old_size, (double)P.lrot_size, (double)P.rrot_size, (double)P.trial.W, P.flops, (double)P.shuffle_bytes, (double)ntasks, (double)P.rounds.size()};
  for (int i = 0; i < n && i < 8; ++i) out[i] = v[i];
  return B2D_OK;
}

int64_t b2d_guess_plan_export(const b2d_ctx* ctx, int what, void* out, int64_t cap) {
  if (!ctx || !ctx->guess.valid) return -1;
  const GuessPlan& P = ctx->guess;
  std::vector<char> buf;
  auto put = [&](const void* p, size_t bytes) { const char* c = (const char*)p; buf.insert(buf.end(), c, c + bytes); };
  switch (what) {
    case 0: put(P.gemm_a.segs.data(), P.gemm_a.segs.size() * sizeof(GSeg)); break;
    case 1: put(P.gemm_a.groups.data(), P.gemm_a.groups.size() * sizeof(GGroup)); break;
    case 10: put(P.gemm_b.segs.data(), P.gemm_b.segs.size() * sizeof(GSeg)); break;
    case 11: put(P.gemm_b.groups.data(), P.gemm_b.groups.size() * sizeof(GGroup)); break;
    case 2: for (const auto& r : P.rounds) put(r.data(), r.size() * sizeof(KronTask)); break;
    case 3: for (const auto& r : P.rounds) { int32_t c = (int32_t)r.size(); put(&c, 4); } break;
    case 4: put(P.gemm_c.segs.data(), P.gemm_c.segs.size() * sizeof(GSeg)); break;
    case 5: put(P.gemm_c.groups.data(), P.gemm_c.groups.size() * sizeof(GGroup)); break;
    case 6:
      put(P.in_old.data(), P.in_old.size() * sizeof(BlockDesc)); put(P.in_lrot.data(), P.in_lrot.size() * sizeof(BlockDesc));
      put(P.in_rrot.data(), P.in_rrot.size() * sizeof(BlockDesc));
      break;
    case 7: {
      int32_t c[4] = {(int32_t)P.in_old.size(), (int32_t)P.in_lrot.size(), (int32_t)P.in_rrot.size(), 0};
      int64_t s[3] = {P.image_size, P.t1_size, P.work_size};
      put(c, sizeof(c)); put(s, sizeof(s));
      break;
    }
    case 8: {
      for (int p = 0; p < P.trial.nblocks(); ++p) {
        BlockDesc d; d.ref_off = P.trial.ref_off[p]; d.dev_off = P.trial.dev_off[p]; d.rows = P.trial.rows[p]; d.cols = P.trial.cols[p]; d.ld = P.trial.ld[p]; d.pad = 0;
        put(&d, sizeof(d));
      }
      break;
    }
    case 9: { int32_t sz[4] = {(int32_t)sizeof(GSeg), (int32_t)sizeof(GGroup), (int32_t)sizeof(KronTask), (int32_t)sizeof(BlockDesc)}; put(sz, sizeof(sz)); break; }
    default: return -1;
  }
  if (out && cap >= (int64_t)buf.size() && !buf.empty()) memcpy(out, buf.data(), buf.size());
  return (int64_t)buf.size();
}

int b2d_guess_transform(b2d_ctx* ctx, const double* old_wave, const double* left_rot, const double* right_rot, int dst_slot, double* trial) {
  NEED_DEVICE();
  GuessPlan& P = ctx->guess;
  if (!P.valid) return fail(ctx, B2D_ERR_ARG, "b2d_guess_transform: call b2d_guess_plan first");
  if ((P.old_size && !old_wave) || (P.lrot_size && !left_rot) || (P.rrot_size && !right_rot)) return fail(ctx, B2D_ERR_ARG, "b2d_guess_transform: null input");
  if (dst_slot < 0 && !trial) return fail(ctx, B2D_ERR_ARG, "b2d_guess_transform: no destination");
  double* dst = nullptr;
  if (dst_slot >= 0) {
    NEED_PLAN(); CHECK_SLOT(dst_slot);
    const PsiLayout& a = ctx->psi; const PsiLayout& b = P.trial;
    if (a.W != b.W || a.Wp != b.Wp || a.bl != b.bl || a.br != b.br || a.rows != b.rows || a.cols != b.cols)
      return fail(ctx, B2D_ERR_ARG, "b2d_guess_transform: the trial vector's sector layout is not the planned big block's");
    dst = user_vec(ctx, dst_slot);
  }
  CU(cudaSetDevice(ctx->device));
  // 1. one upload of the three inputs through the pinned staging buffer, packed into the padded image
  const int64_t n_in = P.old_size + P.lrot_size + P.rrot_size;
  int rc = flush_pending_ops(ctx);
  if (rc) return rc;
  double* pin = pending_room(ctx, (size_t)std::max<int64_t>(n_in, 1), &rc);
  if (!pin) return rc;
  if (P.old_size) memcpy(pin, old_wave, (size_t)P.old_size * 8);
  if (P.lrot_size) memcpy(pin + P.old_size, left_rot, (size_t)P.lrot_size * 8);
  if (P.rrot_size) memcpy(pin + P.old_size + P.lrot_size, right_rot, (size_t)P.rrot_size * 8);
  ctx->pend_used = 0;   // the staging area is consumed below, before anybody else can claim it
  CU(ctx->staging.reserve((size_t)std::max<int64_t>(n_in, 2) * 8));
  CU(cudaMemcpyAsync(ctx->staging.p, pin, (size_t)n_in * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx->guess_image.reserve((size_t)std::max<int64_t>(P.image_size, 16) * 8));
  CU(cudaMemsetAsync(ctx->guess_image.p, 0, (size_t)std::max<int64_t>(P.image_size, 16) * 8, ctx->stream));
  std::vector<BlockDesc> blocks;
  blocks.reserve(P.in_old.size() + P.in_lrot.size() + P.in_rrot.size());
  for (BlockDesc d : P.in_old) blocks.push_back(d);
  for (BlockDesc d : P.in_lrot) { d.ref_off += P.old_size; blocks.push_back(d); }
  for (BlockDesc d : P.in_rrot) { d.ref_off += P.old_size + P.lrot_size; blocks.push_back(d); }
  rc = upload_desc(ctx, ctx->desc_scratch, blocks.data(), blocks.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_pack((const BlockDesc*)ctx->desc_scratch.p, (int)blocks.size(), (const double*)ctx->staging.p, (double*)ctx->guess_image.p, ctx->stream, &ctx->launches));
  // 2. destination and workspace
  if (!dst) {
    CU(ctx->guess_trial.reserve((size_t)std::max<int64_t>(P.trial.Wp, 16) * 8));
    dst = (double*)ctx->guess_trial.p;
  }
  CU(cudaMemsetAsync(dst, 0, (size_t)P.trial.Wp * 8, ctx->stream));
  CU(ctx->work.reserve((size_t)std::max<int64_t>(P.work_size, 16) * 8));
  if (P.work_size > P.t2_off) CU(cudaMemsetAsync((double*)ctx->work.p + P.t2_off, 0, (size_t)(P.work_size - P.t2_off) * 8, ctx->stream));
  // 3. contraction batch a, batch b (one-dot), the shuffle rounds, batch c (two-dot)
  Schedule S1, S3;
  { Chunk c; c.step1 = P.gemm_a; c.step2 = P.gemm_b; c.nterms = 1; c.work = P.work_size; S1.chunks.push_back(std::move(c)); S1.work_max = P.work_size; }
  { Chunk c; c.step1 = P.gemm_c; c.nterms = 1; c.work = P.work_size; S3.chunks.push_back(std::move(c)); S3.work_max = P.work_size; }
  DevSchedule D1, D3;
  struct Release { DevSchedule &a, &b; ~Release() { a.buf.release(); b.buf.release(); } } release_schedules{D1, D3};   // on every return path
  rc = upload_schedule(ctx, S1, D1);
  if (rc) return rc;
  rc = upload_schedule(ctx, S3, D3);
  if (rc) return rc;
  begin_timing(ctx);
  rc = run_schedule(ctx, S1, D1, nullptr, dst, (double*)ctx->guess_image.p);
  if (rc) return rc;
  {
    std::vector<KronTask> tasks;
    for (const auto& r : P.rounds)
      for (KronTask t : r) {
        t.a = (int64_t)(intptr_t)(((t.pad & 2) ? (double*)ctx->guess_image.p : (double*)ctx->work.p) + t.a);   // bit 1: the source is the input image
        t.dst = (int64_t)(intptr_t)(((t.pad & 1) ? dst : (double*)ctx->work.p) + t.dst);                      // bit 0: the task writes the trial vector
        t.pad = 0;
        tasks.push_back(t);
      }
    std::vector<int> round_begin(1, 0);
    for (const auto& r : P.rounds) round_begin.push_back(round_begin.back() + (int)r.size());   // tasks of one round never overlap; the rounds accumulate in stream order
    rc = run_kron_rounds(ctx, tasks, round_begin);
    if (rc) return rc;
  }
  rc = run_schedule(ctx, S3, D3, nullptr, dst, (double*)ctx->guess_image.p);
  end_timing(ctx);
  if (rc) return rc;
  if (trial) {
    std::vector<BlockDesc> tb;
    for (int p = 0; p < P.trial.nblocks(); ++p) {
      BlockDesc d; d.ref_off = P.trial.ref_off[p]; d.dev_off = P.trial.dev_off[p]; d.rows = P.trial.rows[p]; d.cols = P.trial.cols[p]; d.ld = P.trial.ld[p]; d.pad = 0;
      tb.push_back(d);
    }
    rc = upload_desc(ctx, ctx->desc_scratch, tb.data(), tb.size() * sizeof(BlockDesc));
    if (rc) return rc;
    CU(ctx->staging.reserve((size_t)std::max<int64_t>(P.trial.W, 2) * 8));
    CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)tb.size(), dst, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
    CU(cudaMemcpyAsync(trial, ctx->staging.p, (size_t)P.trial.W * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

// statistics of the operator construction since b2d_set_product_stateinfo: out = {products, scatter tasks, algorithmic bytes, rounds of the
// last batched flush}
int b2d_product_stats(const b2d_ctx* ctx, double* out, int n) {
  if (!ctx || !out) return B2D_ERR_ARG;
  const double v[4] = {(double)ctx->kron_nproducts, (double)ctx->kron_ntasks, ctx->kron_bytes, (double)ctx->kron_last_rounds};
  for (int i = 0; i < n && i < 4; ++i) out[i] = v[i];
  return B2D_OK;
}

// A factorised operator written out as the dense sector blocks the reference's Op::build produces (check mode, tests): the scatter
// kernel adds every scaled sub-block at its (row piece, column piece) position; factors of one position go to successive rounds.
static int materialise_factorised(b2d_ctx* ctx, const Side& S, const OpRec& op, DevBuf& buf) {
  CU(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)std::max<int64_t>(op.dev_size, 16) * 8;
  CU(buf.reserve(bytes));
  CU(cudaMemsetAsync(buf.p, 0, bytes, ctx->stream));
  std::vector<KronTask> tasks;
  std::vector<int> round;
  std::map<std::array<int64_t, 3>, int> hits;
  int nrounds = 0;
  for (int i = 0; i < S.nq; ++i)
    for (int j = 0; j < S.nq; ++j) {
      if (!op.allowed[(size_t)i * S.nq + j]) continue;
      const size_t b = (size_t)i * S.nq + j;
      for (int32_t k = op.sub_begin[b]; k < op.sub_begin[b + 1]; ++k) {
        const SubBlock& sb = op.subs[k];
        KronTask kt;
        memset(&kt, 0, sizeof(kt));
        kt.a = (int64_t)(intptr_t)sb.a; kt.b = 0; kt.dst = (int64_t)(intptr_t)((double*)buf.p + op.off[b]); kt.coef = sb.alpha;
        kt.a_rows = sb.m; kt.a_cols = sb.n; kt.lda = sb.lda; kt.a_t = sb.t ? 1 : 0;
        kt.b_rows = kt.b_cols = 1; kt.ldb = 1; kt.row0 = sb.r0; kt.col0 = sb.c0; kt.ldd = pad_ld(S.dims[j]);
        int& h = hits[std::array<int64_t, 3>{kt.dst, (int64_t)kt.row0, (int64_t)kt.col0}];
        tasks.push_back(kt); round.push_back(h);
        nrounds = std::max(nrounds, ++h);
      }
    }
  std::vector<int> count(nrounds + 1, 0);
  for (int r : round) count[r + 1]++;
  for (int r = 0; r < nrounds; ++r) count[r + 1] += count[r];
  std::vector<KronTask> sorted(tasks.size());
  { std::vector<int> at(count.begin(), count.end() - 1); for (size_t i = 0; i < tasks.size(); ++i) sorted[at[round[i]]++] = tasks[i]; }
  return run_kron_rounds(ctx, sorted, count);
}

// Write a factorised operator of a child of the planned big block out densely, in place: afterwards it is an ordinary materialised
// operator (b2d_tensor_multiply, b2d_download_op, ... read the dense blocks).  The sigma schedule of the current plan keeps contracting
// the factors.  Used for the full-size self-check of the benchmark: the same operator pair through both forms.
int b2d_materialise_op(b2d_ctx* ctx, int side, int op_id) {
  NEED_DEVICE();
  if (side < 0 || side > 1 || op_id < 0 || op_id >= (int)ctx->side[side].ops.size()) return fail(ctx, B2D_ERR_ARG, "b2d_materialise_op: bad arguments");
  Side& S = ctx->side[side];
  OpRec& op = S.ops[op_id];
  if (!op.factorised) return B2D_OK;
  DevBuf tmp;
  int rc = materialise_factorised(ctx, S, op, tmp);
  if (rc) { tmp.release(); return rc; }
  if (op.dev_size > 0) {
    double* dst = nullptr;
    cudaError_t e = arena_alloc(ctx, (size_t)op.dev_size * 8, &dst);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, tmp.p, (size_t)op.dev_size * 8, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    tmp.release();
    if (e != cudaSuccess) return fail(ctx, B2D_ERR_CUDA, std::string("b2d_materialise_op: ") + cudaGetErrorString(e));
    op.dev = dst;
    ctx->arena_doubles += op.dev_size;
  } else tmp.release();
  op.factorised = false;
  return B2D_OK;
}

int64_t b2d_product_op_size(const b2d_ctx* ctx, int prod_id) {
  if (!ctx || !ctx->product.set || prod_id < 0 || prod_id >= (int)ctx->product.side.ops.size()) return -1;
  return ctx->product.side.ops[prod_id].packed_size;
}

int b2d_product_op_download(b2d_ctx* ctx, int prod_id, uint8_t* allowed, double* data) {
  NEED_DEVICE();
  if (!ctx->product.set || prod_id < 0 || prod_id >= (int)ctx->product.side.ops.size()) return fail(ctx, B2D_ERR_ARG, "b2d_product_op_download: bad arguments");
  { int frc = finalise_factorised_op(ctx, prod_id); if (frc) return frc; }
  { int frc = flush_product_tasks(ctx); if (frc) return frc; }
  const Side& S = ctx->product.side;
  const OpRec& op = S.ops[prod_id];
  if (allowed) memcpy(allowed, op.allowed.data(), op.allowed.size());
  if (!data || op.packed_size == 0) return B2D_OK;
  const double* dev = op.dev;
  if (op.factorised) {
    int mrc = materialise_factorised(ctx, S, op, ctx->materialised);
    if (mrc) return mrc;
    dev = (const double*)ctx->materialised.p;
  }
  std::vector<BlockDesc> bd = op_blocks(S, op);
  CU(ctx->staging.reserve((size_t)op.packed_size * 8));
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), dev, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(data, ctx->staging.p, (size_t)op.packed_size * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

// ---- device-side shadow of the scratch files ------------------------------------------------------------
// b2d_transform_operators leaves the renormalised left block in its own buffer; b2d_cache_put_rotated moves that buffer (no copy) into
// the cache under a fresh token and b2d_cache_use makes a cached block a child of the next product / big block without any host traffic.
int b2d_cache_put_rotated(b2d_ctx* ctx, uint64_t* token) {
  NEED_DEVICE();
  if (!ctx->have_rotated || !token) return fail(ctx, B2D_ERR_ARG, "b2d_cache_put_rotated: call b2d_transform_operators first");
  CU(cudaSetDevice(ctx->device));
  b2d_ctx::CachedBlock cb;
  cb.side = std::move(ctx->rotated);
  ctx->rotated = Side();
  const double* base = (const double*)ctx->rotated_arena.p;
  int64_t total = 0;
  for (OpRec& op : cb.side.ops) {
    if (op.dev) { op.cache_off = op.dev - base; total = std::max(total, op.cache_off + op.dev_size); op.dev = nullptr; }
    else op.cache_off = -1;
  }
  cb.doubles = total;
  // device budget: option cache_device_mb, or (default, <= 0) automatic - keep the block on the device while at least 35 % of the GPU's
  // memory stays free for the operator arena and the workspaces of the next block iterations
  bool on_device = (double)(ctx->cache_device_doubles + total) * 8.0 <= ctx->cache_device_mb * 1048576.0;
  if (ctx->cache_device_mb <= 0.0) {
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    on_device = (double)free_b >= 0.35 * (double)total_b;
  }
  if (on_device) {
    cb.dev = ctx->rotated_arena;            // the buffer changes owner; b2d_transform_operators allocates a new one next time
    ctx->rotated_arena = DevBuf();
    ctx->cache_device_doubles += total;
  } else {                                   // over budget: spill to pinned host memory
    if (total > 0) {
      if (cudaMallocHost(&cb.pinned, (size_t)total * 8) != cudaSuccess) return fail(ctx, B2D_ERR_CUDA, "b2d_cache_put_rotated: cudaMallocHost failed");
      CU(cudaMemcpyAsync(cb.pinned, base, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
  }
  ctx->have_rotated = false;
  *token = ctx->cache_next_token++;
  ctx->cache[*token] = std::move(cb);
  ++ctx->cache_puts;
  return B2D_OK;
}

int b2d_cache_use(b2d_ctx* ctx, uint64_t token, int side, int is_loop) {
  NEED_DEVICE();
  auto it = ctx->cache.find(token);
  if (it == ctx->cache.end() || side < 0 || side > 1) return fail(ctx, B2D_ERR_ARG, "b2d_cache_use: unknown token");
  CU(cudaSetDevice(ctx->device));
  b2d_ctx::CachedBlock& cb = it->second;
  const double* base = (const double*)cb.dev.p;
  if (!base && cb.doubles > 0) {             // spilled entry: one H2D copy into the arena (freed by the next b2d_reset)
    double* tmp = nullptr;
    CU(arena_alloc(ctx, (size_t)cb.doubles * 8, &tmp));
    CU(cudaMemcpyAsync(tmp, cb.pinned, (size_t)cb.doubles * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->arena_doubles += cb.doubles;
    base = tmp;
  }
  Side s = cb.side;
  for (OpRec& op : s.ops) {
    if (op.cache_off >= 0) op.dev = const_cast<double*>(base) + op.cache_off;
    op.cache_off = -1;
  }
  s.loop = is_loop != 0;
  ctx->side[side] = std::move(s);
  ctx->planned = false;
  ctx->cache_in_use.insert(token);
  ++ctx->cache_hits;
  return B2D_OK;
}

int b2d_cache_block_info(const b2d_ctx* ctx, uint64_t token, int32_t* nq, int32_t* nops, int32_t* nsites) {
  if (!ctx) return B2D_ERR_ARG;
  auto it = ctx->cache.find(token);
  if (it == ctx->cache.end()) return B2D_ERR_ARG;
  if (nq) *nq = it->second.side.nq;
  if (nops) *nops = (int32_t)it->second.side.ops.size();
  if (nsites) *nsites = (int32_t)it->second.side.sites.size();
  return B2D_OK;
}
int b2d_cache_block_sectors(const b2d_ctx* ctx, uint64_t token, int32_t* q, int32_t* dims, int32_t* sites) {
  if (!ctx) return B2D_ERR_ARG;
  auto it = ctx->cache.find(token);
  if (it == ctx->cache.end()) return B2D_ERR_ARG;
  const Side& s = it->second.side;
  if (q) memcpy(q, s.q.data(), s.q.size() * sizeof(int32_t));
  if (dims) memcpy(dims, s.dims.data(), s.dims.size() * sizeof(int32_t));
  if (sites) memcpy(sites, s.sites.data(), s.sites.size() * sizeof(int32_t));
  return B2D_OK;
}
int b2d_cache_op_info(const b2d_ctx* ctx, uint64_t token, int op_id, int32_t* optype, int32_t* norb, int32_t* orbs, int32_t* comp, int64_t* packed_size) {
  if (!ctx) return B2D_ERR_ARG;
  auto it = ctx->cache.find(token);
  if (it == ctx->cache.end() || op_id < 0 || op_id >= (int)it->second.side.ops.size()) return B2D_ERR_ARG;
  const OpRec& op = it->second.side.ops[op_id];
  if (optype) *optype = op.optype;
  if (norb) *norb = op.norb;
  if (orbs) { orbs[0] = op.orbs[0]; orbs[1] = op.orbs[1]; }
  if (comp) *comp = op.comp;
  if (packed_size) *packed_size = op.packed_size;
  return B2D_OK;
}

// packed blocks of one operator of a cached block back on the host (what the scratch file would have held): the binding calls this
// when a host-side code path needs the real matrices after all
int b2d_cache_download_op(b2d_ctx* ctx, uint64_t token, int op_id, uint8_t* allowed, double* data) {
  NEED_DEVICE();
  auto it = ctx->cache.find(token);
  if (it == ctx->cache.end() || op_id < 0 || op_id >= (int)it->second.side.ops.size()) return fail(ctx, B2D_ERR_ARG, "b2d_cache_download_op: bad arguments");
  CU(cudaSetDevice(ctx->device));
  b2d_ctx::CachedBlock& cb = it->second;
  const Side& S = cb.side;
  const OpRec& op = S.ops[op_id];
  if (allowed) memcpy(allowed, op.allowed.data(), op.allowed.size());
  if (!data || op.packed_size == 0) return B2D_OK;
  if (op.cache_off < 0) return fail(ctx, B2D_ERR_ARG, "b2d_cache_download_op: operator was not resident on this rank");
  const int64_t off = op.cache_off;
  const double* dev = nullptr;
  if (cb.dev.p) dev = (const double*)cb.dev.p + off;
  else {   // spilled: stage this operator's padded image on the device
    CU(ctx->materialised.reserve((size_t)op.dev_size * 8));
    CU(cudaMemcpyAsync(ctx->materialised.p, cb.pinned + off, (size_t)op.dev_size * 8, cudaMemcpyHostToDevice, ctx->stream));
    dev = (const double*)ctx->materialised.p;
  }
  std::vector<BlockDesc> bd = op_blocks(S, op);
  CU(ctx->staging.reserve((size_t)op.packed_size * 8));
  int rc = upload_desc(ctx, ctx->desc_scratch, bd.data(), bd.size() * sizeof(BlockDesc));
  if (rc) return rc;
  CU(launch_unpack((const BlockDesc*)ctx->desc_scratch.p, (int)bd.size(), dev, (double*)ctx->staging.p, ctx->stream, &ctx->launches));
  CU(cudaMemcpyAsync(data, ctx->staging.p, (size_t)op.packed_size * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}

int b2d_cache_drop(b2d_ctx* ctx, uint64_t token) {
  if (!ctx) return B2D_ERR_ARG;
  auto it = ctx->cache.find(token);
  if (it == ctx->cache.end()) return B2D_OK;
  if (ctx->has_device) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  if (it->second.dev.p) {
    ctx->cache_device_doubles -= it->second.doubles;
    ctx->spare_bufs.push_back(it->second.dev);   // keep a few: the next renormalised block of this size takes it over
    it->second.dev = DevBuf();
    while (ctx->spare_bufs.size() > 6) {         // bounded: release the smallest
      size_t k = 0;
      for (size_t i = 1; i < ctx->spare_bufs.size(); ++i) if (ctx->spare_bufs[i].cap < ctx->spare_bufs[k].cap) k = i;
      ctx->spare_bufs[k].release();
      ctx->spare_bufs.erase(ctx->spare_bufs.begin() + k);
    }
  }
  if (it->second.pinned) cudaFreeHost(it->second.pinned);
  ctx->cache.erase(it);
  return B2D_OK;
}

int b2d_cache_stats(const b2d_ctx* ctx, double* out, int n) {
  if (!ctx || !out) return B2D_ERR_ARG;
  double spilled = 0;
  for (const auto& kv : ctx->cache) if (!kv.second.dev.p) spilled += (double)kv.second.doubles;
  const double v[6] = {(double)ctx->cache.size(), (double)ctx->cache_device_doubles, spilled, (double)ctx->cache_puts, (double)ctx->cache_hits, (double)ctx->cache_evictions};
  for (int i = 0; i < n && i < 6; ++i) out[i] = v[i];
  return B2D_OK;
}

int b2d_cache_spill(b2d_ctx* ctx, double bytes) {
  NEED_DEVICE();
  CU(cudaSetDevice(ctx->device));
  release_device_memory(bytes > 0 ? (size_t)bytes : 0);
  return B2D_OK;
}

// ---- multi-GPU ----------------------------------------------------------------------------------------------------
int b2d_nccl_unique_id(uint8_t* id128) {
  static Nccl loader;
  std::string err;
  if (!loader.load(err)) { g_create_error = err; return B2D_ERR_NCCL; }
  int r = loader.GetUniqueId(id128);
  if (r != 0) { g_create_error = "ncclGetUniqueId failed"; return B2D_ERR_NCCL; }
  return B2D_OK;
}

int b2d_comm_init(b2d_ctx* ctx, const uint8_t* id128, int rank, int nranks) {
  NEED_DEVICE();
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(ctx, B2D_ERR_ARG, "b2d_comm_init: bad arguments");
  std::string err;
  if (!ctx->nccl.load(err)) return fail(ctx, B2D_ERR_NCCL, err);
  CU(cudaSetDevice(ctx->device));
  Nccl::Id128 id;
  memcpy(id.b, id128, 128);
  int r = ctx->nccl.CommInitRank(&ctx->nccl.comm, nranks, id, rank);
  if (r != 0) return fail(ctx, B2D_ERR_NCCL, std::string("ncclCommInitRank: ") + (ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(r) : "error"));
  return B2D_OK;
}

int b2d_allreduce_slot(b2d_ctx* ctx, int slot) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(slot);
  return allreduce(ctx, user_vec(ctx, slot), ctx->psi.Wp);
}

// ---- measurement --------------------------------------------------------------------------------------------------
int b2d_last_timing(b2d_ctx* ctx, double* out, int n) {
  NEED_DEVICE();
  if (!ctx->timing_valid || !out) return fail(ctx, B2D_ERR_ARG, "nothing timed yet");
  CU(cudaEventSynchronize(ctx->ev[1]));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  double v[4] = {ms, ctx->last_step_ms[0], ctx->last_step_ms[1], 0.0};
  for (int i = 0; i < n && i < 4; ++i) out[i] = v[i];
  return B2D_OK;
}
int b2d_sigma_profile(b2d_ctx* ctx, int src_slot, int dst_slot, double* out) {
  NEED_DEVICE(); NEED_PLAN(); CHECK_SLOT(src_slot); CHECK_SLOT(dst_slot);
  if (!out || src_slot == dst_slot) return fail(ctx, B2D_ERR_ARG, "b2d_sigma_profile: bad arguments");
  CU(cudaSetDevice(ctx->device));
  const bool saved = ctx->phase_timing;
  ctx->phase_timing = true;
  int rc = sigma_dev(ctx, user_vec(ctx, src_slot), user_vec(ctx, dst_slot), false, false);
  ctx->phase_timing = saved;
  if (rc) return rc;
  for (int st = 0; st < 2; ++st)
    for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) {
      double fl = 0.0, pad = 0.0;
      for (const Chunk& ch : ctx->sched.chunks) {
        const GemmBatch& b = st == 0 ? ch.step1 : ch.step2;
        fl += b.class_flops[c]; pad += b.class_padded[c];
      }
      double* o = out + (st * B2D_NUM_TILE_CLASSES + c) * 4;
      o[0] = ctx->class_ms[st][c]; o[1] = fl; o[2] = pad; o[3] = ctx->class_launches[st][c];
    }
  return B2D_OK;
}
int64_t b2d_kernel_launches(const b2d_ctx* ctx) { return ctx ? ctx->launches : -1; }
int b2d_sync(b2d_ctx* ctx) {
  NEED_DEVICE();
  CU(cudaStreamSynchronize(ctx->stream));
  return B2D_OK;
}
void* b2d_stream(b2d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int b2d_measure_fp64_peak(b2d_ctx* ctx, double* dmma_tflops, double* dfma_tflops) {
  NEED_DEVICE();
  CU(cudaSetDevice(ctx->device));
  double a = 0, b = 0;
  CU(measure_fp64(ctx->stream, &a, &b));
  if (dmma_tflops) *dmma_tflops = a;
  if (dfma_tflops) *dfma_tflops = b;
  return B2D_OK;
}

// Level-1 (HBM-bound) kernels of the Davidson iteration timed alone with CUDA events on the context's stream, on the
// wavefunction slots slot0 .. slot0+17 (contents are overwritten).  out[2k] = milliseconds per launch, out[2k+1] = ALGORITHMIC
// bytes per launch (8 bytes x vectors read or written) for k = 0 multi_dot(8 vectors) 1 rotate(8 -> 8)
// 2 residual 3 olsen 4 mgs_step 5 axpy 6 D2D copy (the yardstick).  Consecutive launches rotate over two disjoint sets of
// vectors so that a 126 MB L2 cannot serve the re-reads.
int b2d_measure_level1(b2d_ctx* ctx, int slot0, int reps, double* out) {
  NEED_DEVICE(); NEED_PLAN();
  if (reps < 1 || !out) return fail(ctx, B2D_ERR_ARG, "b2d_measure_level1: bad arguments");
  for (int i = 0; i < 18; ++i) CHECK_SLOT(slot0 + i);
  CU(cudaSetDevice(ctx->device));
  const int64_t n = ctx->psi.Wp;
  cudaStream_t st = ctx->stream;
  int64_t* L = &ctx->launches;
  double* partials = (double*)ctx->partials.p;
  double* sc = (double*)ctx->scalars.p;
  double* misc = sc + 2080;
  double* alpha = sc + 1056;
  {   // a well-conditioned 8x8 rotation (identity) and non-trivial vectors
    std::vector<double> a(1024, 0.0);
    for (int i = 0; i < 8; ++i) a[i * 32 + i] = 1.0;
    CU(cudaMemcpyAsync(alpha, a.data(), 1024 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    for (int i = 0; i < 18; ++i) {
      CU(cudaMemsetAsync(user_vec(ctx, slot0 + i), 0, (size_t)n * 8, st));
      CU(launch_fill_random(user_vec(ctx, slot0 + i), (const BlockDesc*)ctx->psi_blocks.p, ctx->psi.nblocks(), 77 + i, 0.5, st, L));
    }
  }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  auto V = [&](int set, int i) { return user_vec(ctx, slot0 + set * 9 + i); };
  auto run = [&](int k, double vectors, auto&& body) -> int {
    for (int w = 0; w < 2; ++w) { cudaError_t e = body(w & 1); if (e != cudaSuccess) return fail(ctx, B2D_ERR_CUDA, cudaGetErrorString(e)); }
    CU(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; ++r) { cudaError_t e = body(r & 1); if (e != cudaSuccess) return fail(ctx, B2D_ERR_CUDA, cudaGetErrorString(e)); }
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
    out[2 * k] = ms / reps; out[2 * k + 1] = vectors * 8.0 * (double)ctx->psi.W;
    return B2D_OK;
  };
  int rc;
  rc = run(0, 9, [&](int s) { VecList x; for (int i = 0; i < 8; ++i) x.p[i] = V(s, i); return launch_multi_dot(8, x, V(s, 8), n, partials, misc + 8, st, L); });
  if (rc) return rc;
  rc = run(1, 16, [&](int s) { VecList x; for (int i = 0; i < 8; ++i) x.p[i] = V(s, i); return launch_rotate(8, 8, x, alpha, 32, n, st, L); });
  if (rc) return rc;
  rc = run(2, 3, [&](int s) { return launch_residual(V(s, 0), V(s, 1), misc + 8, V(s, 2), n, partials, misc + 4, st, L); });
  if (rc) return rc;
  rc = run(3, 7, [&](int s) { return launch_olsen(V(s, 2), V(s, 1), V(s, 3), misc + 8, n, partials, misc, st, L); });
  if (rc) return rc;
  rc = run(4, 5, [&](int s) { return launch_mgs_step(V(s, 2), V(s, 1), n, partials, misc, st, L); });
  if (rc) return rc;
  rc = run(5, 3, [&](int s) { return launch_axpy(V(s, 4), V(s, 5), nullptr, 1e-3, n, st, L); });
  if (rc) return rc;
  rc = run(6, 2, [&](int s) { return cudaMemcpyAsync(V(s, 6), V(s, 7), (size_t)n * 8, cudaMemcpyDeviceToDevice, st); });
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

}  // extern "C"
