// block_gpu: the UNMODIFIED reference sweep (sanshar/Block 1.1.1, oracle/_ref/libblockref.a) with its hot-path entry points
// re-routed at LINK time (GNU ld --wrap) to libblockb200.so through the C ABI of include/block_b200.h.
//
// This is the reference-side binding of INTEGRATION.md made real and testable: no reference source is edited or copied,
// dmrg.C / sweep.C / renormalise.C / solver.C run as they are (input parsing, block construction, guess wavefunctions, disk
// checkpoints), and only the arithmetic of the path moves to the GPU:
//
//   SpinBlock::diagonalH              spinblock.C:855      (caller solver.C:34)      -> b2d_diagonal
//   Linear::block_davidson            linear.C:179         (caller solver.C:91)      -> b2d_davidson_lower (device-resident Krylov space)
//   SpinBlock::multiplyH              spinblock.C:722      (caller davidson.C:21)    -> b2d_multiplyH_host (only with B2D_DROPIN_DAVIDSON=host)
//   DensityMatrix::makedensitymatrix  density.C:27         (caller renormalise.C:104)-> b2d_make_density (+ b2d_add_onedot_noise)
//   diagonalise_dm                    rotationmat.C:258    (caller renormalise.C:145)-> b2d_diagonalise_dm
//   assign_matrix_by_dm               rotationmat.C:149    (caller renormalise.C:164)-> b2d_select_states + b2d_rotation_download
//   GuessWave::guess_wavefunctions    guess_wavefunction.C:378 (caller solver.C:77)  -> b2d_guess_plan + b2d_guess_transform (B2D_DROPIN_GUESS=device)
//   SpinBlock::transform_operators    save_load_block.C:267(caller sweep.C:279)      -> b2d_transform_operators; the reference's own
//                                     bookkeeping (new StateInfo, core flags, freeing the children) still runs, with its
//                                     MatrixRotate arithmetic (MatrixBLAS.C:553) switched off and the blocks filled from the device.
//
// TEST INFRASTRUCTURE: built by `make -C oracle dropin` into oracle/_ref/block_gpu (it contains reference objects, so it is
// git-ignored like the rest of oracle/_ref and travels to the GPU box with the snapshot); driven by tests/test_gpu_dropin.py
// and bench.py's sweep leg.  There is no CPU fallback: a mode the GPU path does not cover aborts with a message.
//
// Environment:
//   B2D_DEVICE            CUDA device (default 0)
//   B2D_DROPIN_CHECK=1    every hook ALSO runs the reference's CPU function on copies and prints the differences
//   B2D_DROPIN_DAVIDSON   "device" (default): block_davidson on the GPU;  "host": the reference's block_davidson with
//                         multiplyH through b2d_multiplyH_host (host buffers, the e2e form of the sigma call)
//   B2D_DROPIN_TRANSFORM  "device" (default): the transform hook does the block's bookkeeping itself;  "reference": it lets the
//                         reference's transform_operators do it with MatrixRotate switched off (re-builds virtual operators on the CPU)
//   B2D_DROPIN_WORKSPACE_MB   T workspace of the two-step contraction
//   B2D_DROPIN_OPBUILD    "device" (default): a child of the big block that is an enlarged block is built on the GPU from ITS children
//                         (SURVEY N2);  "host": the reference's Op::build constructs it on the CPU and it is uploaded
//   B2D_DROPIN_GUESS      "device" (default): TRANSFORM / TRANSPOSE guesses (GuessWave::transform_previous_wavefunction and its one-dot /
//                         transpose forms) are computed on the GPU (b2d_guess_plan + b2d_guess_transform);  "host": the reference's own
//   B2D_DROPIN_CACHE      "device" (default): renormalised blocks stay on the GPU between block iterations (b2d_cache_*); the host copy
//                         the reference keeps (and writes to its scratch files) holds zeros and, in the first element of every sector block, a
//                         NaN-boxed token that names the device entry - it travels through the reference's own copies / store / restore.
//                         "host": every renormalised operator is downloaded after the rotation and uploaded again (round-1 behaviour; needed
//                         for restartable scratch files).  Check mode and state-specific runs always use "host".
//   B2D_DROPIN_EIG        "host": diagnostic - the density-matrix eigen-decomposition and state selection stay with the reference
//                         (dsyev_), everything else on the GPU: separates eigenvector non-uniqueness from arithmetic differences
//   B2D_DROPIN_OPTIONS    "key=value,..." library options (b2d_set_option), e.g. eig_jacobi_max=512; factorised=0|1 (default: per block iteration by size,
//                         B2D_DROPIN_FACTORISED_MIN_STATES)
//   B2D_DROPIN_STATS      file that receives one line per block iteration (timings, flops, H applications)
//   RANK / WORLD_SIZE / LOCAL_RANK (torchrun) + B2D_NCCL_ID_FILE   several processes, one GPU each, run the SAME sweep: the operator terms of
//                         multiplyH / diagonalH and the noise operators are partitioned over the ranks (b2d_plan(rank, nranks), cost-weighted
//                         ownership) and the partial results all-reduced over NCCL inside the library; everything else is replicated (bit-identical
//                         inputs on every rank - run the host side single-threaded).  Rank 0 writes the NCCL unique id to B2D_NCCL_ID_FILE.
#include <sys/time.h>
#include <unistd.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "wrap_syms_gpu.h"

#include "spinblock.h"
#include "wavefunction.h"
#include "density.h"
#include "rotationmat.h"
#include "solver.h"
#include "linear.h"
#include "davidson.h"
#include "global.h"
#include "input.h"
#include "operatorfunctions.h"
#include "MatrixBLAS.h"
#include "guess_wavefunction.h"

#include "block_b200.h"
#include "guess_binding.hpp"

using namespace SpinAdapted;
using std::string;
using std::vector;

namespace {

double now_s() { timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }
bool env_on(const char* k) { const char* v = getenv(k); return v && *v && strcmp(v, "0") != 0; }

[[noreturn]] void die(const string& msg) {
  fprintf(stderr, "block_gpu: %s\n", msg.c_str());
  fflush(stderr);
  abort();   // the reference's error behaviour (MatrixBLAS.C:491-495)
}

struct OpRef { SparseMatrix* elem; int id; };

// B2D_DROPIN_TIMING=1: wall time per phase of the hooks, printed at exit (diagnostic)
struct PhaseTimes {
  std::map<string, double> t;
  ~PhaseTimes() {
    if (!getenv("B2D_DROPIN_TIMING")) return;
    for (std::map<string, double>::iterator it = t.begin(); it != t.end(); ++it) fprintf(stderr, "B2D_TIMING %-28s %.3f s\n", it->first.c_str(), it->second);
  }
} g_phase;
struct Phase {
  const char* name; double t0;
  explicit Phase(const char* n) : name(n), t0(now_s()) {}
  ~Phase() { g_phase.t[name] += now_s() - t0; }
};

// ---- token that names a device-resident block inside the reference's host copy --------------------------------------
// A quiet NaN with a recognisable payload: anything on the host that consumed the "matrices" numerically would turn into NaN and fail
// loudly instead of silently using zeros.
const uint64_t TOKEN_MAGIC = 0x7FF8B2Dull;                   // bits 63..36 (28 bits): exponent all ones + quiet bit + a recognisable payload
double token_to_double(uint64_t id) { uint64_t b = (TOKEN_MAGIC << 36) | (id & 0xFFFFFFFFFull); double d; memcpy(&d, &b, 8); return d; }
bool double_to_token(double d, uint64_t* id) { uint64_t b; memcpy(&b, &d, 8); if ((b >> 36) != TOKEN_MAGIC) return false; *id = b & 0xFFFFFFFFFull; return true; }
bool cache_enabled() {
  const char* v = getenv("B2D_DROPIN_CACHE");
  if (v && string(v) == "host") return false;
  if (getenv("B2D_DROPIN_CHECK") && strcmp(getenv("B2D_DROPIN_CHECK"), "0") != 0) return false;   // check mode compares real host matrices
  if (dmrginp.setStateSpecific()) return false;               // several blocks with the same sites (one per state)
  if (!dmrginp.direct()) return false;                        // non-direct mode builds every enlarged operator from the host copies
  return true;
}
std::map<vector<int>, uint64_t> g_latest_token;               // sites of a renormalised block -> its newest device entry (older ones are dropped)

// token of a block whose operators live in the device cache, or 0
uint64_t block_token(SpinBlock& b) {
  for (std::map<opTypes, boost::shared_ptr<Op_component_base> >::iterator it = b.ops.begin(); it != b.ops.end(); ++it) {
    Op_component_base& arr = *it->second;
    for (int i = 0; i < arr.get_size(); ++i) {
      vector<boost::shared_ptr<SparseMatrix> > vec = arr.get_local_element(i);
      for (size_t c = 0; c < vec.size(); ++c) {
        SparseMatrix& op = *vec[c];
        if (!op.get_built()) return 0;
        for (int a = 0; a < op.nrows(); ++a)
          for (int q = 0; q < op.ncols(); ++q)
            if (op.allowed(a, q) && op.operator_element(a, q).Storage() > 0) {
              uint64_t id = 0;
              return double_to_token(op.operator_element(a, q).Store()[0], &id) ? id : 0;
            }
      }
    }
  }
  return 0;
}

struct Gpu {
  b2d_ctx* ctx = 0;              // created once, b2d_reset between block iterations
  bool active = false;           // the context currently describes a big block
  const SpinBlock* left = 0;     // children of the big block this context was built for (RenormaliseFrom works on a COPY
  const SpinBlock* right = 0;    // of `big`, renormalise.C:79 `newbig = big`, which shares the children)
  vector<int> lsites, rsites;
  vector<OpRef> left_ops;        // every operator element of the left child, in upload order
  int nslots = 0;
  int64_t W = 0;
  bool rho_on_device = false;    // b2d_make_density ran for this block
  bool rot_on_device = false;    // b2d_select_states ran for this block
  bool in_transform = false;     // MatrixRotate is switched off
  // statistics of the current block iteration
  double t_build = 0;            // inside the reference's own Op::build for direct-mode virtual operators (host side)
  double t_guess = 0;
  double t_upload = 0, t_diag = 0, t_dav = 0, t_rho = 0, t_eig = 0, t_rot = 0, dav_dev_ms = 0, flops = 0;
  int nmult = 0, call = -1;
  long long launch0 = 0;         // kernel-launch counter of the (reused) context when this block iteration began
  bool dirty = false;            // statistics of this context not written yet
  bool integrals_set = false;    // b2d_set_integrals done (kept across b2d_reset)
  int children_on_device = 0;    // children of big blocks built on the device so far
  int rank = 0, world = 1;       // term partition over processes (one GPU each)
  int cache_uses = 0;            // renormalised blocks taken from the device cache instead of being uploaded
  // check mode: CPU results kept between hooks
  SparseMatrix* chk_transform = 0;
  vector<DiagonalMatrix> chk_eigs;
  DensityMatrix chk_vecs;
} g;

// The CUDA context (driver initialisation, streams, pinned buffers: 1-2 s) is created on a helper thread while the reference reads its
// input and builds its first blocks; the first hook joins it.
std::thread* g_early_thread = 0;
b2d_ctx* g_early_ctx = 0;
int g_early_rc = 0;
int dropin_device() {
  const int world = getenv("WORLD_SIZE") ? atoi(getenv("WORLD_SIZE")) : 1;
  return getenv("B2D_DEVICE") ? atoi(getenv("B2D_DEVICE")) : (world > 1 && getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : 0);
}
__attribute__((constructor)) void early_create() {
  if (getenv("B2D_DROPIN_NO_EARLY_INIT")) return;
  g_early_thread = new std::thread([]() { g_early_rc = b2d_create(dropin_device(), &g_early_ctx); });
}

void ck(int rc, const char* what) {
  if (rc) die(string(what) + ": " + b2d_last_error(g.ctx));
}

void flatten(const SparseMatrix& w, vector<double>& out) {   // Wavefunction::FlattenInto order, wavefunction.C:167-186
  out.clear();
  for (int l = 0; l < w.nrows(); ++l)
    for (int r = 0; r < w.ncols(); ++r)
      if (w.allowed(l, r)) {
        const Matrix& m = w.operator_element(l, r);
        out.insert(out.end(), m.Store(), m.Store() + m.Storage());
      }
}
void collect(SparseMatrix& w, const vector<double>& in) {     // Wavefunction::CollectFrom, wavefunction.C:188-203
  size_t off = 0;
  for (int l = 0; l < w.nrows(); ++l)
    for (int r = 0; r < w.ncols(); ++r)
      if (w.allowed(l, r)) {
        Matrix& m = w.operator_element(l, r);
        if (off + m.Storage() > in.size()) die("collect: size mismatch");
        memcpy(m.Store(), in.data() + off, sizeof(double) * m.Storage());
        off += m.Storage();
      }
  if (off != in.size()) die("collect: size mismatch");
}

double g_t_start = now_s();   // process start (static initialisation)
void write_stats();
void release() {
  if (g.ctx && g.dirty) write_stats();   // a context that is replaced before transform_operators (one-dot, dot on the environment side)
  if (g.ctx && b2d_reset(g.ctx)) die(string("b2d_reset: ") + b2d_last_error(g.ctx));   // one context for the whole run
  g.dirty = false;
  g.active = false; g.left = g.right = 0; g.left_ops.clear(); g.nslots = 0; g.rho_on_device = g.rot_on_device = false;
}

void write_stats() {
  const char* path = getenv("B2D_DROPIN_STATS");
  if (!path || !g.ctx || !g.active) return;
  FILE* f = fopen(path, "a");
  if (!f) return;
  fprintf(f, "call=%d lsites=%d rsites=%d W=%lld sigma_flops=%.6e n_multiply=%d host_op_build_s=%.6f upload_s=%.6f diag_s=%.6f davidson_s=%.6f davidson_dev_ms=%.3f "
             "density_s=%.6f eig_s=%.6f rotate_s=%.6f launches=%lld children_built_on_device=%d guess_s=%.6f cache_uses=%d\n",
          g.call, (int)g.lsites.size(), (int)g.rsites.size(), (long long)g.W, g.flops, g.nmult, g.t_build, g.t_upload, g.t_diag, g.t_dav, g.dav_dev_ms, g.t_rho,
          g.t_eig, g.t_rot, (long long)(b2d_kernel_launches(g.ctx) - g.launch0), g.children_on_device, g.t_guess, g.cache_uses);
  fclose(f);
}

// B2D_DROPIN_STOP_AFTER=<n>: a bounded run (profiles of configurations whose whole sweep does not fit the GPU budget) - after the n-th block
// iteration's statistics are written the process reports the cache occupancy and ends; never set in the parity tests
void maybe_stop() {
  static const int stop_after = getenv("B2D_DROPIN_STOP_AFTER") ? atoi(getenv("B2D_DROPIN_STOP_AFTER")) : 0;
  if (getenv("B2D_DROPIN_TIMING") && g.ctx) {
    double cs[5] = {0, 0, 0, 0, 0};
    b2d_cache_stats(g.ctx, cs, 5);
    fprintf(stderr, "B2D_PROGRESS call=%d t=%.1f s cache entries=%d device=%.2f GB pinned=%.2f GB\n", g.call, now_s() - g_t_start, (int)cs[0], cs[1] * 8e-9, cs[2] * 8e-9);
  }
  if (stop_after > 0 && g.call >= stop_after) {
    fprintf(stderr, "B2D_DROPIN_STOP_AFTER=%d reached: bounded run ends here\n", stop_after);
    fflush(stdout); fflush(stderr);
    _exit(0);
  }
}

void upload_block(int side, SpinBlock& b, vector<OpRef>* keep) {
  const StateInfo& s = b.get_stateInfo();
  int nq = (int)s.quanta.size();
  vector<int32_t> q(3 * nq), dims(nq);
  for (int i = 0; i < nq; ++i) {
    q[3 * i] = s.quanta[i].get_n(); q[3 * i + 1] = s.quanta[i].get_s().getirrep(); q[3 * i + 2] = s.quanta[i].get_symm().getirrep();
    dims[i] = s.quantaStates[i];
  }
  vector<int32_t> sites(b.get_sites().begin(), b.get_sites().end());
  if (uint64_t tok = block_token(b)) {
    // the block's operators never left the GPU (its host matrices hold the token): make the cached entry this child
    int32_t cnq = 0, cnops = 0, cns = 0;
    if (b2d_cache_block_info(g.ctx, tok, &cnq, &cnops, &cns)) die("device block cache: unknown token in a host block (scratch files of another run? use B2D_DROPIN_CACHE=host)");
    vector<int32_t> cq(3 * (size_t)cnq), cdims(cnq), csites(cns);
    ck(b2d_cache_block_sectors(g.ctx, tok, cq.data(), cdims.data(), csites.data()), "b2d_cache_block_sectors");
    if (cnq != nq || cq != q || cdims != dims || csites != sites) die("device block cache: the cached block's StateInfo differs from the host block's");
    { Phase ph("cache_use"); ck(b2d_cache_use(g.ctx, tok, side, b.is_loopblock() ? 1 : 0), "b2d_cache_use"); }
    Phase ph2("cache_verify");
    int k = 0;
    for (std::map<opTypes, boost::shared_ptr<Op_component_base> >::iterator it = b.ops.begin(); it != b.ops.end(); ++it) {
      Op_component_base& arr = *it->second;
      for (int i = 0; i < arr.get_size(); ++i) {
        vector<boost::shared_ptr<SparseMatrix> > vec = arr.get_local_element(i);
        for (size_t c = 0; c < vec.size(); ++c, ++k) {
          int32_t ty = -1, norb = -1, orbs[2] = {-1, -1}, comp = -1;
          if (b2d_cache_op_info(g.ctx, tok, k, &ty, &norb, orbs, &comp, 0)) die("device block cache: fewer operators than the host block");
          const std::vector<int>& ho = vec[c]->get_orbs();
          bool same = ty == (int)it->first && norb == (int)ho.size() && comp == (int)c;
          for (int t = 0; same && t < norb; ++t) same = orbs[t] == ho[t];
          if (!same) die("device block cache: operator order differs from the host block's");
          if (keep) keep->push_back(OpRef{vec[c].get(), k});
        }
      }
    }
    if (k != cnops) die("device block cache: operator count differs from the host block's");
    ++g.cache_uses;
    return;
  }
  ck(b2d_set_block(g.ctx, side, nq, q.data(), dims.data(), b.is_loopblock() ? 1 : 0, (int)sites.size(), sites.data()), "b2d_set_block");
  vector<uint8_t> allowed((size_t)nq * nq);
  vector<const double*> blocks;
  for (std::map<opTypes, boost::shared_ptr<Op_component_base> >::iterator it = b.ops.begin(); it != b.ops.end(); ++it) {
    Op_component_base& arr = *it->second;
    for (int i = 0; i < arr.get_size(); ++i) {
      vector<boost::shared_ptr<SparseMatrix> > vec = arr.get_local_element(i);
      for (size_t c = 0; c < vec.size(); ++c) {
        // direct-mode virtual operators are built here by the reference's own Op::build (host side of the seam)
        double tb = now_s();
        boost::shared_ptr<SparseMatrix> rep = vec[c]->getworkingrepresentation(&b);
        g.t_build += now_s() - tb;
        SparseMatrix& op = *rep;
        if (op.get_deltaQuantum_size() != 1) die("operator with several deltaQuantum components (non spin-adapted / BCS run): not covered");
        if (op.nrows() != nq || op.ncols() != nq) die("operator shape does not match the block's StateInfo");
        int32_t orbs[2] = {-1, -1};
        int norb = (int)op.get_orbs().size();
        if (norb > 2) die("operator with more than two orbital indices on the sweep path");
        for (int k = 0; k < norb; ++k) orbs[k] = op.get_orbs()[k];
        SpinQuantum dq = op.get_deltaQuantum(0);
        int32_t dqv[3] = {dq.get_n(), dq.get_s().getirrep(), dq.get_symm().getirrep()};
        blocks.clear();
        for (int a = 0; a < nq; ++a)
          for (int bq = 0; bq < nq; ++bq) {
            bool al = op.allowed(a, bq);
            allowed[(size_t)a * nq + bq] = al ? 1 : 0;
            if (al) blocks.push_back(op.operator_element(a, bq).Store());   // newmat Matrix: row-major, contiguous
          }
        if (blocks.empty()) blocks.push_back(0);
        int id = -1;
        ck(b2d_add_op_blocks(g.ctx, side, (int)it->first, norb, orbs, (int)c, dqv, op.get_fermion() ? 1 : 0, allowed.data(), blocks.data(), &id), "b2d_add_op_blocks");
        if (keep) keep->push_back(OpRef{vec[c].get(), id});
      }
    }
  }
}

bool opbuild_on_device() { const char* v = getenv("B2D_DROPIN_OPBUILD"); return !(v && string(v) == "host"); }

void set_integrals_once(const SpinBlock& big) {
  if (g.integrals_set) return;
  const int idx = big.get_integralIndex();
  const int n = (int)dmrginp.spin_orbs_symmetry().size() / 2;
  vector<double> h1((size_t)n * n), h2((size_t)n * n * n * n);
  vector<int32_t> irr(n);
  for (int i = 0; i < n; ++i) {
    irr[i] = SymmetryOfSpatialOrb(i).getirrep();
    for (int j = 0; j < n; ++j) h1[(size_t)i * n + j] = v_1[idx](2 * i, 2 * j);
  }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) for (int k = 0; k < n; ++k) for (int l = 0; l < n; ++l)
    h2[(((size_t)i * n + j) * n + k) * n + l] = v_2[idx](2 * i, 2 * j, 2 * k, 2 * l);
  ck(b2d_set_integrals(g.ctx, n, h1.data(), h2.data(), irr.data(), dmrginp.oneindex_screen_tol(), dmrginp.twoindex_screen_tol()), "b2d_set_integrals");
  g.integrals_set = true;
}

// a child of the big block that is itself an enlarged block (renormalised block x dot) is NOT built by the reference's
// Op::build on the CPU and uploaded; its two children (M-state operators + the one-site dot: ~16x less data) are uploaded and every
// operator of the child is built on the device (b2d_build_enlarged_op: host planner + kron_scatter_kernel).  Returns false - nothing
// done - when the block is not such a product or carries operator types outside the energy sweep: the caller uploads it as it is.
bool build_child_on_device(int slot, SpinBlock& b, vector<OpRef>* keep) {
  if (!b.get_leftBlock() || !b.get_rightBlock()) return false;
  if (b.get_leftBlock()->get_sites().empty() || b.get_rightBlock()->get_sites().empty()) return false;
  const StateInfo& si = b.get_stateInfo();
  if (!si.hasCollectedQuanta || !si.unCollectedStateInfo || !si.leftStateInfo || !si.rightStateInfo) return false;
  SpinBlock& cl = *b.get_leftBlock();
  SpinBlock& cr = *b.get_rightBlock();
  if (si.leftStateInfo->quanta.size() != cl.get_stateInfo().quanta.size() || si.rightStateInfo->quanta.size() != cr.get_stateInfo().quanta.size()) return false;
  if (si.leftStateInfo->quantaStates != cl.get_stateInfo().quantaStates || si.rightStateInfo->quantaStates != cr.get_stateInfo().quantaStates) return false;
  static const int known[] = {HAM, CRE, CRE_CRE, DES_DESCOMP, CRE_DES, CRE_DESCOMP, CRE_CRE_DESCOMP, OVERLAP};
  for (std::map<opTypes, boost::shared_ptr<Op_component_base> >::iterator it = b.ops.begin(); it != b.ops.end(); ++it) {
    bool ok = false;
    for (int t : known) ok = ok || t == (int)it->first;
    if (!ok) return false;
  }
  double tb0 = g.t_build;
  upload_block(0, cl, 0);          // the grandchildren: renormalised operators (core) and the dot
  upload_block(1, cr, 0);
  g.t_build = tb0;                 // getworkingrepresentation of core operators does not build anything
  int nq = (int)si.quanta.size();
  vector<int32_t> q(3 * nq), dims(nq), begin(1, 0), flat;
  for (int i = 0; i < nq; ++i) {
    q[3 * i] = si.quanta[i].get_n(); q[3 * i + 1] = si.quanta[i].get_s().getirrep(); q[3 * i + 2] = si.quanta[i].get_symm().getirrep();
    dims[i] = si.quantaStates[i];
    flat.insert(flat.end(), si.oldToNewState[i].begin(), si.oldToNewState[i].end());
    begin.push_back((int32_t)flat.size());
  }
  vector<int32_t> lmap(si.leftUnMapQuanta.begin(), si.leftUnMapQuanta.end()), rmap(si.rightUnMapQuanta.begin(), si.rightUnMapQuanta.end());
  vector<int32_t> unc(si.unCollectedStateInfo->quantaStates.begin(), si.unCollectedStateInfo->quantaStates.end());
  ck(b2d_set_product_stateinfo(g.ctx, nq, q.data(), dims.data(), (int)unc.size(), lmap.data(), rmap.data(), unc.data(), begin.data(), flat.data()),
     "b2d_set_product_stateinfo");
  const int hub = dmrginp.hamiltonian() == HUBBARD ? 1 : 0;
  const bool check = env_on("B2D_DROPIN_CHECK");
  double worst = 0, scale = 0;
  int nops = 0;
  vector<uint8_t> allowed((size_t)nq * nq);
  vector<double> data;
  for (std::map<opTypes, boost::shared_ptr<Op_component_base> >::iterator it = b.ops.begin(); it != b.ops.end(); ++it) {
    Op_component_base& arr = *it->second;
    for (int i = 0; i < arr.get_size(); ++i) {
      vector<boost::shared_ptr<SparseMatrix> > vec = arr.get_local_element(i);
      for (size_t c = 0; c < vec.size(); ++c) {
        SparseMatrix& op = *vec[c];
        if (op.get_deltaQuantum_size() != 1) die("operator with several deltaQuantum components (non spin-adapted / BCS run): not covered");
        int32_t orbs[2] = {-1, -1};
        int norb = (int)op.get_orbs().size();
        if (norb > 2) die("operator with more than two orbital indices on the sweep path");
        for (int k = 0; k < norb; ++k) orbs[k] = op.get_orbs()[k];
        SpinQuantum dq = op.get_deltaQuantum(0);
        int32_t dqv[3] = {dq.get_n(), dq.get_s().getirrep(), dq.get_symm().getirrep()};
        int id = -1;
        ck(b2d_build_enlarged_op(g.ctx, (int)it->first, norb, orbs, (int)c, dqv, op.get_fermion() ? 1 : 0, hub, &id), "b2d_build_enlarged_op");
        if (keep) keep->push_back(OpRef{vec[c].get(), id});
        ++nops;
        if (check) {   // the reference's own Op::build on the CPU
          boost::shared_ptr<SparseMatrix> rep = vec[c]->getworkingrepresentation(&b);
          int64_t n = b2d_product_op_size(g.ctx, id);
          data.resize((size_t)std::max<int64_t>(n, 1));
          ck(b2d_product_op_download(g.ctx, id, allowed.data(), data.data()), "b2d_product_op_download");
          size_t off = 0;
          for (int a = 0; a < nq; ++a)
            for (int bq = 0; bq < nq; ++bq) {
              if ((allowed[(size_t)a * nq + bq] != 0) != (bool)rep->allowed(a, bq)) die("device-built operator: allowed mask differs from the reference's");
              if (!rep->allowed(a, bq)) continue;
              const Matrix& m = rep->operator_element(a, bq);
              for (int e = 0; e < m.Storage(); ++e) { worst = std::max(worst, fabs(m.Store()[e] - data[off + e])); scale = std::max(scale, fabs(m.Store()[e])); }
              off += m.Storage();
            }
        }
      }
    }
  }
  if (check) fprintf(stderr, "B2D_CHECK call=%d opbuild side=%d ops=%d max_abs_diff=%.3e (max |O| %.3e)\n", g.call + 1, slot, nops, worst, scale);
  vector<int32_t> sites(b.get_sites().begin(), b.get_sites().end());
  ck(b2d_stash_product(g.ctx, slot, b.is_loopblock() ? 1 : 0, (int)sites.size(), sites.data()), "b2d_stash_product");
  ++g.children_on_device;
  return true;
}

// one context per big block (one block iteration); built at the first hook that sees the block (diagonalH, solver.C:34)
void ensure_ctx(const SpinBlock& big_c) {
  SpinBlock& big = const_cast<SpinBlock&>(big_c);
  if (!big.get_leftBlock() || !big.get_rightBlock()) die("big block without children");
  if (g.ctx && g.active && g.left == big.get_leftBlock() && g.right == big.get_rightBlock() && g.lsites == big.get_leftBlock()->get_sites() &&
      g.rsites == big.get_rightBlock()->get_sites())
    return;
  release();
  if (!dmrginp.spinAdapted()) die("non spin-adapted run: not covered by the GPU path");
  if (dmrginp.hamiltonian() != QUANTUM_CHEMISTRY && dmrginp.hamiltonian() != HUBBARD) die("Hamiltonian type not covered by the GPU path");
  double t0 = now_s();
  g.t_build = 0;
  g.world = getenv("WORLD_SIZE") ? atoi(getenv("WORLD_SIZE")) : 1;
  g.rank = getenv("RANK") ? atoi(getenv("RANK")) : 0;
  int dev = dropin_device();
  if (!g.ctx) {
    if (g_early_thread) {
      g_early_thread->join();
      delete g_early_thread; g_early_thread = 0;
      if (g_early_rc) die(string("b2d_create: ") + b2d_last_error(0));
      g.ctx = g_early_ctx;
    } else if (b2d_create(dev, &g.ctx)) die(string("b2d_create: ") + b2d_last_error(0));
    if (g.world > 1) {   // distribute.C's boost::mpi split of the operator terms -> one process per GPU, NCCL all-reduce of the partial sigma
      const char* idf = getenv("B2D_NCCL_ID_FILE");
      if (!idf) die("WORLD_SIZE > 1 needs B2D_NCCL_ID_FILE (a path every rank can read: rank 0 writes the NCCL unique id there)");
      uint8_t id[128];
      if (g.rank == 0) {
        if (b2d_nccl_unique_id(id)) die(string("b2d_nccl_unique_id: ") + b2d_last_error(0));
        string tmp = string(idf) + ".tmp";
        FILE* f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(id, 1, 128, f) != 128) die("cannot write B2D_NCCL_ID_FILE");
        fclose(f);
        rename(tmp.c_str(), idf);
      } else {
        double t0w = now_s();
        FILE* f = 0;
        while (!(f = fopen(idf, "rb"))) { if (now_s() - t0w > 300) die("timed out waiting for B2D_NCCL_ID_FILE"); usleep(20000); }
        if (fread(id, 1, 128, f) != 128) die("short read of B2D_NCCL_ID_FILE");
        fclose(f);
      }
      ck(b2d_comm_init(g.ctx, id, g.rank, g.world), "b2d_comm_init");
      ck(b2d_set_option(g.ctx, "balance_terms", 1.0), "b2d_set_option");
      // every rank holds the whole block here: the eigen-decomposition (sectors) and the operator rotation (operators) are divided too
      if (!getenv("B2D_DROPIN_REPLICATE_RENORM")) ck(b2d_set_option(g.ctx, "partition_renormalisation", 1.0), "b2d_set_option");
    }
  }
  g.active = true;
  g.launch0 = b2d_kernel_launches(g.ctx);
  if (getenv("B2D_DROPIN_WORKSPACE_MB")) ck(b2d_set_option(g.ctx, "workspace_mb", atof(getenv("B2D_DROPIN_WORKSPACE_MB"))), "b2d_set_option");
  // default of the drop-in, per block iteration: FACTORISED enlarged-block operators (no enlarged operator is materialised; 16 x
  // smaller in memory, measured faster from M ~ 1000 on) when an enlarged child has at least B2D_DROPIN_FACTORISED_MIN_STATES states
  // (default 3000), materialised ones below (less host planning per block iteration).  B2D_DROPIN_OPTIONS="factorised=0|1" forces one form.
  {
    static const int min_states = getenv("B2D_DROPIN_FACTORISED_MIN_STATES") ? atoi(getenv("B2D_DROPIN_FACTORISED_MIN_STATES")) : 3000;
    const int states = std::max(big.get_leftBlock()->get_stateInfo().totalStates, big.get_rightBlock()->get_stateInfo().totalStates);
    ck(b2d_set_option(g.ctx, "factorised", states >= min_states ? 1.0 : 0.0), "b2d_set_option");
  }
  if (getenv("B2D_DROPIN_OPTIONS")) {   // "key=value,key=value" -> b2d_set_option
    string all = getenv("B2D_DROPIN_OPTIONS");
    size_t pos = 0;
    while (pos < all.size()) {
      size_t end = all.find(',', pos); if (end == string::npos) end = all.size();
      string kv = all.substr(pos, end - pos);
      size_t eq = kv.find('=');
      if (eq != string::npos) ck(b2d_set_option(g.ctx, kv.substr(0, eq).c_str(), atof(kv.substr(eq + 1).c_str())), "b2d_set_option");
      pos = end + 1;
    }
  }
  g.left = big.get_leftBlock(); g.right = big.get_rightBlock(); g.lsites = big.get_leftBlock()->get_sites(); g.rsites = big.get_rightBlock()->get_sites();
  // each child: built on the device from ITS children where possible, otherwise the reference builds it and it is uploaded
  for (int slot = 0; slot < 2; ++slot) {
    SpinBlock& child = slot == 0 ? *big.get_leftBlock() : *big.get_rightBlock();
    vector<OpRef>* keep = slot == 0 ? &g.left_ops : 0;
    bool on_device = false;
    if (opbuild_on_device()) {
      set_integrals_once(big);
      on_device = build_child_on_device(slot, child, keep);
    }
    if (!on_device) {
      if (!block_token(child) && ((child.get_leftBlock() && block_token(*child.get_leftBlock())) || (child.get_rightBlock() && block_token(*child.get_rightBlock()))))
        die("a block that cannot be built on the device has device-cached children (their host matrices hold tokens): run with B2D_DROPIN_CACHE=host");
      upload_block(0, child, keep);
      ck(b2d_stash_side(g.ctx, slot, 0), "b2d_stash_side");
    }
  }
  ck(b2d_assemble_big(g.ctx), "b2d_assemble_big");
  SpinQuantum tq = dmrginp.effective_molecule_quantum();
  int32_t dq[3] = {tq.get_n(), tq.get_s().getirrep(), tq.get_symm().getirrep()};
  int norbs = (int)dmrginp.spin_orbs_symmetry().size() / 2;
  ck(b2d_plan(g.ctx, dq, coreEnergy[big.get_integralIndex()], dmrginp.hamiltonian() == HUBBARD ? 1 : 0, norbs, g.rank, g.world), "b2d_plan");
  g.W = b2d_psi_size(g.ctx);
  g.flops = b2d_sigma_flops(g.ctx, 1);
  g.t_upload = now_s() - t0 - g.t_build;
  g.t_diag = g.t_dav = g.t_rho = g.t_eig = g.t_rot = g.dav_dev_ms = g.t_guess = 0; g.nmult = 0;
  g.dirty = true;
  ++g.call;
}

void need_slots(int n) {
  if (n > g.nslots) { ck(b2d_vec_reserve(g.ctx, n), "b2d_vec_reserve"); g.nslots = n; }
}

void upload_wave(int slot, const SparseMatrix& w) {
  vector<double> flat; flatten(w, flat);
  if ((int64_t)flat.size() != g.W) die("wavefunction size differs from the planned psi layout");
  ck(b2d_vec_upload(g.ctx, slot, flat.data()), "b2d_vec_upload");
}
void download_wave(int slot, SparseMatrix& w) {
  vector<double> flat((size_t)g.W);
  ck(b2d_vec_download(g.ctx, slot, flat.data()), "b2d_vec_download");
  collect(w, flat);
}

double max_abs_diff(const SparseMatrix& a, const SparseMatrix& b, double* scale = 0) {
  vector<double> x, y; flatten(a, x); flatten(b, y);
  if (x.size() != y.size()) return 1e300;
  double d = 0, s = 0;
  for (size_t i = 0; i < x.size(); ++i) { d = std::max(d, fabs(x[i] - y[i])); s = std::max(s, fabs(y[i])); }
  if (scale) *scale = s;
  return d;
}

const int DIAG_SLOT = 0, ROOT_SLOT0 = 1;

}  // namespace

// ---------------- wrapped entry points ----------------
namespace SpinAdapted {

// ---- SpinBlock::diagonalH ----
void real_diagonalH(const SpinBlock* self, DiagonalMatrix& e) asm("__real_" SYM_diagonalH);
void wrap_diagonalH(const SpinBlock* self, DiagonalMatrix& e) asm("__wrap_" SYM_diagonalH);
void wrap_diagonalH(const SpinBlock* self, DiagonalMatrix& e) {
  ensure_ctx(*self);
  double t0 = now_s();
  if ((int64_t)e.Ncols() != g.W) die("diagonalH: DiagonalMatrix length differs from the psi layout");
  need_slots(ROOT_SLOT0 + 1);
  ck(b2d_diagonal(g.ctx, DIAG_SLOT), "b2d_diagonal");
  ck(b2d_vec_download(g.ctx, DIAG_SLOT, e.Store()), "b2d_vec_download");
  g.t_diag += now_s() - t0;
  if (env_on("B2D_DROPIN_CHECK")) {
    DiagonalMatrix e2; e2.ReSize(e.Ncols()); e2 = 0;
    real_diagonalH(self, e2);
    double d = 0; for (int i = 0; i < e.Ncols(); ++i) d = std::max(d, fabs(e.element(i) - e2.element(i)));
    fprintf(stderr, "B2D_CHECK call=%d diagonalH max_abs_diff=%.3e\n", g.call, d);
  }
}

// ---- GuessWave::guess_wavefunctions (vector form, solver.C:77) ----
// SURVEY N1.  For a TRANSFORM or TRANSPOSE guess (two-dot :524-636 / :55-84, one-dot :832-936 / :140-198) the reference loads the previous wavefunction and two rotation matrices from its
// scratch files and runs TransformLeftBlock / onedot_shufflesysdot / TransformRightBlock on the CPU (guess_wavefunction.C:524-636).  Here
// the same files are loaded the same way, the StateInfo tables the transform reads are handed to b2d_guess_plan, and the arithmetic runs
// on the device.  BASIC / TRANSPOSE guesses go to the reference's own function.
void real_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) asm("__real_" SYM_guess_wavefunctions);
void wrap_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) asm("__wrap_" SYM_guess_wavefunctions);
void wrap_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) {
  const char* mode = getenv("B2D_DROPIN_GUESS");
  if ((mode && string(mode) == "host") || !g.ctx || gw == BASIC) { real_guess(solution, e, big, gw, onedot, transpose_guess_wave, additional_noise, currentState); return; }
  double t0 = now_s();
  // every root must be one of the covered forms, otherwise the whole call goes to the reference (it owns the loop over roots)
  vector<b2d_binding::GuessBinding> B(solution.size());
  for (size_t i = 0; i < solution.size(); ++i) {
    const int state = (dmrginp.setStateSpecific() || dmrginp.calc_type() == COMPRESS || dmrginp.calc_type() == MPS_NEVPT) ? currentState : (int)i;   // :383
    if (!b2d_binding::make_guess_binding(B[i], big, gw, onedot, transpose_guess_wave, state)) {
      real_guess(solution, e, big, gw, onedot, transpose_guess_wave, additional_noise, currentState);
      return;
    }
  }
  for (size_t i = 0; i < solution.size(); ++i) {
    solution[i].initialise(dmrginp.effective_molecule_quantum_vec(), &big, onedot);     // guess_wavefunction.C:282
    double info[8];
    ck(b2d_guess_plan(g.ctx, &B[i].d, info, 8), "b2d_guess_plan");
    vector<double> flat((size_t)info[3]);
    if ((int64_t)flat.size() != g.W) die("guess transform: trial vector length differs from the psi layout");
    ck(b2d_guess_transform(g.ctx, B[i].old.data(), B[i].lrot.data(), B[i].rrot.data(), -1, flat.data()), "b2d_guess_transform");
    collect(solution[i], flat);
  }
  if (env_on("B2D_DROPIN_CHECK")) {   // the reference's own transform of every root (its loop maps root i to state i itself)
    vector<Wavefunction> ref(solution.size());
    real_guess(ref, e, big, gw, onedot, transpose_guess_wave, additional_noise, currentState);
    for (size_t i = 0; i < solution.size(); ++i) {
      vector<double> rf, gf; flatten(ref[i], rf); flatten(solution[i], gf);
      double worst = 0, scale = 0;
      for (size_t k = 0; k < rf.size() && k < gf.size(); ++k) { worst = std::max(worst, fabs(rf[k] - gf[k])); scale = std::max(scale, fabs(rf[k])); }
      fprintf(stderr, "B2D_CHECK call=%d guess_transform mode=%d root=%d max_abs_diff=%.3e (max |psi| %.3e)\n", g.call, (int)B[i].d.mode, (int)i, worst, scale);
    }
  }
  g.t_guess += now_s() - t0;
}

// ---- SpinBlock::multiplyH ----
void real_multiplyH(const SpinBlock* self, Wavefunction& c, Wavefunction* v, int num_threads) asm("__real_" SYM_multiplyH);
void wrap_multiplyH(const SpinBlock* self, Wavefunction& c, Wavefunction* v, int num_threads) asm("__wrap_" SYM_multiplyH);
void wrap_multiplyH(const SpinBlock* self, Wavefunction& c, Wavefunction* v, int num_threads) {
  ensure_ctx(*self);
  vector<double> cf, vf; flatten(c, cf); flatten(*v, vf);
  if ((int64_t)cf.size() != g.W || (int64_t)vf.size() != g.W) die("multiplyH: wavefunction size differs from the psi layout");
  need_slots(ROOT_SLOT0 + 1);
  ck(b2d_multiplyH_host(g.ctx, cf.data(), vf.data(), 1), "b2d_multiplyH_host");   // v += H c (linear.C:239-253)
  ++g.nmult;
  if (env_on("B2D_DROPIN_CHECK")) {
    Wavefunction v2 = *v;
    real_multiplyH(self, c, &v2, num_threads);
    collect(*v, vf);
    double s; double d = max_abs_diff(*v, v2, &s);
    fprintf(stderr, "B2D_CHECK call=%d multiplyH max_abs_diff=%.3e (max |sigma| %.3e)\n", g.call, d, s);
    return;
  }
  collect(*v, vf);
}

// ---- Linear::block_davidson ----
void real_davidson(vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp, Davidson_functor& h_multiply, bool& useprecond,
                   int currentRoot, vector<Wavefunction>& lowerStates) asm("__real_" SYM_block_davidson);
void wrap_davidson(vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp, Davidson_functor& h_multiply, bool& useprecond,
                   int currentRoot, vector<Wavefunction>& lowerStates) asm("__wrap_" SYM_block_davidson);
void wrap_davidson(vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp, Davidson_functor& h_multiply, bool& useprecond,
                   int currentRoot, vector<Wavefunction>& lowerStates) {
  const char* mode = getenv("B2D_DROPIN_DAVIDSON");
  // outside the device solver's limits (b2d_davidson: nroots <= deflation_min < deflation_max <= 30) the reference's own block_davidson
  // runs with every H application on the GPU - a valid Block input never aborts because of the subspace size
  const bool beyond_limits = dmrginp.deflation_max_size() > 30 || (int)b.size() > dmrginp.deflation_min_size() || dmrginp.deflation_max_size() <= dmrginp.deflation_min_size();
  if ((mode && string(mode) == "host") || beyond_limits) {   // the reference's Davidson; every H application crosses the boundary with host buffers
    double t0 = now_s();
    real_davidson(b, h_diag, normtol, warmUp, h_multiply, useprecond, currentRoot, lowerStates);
    g.t_dav += now_s() - t0;
    return;
  }
  ensure_ctx(h_multiply.get_block());
  if (!useprecond) die("block_davidson without the Olsen preconditioner: not covered");
  int nroots = (int)b.size(), nlow = (int)lowerStates.size();
  if ((int64_t)h_diag.Ncols() != g.W) die("block_davidson: h_diag length differs from the psi layout");
  vector<Wavefunction> b_in; DiagonalMatrix h_in;
  bool check = env_on("B2D_DROPIN_CHECK");
  if (check) { b_in = b; h_in = h_diag; }
  double t0 = now_s();
  need_slots(ROOT_SLOT0 + nroots + nlow);
  ck(b2d_vec_upload(g.ctx, DIAG_SLOT, h_diag.Store()), "b2d_vec_upload(diag)");
  for (int i = 0; i < nroots; ++i) upload_wave(ROOT_SLOT0 + i, b[i]);
  for (int i = 0; i < nlow; ++i) upload_wave(ROOT_SLOT0 + nroots + i, lowerStates[i]);
  vector<double> ev(nroots);
  int nm = 0; double res = 0;
  ck(b2d_davidson_lower(g.ctx, nroots, ROOT_SLOT0, DIAG_SLOT, normtol, dmrginp.deflation_min_size(), dmrginp.deflation_max_size(), nlow,
                        ROOT_SLOT0 + nroots, ev.data(), &nm, &res), "b2d_davidson");
  double tm[4] = {0, 0, 0, 0};
  b2d_last_timing(g.ctx, tm, 4);
  g.dav_dev_ms += tm[0];
  b.resize(nroots);
  for (int i = 0; i < nroots; ++i) download_wave(ROOT_SLOT0 + i, b[i]);
  for (int i = 0; i < std::min(nroots, h_diag.Ncols()); ++i) h_diag.element(i) = ev[i];   // linear.C:344-345
  g.nmult += nm;
  g.t_dav += now_s() - t0;
  if (check) {
    real_davidson(b_in, h_in, normtol, warmUp, h_multiply, useprecond, currentRoot, lowerStates);
    for (int i = 0; i < nroots; ++i) {
      double ov = DotProduct(b[i], b_in[i]);
      fprintf(stderr, "B2D_CHECK call=%d davidson root=%d E_gpu=%.12f E_cpu=%.12f dE=%.3e |<gpu|cpu>|=%.12f n_multiply_gpu=%d\n", g.call, i, ev[i],
              h_in.element(i), ev[i] - h_in.element(i), fabs(ov), nm);
    }
  }
}

// ---- DensityMatrix::makedensitymatrix ----
void real_makedm(DensityMatrix* self, const vector<Wavefunction>& ws, SpinBlock& big, const vector<double>& wts, const double noise, const double add_noise, bool warmup) asm("__real_" SYM_makedensitymatrix);
void wrap_makedm(DensityMatrix* self, const vector<Wavefunction>& ws, SpinBlock& big, const vector<double>& wts, const double noise, const double add_noise, bool warmup) asm("__wrap_" SYM_makedensitymatrix);
void wrap_makedm(DensityMatrix* self, const vector<Wavefunction>& ws, SpinBlock& big, const vector<double>& wts, const double noise, const double add_noise, bool warmup) {
  ensure_ctx(big);
  DensityMatrix chk;
  bool check = env_on("B2D_DROPIN_CHECK");
  if (check && add_noise > NUMERICAL_ZERO) check = false;   // the CPU copy would consume the rand() stream a second time
  if (check) { chk = *self; real_makedm(&chk, ws, big, wts, noise, add_noise, warmup); }
  double t0 = now_s();
  int nroots = (int)wts.size();     // density.C:31: one term per weight
  if ((int)ws.size() < nroots) die("makedensitymatrix: fewer wavefunctions than weights");
  need_slots(ROOT_SLOT0 + std::max(nroots, (int)ws.size()));
  for (size_t i = 0; i < ws.size(); ++i) upload_wave(ROOT_SLOT0 + (int)i, ws[i]);
  ck(b2d_make_density(g.ctx, nroots, ROOT_SLOT0, wts.data()), "b2d_make_density");
  if (noise > NUMERICAL_ZERO) ck(b2d_add_onedot_noise(g.ctx, (int)ws.size(), ROOT_SLOT0, noise), "b2d_add_onedot_noise");   // density.C:40-60
  if (add_noise > NUMERICAL_ZERO) {
    // DensityMatrix::add_twodot_noise (density.C:92-165), once per root with (1.0*noise)/nroots (density.C:71-80): random
    // wavefunctions in the 12 neighbouring (N +- 1|2, S +- 1|2) targets, drawn HERE with the reference's own
    // Wavefunction::Randomise so that the glibc rand() stream is the unmodified run's; the products w w^T run on the device
    if (dmrginp.hamiltonian() == BCS) die("BCS: not covered");
    const int N = dmrginp.total_particle_number(), S = dmrginp.total_spin_number().getirrep();
    const IrrepSpace& sym = dmrginp.total_symmetry_number();
    vector<SpinQuantum> toadd;
    const int one[2] = {+1, -1}, two[3] = {+2, 0, -2};
    for (int k = 0; k < 2; ++k) toadd.push_back(SpinQuantum(N + one[k], SpinSpace(S + 1), sym));
    if (S >= 1) for (int k = 0; k < 2; ++k) toadd.push_back(SpinQuantum(N + one[k], SpinSpace(S - 1), sym));
    for (int k = 0; k < 3; ++k) toadd.push_back(SpinQuantum(N + two[k], SpinSpace(S + 2), sym));
    toadd.push_back(SpinQuantum(N + 2, SpinSpace(S), sym));
    toadd.push_back(SpinQuantum(N - 2, SpinSpace(S), sym));
    if (S >= 2) for (int k = 0; k < 3; ++k) toadd.push_back(SpinQuantum(N + two[k], SpinSpace(S - 2), sym));
    vector<double> flat;
    for (size_t root = 0; root < ws.size(); ++root)
      for (size_t q = 0; q < toadd.size(); ++q) {
        Wavefunction w;
        w.initialise(toadd[q], &big, false);
        w.Randomise();
        double nrm = DotProduct(w, w);
        if (fabs(nrm) > NUMERICAL_ZERO) {
          Scale(1. / sqrt(nrm), w);
          flatten(w, flat);
          int32_t dq[3] = {toadd[q].get_n(), toadd[q].get_s().getirrep(), toadd[q].get_symm().getirrep()};
          if ((int64_t)flat.size() != b2d_wavefunction_size(g.ctx, dq)) die("add_twodot_noise: noise wavefunction size differs from the device layout");
          ck(b2d_add_wavefunction_density(g.ctx, dq, flat.data(), ((1.0 * noise) / ws.size()) / toadd.size()), "b2d_add_wavefunction_density");
        }
        w.CleanUp();
      }
  }
  vector<double> rho((size_t)b2d_density_size(g.ctx));
  ck(b2d_density_download(g.ctx, rho.data()), "b2d_density_download");
  size_t off = 0;
  for (int q = 0; q < self->nrows(); ++q) {
    if (!self->allowed(q, q)) die("density matrix without a diagonal block");
    Matrix& m = self->operator_element(q, q);
    if (off + m.Storage() > rho.size()) die("density matrix size mismatch");
    for (int k = 0; k < m.Storage(); ++k) m.Store()[k] += rho[off + k];   // MultiplyProduct accumulates (operatorfunctions.C:630-650)
    off += m.Storage();
  }
  if (off != rho.size()) die("density matrix size mismatch");
  g.rho_on_device = true;
  g.t_rho += now_s() - t0;
  if (check) {
    double s; double d = max_abs_diff(*self, chk, &s);
    fprintf(stderr, "B2D_CHECK call=%d makedensitymatrix noise=%.1e max_abs_diff=%.3e (max |rho| %.3e)\n", g.call, noise, d, s);
  }
}

// ---- diagonalise_dm ----
void real_diagdm(SparseMatrix& traced, SparseMatrix& transform, vector<DiagonalMatrix>& eigs) asm("__real_" SYM_diagonalise_dm);
void wrap_diagdm(SparseMatrix& traced, SparseMatrix& transform, vector<DiagonalMatrix>& eigs) asm("__wrap_" SYM_diagonalise_dm);
bool eig_on_host() { const char* v = getenv("B2D_DROPIN_EIG"); return v && string(v) == "host"; }
void wrap_diagdm(SparseMatrix& traced, SparseMatrix& transform, vector<DiagonalMatrix>& eigs) {
  if (eig_on_host()) { real_diagdm(traced, transform, eigs); return; }   // diagnostic: dsyev_ eigenvectors, everything else on the GPU
  if (!g.ctx || !g.active || !g.rho_on_device) die("diagonalise_dm outside a GPU block iteration (module not covered by the GPU path)");
  bool check = env_on("B2D_DROPIN_CHECK");
  if (check) {
    g.chk_vecs = static_cast<DensityMatrix&>(transform);
    g.chk_eigs.clear();
    real_diagdm(traced, g.chk_vecs, g.chk_eigs);
  }
  double t0 = now_s();
  // honour the argument: diagonalise the matrix we were handed (it is the one the makedensitymatrix hook produced)
  vector<double> rho;
  for (int q = 0; q < traced.nrows(); ++q) { const Matrix& m = traced.operator_element(q, q); rho.insert(rho.end(), m.Store(), m.Store() + m.Storage()); }
  if ((int64_t)rho.size() != b2d_density_size(g.ctx)) die("diagonalise_dm: density matrix size mismatch");
  ck(b2d_density_upload(g.ctx, rho.data()), "b2d_density_upload");
  int nq = traced.nrows();
  size_t total = 0; for (int q = 0; q < nq; ++q) total += traced.operator_element(q, q).Nrows();
  vector<double> ev(total);
  ck(b2d_diagonalise_dm(g.ctx, ev.data()), "b2d_diagonalise_dm");
  eigs.resize(nq);
  size_t off = 0;
  for (int q = 0; q < nq; ++q) {
    int d = traced.operator_element(q, q).Nrows();
    DiagonalMatrix w(d);
    for (int i = 0; i < d; ++i) w.element(i, i) = ev[off + i];   // ascending, < 1e-14 already clamped (rotationmat.C:274-276)
    eigs[q] = w;
    off += d;
  }
  g.t_eig += now_s() - t0;
  if (check) {
    double d = 0, rel_cut = 0; int near_cut = 0;
    for (int q = 0; q < nq; ++q) for (int i = 0; i < eigs[q].Nrows(); ++i) {
      double a = eigs[q].element(i, i), b = g.chk_eigs[q].element(i, i);
      d = std::max(d, fabs(a - b));
      if (b > 1e-14 && b < 1e-12) { ++near_cut; rel_cut = std::max(rel_cut, fabs(a - b) / b); }   // the 1e-13 keep threshold (rotationmat.C:161)
    }
    fprintf(stderr, "B2D_CHECK call=%d diagonalise_dm eigenvalue max_abs_diff=%.3e near_cut=%d near_cut_max_rel_diff=%.3e\n", g.call, d, near_cut, rel_cut);
  }
}

// ---- assign_matrix_by_dm ----
double real_assign(vector<Matrix>& rot, vector<DiagonalMatrix>& eigs, SparseMatrix& transform, vector<std::pair<int, int> >& inorder, vector<vector<int> >& byq,
                   int nbydm, int nbyq, int lsize, int rsize) asm("__real_" SYM_assign_matrix_by_dm);
double wrap_assign(vector<Matrix>& rot, vector<DiagonalMatrix>& eigs, SparseMatrix& transform, vector<std::pair<int, int> >& inorder, vector<vector<int> >& byq,
                   int nbydm, int nbyq, int lsize, int rsize) asm("__wrap_" SYM_assign_matrix_by_dm);
double wrap_assign(vector<Matrix>& rot, vector<DiagonalMatrix>& eigs, SparseMatrix& transform, vector<std::pair<int, int> >& inorder, vector<vector<int> >& byq,
                   int nbydm, int nbyq, int lsize, int rsize) {
  if (eig_on_host()) return real_assign(rot, eigs, transform, inorder, byq, nbydm, nbyq, lsize, rsize);
  if (!g.ctx || !g.active || !g.rho_on_device) die("assign_matrix_by_dm outside a GPU block iteration (module not covered by the GPU path)");
  if (nbyq != 0) die("keptqstates != 0: not covered (sweep_params.C:80 always passes 0)");
  if (dmrginp.do_pdm()) die("do_pdm keeps zero-weight states (rotationmat.C:161): not covered by the GPU path");
  double t0 = now_s();
  int nq = (int)eigs.size();
  vector<int32_t> kept(nq, 0);
  double discarded = 0;
  // nbydm = min(#eigenvalues, keptstates) (renormalise.C:156): the library applies the same min
  ck(b2d_select_states(g.ctx, nbydm, kept.data(), &discarded), "b2d_select_states");
  vector<double> flat((size_t)std::max<int64_t>(b2d_rotation_size(g.ctx), 1));
  ck(b2d_rotation_download(g.ctx, flat.data()), "b2d_rotation_download");
  rot.clear(); rot.resize(nq);
  size_t off = 0;
  for (int q = 0; q < nq; ++q) {
    int d = eigs[q].Nrows();
    if (kept[q] == 0) continue;
    rot[q].ReSize(d, kept[q]);
    memcpy(rot[q].Store(), flat.data() + off, sizeof(double) * (size_t)d * kept[q]);
    off += (size_t)d * kept[q];
  }
  g.rot_on_device = true;
  g.t_eig += now_s() - t0;
  if (env_on("B2D_DROPIN_CHECK")) {
    vector<Matrix> rot2;
    vector<std::pair<int, int> > in2; vector<vector<int> > by2;
    sort_weights(g.chk_eigs, in2, by2);
    double d2 = real_assign(rot2, g.chk_eigs, g.chk_vecs, in2, by2, std::min((int)in2.size(), nbydm), 0, lsize, rsize);
    int mism = 0; double proj = 0;
    for (int q = 0; q < nq; ++q) {
      if (rot2[q].Ncols() != rot[q].Ncols()) { ++mism; continue; }
      if (rot[q].Ncols() == 0) continue;
      Matrix pa = rot[q] * rot[q].t(), pb = rot2[q] * rot2[q].t();   // projectors: sign / degenerate-rotation invariant
      for (int k = 0; k < pa.Storage(); ++k) proj = std::max(proj, fabs(pa.Store()[k] - pb.Store()[k]));
    }
    fprintf(stderr, "B2D_CHECK call=%d select_states sectors_with_different_kept_count=%d discarded_gpu=%.6e discarded_cpu=%.6e projector_max_abs_diff=%.3e\n",
            g.call, mism, discarded, d2, proj);
  }
  return discarded;
}

// ---- MatrixRotate (switched off while the transform hook runs the reference's bookkeeping) ----
void real_rotate(const Matrix& a, const Matrix& b, const Matrix& c, Matrix& d) asm("__real_" SYM_MatrixRotate);
void wrap_rotate(const Matrix& a, const Matrix& b, const Matrix& c, Matrix& d) asm("__wrap_" SYM_MatrixRotate);
void wrap_rotate(const Matrix& a, const Matrix& b, const Matrix& c, Matrix& d) {
  if (g.in_transform && !env_on("B2D_DROPIN_CHECK")) return;
  real_rotate(a, b, c, d);
}

// ---- SpinBlock::transform_operators ----
void real_transform(SpinBlock* self, vector<Matrix>& rot) asm("__real_" SYM_transform_operators);
void wrap_transform(SpinBlock* self, vector<Matrix>& rot) asm("__wrap_" SYM_transform_operators);
void wrap_transform(SpinBlock* self, vector<Matrix>& rot) {
  if (!g.ctx || !g.active || g.lsites != self->get_sites())
    die("transform_operators on a block that was not the left child of the last GPU block iteration (warm-up / one-dot tail: not covered)");
  double t0 = now_s();
  int nq = (int)rot.size();
  // honour the argument: rotate with the matrices we were handed
  vector<int32_t> kept(nq);
  vector<double> flat;
  for (int q = 0; q < nq; ++q) {
    kept[q] = rot[q].Ncols();
    if (kept[q]) flat.insert(flat.end(), rot[q].Store(), rot[q].Store() + rot[q].Storage());
  }
  if (flat.empty()) flat.push_back(0.0);
  { Phase ph("rot_upload"); ck(b2d_rotation_upload(g.ctx, kept.data(), flat.data()), "b2d_rotation_upload"); }
  { Phase ph("transform_device"); ck(b2d_transform_operators(g.ctx), "b2d_transform_operators"); }
  bool check = env_on("B2D_DROPIN_CHECK");
  const char* tmode = getenv("B2D_DROPIN_TRANSFORM");
  if (check || (tmode && string(tmode) == "reference")) {
    // the reference's own bookkeeping (new StateInfo, allocation, core flags, freeing the children) with MatrixRotate switched
    // off.  It also re-BUILDS every virtual operator on the CPU (BaseOperator.C:369-375) only to ignore it: kept for check mode.
    g.in_transform = true;
    real_transform(self, rot);
    g.in_transform = false;
  } else {
    // the same bookkeeping done here (save_load_block.C:270-316), without building anything on the CPU: the un-rotated
    // operators are already on the device
    Phase ph("transform_host_bookkeeping");
    StateInfo before = self->braStateInfo;
    vector<SpinQuantum> nquanta; vector<int> nstates, nmap;
    for (int q = 0; q < nq; ++q)
      if (kept[q] != 0) { nquanta.push_back(before.quanta[q]); nstates.push_back(kept[q]); nmap.push_back(q); }
    StateInfo after(nquanta, nstates, nmap);
    const bool token_only = !check && !(tmode && string(tmode) == "reference") && cache_enabled();
    for (size_t k = 0; k < g.left_ops.size(); ++k) {
      SparseMatrix& op = *g.left_ops[k].elem;
      if (token_only) {
        // the matrices stay on the device, so the host copy needs the allowed mask (SparseMatrix::allocate, BaseOperator.C:123-145)
        // but no storage: every allowed block is a 1 x 1 matrix that will carry the cache token.  The reference's store / restore / copies
        // then move a few bytes per block instead of the renormalised operators, and nothing on the host can mistake them for data
        // (any arithmetic on a token is NaN).
        const int n = (int)after.quanta.size();
        op.resize(n, n);
        for (int a = 0; a < n; ++a)
          for (int b = 0; b < n; ++b) {
            bool al = false;
            for (int q = 0; q < op.get_deltaQuantum_size() && !al; ++q) al = after.quanta[a].allow(op.get_deltaQuantum(q), after.quanta[b]);
            op.allowed(a, b) = al;
            if (al) { op.operator_element(a, b).ReSize(1, 1); op.operator_element(a, b).Store()[0] = 0.0; }
          }
      } else {
        op.allocate(after);       // allowed mask + zeroed blocks on the retained sectors (BaseOperator.C:123-145)
      }
      op.set_built() = true;
    }
    self->braStateInfo = after;
    self->braStateInfo.AllocatePreviousStateInfo();
    *self->braStateInfo.previousStateInfo = before;
    self->ketStateInfo = self->braStateInfo;
    for (std::map<opTypes, boost::shared_ptr<Op_component_base> >::iterator it = self->ops.begin(); it != self->ops.end(); ++it)
      if (!it->second->is_core()) it->second->set_core(true);
    self->direct = false;
    if (self->leftBlock) self->leftBlock->clear();
    if (self->rightBlock) self->rightBlock->clear();
  }
  int nnew = b2d_rotated_num_sectors(g.ctx);
  if (nnew != (int)self->get_stateInfo().quanta.size()) die("transform_operators: retained sector count differs from the reference's StateInfo");
  if (!check && !(tmode && string(tmode) == "reference") && cache_enabled()) {
    // the rotated operators stay on the device; the host copy keeps its zeroed blocks and gets the entry's token in the first
    // element of every sector block (SpinBlock::store / restore and the reference's copies carry it along)
    uint64_t tok = 0;
    { Phase ph("cache_put"); ck(b2d_cache_put_rotated(g.ctx, &tok), "b2d_cache_put_rotated"); }
    Phase ph3("cache_tag_drop");
    const double tag = token_to_double(tok);
    for (size_t k = 0; k < g.left_ops.size(); ++k) {
      SparseMatrix& op = *g.left_ops[k].elem;
      if ((int)k != g.left_ops[k].id) die("transform_operators: operator ids are not in upload order");
      for (int a = 0; a < nnew; ++a)
        for (int b = 0; b < nnew; ++b)
          if (op.allowed(a, b) && op.operator_element(a, b).Storage() > 0) op.operator_element(a, b).Store()[0] = tag;
    }
    std::map<vector<int>, uint64_t>::iterator old = g_latest_token.find(self->get_sites());
    if (old != g_latest_token.end()) ck(b2d_cache_drop(g.ctx, old->second), "b2d_cache_drop");   // the block this one replaces (same sites, previous sweep)
    g_latest_token[self->get_sites()] = tok;
    g.t_rot += now_s() - t0;
    write_stats();
    g.dirty = false;
    release();
    maybe_stop();
    return;
  }
  double worst = 0, scale = 0;
  vector<uint8_t> allowed((size_t)nnew * nnew);
  vector<double> data((size_t)std::max<int64_t>(b2d_rotated_total_size(g.ctx), 1));
  { Phase ph("rotated_download"); ck(b2d_rotated_download_all(g.ctx, data.data()), "b2d_rotated_download_all"); }   // one device pass + one copy for every operator
  Phase ph4("rotated_copy_to_host_blocks");
  size_t off = 0;
  for (size_t k = 0; k < g.left_ops.size(); ++k) {
    SparseMatrix& op = *g.left_ops[k].elem;
    int id = g.left_ops[k].id;
    if ((int)k != id) die("transform_operators: operator ids are not in upload order");
    if (op.nrows() != nnew || op.ncols() != nnew) die("transform_operators: an operator was not re-allocated");
    ck(b2d_rotated_op_download(g.ctx, id, allowed.data(), 0), "b2d_rotated_op_download(mask)");
    size_t begin = off;
    for (int a = 0; a < nnew; ++a)
      for (int b = 0; b < nnew; ++b) {
        bool al = op.allowed(a, b);
        if (al != (allowed[(size_t)a * nnew + b] != 0)) die("transform_operators: allowed mask differs from SparseMatrix::allocate");
        if (!al) continue;
        Matrix& m = op.operator_element(a, b);
        if (off + m.Storage() > data.size()) die("transform_operators: rotated operator size mismatch");
        if (check) for (int i = 0; i < m.Storage(); ++i) { worst = std::max(worst, fabs(m.Store()[i] - data[off + i])); scale = std::max(scale, fabs(m.Store()[i])); }
        memcpy(m.Store(), data.data() + off, sizeof(double) * m.Storage());
        off += m.Storage();
      }
    if ((int64_t)(off - begin) != b2d_rotated_op_size(g.ctx, id)) die("transform_operators: rotated operator size mismatch");
  }
  g.t_rot += now_s() - t0;
  if (check) fprintf(stderr, "B2D_CHECK call=%d transform_operators ops=%d max_abs_diff=%.3e (max |O'| %.3e)\n", g.call, (int)g.left_ops.size(), worst, scale);
  write_stats();
  g.dirty = false;
  release();
  maybe_stop();
}

// ---- SpinBlock::RenormaliseFrom: only guards the modes the GPU path does not cover ----
void real_RenormaliseFrom(SpinBlock* self, vector<double>& energies, vector<double>& spins, double& error, vector<Matrix>& rotateMatrix,
                          const int keptstates, const int keptqstates, const double tol, SpinBlock& big, const guessWaveTypes& gw,
                          const double noise, const double additional_noise, const bool& onedot, SpinBlock& System, SpinBlock& sysDot,
                          SpinBlock& environment, const bool& dot_with_sys, const bool& warmUp, int sweepiter, int currentRoot,
                          vector<Wavefunction>& lowerStates, DensityMatrix* rdm) asm("__real_" SYM_RenormaliseFrom);
void wrap_RenormaliseFrom(SpinBlock* self, vector<double>& energies, vector<double>& spins, double& error, vector<Matrix>& rotateMatrix,
                          const int keptstates, const int keptqstates, const double tol, SpinBlock& big, const guessWaveTypes& gw,
                          const double noise, const double additional_noise, const bool& onedot, SpinBlock& System, SpinBlock& sysDot,
                          SpinBlock& environment, const bool& dot_with_sys, const bool& warmUp, int sweepiter, int currentRoot,
                          vector<Wavefunction>& lowerStates, DensityMatrix* rdm) asm("__wrap_" SYM_RenormaliseFrom);
void wrap_RenormaliseFrom(SpinBlock* self, vector<double>& energies, vector<double>& spins, double& error, vector<Matrix>& rotateMatrix,
                          const int keptstates, const int keptqstates, const double tol, SpinBlock& big, const guessWaveTypes& gw,
                          const double noise, const double additional_noise, const bool& onedot, SpinBlock& System, SpinBlock& sysDot,
                          SpinBlock& environment, const bool& dot_with_sys, const bool& warmUp, int sweepiter, int currentRoot,
                          vector<Wavefunction>& lowerStates, DensityMatrix* rdm) {
  // one-dot step with the dot on the environment side (renormalise.C:64-79): the solve runs on system x (dot+environment), the
  // reference reshuffles the solutions on the host (onedot_shufflesysdot) and the density matrix / rotation then run on
  // (system+dot) x environment: two device contexts in one call, created by whichever hook first sees each big block
  if (dmrginp.solve_method() != DAVIDSON) die("solver other than Davidson: not covered");
  real_RenormaliseFrom(self, energies, spins, error, rotateMatrix, keptstates, keptqstates, tol, big, gw, noise, additional_noise, onedot, System, sysDot,
                       environment, dot_with_sys, warmUp, sweepiter, currentRoot, lowerStates, rdm);
}

}  // namespace SpinAdapted
