// Fixture generator and CPU-baseline timer for the oracle.  TEST INFRASTRUCTURE ONLY.
//
// This translation unit is OUR code.  It is linked together with the unmodified reference objects and uses GNU ld's
// --wrap to interpose on the reference's hot-path entry points (no reference source is edited or copied):
//
//   SpinBlock::RenormaliseFrom      renormalise.C:39      (caller sweep.C:263)
//   Solver::solve_wavefunction      solver.C:21           (caller renormalise.C:57)
//   Linear::block_davidson          linear.C:179          (caller solver.C:91)
//   DensityMatrix::makedensitymatrix density.C:27         (caller renormalise.C:104)
//   SpinBlock::transform_operators  save_load_block.C:267 (caller sweep.C:279)
//   SpinBlock::multiplyH            spinblock.C:722       (caller davidson.C:21)
//   GuessWave::guess_wavefunctions  guess_wavefunction.C:378 (caller solver.C:77)   [ORACLE_DUMP_GUESS: SURVEY N1 fixtures]
//
// Environment:
//   ORACLE_DUMP_DIR    directory for site<k>.bin records (unset => no dumps)
//   ORACLE_DUMP_CALLS  comma list of RenormaliseFrom call numbers (0-based) to dump, or "all"
//   ORACLE_TIMING      if set, print per-call sigma timing/flop lines ("ORACLE_SIGMA ...") to stderr
//
// Record format (little endian): repeated { u32 name_len, name, u8 dtype (0=i32,1=f64), u32 ndim, u64 dims[ndim], data }.
#include <sys/time.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "wrap_syms.h"

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
#include "guess_wavefunction.h"

using namespace SpinAdapted;
using std::string;
using std::vector;

namespace {

double now_s() { timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }

struct Dumper {
  std::ofstream f;
  bool open(const string& path, bool append) {
    f.open(path.c_str(), std::ios::binary | (append ? std::ios::app : std::ios::trunc));
    return f.good();
  }
  void head(const string& name, uint8_t dtype, const vector<uint64_t>& dims) {
    uint32_t n = name.size(); f.write((char*)&n, 4); f.write(name.data(), n);
    f.write((char*)&dtype, 1);
    uint32_t nd = dims.size(); f.write((char*)&nd, 4);
    for (uint64_t d : dims) f.write((char*)&d, 8);
  }
  void ints(const string& name, const vector<int>& v, vector<uint64_t> dims = vector<uint64_t>()) {
    if (dims.empty()) dims.push_back(v.size());
    head(name, 0, dims);
    vector<int32_t> t(v.begin(), v.end());
    f.write((char*)t.data(), 4 * t.size());
  }
  void dbls(const string& name, const vector<double>& v, vector<uint64_t> dims = vector<uint64_t>()) {
    if (dims.empty()) dims.push_back(v.size());
    head(name, 1, dims);
    f.write((char*)v.data(), 8 * v.size());
  }
};

int g_call = -1;             // RenormaliseFrom call counter
bool g_dump_this = false;    // dump the current call?
string g_path;
double g_sigma_time = 0.0;
long g_sigma_calls = 0;

bool want_dump(int call) {
  const char* dir = getenv("ORACLE_DUMP_DIR");
  if (!dir) return false;
  const char* calls = getenv("ORACLE_DUMP_CALLS");
  if (!calls || string(calls) == "all") return true;
  std::stringstream ss(calls); string tok;
  while (std::getline(ss, tok, ',')) if (atoi(tok.c_str()) == call) return true;
  return false;
}

void flatten(const SparseMatrix& w, vector<double>& out) {
  out.clear();
  for (int l = 0; l < w.nrows(); ++l)
    for (int r = 0; r < w.ncols(); ++r)
      if (w.allowed(l, r)) {
        const Matrix& m = w.operator_element(l, r);
        out.insert(out.end(), m.Store(), m.Store() + m.Storage());
      }
}

void dump_stateinfo(Dumper& d, const string& p, const StateInfo& s) {
  vector<int> q;
  for (size_t i = 0; i < s.quanta.size(); ++i) {
    q.push_back(s.quanta[i].get_n()); q.push_back(s.quanta[i].get_s().getirrep()); q.push_back(s.quanta[i].get_symm().getirrep());
  }
  d.ints(p + "q", q, {s.quanta.size(), 3});
  d.ints(p + "dims", s.quantaStates);
}

void dump_op(Dumper& d, const string& p, int optype, bool core, int comp, SparseMatrix& op) {
  vector<int> meta;
  meta.push_back(optype); meta.push_back(core); meta.push_back((int)op.get_orbs().size());
  meta.push_back(op.get_orbs(0)); meta.push_back(op.get_orbs(1)); meta.push_back(comp);
  SpinQuantum dq = op.get_deltaQuantum(0);
  meta.push_back(dq.get_n()); meta.push_back(dq.get_s().getirrep()); meta.push_back(dq.get_symm().getirrep());
  meta.push_back(op.get_fermion()); meta.push_back(op.get_deltaQuantum_size());
  d.ints(p + "meta", meta);
  vector<int> allowed; vector<double> data;
  for (int i = 0; i < op.nrows(); ++i)
    for (int j = 0; j < op.ncols(); ++j) {
      allowed.push_back(op.allowed(i, j) ? 1 : 0);
      if (op.allowed(i, j)) { const Matrix& m = op.operator_element(i, j); data.insert(data.end(), m.Store(), m.Store() + m.Storage()); }
    }
  d.ints(p + "allowed", allowed, {(uint64_t)op.nrows(), (uint64_t)op.ncols()});
  d.dbls(p + "data", data);
}

// every operator array of a block, each element expanded to its working representation
// (direct-mode virtual operators are BUILT here by the reference's own Op::build)
void dump_block(Dumper& d, const string& p, SpinBlock& b, const std::set<int>* only = 0) {
  d.ints(p + "sites", b.get_sites());
  d.ints(p + "flags", vector<int>{b.is_loopblock(), b.is_direct(), b.get_integralIndex()});
  dump_stateinfo(d, p, b.get_stateInfo());
  int m = 0;
  for (std::map<opTypes, boost::shared_ptr<Op_component_base> >::iterator it = b.ops.begin(); it != b.ops.end(); ++it) {
    if (only && !only->count((int)it->first)) continue;
    Op_component_base& arr = *it->second;
    for (int i = 0; i < arr.get_size(); ++i) {
      vector<boost::shared_ptr<SparseMatrix> > vec = arr.get_local_element(i);
      for (size_t c = 0; c < vec.size(); ++c) {
        boost::shared_ptr<SparseMatrix> rep = vec[c]->getworkingrepresentation(&b);
        std::ostringstream nm; nm << p << "op" << m++ << ".";
        dump_op(d, nm.str(), (int)it->first, arr.is_core(), (int)c, *rep);
      }
    }
  }
  d.ints(p + "nops", vector<int>{m});
}

// a deterministic non-trivial wavefunction ("rpsi")
void fill_lcg(SparseMatrix& w, uint64_t st = 0x9E3779B97F4A7C15ull) {
  for (int l = 0; l < w.nrows(); ++l) for (int r = 0; r < w.ncols(); ++r) if (w.allowed(l, r)) {
    Matrix& m = w.operator_element(l, r);
    for (int k = 0; k < m.Storage(); ++k) { st = st * 6364136223846793005ull + 1442695040888963407ull; m.Store()[k] = ((st >> 11) * (1.0 / 9007199254740992.0)) - 0.5; }
  }
}

const std::set<int>& hot_optypes() {
  static std::set<int> s;
  if (s.empty()) { int t[] = {HAM, CRE, CRE_CRE, DES_DESCOMP, CRE_DES, CRE_DESCOMP, CRE_CRE_DESCOMP, OVERLAP}; s.insert(t, t + 8); }
  return s;
}


}  // namespace

// ---------------- wrapped entry points ----------------
namespace SpinAdapted {

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
  ++g_call;
  g_dump_this = want_dump(g_call) && !(onedot && !dot_with_sys);
  double t0 = now_s(); double s0 = g_sigma_time; long c0 = g_sigma_calls;
  if (g_dump_this) {
    std::ostringstream p; p << getenv("ORACLE_DUMP_DIR") << "/site" << g_call << ".bin";
    g_path = p.str();
    Dumper d; d.open(g_path, false);
    SpinQuantum tq = dmrginp.effective_molecule_quantum();
    d.ints("meta", vector<int>{g_call, big.get_leftBlock()->get_sites()[0] == 0, onedot, dot_with_sys, dmrginp.nroots(sweepiter), keptstates,
                               sweepiter, (int)dmrginp.hamiltonian(), dmrginp.spinAdapted(), big.get_integralIndex(), keptqstates, (int)warmUp,
                               (int)gw, currentRoot});
    d.dbls("meta_f", vector<double>{tol, noise, additional_noise, coreEnergy[big.get_integralIndex()]});
    d.dbls("weights", dmrginp.weights(sweepiter));
    d.ints("psi_dq", vector<int>{tq.get_n(), tq.get_s().getirrep(), tq.get_symm().getirrep()});
    string symname = SpinAdapted::sym; d.ints("sym", vector<int>(symname.begin(), symname.end()));
    d.ints("spin_orbs_symmetry", dmrginp.spin_orbs_symmetry());
    dump_block(d, "L.", *big.get_leftBlock(), &hot_optypes());
    dump_block(d, "R.", *big.get_rightBlock(), &hot_optypes());
    const StateInfo& bs = big.get_stateInfo();
    dump_stateinfo(d, "big.", bs);
    d.ints("big.lmap", bs.leftUnMapQuanta); d.ints("big.rmap", bs.rightUnMapQuanta); d.ints("big.unblocked", bs.unBlockedIndex);
    if (getenv("ORACLE_DUMP_CHILDREN") && big.get_leftBlock()->get_leftBlock() && big.get_leftBlock()->get_rightBlock()) {
      // for the operator-construction oracle: the two children of the enlarged left block with every operator they
      // carry, and the product StateInfo maps of the enlarged block (TensorProduct / TensorTrace, operatorfunctions.C:19-254)
      SpinBlock& nl = *big.get_leftBlock();
      dump_block(d, "LL.", *nl.get_leftBlock());
      dump_block(d, "LR.", *nl.get_rightBlock());
      const StateInfo& si = nl.get_stateInfo();
      d.ints("L.si.lmap", si.leftUnMapQuanta); d.ints("L.si.rmap", si.rightUnMapQuanta);
      d.ints("L.si.uncollected_dims", si.unCollectedStateInfo->quantaStates);
      vector<int> o2n, o2n_begin(1, 0);
      for (size_t q = 0; q < si.oldToNewState.size(); ++q) { o2n.insert(o2n.end(), si.oldToNewState[q].begin(), si.oldToNewState[q].end()); o2n_begin.push_back((int)o2n.size()); }
      d.ints("L.si.old_to_new", o2n); d.ints("L.si.old_to_new_begin", o2n_begin);
      d.ints("L.si.left_is_LL", vector<int>{si.leftStateInfo == &nl.get_leftBlock()->get_stateInfo() ? 1 : 0});
      dump_block(d, "LA.", nl);    // the enlarged block with EVERY operator type (the hot-path subset is "L.")
      // integrals as the reference's own accessors return them (reordered orbitals), spatial indices: v1[i][j] = v_1(2i,2j),
      // v2[i][j][k][l] = v_2(2i,2j,2k,2l)
      const int idx = big.get_integralIndex();
      const int n = (int)dmrginp.spin_orbs_symmetry().size() / 2;
      vector<double> h1((size_t)n * n), h2((size_t)n * n * n * n);
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) h1[(size_t)i * n + j] = v_1[idx](2 * i, 2 * j);
      for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) for (int k = 0; k < n; ++k) for (int l = 0; l < n; ++l)
        h2[(((size_t)i * n + j) * n + k) * n + l] = v_2[idx](2 * i, 2 * j, 2 * k, 2 * l);
      d.dbls("v1", h1, {(uint64_t)n, (uint64_t)n});
      d.dbls("v2", h2, {(uint64_t)n, (uint64_t)n, (uint64_t)n, (uint64_t)n});
      d.dbls("screen_tol", vector<double>{dmrginp.oneindex_screen_tol(), dmrginp.twoindex_screen_tol()});
    }
    Wavefunction w; w.initialise(dmrginp.effective_molecule_quantum_vec(), &big, onedot);
    vector<int> allowed;
    for (int l = 0; l < w.nrows(); ++l) for (int r = 0; r < w.ncols(); ++r) allowed.push_back(w.allowed(l, r) ? 1 : 0);
    d.ints("psi_allowed", allowed, {(uint64_t)w.nrows(), (uint64_t)w.ncols()});
    // a deterministic non-trivial vector and the reference's sigma for it
    fill_lcg(w);
    Wavefunction v = w; v.Clear();
    big.multiplyH(w, &v, 1);
    vector<double> flat; flatten(w, flat); d.dbls("rpsi", flat); flatten(v, flat); d.dbls("rsigma", flat);
  }
  real_RenormaliseFrom(self, energies, spins, error, rotateMatrix, keptstates, keptqstates, tol, big, gw, noise, additional_noise, onedot,
                       System, sysDot, environment, dot_with_sys, warmUp, sweepiter, currentRoot, lowerStates, rdm);
  if (g_dump_this) {
    Dumper d; d.open(g_path, true);
    d.dbls("energies", energies); d.dbls("error", vector<double>{error});
    vector<int> shape; vector<double> data;
    for (size_t q = 0; q < rotateMatrix.size(); ++q) {
      shape.push_back(rotateMatrix[q].Nrows()); shape.push_back(rotateMatrix[q].Ncols());
      if (rotateMatrix[q].Ncols()) data.insert(data.end(), rotateMatrix[q].Store(), rotateMatrix[q].Store() + rotateMatrix[q].Storage());
    }
    d.ints("rot.shape", shape, {rotateMatrix.size(), 2}); d.dbls("rot.data", data);
  }
  if (getenv("ORACLE_TIMING"))
    fprintf(stderr, "ORACLE_SITE call=%d renorm_s=%.6f sigma_s=%.6f sigma_calls=%ld\n", g_call, now_s() - t0, g_sigma_time - s0, g_sigma_calls - c0);
}

void real_solve(vector<Wavefunction>& solution, vector<double>& energies, SpinBlock& big, const double tol, const guessWaveTypes& gw, const bool& onedot,
                const bool& dot_with_sys, const bool& warmUp, double additional_noise, int currentRoot, vector<Wavefunction>& lowerStates) asm("__real_" SYM_solve_wavefunction);
void wrap_solve(vector<Wavefunction>& solution, vector<double>& energies, SpinBlock& big, const double tol, const guessWaveTypes& gw, const bool& onedot,
                const bool& dot_with_sys, const bool& warmUp, double additional_noise, int currentRoot, vector<Wavefunction>& lowerStates) asm("__wrap_" SYM_solve_wavefunction);
void wrap_solve(vector<Wavefunction>& solution, vector<double>& energies, SpinBlock& big, const double tol, const guessWaveTypes& gw, const bool& onedot,
                const bool& dot_with_sys, const bool& warmUp, double additional_noise, int currentRoot, vector<Wavefunction>& lowerStates) {
  real_solve(solution, energies, big, tol, gw, onedot, dot_with_sys, warmUp, additional_noise, currentRoot, lowerStates);
  if (!g_dump_this) return;
  Dumper d; d.open(g_path, true);
  DiagonalMatrix e; e.ReSize(big.get_stateInfo().totalStates); e = 0;
  big.diagonalH(e);
  d.dbls("diag", vector<double>(e.Store(), e.Store() + e.Storage()));
  vector<double> flat;
  for (size_t i = 0; i < solution.size(); ++i) {
    std::ostringstream a, b; a << "psi" << i; b << "sigma" << i;
    flatten(solution[i], flat); d.dbls(a.str(), flat);
    Wavefunction v = solution[i]; v.Clear();
    big.multiplyH(solution[i], &v, 1);
    flatten(v, flat); d.dbls(b.str(), flat);
  }
  d.dbls("solve_energies", energies);
}

void real_davidson(vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp, Davidson_functor& h_multiply, bool& useprecond,
                   int currentRoot, vector<Wavefunction>& lowerStates) asm("__real_" SYM_block_davidson);
void wrap_davidson(vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp, Davidson_functor& h_multiply, bool& useprecond,
                   int currentRoot, vector<Wavefunction>& lowerStates) asm("__wrap_" SYM_block_davidson);
void wrap_davidson(vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp, Davidson_functor& h_multiply, bool& useprecond,
                   int currentRoot, vector<Wavefunction>& lowerStates) {
  long c0 = g_sigma_calls;
  DiagonalMatrix diag_in;
  if (g_dump_this) {
    diag_in = h_diag;
    Dumper d; d.open(g_path, true);
    vector<double> flat;
    for (size_t i = 0; i < b.size(); ++i) { std::ostringstream a; a << "guess" << i; flatten(b[i], flat); d.dbls(a.str(), flat); }
    d.dbls("dav_tol", vector<double>{normtol});
    d.ints("dav_in", vector<int>{(int)b.size(), (int)useprecond, currentRoot, (int)lowerStates.size(), dmrginp.deflation_min_size(), dmrginp.deflation_max_size()});
  }
  real_davidson(b, h_diag, normtol, warmUp, h_multiply, useprecond, currentRoot, lowerStates);
  if (g_dump_this) {
    Dumper d; d.open(g_path, true);
    vector<double> ev; for (size_t i = 0; i < b.size() && (int)i < h_diag.Ncols(); ++i) ev.push_back(h_diag.element(i));
    d.dbls("dav_evals", ev);
    d.ints("dav_out", vector<int>{(int)(g_sigma_calls - c0), (int)b.size()});
    // state-specific form (lowerStates, linear.C:201-208,311-317,369-375): one more solve of the reference's own block_davidson
    // for ONE root with a lower state.  Lower state = a deterministic pseudo-random, unnormalised vector; guess = converged root 0
    // + 1% of another such vector, so that the solve takes a handful of iterations (a long Krylov run from a random guess amplifies
    // rounding differences and cannot be compared bit for bit).  Pins the three lower-state projections.
    if (lowerStates.empty()) {
      long c1 = g_sigma_calls;
      vector<Wavefunction> lo(1, b[0]);
      fill_lcg(lo[0], 0xD1B54A32D192ED03ull);
      vector<Wavefunction> g1(1, b[0]);
      Wavefunction pert = b[0];
      fill_lcg(pert);
      ScaleAdd(0.01, pert, g1[0]);
      vector<double> flat;
      flatten(lo[0], flat); d.dbls("ss_lower", flat);
      flatten(g1[0], flat); d.dbls("ss_guess", flat);
      DiagonalMatrix hd = diag_in;
      bool up = useprecond;
      real_davidson(g1, hd, normtol, warmUp, h_multiply, up, 1, lo);
      d.dbls("ss_eval", vector<double>{hd.element(0)});
      flatten(g1[0], flat); d.dbls("ss_psi", flat);
      d.ints("ss_nmult", vector<int>{(int)(g_sigma_calls - c1)});
    }
  }
}

void real_makedm(DensityMatrix* self, const vector<Wavefunction>& ws, SpinBlock& big, const vector<double>& wts, const double noise, const double add_noise, bool warmup) asm("__real_" SYM_makedensitymatrix);
void wrap_makedm(DensityMatrix* self, const vector<Wavefunction>& ws, SpinBlock& big, const vector<double>& wts, const double noise, const double add_noise, bool warmup) asm("__wrap_" SYM_makedensitymatrix);
void wrap_makedm(DensityMatrix* self, const vector<Wavefunction>& ws, SpinBlock& big, const vector<double>& wts, const double noise, const double add_noise, bool warmup) {
  real_makedm(self, ws, big, wts, noise, add_noise, warmup);
  if (!g_dump_this) return;
  Dumper d; d.open(g_path, true);
  vector<double> data;
  for (int q = 0; q < self->nrows(); ++q) if (self->allowed(q, q)) { const Matrix& m = self->operator_element(q, q); data.insert(data.end(), m.Store(), m.Store() + m.Storage()); }
  d.dbls("rdm.data", data);
  d.dbls("rdm.args", vector<double>{noise, add_noise, (double)warmup});
}

void real_transform(SpinBlock* self, vector<Matrix>& rot) asm("__real_" SYM_transform_operators);
void wrap_transform(SpinBlock* self, vector<Matrix>& rot) asm("__wrap_" SYM_transform_operators);
void wrap_transform(SpinBlock* self, vector<Matrix>& rot) {
  double t0 = now_s();
  if (g_dump_this && getenv("ORACLE_DUMP_FULLOPS")) {
    // the un-rotated operators this block will carry forward (everything, incl. the non-hot-path types)
    Dumper d; d.open(g_path, true);
    dump_block(d, "U.", *self);
  }
  real_transform(self, rot);
  if (getenv("ORACLE_TIMING")) fprintf(stderr, "ORACLE_ROTATE call=%d rotate_s=%.6f\n", g_call, now_s() - t0);
  if (!g_dump_this) return;
  Dumper d; d.open(g_path, true);
  dump_block(d, "N.", *self);
}

// the guess-wavefunction transform of a two-dot step (GuessWave::transform_previous_wavefunction, guess_wavefunction.C:524-636).
// The wrapped caller-facing entry point is guess_wavefunctions (called across translation units from solver.C:77); the inputs the
// reference loads from its scratch files inside (previous wavefunction + its StateInfo tree, the two rotation matrices) are loaded
// here the same way and dumped together with every StateInfo table the transform reads; the output is the reference's own trial vector.
void dump_si_tables(Dumper& d, const string& p, const StateInfo& s) {
  dump_stateinfo(d, p, s);
  d.ints(p + "new_quanta_map", s.newQuantaMap);
  if (s.hasCollectedQuanta && s.unCollectedStateInfo) {
    const StateInfo& u = *s.unCollectedStateInfo;
    dump_stateinfo(d, p + "unc.", u);
    d.ints(p + "unc.lmap", u.leftUnMapQuanta); d.ints(p + "unc.rmap", u.rightUnMapQuanta);
    vector<int> o2n, begin(1, 0);
    for (size_t q = 0; q < s.oldToNewState.size(); ++q) { o2n.insert(o2n.end(), s.oldToNewState[q].begin(), s.oldToNewState[q].end()); begin.push_back((int)o2n.size()); }
    d.ints(p + "old_to_new", o2n); d.ints(p + "old_to_new_begin", begin);
  }
}
void dump_rotation(Dumper& d, const string& p, const vector<Matrix>& rot) {
  vector<int> shape; vector<double> data;
  for (size_t q = 0; q < rot.size(); ++q) {
    shape.push_back(rot[q].Nrows()); shape.push_back(rot[q].Ncols());
    if (rot[q].Ncols()) data.insert(data.end(), rot[q].Store(), rot[q].Store() + rot[q].Storage());
  }
  d.ints(p + "shape", shape, {rot.size(), 2}); d.dbls(p + "data", data);
}

void real_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) asm("__real_" SYM_guess_wavefunctions);
void wrap_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) asm("__wrap_" SYM_guess_wavefunctions);
void wrap_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) {
  real_guess(solution, e, big, gw, onedot, transpose_guess_wave, additional_noise, currentState);
  if (getenv("ORACLE_DUMP_GUESS") && gw == TRANSFORM && onedot && want_dump(g_call) && big.get_leftBlock()->get_leftBlock()) {
    // one-dot branch (guess_wavefunction.C:600-628 -> onedot_transform_wavefunction :832-936): a file of its own, because the
    // RenormaliseFrom record is not written for one-dot steps with the dot on the environment side
    std::ostringstream fp; fp << getenv("ORACLE_DUMP_DIR") << "/guess1dot_" << g_call << ".bin";
    Dumper d; d.open(fp.str(), false);
    const int nroots = (int)solution.size();
    d.ints("meta", vector<int>{g_call, big.get_leftBlock()->get_sites()[0] == 0, 1, (int)transpose_guess_wave});
    d.ints("gw.nroots", vector<int>{nroots, (int)transpose_guess_wave});
    const StateInfo& bs = big.get_stateInfo();
    for (int i = 0; i < nroots; ++i) {
      const int state = (dmrginp.setStateSpecific() || dmrginp.calc_type() == COMPRESS || dmrginp.calc_type() == MPS_NEVPT) ? currentState : i;
      std::ostringstream pp; pp << "gw" << i << ".";
      const string p = pp.str();
      StateInfo oldSI; Wavefunction oldWave; vector<Matrix> lrot, rrot;
      if (transpose_guess_wave) {
        oldWave.LoadWavefunctionInfo(oldSI, big.get_leftBlock()->get_leftBlock()->get_sites(), state);
        LoadRotationMatrix(big.get_leftBlock()->get_leftBlock()->get_sites(), lrot, state);
        vector<int> rotsites = big.get_rightBlock()->get_sites();
        rotsites.insert(rotsites.end(), big.get_leftBlock()->get_rightBlock()->get_sites().begin(), big.get_leftBlock()->get_rightBlock()->get_sites().end());
        std::sort(rotsites.begin(), rotsites.end());
        LoadRotationMatrix(rotsites, rrot, state);
      } else {
        oldWave.LoadWavefunctionInfo(oldSI, big.get_leftBlock()->get_sites(), state);
        LoadRotationMatrix(big.get_leftBlock()->get_sites(), lrot, state);
        LoadRotationMatrix(big.get_rightBlock()->get_sites(), rrot, state);
      }
      SpinQuantum dq = oldWave.get_deltaQuantum(0);
      d.ints(p + "dq", vector<int>{dq.get_n(), dq.get_s().getirrep(), dq.get_symm().getirrep(), (int)oldWave.get_deltaQuantum_size()});
      dump_si_tables(d, p + "left.", *bs.leftStateInfo);
      dump_si_tables(d, p + "right.", *bs.rightStateInfo);
      dump_si_tables(d, p + "oldleft.", *oldSI.leftStateInfo);
      dump_si_tables(d, p + "oldcol.", *oldSI.rightStateInfo);
      if (transpose_guess_wave) {
        dump_si_tables(d, p + "sys.", *bs.leftStateInfo->leftStateInfo);
        dump_si_tables(d, p + "dot.", *bs.leftStateInfo->rightStateInfo);
        StateInfo newenv;     // guess_wavefunction.C:853-856
        TensorProduct(*(bs.rightStateInfo), *(bs.leftStateInfo->rightStateInfo), newenv, NO_PARTICLE_SPIN_NUMBER_CONSTRAINT);
        newenv.CollectQuanta();
        dump_si_tables(d, p + "newenv.", newenv);
      }
      vector<int> allowed; vector<double> data;
      for (int a = 0; a < oldWave.nrows(); ++a) for (int b = 0; b < oldWave.ncols(); ++b) {
        allowed.push_back(oldWave.allowed(a, b) ? 1 : 0);
        if (oldWave.allowed(a, b)) { const Matrix& m = oldWave.operator_element(a, b); data.insert(data.end(), m.Store(), m.Store() + m.Storage()); }
      }
      d.ints(p + "old.allowed", allowed, {(uint64_t)oldWave.nrows(), (uint64_t)oldWave.ncols()});
      d.dbls(p + "old.data", data);
      dump_rotation(d, p + "lrot.", lrot);
      dump_rotation(d, p + "rrot.", rrot);
      vector<int> tallowed;
      for (int l = 0; l < solution[i].nrows(); ++l) for (int r = 0; r < solution[i].ncols(); ++r) tallowed.push_back(solution[i].allowed(l, r) ? 1 : 0);
      d.ints(p + "trial.allowed", tallowed, {(uint64_t)solution[i].nrows(), (uint64_t)solution[i].ncols()});
      vector<double> flat; flatten(solution[i], flat); d.dbls(p + "trial", flat);
      oldSI.Free();
    }
    return;
  }
  if (getenv("ORACLE_DUMP_GUESS") && gw == TRANSPOSE && onedot && want_dump(g_call) && big.get_leftBlock()->get_rightBlock()) {
    // first block iteration of a one-dot sweep: transpose_previous_wavefunction (guess_wavefunction.C:100-112) -> onedot_transpose_wavefunction (:140-198)
    const int nroots = (int)solution.size();
    const StateInfo& bs = big.get_stateInfo();
    std::ostringstream fp; fp << getenv("ORACLE_DUMP_DIR") << "/guessT1_" << g_call << ".bin";
    Dumper d; d.open(fp.str(), false);
    d.ints("meta", vector<int>{g_call, big.get_leftBlock()->get_sites()[0] == 0, 1, 4});
    d.ints("gw.nroots", vector<int>{nroots, 4});
    vector<int> wfsites = big.get_rightBlock()->get_sites();
    wfsites.insert(wfsites.end(), big.get_leftBlock()->get_rightBlock()->get_sites().begin(), big.get_leftBlock()->get_rightBlock()->get_sites().end());
    std::sort(wfsites.begin(), wfsites.end());
    for (int i = 0; i < nroots; ++i) {
      const int state = (dmrginp.setStateSpecific() || dmrginp.calc_type() == COMPRESS || dmrginp.calc_type() == MPS_NEVPT) ? currentState : i;
      StateInfo oldSI; Wavefunction oldWave;
      oldWave.LoadWavefunctionInfo(oldSI, wfsites, state);
      std::ostringstream pp; pp << "gw" << i << ".";
      const string p = pp.str();
      SpinQuantum dq = oldWave.get_deltaQuantum(0);
      d.ints(p + "dq", vector<int>{dq.get_n(), dq.get_s().getirrep(), dq.get_symm().getirrep(), (int)oldWave.get_deltaQuantum_size()});
      dump_si_tables(d, p + "left.", *bs.leftStateInfo);
      dump_si_tables(d, p + "sys.", *bs.leftStateInfo->leftStateInfo);
      dump_si_tables(d, p + "dot.", *bs.leftStateInfo->rightStateInfo);
      dump_si_tables(d, p + "right.", *bs.rightStateInfo);
      dump_si_tables(d, p + "oldleft.", *oldSI.leftStateInfo);
      dump_si_tables(d, p + "oldsys.", *oldSI.leftStateInfo->leftStateInfo);
      dump_si_tables(d, p + "olddot.", *oldSI.leftStateInfo->rightStateInfo);
      dump_si_tables(d, p + "oldcol.", *oldSI.rightStateInfo);
      vector<int> allowed; vector<double> data;
      for (int a = 0; a < oldWave.nrows(); ++a) for (int b = 0; b < oldWave.ncols(); ++b) {
        allowed.push_back(oldWave.allowed(a, b) ? 1 : 0);
        if (oldWave.allowed(a, b)) { const Matrix& m = oldWave.operator_element(a, b); data.insert(data.end(), m.Store(), m.Store() + m.Storage()); }
      }
      d.ints(p + "old.allowed", allowed, {(uint64_t)oldWave.nrows(), (uint64_t)oldWave.ncols()});
      d.dbls(p + "old.data", data);
      vector<double> flat; flatten(solution[i], flat); d.dbls(p + "trial", flat);
      oldSI.Free();
    }
    return;
  }
  if (getenv("ORACLE_DUMP_GUESS") && gw == TRANSPOSE && !onedot && want_dump(g_call)) {
    // first block iteration of a sweep: GuessWave::transpose_previous_wavefunction (guess_wavefunction.C:55-84), two-dot to two-dot
    const int nroots = (int)solution.size();
    const StateInfo& bs = big.get_stateInfo();
    std::ostringstream fp; fp << getenv("ORACLE_DUMP_DIR") << "/guessT_" << g_call << ".bin";
    Dumper d; bool opened = false;
    for (int i = 0; i < nroots; ++i) {
      const int state = (dmrginp.setStateSpecific() || dmrginp.calc_type() == COMPRESS || dmrginp.calc_type() == MPS_NEVPT) ? currentState : i;
      StateInfo oldSI; Wavefunction oldWave;
      oldWave.LoadWavefunctionInfo(oldSI, big.get_rightBlock()->get_sites(), state);
      if (oldWave.get_onedot()) { oldSI.Free(); return; }     // one-dot -> two-dot switch: not dumped
      if (!opened) { d.open(fp.str(), false); opened = true; d.ints("meta", vector<int>{g_call, big.get_leftBlock()->get_sites()[0] == 0, 0, 3}); d.ints("gw.nroots", vector<int>{nroots, 3}); }
      std::ostringstream pp; pp << "gw" << i << ".";
      const string p = pp.str();
      SpinQuantum dq = oldWave.get_deltaQuantum(0);
      d.ints(p + "dq", vector<int>{dq.get_n(), dq.get_s().getirrep(), dq.get_symm().getirrep(), (int)oldWave.get_deltaQuantum_size()});
      dump_si_tables(d, p + "left.", *bs.leftStateInfo);
      dump_si_tables(d, p + "right.", *bs.rightStateInfo);
      dump_si_tables(d, p + "oldleft.", *oldSI.leftStateInfo);
      dump_si_tables(d, p + "oldcol.", *oldSI.rightStateInfo);
      vector<int> allowed; vector<double> data;
      for (int a = 0; a < oldWave.nrows(); ++a) for (int b = 0; b < oldWave.ncols(); ++b) {
        allowed.push_back(oldWave.allowed(a, b) ? 1 : 0);
        if (oldWave.allowed(a, b)) { const Matrix& m = oldWave.operator_element(a, b); data.insert(data.end(), m.Store(), m.Store() + m.Storage()); }
      }
      d.ints(p + "old.allowed", allowed, {(uint64_t)oldWave.nrows(), (uint64_t)oldWave.ncols()});
      d.dbls(p + "old.data", data);
      vector<double> flat; flatten(solution[i], flat); d.dbls(p + "trial", flat);
      oldSI.Free();
    }
    return;
  }
  if (!g_dump_this || !getenv("ORACLE_DUMP_GUESS") || gw != TRANSFORM || onedot) return;
  Dumper d; d.open(g_path, true);
  const int nroots = (int)solution.size();
  d.ints("gw.nroots", vector<int>{nroots, (int)transpose_guess_wave});
  for (int i = 0; i < nroots; ++i) {
    const int state = (dmrginp.setStateSpecific() || dmrginp.calc_type() == COMPRESS || dmrginp.calc_type() == MPS_NEVPT) ? currentState : i;
    std::ostringstream pp; pp << "gw" << i << ".";
    const string p = pp.str();
    StateInfo oldSI;
    Wavefunction oldWave;
    vector<Matrix> lrot, rrot;
    oldWave.LoadWavefunctionInfo(oldSI, big.get_leftBlock()->get_leftBlock()->get_sites(), state);
    LoadRotationMatrix(big.get_leftBlock()->get_leftBlock()->get_sites(), lrot, state);
    LoadRotationMatrix(big.get_rightBlock()->get_sites(), rrot, state);
    const StateInfo& bs = big.get_stateInfo();
    SpinQuantum dq = oldWave.get_deltaQuantum(0);
    d.ints(p + "dq", vector<int>{dq.get_n(), dq.get_s().getirrep(), dq.get_symm().getirrep(), (int)oldWave.get_deltaQuantum_size()});
    dump_si_tables(d, p + "sys.", *bs.leftStateInfo->leftStateInfo);      // S': the renormalised system block
    dump_si_tables(d, p + "dot.", *bs.leftStateInfo->rightStateInfo);     // the new system dot
    dump_si_tables(d, p + "left.", *bs.leftStateInfo);                    // S' (x) dot, collected (+ its uncollected tables)
    dump_si_tables(d, p + "right.", *bs.rightStateInfo);                  // the new environment side
    dump_si_tables(d, p + "oldleft.", *oldSI.leftStateInfo);              // row space of the previous wavefunction
    dump_si_tables(d, p + "oldright.", *oldSI.rightStateInfo);            // its column space: E_old (x) dot, collected
    dump_si_tables(d, p + "env.", *oldSI.rightStateInfo->leftStateInfo);  // E_old: the renormalised environment the rotation R produced
    dump_si_tables(d, p + "olddot.", *oldSI.rightStateInfo->rightStateInfo);
    vector<int> allowed; vector<double> data;
    for (int a = 0; a < oldWave.nrows(); ++a) for (int b = 0; b < oldWave.ncols(); ++b) {
      allowed.push_back(oldWave.allowed(a, b) ? 1 : 0);
      if (oldWave.allowed(a, b)) { const Matrix& m = oldWave.operator_element(a, b); data.insert(data.end(), m.Store(), m.Store() + m.Storage()); }
    }
    d.ints(p + "old.allowed", allowed, {(uint64_t)oldWave.nrows(), (uint64_t)oldWave.ncols()});
    d.dbls(p + "old.data", data);
    dump_rotation(d, p + "lrot.", lrot);
    dump_rotation(d, p + "rrot.", rrot);
    vector<int> tallowed;
    for (int l = 0; l < solution[i].nrows(); ++l) for (int r = 0; r < solution[i].ncols(); ++r) tallowed.push_back(solution[i].allowed(l, r) ? 1 : 0);
    d.ints(p + "trial.allowed", tallowed, {(uint64_t)solution[i].nrows(), (uint64_t)solution[i].ncols()});
    vector<double> flat; flatten(solution[i], flat); d.dbls(p + "trial", flat);
    oldSI.Free();
  }
}

// ORACLE_TIMING: where the reference spends its time outside the hot path (scratch files)
double g_store_s = 0, g_restore_s = 0; long g_store_n = 0, g_restore_n = 0;
struct IoReport { ~IoReport() { if (getenv("ORACLE_TIMING")) fprintf(stderr, "ORACLE_IO store_s=%.3f (%ld calls) restore_s=%.3f (%ld calls)\n", g_store_s, g_store_n, g_restore_s, g_restore_n); } } g_io_report;
void real_store(bool forward, const vector<int>& sites, SpinBlock& b, int left, int right, char* name) asm("__real_" SYM_store);
void wrap_store(bool forward, const vector<int>& sites, SpinBlock& b, int left, int right, char* name) asm("__wrap_" SYM_store);
void wrap_store(bool forward, const vector<int>& sites, SpinBlock& b, int left, int right, char* name) {
  double t0 = now_s(); real_store(forward, sites, b, left, right, name); g_store_s += now_s() - t0; ++g_store_n;
}
string real_restore(bool forward, const vector<int>& sites, SpinBlock& b, int left, int right, char* name) asm("__real_" SYM_restore);
string wrap_restore(bool forward, const vector<int>& sites, SpinBlock& b, int left, int right, char* name) asm("__wrap_" SYM_restore);
string wrap_restore(bool forward, const vector<int>& sites, SpinBlock& b, int left, int right, char* name) {
  double t0 = now_s(); string r = real_restore(forward, sites, b, left, right, name); g_restore_s += now_s() - t0; ++g_restore_n; return r;
}

void real_multiplyH(const SpinBlock* self, Wavefunction& c, Wavefunction* v, int num_threads) asm("__real_" SYM_multiplyH);
void wrap_multiplyH(const SpinBlock* self, Wavefunction& c, Wavefunction* v, int num_threads) asm("__wrap_" SYM_multiplyH);
void wrap_multiplyH(const SpinBlock* self, Wavefunction& c, Wavefunction* v, int num_threads) {
  double t0 = now_s();
  real_multiplyH(self, c, v, num_threads);
  g_sigma_time += now_s() - t0;
  ++g_sigma_calls;
}

}  // namespace SpinAdapted
