// C++ host mirror of the reference's entry points for the sweep hot path, on top of the C ABI (include/block_b200.h).
//
// A Block maintainer who swaps the CPU path for the GPU one keeps calling the same names with the same argument meaning
// (reference file:line under the reference root):
//     SpinBlock::multiplyH(Wavefunction& c, Wavefunction* v, int num_threads) const          spinblock.h:235
//     SpinBlock::diagonalH(DiagonalMatrix& e) const                                           spinblock.h:240
//     SpinBlock::RenormaliseFrom(energies, spins, error, rotateMatrix, keptstates, ...)       spinblock.h:247-251
//     SpinBlock::transform_operators(std::vector<Matrix>& rotateMatrix)                       spinblock.h:253
//     operatorfunctions::TensorMultiply(ablock, a, b, cblock, c, v, opQ, scale)               operatorfunctions.h:46
//     Linear::block_davidson(b, h_diag, normtol, warmUp, h_multiply, useprecond, currentRoot, lowerStates)   linear.h:28
// Storage types are deliberately the plain shapes of the reference's (newmat Matrix = row-major doubles,
// SparseMatrix = allowed mask + one Matrix per allowed sector pair, Wavefunction : SparseMatrix, StateInfo = quanta +
// quantaStates) so that the adapter in INTEGRATION.md is a field-by-field copy.  All arithmetic happens in
// libblockb200.so on the GPU; errors print the library's message and abort(), like the reference.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/block_b200.h"

namespace b2d_host {

struct SpinQuantum {   // SpinQuantum.h: particle number, 2S, irrep (abelian)
  int n = 0, s = 0, irrep = 0;
  SpinQuantum() {}
  SpinQuantum(int n_, int s_, int irrep_) : n(n_), s(s_), irrep(irrep_) {}
  bool operator==(const SpinQuantum& o) const { return n == o.n && s == o.s && irrep == o.irrep; }
  // SpinQuantum::allow(s1, s2) SpinQuantum.h:77: *this is in s1 + s2
  bool allow(const SpinQuantum& s1, const SpinQuantum& s2) const;
};

struct StateInfo {   // StateInfo.h:113-147, the part the hot path reads
  std::vector<SpinQuantum> quanta;
  std::vector<int> quantaStates;
  int totalStates() const;
};

struct Matrix {   // newmat Matrix: row-major, Store()/Storage() (newmat.h:455-456)
  int nrows = 0, ncols = 0;
  std::vector<double> store;
  void ReSize(int r, int c) { nrows = r; ncols = c; store.assign((size_t)r * c, 0.0); }
  int Nrows() const { return nrows; }
  int Ncols() const { return ncols; }
  double* Store() { return store.data(); }
  const double* Store() const { return store.data(); }
  int Storage() const { return (int)store.size(); }
  double& element(int i, int j) { return store[(size_t)i * ncols + j]; }
  double element(int i, int j) const { return store[(size_t)i * ncols + j]; }
};
typedef std::vector<double> DiagonalMatrix;

enum opTypes { HAM = 0, CRE = 1, CRE_CRE = 2, DES_DESCOMP = 3, CRE_DES = 4, CRE_DESCOMP = 5, CRE_CRE_DESCOMP = 6, OVERLAP = 13 };   // BaseOperator.h:35-44

class SparseMatrix {   // BaseOperator.h:75-243
 public:
  SparseMatrix() {}
  virtual ~SparseMatrix() {}
  opTypes optype = HAM;
  std::vector<int> orbs;
  int comp = 0;   // index inside the vector over spin components (op_components.C:172-184)
  std::vector<SpinQuantum> deltaQuantum = std::vector<SpinQuantum>(1);
  bool fermion = false;
  int nrows() const { return nr; }
  int ncols() const { return nc; }
  void resize(int r, int c);
  char& allowed(int i, int j) { return allowedQuantaMatrix[(size_t)i * nc + j]; }
  const char& allowed(int i, int j) const { return allowedQuantaMatrix[(size_t)i * nc + j]; }
  Matrix& operator_element(int i, int j) { return operatorMatrix[(size_t)i * nc + j]; }
  const Matrix& operator_element(int i, int j) const { return operatorMatrix[(size_t)i * nc + j]; }
  Matrix& operator()(int i, int j) { return operator_element(i, j); }
  SpinQuantum get_deltaQuantum(int i = 0) const { return deltaQuantum[i]; }
  bool get_fermion() const { return fermion; }
  virtual char conjugacy() const { return 'n'; }
  virtual const SparseMatrix& stored() const { return *this; }
  // SparseMatrix::allocate(sr, sc) BaseOperator.C:123-145
  void allocate(const StateInfo& sr, const StateInfo& sc);
  void allocate(const StateInfo& s) { allocate(s, s); }
  // allowed blocks concatenated (i outer, j inner): the host layout of b2d_add_op / b2d_vec_upload
  void FlattenInto(std::vector<double>& flat) const;
  void CollectFrom(const std::vector<double>& flat);
  int64_t packed_size() const;

 protected:
  int nr = 0, nc = 0;
  std::vector<char> allowedQuantaMatrix;
  std::vector<Matrix> operatorMatrix;
};

class Transposeview : public SparseMatrix {   // BaseOperator.h:208-243: a view, never materialised
 public:
  explicit Transposeview(const SparseMatrix& op) : op_(op) {}
  char conjugacy() const override { return op_.conjugacy() == 'n' ? 't' : 'n'; }
  const SparseMatrix& stored() const override { return op_.stored(); }

 private:
  const SparseMatrix& op_;
};

class SpinBlock;

class Wavefunction : public SparseMatrix {   // wavefunction.h:16-56
 public:
  Wavefunction() {}
  Wavefunction(const SpinQuantum& dQ, const SpinBlock* big, bool onedot) { initialise(dQ, big, onedot); }
  // Wavefunction::initialise wavefunction.C:18-56
  void initialise(const SpinQuantum& dQ, const SpinBlock* big, bool onedot);
  void Clear();
  bool onedot = false;
};

struct Davidson_functor {   // davidson.h:15-19
  virtual ~Davidson_functor() {}
  virtual void operator()(Wavefunction& c, Wavefunction& v) = 0;
  virtual const SpinBlock& get_block() = 0;
};

enum guessWaveTypes { BASIC, TRANSFORM, TRANSPOSE };   // guess_wavefunction.h
enum hamTypes { QUANTUM_CHEMISTRY, HUBBARD };            // input.h

struct DeviceOptions {
  int device = 0;            // CUDA device of this process (one process per GPU)
  int rank = 0, nranks = 1;  // share of the operator terms (distribute.C / para_array.h ownership)
  double core_energy = 0.0;  // coreEnergy[integralIndex]
  hamTypes ham = QUANTUM_CHEMISTRY;
  int norbs = 0;             // number of spatial orbitals (trimap_2d length)
  int deflation_min = 2, deflation_max = 20;   // input.C:116
};

class SpinBlock {   // spinblock.h:24
 public:
  SpinBlock() {}
  ~SpinBlock();
  SpinBlock(const SpinBlock&) = delete;
  SpinBlock& operator=(const SpinBlock&) = delete;

  // ---- a child block: StateInfo, sites and operator arrays (Op_component<Op>, flattened) ----
  StateInfo stateInfo;
  std::vector<int> sites;
  bool loopblock = false;
  std::vector<std::shared_ptr<SparseMatrix>> ops;
  const StateInfo& get_stateInfo() const { return stateInfo; }
  const std::vector<int>& get_sites() const { return sites; }
  bool is_loopblock() const { return loopblock; }

  // ---- the big block: InitBigBlock (initblocks.C) hands over both children; operators are uploaded once ----
  void set_big_block(SpinBlock* left, SpinBlock* right, const SpinQuantum& target, const DeviceOptions& opt);
  SpinBlock* get_leftBlock() const { return leftBlock; }
  SpinBlock* get_rightBlock() const { return rightBlock; }
  const SpinQuantum& get_target() const { return target; }

  // ---- the reference's entry points ----
  void multiplyH(Wavefunction& c, Wavefunction* v, int num_threads) const;                       // spinblock.C:722
  void diagonalH(DiagonalMatrix& e) const;                                                        // spinblock.C:855
  void RenormaliseFrom(std::vector<double>& energies, std::vector<double>& spins, double& error, std::vector<Matrix>& rotateMatrix,
                       const int keptstates, const int keptqstates, const double tol, SpinBlock& big, const guessWaveTypes& guesswavetype,
                       const double noise, const double additional_noise, const bool& onedot, SpinBlock& system, SpinBlock& sysDot,
                       SpinBlock& environment, const bool& dot_with_sys, const bool& warmUp, int sweepiter, int currentRoot,
                       std::vector<Wavefunction>& lowerStates, std::vector<Wavefunction>* solution = nullptr,
                       const std::vector<double>* weights = nullptr);                                 // renormalise.C:39
  void transform_operators(std::vector<Matrix>& rotateMatrix);                                    // save_load_block.C:267

  // plumbing for operatorfunctions / Linear
  b2d_ctx* context() const { return ctx; }
  int op_id(int side, const SparseMatrix* op) const;
  int64_t psi_size() const;
  const DeviceOptions& options() const { return opt; }

 private:
  SpinBlock* leftBlock = nullptr;
  SpinBlock* rightBlock = nullptr;
  SpinBlock* parent = nullptr;   // the big block this block is the left child of
  SpinQuantum target;
  DeviceOptions opt;
  b2d_ctx* ctx = nullptr;
  std::vector<const SparseMatrix*> registered[2];
};

namespace operatorfunctions {
// operatorfunctions.C:485-537: v += scale (a x b) c; `a` belongs to ablock, which is one of cblock's children
void TensorMultiply(const SpinBlock* ablock, const SparseMatrix& a, const SparseMatrix& b, const SpinBlock* cblock, Wavefunction& c,
                    Wavefunction& v, const SpinQuantum opQ, double scale);
}

namespace Linear {
// linear.C:179-385, state-averaged (lowerStates empty) or state-specific (lowerStates = the lower roots, already orthogonalised
// among themselves by the caller, solver.C:79-86).  The Krylov space lives on the device: the functor is only asked for its
// block (no callbacks into the host during the solve).
void block_davidson(std::vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp, Davidson_functor& h_multiply,
                    bool& useprecond, int currentRoot, std::vector<Wavefunction>& lowerStates);
}

}  // namespace b2d_host
