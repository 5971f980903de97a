// Implementation of the C++ host mirror (b2d_host.hpp): marshalling between the reference-shaped containers and the flat
// buffers of the C ABI.  No arithmetic here.
#include "b2d_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace b2d_host {

namespace {
[[noreturn]] void die(const b2d_ctx* ctx, const char* where) {
  // the reference aborts on every error (abort()/exit/assert); so does its drop-in
  fprintf(stderr, "block_b200: %s failed: %s\n", where, b2d_last_error(ctx));
  abort();
}
#define B2D_CK(ctx, call) do { if ((call) != B2D_OK) die(ctx, #call); } while (0)

void block_tables(const SpinBlock& b, std::vector<int32_t>& q, std::vector<int32_t>& dims) {
  const StateInfo& s = b.get_stateInfo();
  q.clear(); dims.clear();
  for (size_t i = 0; i < s.quanta.size(); ++i) {
    q.push_back(s.quanta[i].n); q.push_back(s.quanta[i].s); q.push_back(s.quanta[i].irrep);
    dims.push_back(s.quantaStates[i]);
  }
}
}  // namespace

bool SpinQuantum::allow(const SpinQuantum& s1, const SpinQuantum& s2) const {
  if (n != s1.n + s2.n) return false;
  if (irrep != (s1.irrep ^ s2.irrep)) return false;                       // abelian groups: Symmetry::add = XOR table
  const int lo = std::abs(s1.s - s2.s), hi = s1.s + s2.s;                  // SpinSpace::operator+= SpinSpace.C:24-35
  return s >= lo && s <= hi && ((s - lo) % 2 == 0);
}

int StateInfo::totalStates() const {
  int t = 0;
  for (int d : quantaStates) t += d;
  return t;
}

void SparseMatrix::resize(int r, int c) {
  nr = r; nc = c;
  allowedQuantaMatrix.assign((size_t)r * c, 0);
  operatorMatrix.assign((size_t)r * c, Matrix());
}

void SparseMatrix::allocate(const StateInfo& sr, const StateInfo& sc) {
  resize((int)sr.quanta.size(), (int)sc.quanta.size());
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j) {
      bool ok = false;
      for (const SpinQuantum& dq : deltaQuantum) ok = ok || sr.quanta[i].allow(dq, sc.quanta[j]);
      allowed(i, j) = ok;
      if (ok) operator_element(i, j).ReSize(sr.quantaStates[i], sc.quantaStates[j]);
    }
}

int64_t SparseMatrix::packed_size() const {
  int64_t n = 0;
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
      if (allowed(i, j)) n += operator_element(i, j).Storage();
  return n;
}

void SparseMatrix::FlattenInto(std::vector<double>& flat) const {
  flat.clear();
  flat.reserve((size_t)packed_size());
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
      if (allowed(i, j)) {
        const Matrix& m = operator_element(i, j);
        flat.insert(flat.end(), m.Store(), m.Store() + m.Storage());
      }
}

void SparseMatrix::CollectFrom(const std::vector<double>& flat) {
  size_t off = 0;
  for (int i = 0; i < nr; ++i)
    for (int j = 0; j < nc; ++j)
      if (allowed(i, j)) {
        Matrix& m = operator_element(i, j);
        std::memcpy(m.Store(), flat.data() + off, sizeof(double) * m.Storage());
        off += m.Storage();
      }
}

void Wavefunction::initialise(const SpinQuantum& dQ, const SpinBlock* big, bool onedot_) {
  onedot = onedot_;
  deltaQuantum.assign(1, dQ);
  const StateInfo& sl = big->get_leftBlock()->get_stateInfo();
  const StateInfo& sr = big->get_rightBlock()->get_stateInfo();
  resize((int)sl.quanta.size(), (int)sr.quanta.size());
  for (int l = 0; l < nr; ++l)
    for (int r = 0; r < nc; ++r) {
      allowed(l, r) = dQ.allow(sl.quanta[l], sr.quanta[r]);
      if (allowed(l, r)) operator_element(l, r).ReSize(sl.quantaStates[l], sr.quantaStates[r]);
    }
}

void Wavefunction::Clear() {
  for (Matrix& m : operatorMatrix) std::fill(m.store.begin(), m.store.end(), 0.0);
}

SpinBlock::~SpinBlock() {
  if (ctx) b2d_destroy(ctx);
}

void SpinBlock::set_big_block(SpinBlock* left, SpinBlock* right, const SpinQuantum& target_, const DeviceOptions& o) {
  leftBlock = left; rightBlock = right; target = target_; opt = o;
  left->parent = this;
  if (ctx) { b2d_destroy(ctx); ctx = nullptr; }
  if (b2d_create(o.device, &ctx) != B2D_OK) die(nullptr, "b2d_create");
  SpinBlock* child[2] = {left, right};
  std::vector<int32_t> q, dims;
  std::vector<double> flat;
  for (int side = 0; side < 2; ++side) {
    block_tables(*child[side], q, dims);
    std::vector<int32_t> sites(child[side]->sites.begin(), child[side]->sites.end());
    B2D_CK(ctx, b2d_set_block(ctx, side, (int)dims.size(), q.data(), dims.data(), child[side]->loopblock ? 1 : 0, (int)sites.size(), sites.data()));
    registered[side].clear();
    for (const std::shared_ptr<SparseMatrix>& op : child[side]->ops) {
      int32_t orbs[2] = {-1, -1};
      for (size_t k = 0; k < op->orbs.size() && k < 2; ++k) orbs[k] = op->orbs[k];
      const SpinQuantum dq0 = op->get_deltaQuantum(0);
      int32_t dq[3] = {dq0.n, dq0.s, dq0.irrep};
      std::vector<uint8_t> allowed((size_t)op->nrows() * op->ncols());
      for (int i = 0; i < op->nrows(); ++i)
        for (int j = 0; j < op->ncols(); ++j) allowed[(size_t)i * op->ncols() + j] = op->allowed(i, j) ? 1 : 0;
      op->FlattenInto(flat);
      int id = -1;
      B2D_CK(ctx, b2d_add_op(ctx, side, (int)op->optype, (int)op->orbs.size(), orbs, op->comp, dq, op->fermion ? 1 : 0, allowed.data(),
                             flat.empty() ? nullptr : flat.data(), &id));
      registered[side].push_back(op.get());
    }
  }
  int32_t tq[3] = {target.n, target.s, target.irrep};
  int norbs = o.norbs > 0 ? o.norbs : (int)(left->sites.size() + right->sites.size());
  B2D_CK(ctx, b2d_plan(ctx, tq, o.core_energy, o.ham == HUBBARD ? 1 : 0, norbs, o.rank, o.nranks));
}

int SpinBlock::op_id(int side, const SparseMatrix* op) const {
  for (size_t k = 0; k < registered[side].size(); ++k)
    if (registered[side][k] == op) return (int)k;
  return -1;
}

int64_t SpinBlock::psi_size() const { return b2d_psi_size(ctx); }

void SpinBlock::multiplyH(Wavefunction& c, Wavefunction* v, int /*num_threads*/) const {
  std::vector<double> cf, vf;
  c.FlattenInto(cf);
  v->FlattenInto(vf);
  B2D_CK(ctx, b2d_multiplyH_host(ctx, cf.data(), vf.data(), 1));   // v += H c: the caller cleared v (linear.C:239-240)
  v->CollectFrom(vf);
}

void SpinBlock::diagonalH(DiagonalMatrix& e) const {
  B2D_CK(ctx, b2d_vec_reserve(ctx, 1));
  B2D_CK(ctx, b2d_diagonal(ctx, 0));
  e.assign((size_t)psi_size(), 0.0);
  B2D_CK(ctx, b2d_vec_download(ctx, 0, e.data()));
}

void SpinBlock::RenormaliseFrom(std::vector<double>& energies, std::vector<double>& spins, double& error, std::vector<Matrix>& rotateMatrix,
                                const int keptstates, const int keptqstates, const double tol, SpinBlock& big, const guessWaveTypes&,
                                const double noise, const double additional_noise, const bool& onedot, SpinBlock&, SpinBlock&, SpinBlock&,
                                const bool&, const bool&, int, int currentRoot, std::vector<Wavefunction>& lowerStates,
                                std::vector<Wavefunction>* solution, const std::vector<double>* weights) {
  b2d_ctx* c = big.context();
  if (!c || big.get_leftBlock() != this) { fprintf(stderr, "block_b200: RenormaliseFrom: `this` must be the left child of `big`\n"); abort(); }
  if (additional_noise != 0.0 || onedot || keptqstates != 0 || currentRoot >= 0 || !lowerStates.empty() || !solution || solution->empty()) {
    fprintf(stderr, "block_b200: RenormaliseFrom: only the two-dot, state-averaged form with caller-supplied guesses and without additional (random) noise is on the GPU path yet\n");
    abort();
  }
  const int nroots = (int)solution->size();
  std::vector<double> w = weights ? *weights : std::vector<double>(nroots, 1.0 / nroots);
  B2D_CK(c, b2d_vec_reserve(c, nroots + 1));
  std::vector<double> flat;
  for (int i = 0; i < nroots; ++i) {
    (*solution)[i].FlattenInto(flat);
    B2D_CK(c, b2d_vec_upload(c, i, flat.data()));
  }
  const StateInfo& sl = get_stateInfo();
  std::vector<int32_t> kept(sl.quanta.size(), 0);
  energies.assign(nroots, 0.0);
  spins.assign(nroots, 0.0);
  int nmult = 0;
  B2D_CK(c, b2d_renormalise_from(c, nroots, 0, w.data(), tol, keptstates, big.options().deflation_min, big.options().deflation_max, noise, energies.data(),
                                 kept.data(), &error, &nmult));
  flat.assign((size_t)big.psi_size(), 0.0);
  for (int i = 0; i < nroots; ++i) {
    B2D_CK(c, b2d_vec_download(c, i, flat.data()));
    (*solution)[i].CollectFrom(flat);
  }
  // rotateMatrix[q]: d_q x kept_q, Ncols() == 0 for a dropped sector (rotationmat.C:149-256)
  std::vector<double> rot((size_t)std::max<int64_t>(b2d_rotation_size(c), 1));
  B2D_CK(c, b2d_rotation_download(c, rot.data()));
  rotateMatrix.assign(sl.quanta.size(), Matrix());
  size_t off = 0;
  for (size_t q = 0; q < sl.quanta.size(); ++q) {
    if (kept[q] == 0) { rotateMatrix[q].nrows = sl.quantaStates[q]; continue; }
    rotateMatrix[q].ReSize(sl.quantaStates[q], kept[q]);
    std::memcpy(rotateMatrix[q].Store(), rot.data() + off, sizeof(double) * rotateMatrix[q].Storage());
    off += rotateMatrix[q].Storage();
  }
}

void SpinBlock::transform_operators(std::vector<Matrix>& rotateMatrix) {
  // `this` is the left child of a big block that holds the device context
  SpinBlock* big = parent;
  if (!big || !big->context()) { fprintf(stderr, "block_b200: transform_operators: block is not the left child of a device-resident big block\n"); abort(); }
  b2d_ctx* c = big->context();
  std::vector<int32_t> kept(rotateMatrix.size());
  std::vector<double> rot;
  for (size_t q = 0; q < rotateMatrix.size(); ++q) {
    kept[q] = rotateMatrix[q].Ncols();
    rot.insert(rot.end(), rotateMatrix[q].store.begin(), rotateMatrix[q].store.end());
  }
  rot.push_back(0.0);
  B2D_CK(c, b2d_rotation_upload(c, kept.data(), rot.data()));
  B2D_CK(c, b2d_transform_operators(c));
  // new StateInfo: sectors with >= 1 kept state (save_load_block.C:270-283); operators replaced by their rotated blocks
  const int nq = b2d_rotated_num_sectors(c);
  std::vector<int32_t> old(nq), dims(nq);
  B2D_CK(c, b2d_rotated_sectors(c, old.data(), dims.data()));
  StateInfo ns;
  for (int a = 0; a < nq; ++a) { ns.quanta.push_back(stateInfo.quanta[old[a]]); ns.quantaStates.push_back(dims[a]); }
  std::vector<uint8_t> allowed((size_t)nq * nq);
  std::vector<double> flat;
  for (size_t k = 0; k < ops.size(); ++k) {
    SparseMatrix& op = *ops[k];
    const int id = big->op_id(0, &op);
    flat.assign((size_t)std::max<int64_t>(b2d_rotated_op_size(c, id), 1), 0.0);
    B2D_CK(c, b2d_rotated_op_download(c, id, allowed.data(), flat.data()));
    op.resize(nq, nq);
    for (int a = 0; a < nq; ++a)
      for (int b = 0; b < nq; ++b)
        if (allowed[(size_t)a * nq + b]) { op.allowed(a, b) = 1; op.operator_element(a, b).ReSize(dims[a], dims[b]); }
    op.CollectFrom(flat);
  }
  stateInfo = ns;
}

namespace operatorfunctions {
void TensorMultiply(const SpinBlock* ablock, const SparseMatrix& a, const SparseMatrix& b, const SpinBlock* cblock, Wavefunction& c,
                    Wavefunction& v, const SpinQuantum opQ, double scale) {
  b2d_ctx* ctx = cblock->context();
  const bool a_left = cblock->get_leftBlock() == ablock;                    // operatorfunctions.C:496-501
  const SparseMatrix& lop = a_left ? a : b;
  const SparseMatrix& rop = a_left ? b : a;
  const int lid = cblock->op_id(0, &lop.stored()), rid = cblock->op_id(1, &rop.stored());
  if (lid < 0 || rid < 0) { fprintf(stderr, "block_b200: TensorMultiply: operator is not registered with the big block\n"); abort(); }
  const int flags = (lop.conjugacy() == 't' ? 1 : 0) | (rop.conjugacy() == 't' ? 2 : 0);
  std::vector<double> cf, vf;
  c.FlattenInto(cf);
  v.FlattenInto(vf);
  B2D_CK(ctx, b2d_vec_reserve(ctx, 2));
  B2D_CK(ctx, b2d_vec_upload(ctx, 0, cf.data()));
  B2D_CK(ctx, b2d_vec_upload(ctx, 1, vf.data()));
  B2D_CK(ctx, b2d_tensor_multiply(ctx, lid, rid, flags, opQ.s, scale, 0, 1));
  B2D_CK(ctx, b2d_vec_download(ctx, 1, vf.data()));
  v.CollectFrom(vf);
}
}  // namespace operatorfunctions

namespace Linear {
void block_davidson(std::vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool&, Davidson_functor& h_multiply, bool&,
                    int /*currentRoot*/, std::vector<Wavefunction>& lowerStates) {
  const SpinBlock& big = h_multiply.get_block();
  b2d_ctx* ctx = big.context();
  const int n = (int)b.size(), nlow = (int)lowerStates.size();   // lowerStates non-empty: the state-specific form (linear.C:201-208,311-317,369-375)
  B2D_CK(ctx, b2d_vec_reserve(ctx, n + 1 + nlow));
  std::vector<double> flat;
  for (int i = 0; i < n; ++i) {
    b[i].FlattenInto(flat);
    B2D_CK(ctx, b2d_vec_upload(ctx, i, flat.data()));
  }
  B2D_CK(ctx, b2d_vec_upload(ctx, n, h_diag.data()));
  for (int i = 0; i < nlow; ++i) {
    lowerStates[i].FlattenInto(flat);
    B2D_CK(ctx, b2d_vec_upload(ctx, n + 1 + i, flat.data()));
  }
  std::vector<double> evals(n);
  int nmult = 0;
  double res = 0.0;
  B2D_CK(ctx, b2d_davidson_lower(ctx, n, 0, n, normtol, big.options().deflation_min, big.options().deflation_max, nlow, n + 1, evals.data(), &nmult,
                                 &res));
  flat.assign((size_t)big.psi_size(), 0.0);
  for (int i = 0; i < n; ++i) {
    B2D_CK(ctx, b2d_vec_download(ctx, i, flat.data()));
    b[i].CollectFrom(flat);
    h_diag[i] = evals[i];                                                  // linear.C:344-345
  }
}
}  // namespace Linear

}  // namespace b2d_host
