// Host planner for the construction of ENLARGED-BLOCK operators: which products of the two children's
// operators, with which scalar factors, make up an operator of the enlarged block.  Integer / scalar work only; the products
// themselves run on the device (b2d_product_op_accumulate -> kron_scatter_kernel).
//
// Reference (file:line under the reference root):
//   TensorOp (spin-coupled strings of spin-orbital c / d operators)           tensor_operator.h:27-289
//   SparseMatrix::calcCompfactor                                             Operators.C:87-137
//   Cre / CreDes / CreCre / Overlap ::build                                  Operators.C:453-487, 624-691, 847-900, 2597-2625
//   CreDesComp / DesDesComp ::build                                          Operators.C:1059-1194, 1472-1616
//   CreCreDesComp::build -> opxop::cxcdcomp / dxcccomp (operator forms)      Operators.C:1905-1966, opxop.C:376-532
//   Ham::build -> opxop::cxcddcomp / cdxcdcomp / ddxcccomp (operator forms)  Operators.C:2322-2395, opxop.C:22-143
// Spin-adapted, abelian point group (irreps are bit patterns, product = XOR), energy sweep (no explicit DES operator arrays).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "plan.hpp"

namespace b2d {

struct Integrals {
  int n = 0;                       // spatial orbitals
  std::vector<double> h1, h2;      // v_1(2i,2j) and v_2(2i,2j,2k,2l) as the reference's accessors return them (reordered orbitals)
  std::vector<int> irrep;          // per spatial orbital
  double one_tol = 1e-15, two_tol = 1e-15;
  bool set = false;
  // IntegralMatrix.C:32-46, 312-333 (rhf): zero unless the spins of (i,k) and of (j,l) match
  double v1(int i, int j) const { return ((i & 1) != (j & 1)) ? 0.0 : h1[(size_t)(i / 2) * n + j / 2]; }
  double v2(int i, int j, int k, int l) const {
    if ((i & 1) != (k & 1) || (j & 1) != (l & 1)) return 0.0;
    return h2[(((size_t)(i / 2) * n + j / 2) * n + k / 2) * n + l / 2];
  }
};

struct TensorOpStr {   // tensor_operator.h:27-48
  int spin = 0, irrep = 0;
  bool empty = true;
  std::vector<std::vector<double>> szops;     // index (spin - sz) / 2
  std::vector<std::vector<int>> opindices;
  std::vector<int> optypes;
  TensorOpStr() {}
  TensorOpStr(int k, int sign, const Integrals& ints) {   // :50-126
    const int K = 2 * k;
    empty = false; spin = 1; irrep = ints.irrep[k];
    optypes.assign(1, sign);
    szops = {{sign * 1.0, 0.0}, {0.0, 1.0}};
    if (sign < 0) opindices = {{K + 1}, {K + 0}}; else opindices = {{K + 0}, {K + 1}};
  }
  TensorOpStr product(const TensorOpStr& o, int pspin, int pirrep, AngMom& am) const {   // :196-289 (`identical` forced false)
    TensorOpStr r;
    if (pspin < std::abs(spin - o.spin) || pspin > spin + o.spin) throw std::runtime_error("TensorOp::product: spins do not couple");
    if ((irrep ^ o.irrep) != pirrep) return r;
    r.empty = false;
    r.optypes = optypes; r.optypes.insert(r.optypes.end(), o.optypes.begin(), o.optypes.end());
    for (const auto& a : opindices)
      for (const auto& b : o.opindices) { std::vector<int> v = a; v.insert(v.end(), b.begin(), b.end()); r.opindices.push_back(v); }
    const size_t n2 = o.opindices.size();
    r.szops.assign(pspin + 1, std::vector<double>(r.opindices.size(), 0.0));
    for (int sz = pspin; sz >= -pspin; sz -= 2)
      for (int sz1 = spin; sz1 >= -spin; sz1 -= 2)
        for (int sz2 = o.spin; sz2 >= -o.spin; sz2 -= 2) {
          const double cleb = am.clebsch(spin, sz1, o.spin, sz2, pspin, sz);
          if (std::fabs(cleb) <= 1e-14) continue;
          const auto& v1 = szops[(spin - sz1) / 2];
          const auto& v2 = o.szops[(o.spin - sz2) / 2];
          auto& dst = r.szops[(pspin - sz) / 2];
          for (size_t i = 0; i < v1.size(); ++i)
            for (size_t j = 0; j < v2.size(); ++j) dst[i * n2 + j] += cleb * v1[i] * v2[j];
        }
    r.spin = pspin; r.irrep = pirrep;
    return r;
  }
};

enum CompKind { COMP_CD, COMP_DD };

// SparseMatrix::calcCompfactor(op1, op2, comp, v_2, integralIndex) Operators.C:87-137
inline double calc_compfactor(const TensorOpStr& op1, const TensorOpStr& op2, CompKind comp, const Integrals& I, AngMom& am) {
  double factor = 0.0;
  const auto& c1 = op1.szops[0];
  for (int sz2 = -op2.spin; sz2 <= op2.spin; sz2 += 2) {
    const auto& c2 = op2.szops[(op2.spin - sz2) / 2];
    double cleb = am.clebsch(op1.spin, op1.spin, op2.spin, sz2, 0, 0);
    if ((op1.irrep ^ op2.irrep) != 0) cleb = 0.0;
    if (std::fabs(cleb) <= 1e-14) continue;
    for (size_t i1 = 0; i1 < c1.size(); ++i1)
      for (size_t i2 = 0; i2 < c2.size(); ++i2) {
        if (c1[i1] == 0.0 || c2[i2] == 0.0) continue;
        const auto& a = op1.opindices[i1];
        const auto& b = op2.opindices[i2];
        double t;
        if (comp == COMP_CD)
          t = 0.5 * (-I.v2(a[0], b[0], b[1], a[1]) - I.v2(b[0], a[0], a[1], b[1]) + I.v2(b[0], a[0], b[1], a[1]) + I.v2(a[0], b[0], a[1], b[1]));
        else
          t = 0.5 * I.v2(a[0], a[1], b[1], b[0]);
        factor += t * c1[i1] * c2[i2] / cleb;
      }
    break;   // `found`: only the first Sz component with a non-vanishing coupling
  }
  return factor;
}

struct ProductCall {
  int lop, rop;     // operator ids on the left / right child; -1 = identity (TensorTrace)
  bool lt, rt;      // Transposeview flags
  double scale;
};

class OpBuildPlanner {
 public:
  OpBuildPlanner(const Side& L, const Side& R, const Integrals& ints, bool hubbard, AngMom& am) : L_(L), R_(R), I_(ints), hubbard_(hubbard), am_(am) {}

  // products of the enlarged-block operator (optype, orbs, dq); throws if the children cannot supply it
  std::vector<ProductCall> plan(int optype, const int* orbs, const int* dq) {
    out_.clear();
    switch (optype) {
      case OP_OVERLAP: overlap(); break;
      case OP_CRE: cre(orbs[0], dq); break;
      case OP_CRE_DES: credes(orbs[0], orbs[1], dq); break;
      case OP_CRE_CRE: crecre(orbs[0], orbs[1], dq); break;
      case OP_CRE_DESCOMP: twoindex_comp(OP_CRE_DESCOMP, orbs[0], orbs[1], dq); break;
      case OP_DES_DESCOMP: twoindex_comp(OP_DES_DESCOMP, orbs[0], orbs[1], dq); break;
      case OP_CRE_CRE_DESCOMP: crecredescomp(orbs[0], dq); break;
      case OP_HAM: ham(); break;
      default: throw std::runtime_error("operator type " + std::to_string(optype) + " is not part of the energy sweep");
    }
    return out_;
  }

 private:
  const Side& L_;
  const Side& R_;
  const Integrals& I_;
  bool hubbard_;
  AngMom& am_;
  std::vector<ProductCall> out_;

  static int find(const Side& s, int optype, const int* orbs, int norb, const int* dq) {   // get_op_rep(optype, deltaQuantum, i[, j])
    for (size_t m = 0; m < s.ops.size(); ++m) {
      const OpRec& o = s.ops[m];
      if (o.optype != optype || o.norb != norb) continue;
      bool same = true;
      for (int k = 0; k < norb; ++k) same = same && o.orbs[k] == orbs[k];
      if (!same) continue;
      if (dq && (o.dq[0] != dq[0] || o.dq[1] != dq[1] || o.dq[2] != dq[2])) continue;
      return (int)m;
    }
    return -1;
  }
  static bool has_type(const Side& s, int optype) {
    for (const OpRec& o : s.ops) if (o.optype == optype) return true;
    return false;
  }
  int overlap_of(const Side& s) const {
    int none[2] = {-1, -1};
    int id = find(s, OP_OVERLAP, none, 0, nullptr);
    if (id < 0) throw std::runtime_error("child block without an OVERLAP operator");
    return id;
  }
  int cre_of(const Side& s, int i) const { int o[2] = {i, -1}; return find(s, OP_CRE, o, 1, nullptr); }
  void emit(int lop, bool lt, int rop, bool rt, double scale) { out_.push_back(ProductCall{lop, rop, lt, rt, scale}); }
  static void negq(const int* q, int* o) { o[0] = -q[0]; o[1] = q[1]; o[2] = q[2]; }

  // op on the LEFT child: TensorTrace if the right child is empty, else TensorProduct with its OVERLAP (Operators.C:466-472, ...)
  void left_product_or_trace(int lop) {
    if (R_.sites.empty()) emit(lop, false, -1, false, 1.0);
    else emit(lop, false, overlap_of(R_), false, 1.0);
  }
  void overlap() {                                                            // Operators.C:2597-2625
    if (R_.sites.empty()) emit(overlap_of(L_), false, -1, false, 1.0);
    else emit(overlap_of(L_), false, overlap_of(R_), false, 1.0);
  }
  void cre(int i, const int* dq) {                                            // :453-487
    int o[2] = {i, -1};
    int l = find(L_, OP_CRE, o, 1, dq), r = find(R_, OP_CRE, o, 1, dq);
    if (cre_of(L_, i) >= 0 && l >= 0) left_product_or_trace(l);
    else if (r >= 0) emit(overlap_of(L_), false, r, false, 1.0);
    else throw std::runtime_error("Cre::build: orbital on neither child");
  }
  void credes(int i, int j, const int* dq) {                                  // :624-691 (Transposeview branches)
    int o[2] = {i, j};
    int any_l = find(L_, OP_CRE_DES, o, 2, nullptr), any_r = find(R_, OP_CRE_DES, o, 2, nullptr);
    if (any_l >= 0) { int l = find(L_, OP_CRE_DES, o, 2, dq); if (l < 0) throw std::runtime_error("CreDes::build: component missing"); left_product_or_trace(l); return; }
    if (any_r >= 0) { int r = find(R_, OP_CRE_DES, o, 2, dq); if (r < 0) throw std::runtime_error("CreDes::build: component missing"); emit(overlap_of(L_), false, r, false, 1.0); return; }
    if (cre_of(L_, i) >= 0) { emit(cre_of(L_, i), false, need(cre_of(R_, j)), true, 1.0); return; }
    if (cre_of(R_, i) >= 0) {
      const OpRec& op1 = R_.ops[cre_of(R_, i)];
      const OpRec& op2 = L_.ops[need(cre_of(L_, j))];
      int n2[3]; negq(op2.dq, n2);
      emit(cre_of(L_, j), true, cre_of(R_, i), false, am_.commute_parity(op1.dq, n2, dq));
      return;
    }
    throw std::runtime_error("CreDes::build: orbitals not available");
  }
  void crecre(int i, int j, const int* dq) {                                  // :847-900
    int o[2] = {i, j};
    int any_l = find(L_, OP_CRE_CRE, o, 2, nullptr), any_r = find(R_, OP_CRE_CRE, o, 2, nullptr);
    if (any_l >= 0) { int l = find(L_, OP_CRE_CRE, o, 2, dq); if (l < 0) throw std::runtime_error("CreCre::build: component missing"); left_product_or_trace(l); return; }
    if (any_r >= 0) { int r = find(R_, OP_CRE_CRE, o, 2, dq); if (r < 0) throw std::runtime_error("CreCre::build: component missing"); emit(overlap_of(L_), false, r, false, 1.0); return; }
    if (cre_of(L_, i) >= 0) { emit(cre_of(L_, i), false, need(cre_of(R_, j)), false, 1.0); return; }
    if (cre_of(R_, i) >= 0) {
      const OpRec& op1 = R_.ops[cre_of(R_, i)];
      const OpRec& op2 = L_.ops[need(cre_of(L_, j))];
      emit(cre_of(L_, j), false, cre_of(R_, i), false, am_.commute_parity(op1.dq, op2.dq, dq));
      return;
    }
    throw std::runtime_error("CreCre::build: orbitals not available");
  }
  static int need(int id) { if (id < 0) throw std::runtime_error("operator construction: a CRE operator is missing on a child"); return id; }

  // comp(L) x 1 and 1 x comp(R) (Operators.C:1076-1101); false if the right child is the empty dummy block
  bool carry_over(int optype, const int* orbs, int norb, const int* dq) {
    if (find(L_, optype, orbs, norb, nullptr) >= 0) left_product_or_trace(need(find(L_, optype, orbs, norb, dq)));
    if (R_.sites.empty()) return false;
    if (find(R_, optype, orbs, norb, nullptr) >= 0) emit(overlap_of(L_), false, need(find(R_, optype, orbs, norb, dq)), false, 1.0);
    return true;
  }
  void twoindex_comp(int optype, int i, int j, const int* dq) {               // CreDesComp :1059-1194, DesDesComp :1472-1616
    if (!I_.set) throw std::runtime_error("complementary operators need the integrals (b2d_set_integrals)");
    int o[2] = {i, j};
    const int spin = dq[1], sym = dq[2];
    if (!carry_over(optype, o, 2, dq)) return;
    if (optype == OP_CRE_DESCOMP) {
      TensorOpStr CD1 = TensorOpStr(i, 1, I_).product(TensorOpStr(j, -1, I_), spin, sym, am_);
      for (int k : L_.sites)
        for (int l : R_.sites) {
          const bool have = cre_of(L_, k) >= 0 && cre_of(R_, l) >= 0;
          TensorOpStr CD2 = TensorOpStr(k, 1, I_).product(TensorOpStr(l, -1, I_), spin, sym, am_);
          if (!CD2.empty) {
            double s = calc_compfactor(CD1, CD2, COMP_CD, I_, am_);
            if (have && std::fabs(s) > I_.two_tol) emit(cre_of(L_, k), false, cre_of(R_, l), true, s);          // c+_k(L) x d_l(R)
          }
          CD2 = TensorOpStr(l, 1, I_).product(TensorOpStr(k, -1, I_), spin, sym, am_);
          if (!CD2.empty) {
            double s = calc_compfactor(CD1, CD2, COMP_CD, I_, am_);
            if (have && std::fabs(s) > I_.two_tol) {                                                              // c+_l(R) x d_k(L)
              const OpRec& op1 = R_.ops[cre_of(R_, l)];
              int n2[3]; negq(L_.ops[cre_of(L_, k)].dq, n2);
              emit(cre_of(L_, k), true, cre_of(R_, l), false, s * am_.commute_parity(op1.dq, n2, dq));
            }
          }
        }
    } else {
      TensorOpStr CC1 = TensorOpStr(i, 1, I_).product(TensorOpStr(j, 1, I_), spin, sym, am_);
      for (int k : L_.sites)
        for (int l : R_.sites) {
          TensorOpStr DD2 = TensorOpStr(k, -1, I_).product(TensorOpStr(l, -1, I_), spin, sym, am_);
          if (DD2.empty) continue;
          double s = calc_compfactor(CC1, DD2, COMP_DD, I_, am_);
          double s2 = calc_compfactor(CC1, TensorOpStr(l, -1, I_).product(TensorOpStr(k, -1, I_), spin, sym, am_), COMP_DD, I_, am_);
          if (cre_of(L_, k) >= 0 && cre_of(R_, l) >= 0 && std::fabs(s) + std::fabs(s2) > I_.two_tol) {
            int n1[3], n2[3]; negq(L_.ops[cre_of(L_, k)].dq, n1); negq(R_.ops[cre_of(R_, l)].dq, n2);
            s += am_.commute_parity(n1, n2, dq) * s2;
            if (std::fabs(s) > I_.two_tol) emit(cre_of(L_, k), true, cre_of(R_, l), true, s);                    // d_k(L) x d_l(R)
          }
        }
    }
  }

  double recoupling(int j2, int j1, int j21, int phase_twice) {              // opxop.C:395, 421, 471, 504
    return std::pow(-1.0, (double)(phase_twice / 2)) * am_.six_j(j2, j1, j21, 1, 0, j2) * std::sqrt((j21 + 1.0) * (j2 + 1.0));
  }
  // components (all deltaQuanta) of a two-index operator on one child, in storage order
  static std::vector<int> comps(const Side& s, int optype, int a, int b) {
    std::vector<int> v;
    for (size_t m = 0; m < s.ops.size(); ++m)
      if (s.ops[m].optype == optype && s.ops[m].norb == 2 && s.ops[m].orbs[0] == a && s.ops[m].orbs[1] == b) v.push_back((int)m);
    return v;
  }
  void emit_holder(bool holder_is_left, int holder_op, bool holder_t, int other_op, bool other_t, double scale) {
    if (holder_is_left) emit(holder_op, holder_t, other_op, other_t, scale);
    else emit(other_op, other_t, holder_op, holder_t, scale);
  }
  // opxop::cxcdcomp (operator form) opxop.C:376-445: c+_j (one child) x CDcomp_{jI} or its transpose (the child that holds the comps)
  void cxcdcomp(bool holder_is_left, int op1_id, int I, const int* cdq) {
    const Side& holder = holder_is_left ? L_ : R_;
    const Side& src = holder_is_left ? R_ : L_;
    const OpRec& op1 = src.ops[op1_id];
    const int j = op1.orbs[0], j1 = op1.dq[1], j21 = cdq[1];
    if (j >= I) {
      for (int id : comps(holder, OP_CRE_DESCOMP, j, I)) {
        const OpRec& op2 = holder.ops[id];
        double f = recoupling(op2.dq[1], j1, j21, 2 + op2.dq[1]);
        if (!holder_is_left) f *= am_.commute_parity(op1.dq, op2.dq, cdq);
        emit_holder(holder_is_left, id, false, op1_id, false, f);
      }
    } else {
      for (int id : comps(holder, OP_CRE_DESCOMP, I, j)) {
        const OpRec& op2 = holder.ops[id];
        double f = recoupling(op2.dq[1], j1, j21, 1 + 1 + 0 + op2.dq[1]);
        if (!holder_is_left) { int n2[3]; negq(op2.dq, n2); f *= am_.commute_parity(op1.dq, n2, cdq); }
        if (op2.dq[1] == 2) f = -f;                                            // TensorOp::getTransposeFactorCD, abelian
        emit_holder(holder_is_left, id, true, op1_id, false, f);
      }
    }
  }
  // opxop::dxcccomp (operator form, no DES arrays) opxop.C:447-532: d_j (one child) x DDcomp_{kj}^T (the holder), scale 2
  void dxcccomp(bool holder_is_left, int op1_id, int K, const int* cdq, double scale) {
    const Side& holder = holder_is_left ? L_ : R_;
    const Side& src = holder_is_left ? R_ : L_;
    const OpRec& op1 = src.ops[op1_id];
    int k = K, i = op1.orbs[0];
    bool transpose = false;
    if (k < i) { k = i; i = K; transpose = true; }
    const int iq[3] = {1, 1, I_.irrep[i]}, kq[3] = {1, 1, I_.irrep[k]};
    int niq[3]; negq(iq, niq);
    for (int id : comps(holder, OP_DES_DESCOMP, k, i)) {
      const OpRec& op2 = holder.ops[id];
      int topq[3]; negq(op2.dq, topq);
      double f = recoupling(op2.dq[1], op1.dq[1], cdq[1], 2 + op2.dq[1]);
      if (op2.dq[1] == 0) f = -f;                                              // TensorOp::getTransposeFactorDD, abelian
      if (transpose) f *= am_.commute_parity(iq, kq, topq);
      if (!holder_is_left) f *= am_.commute_parity(niq, topq, kq);             // loop block is the left child
      emit_holder(holder_is_left, id, true, op1_id, true, f * scale);
    }
  }
  void crecredescomp(int k, const int* dq) {                                   // Operators.C:1905-1966
    if (!I_.set) throw std::runtime_error("complementary operators need the integrals (orbital irreps; b2d_set_integrals)");
    int o[2] = {k, -1};
    if (!carry_over(OP_CRE_CRE_DESCOMP, o, 1, dq) || hubbard_) return;
    const bool loop_is_left = L_.loop;                                         // assignloopblock BaseOperator.C:298-304
    const Side& loopb = loop_is_left ? L_ : R_;
    const Side& otherb = loop_is_left ? R_ : L_;
    if (!has_type(loopb, OP_CRE_DESCOMP)) return;
    for (int pass = 0; pass < 2; ++pass) {
      const Side& src = pass == 0 ? loopb : otherb;
      const bool holder_is_left = pass == 0 ? !loop_is_left : loop_is_left;
      for (size_t m = 0; m < src.ops.size(); ++m) {
        if (src.ops[m].optype != OP_CRE) continue;
        cxcdcomp(holder_is_left, (int)m, k, dq);
        dxcccomp(holder_is_left, (int)m, k, dq, 2.0);                          // 2.0: CCcomp_ij = -CCcomp_ji
      }
    }
  }
  void ham() {                                                                 // Operators.C:2322-2395 = the pairs of multiplyH
    int norbs = std::max(1, I_.n);
    for (const Term& t : enumerate_terms(L_, R_, 0.0, hubbard_, norbs, 1, am_)) {
      if (t.kind == TERM_CORE) continue;
      emit(t.lop, t.lt, t.rop, t.rt, t.scale);
    }
  }
};

}  // namespace b2d
