// Host-side planning of the sweep hot path: sector tables, packed device layouts, the term list of
// SpinBlock::multiplyH and the grouped-contraction schedule the sm_100a kernels execute.  Pure integer/scalar work,
// no CUDA: this is what the CPU tests exercise through a planning-only context.
//
// Reference behaviour restated here (file:line under the reference root):
//   psi layout            Wavefunction::initialise wavefunction.C:18-56, FlattenInto :167-186
//   operator allocation   SparseMatrix::allocate BaseOperator.C:123-145 (allowed mask is passed in by the caller)
//   term list             SpinBlock::multiplyH spinblock.C:722-789, opxop::{cxcddcomp,cdxcdcomp,ddxcccomp} opxop.C:155-285
//   per-GEMM factors      operatorfunctions::TensorMultiply operatorfunctions.C:485-537
//   diag(H)               SpinBlock::diagonalH spinblock.C:855-899, operatorfunctions.C:653-762, opxop.C:295-365
//   term ownership        processorindex para_array.h:33-42, trimap_2d para_array.h:360-383
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "angmom.hpp"
#include "gemm_desc.h"

namespace b2d {

constexpr int LD_ALIGN = 2;    // leading dimensions are multiples of 2 doubles: every row starts 16-byte aligned
constexpr int BLK_ALIGN = 16;  // sector blocks start on 128-byte boundaries

inline int pad_ld(int n) { return (n + LD_ALIGN - 1) / LD_ALIGN * LD_ALIGN; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

enum { OP_HAM = 0, OP_CRE = 1, OP_CRE_CRE = 2, OP_DES_DESCOMP = 3, OP_CRE_DES = 4, OP_CRE_DESCOMP = 5, OP_CRE_CRE_DESCOMP = 6, OP_OVERLAP = 13 };

// One (row piece, column piece) contribution to a sector block of a FACTORISED enlarged-block operator (see OpRec::factorised):
//   block[r0 : r0 + m, c0 : c0 + n] += alpha * op(A),   op(A) = A (stored m x n) or A^T (stored n x m, t = true), leading dimension lda
// A is a sector block of an operator of the renormalised child (or that product's identity block); the 1 x 1 element of the dot operator,
// the 9j coefficient, Transposeview scalings and the fermion sign of operatorfunctions.C:205-218 are folded into alpha.
struct SubBlock {
  int32_t r0, c0, m, n;
  const double* a;
  int32_t lda;
  bool t;
  double alpha;
};

struct OpRec {
  int optype = 0, norb = 0, orbs[2] = {-1, -1}, comp = 0;
  int dq[3] = {0, 0, 0};
  bool fermion = false;
  std::vector<uint8_t> allowed;   // nq x nq
  std::vector<int64_t> off;       // nq x nq: offset (doubles) of block (i,j) from the operator's device base, -1 if absent
  int64_t packed_size = 0;        // host (reference) layout
  int64_t dev_size = 0;           // padded device layout
  double* dev = nullptr;
  bool pending = false;           // selected for allocation by the current b2d_plan
  int64_t cache_off = -1;         // inside a cached block (b2d_cache_*): offset of the operator from the entry's buffer, -1 if not resident
  // Factorised form (operators of an enlarged block S (x) dot, SURVEY.md 7 "hard parts"): the block is never materialised; every allowed
  // sector block (i,j) is the list subs[sub_begin[i * nq + j] .. sub_begin[i * nq + j + 1]) of scaled sub-blocks of the renormalised child's
  // operators.  `off` / `dev_size` keep describing the materialised layout (used when the operator is materialised on request).
  bool factorised = false;
  std::vector<int32_t> sub_begin;
  std::vector<SubBlock> subs;
  std::vector<double> host;       // host copy of a SMALL operator's packed blocks (the one-site dot: its 1 x 1 elements become alphas)
  bool resident() const { return dev != nullptr || factorised || dev_size == 0; }
};

struct Side {
  int nq = 0;
  std::vector<int> q;      // nq x 3
  std::vector<int> dims;
  bool loop = false;
  std::vector<int> sites;
  std::vector<OpRec> ops;
  // un-collected pieces of every sector (offset, size) when the block is a product S (x) dot (StateInfo::oldToNewState order); empty
  // for a block that was uploaded as it is
  std::vector<std::vector<std::pair<int, int>>> pieces;
  const int* quantum(int i) const { return &q[3 * i]; }
  int find(int optype, const int* orbs, int norb, int comp) const {
    for (size_t m = 0; m < ops.size(); ++m) {
      const OpRec& o = ops[m];
      if (o.optype != optype || o.comp != comp || o.norb != norb) continue;
      bool same = true;
      for (int k = 0; k < norb; ++k) same = same && o.orbs[k] == orbs[k];
      if (same) return (int)m;
    }
    return -1;
  }
};

// lay an operator's allowed blocks out in the padded device arena order ((i,j) row-major like the host packing)
inline void layout_op(const Side& s, OpRec& op) {
  op.off.assign((size_t)s.nq * s.nq, -1);
  int64_t dev = 0, packed = 0;
  for (int i = 0; i < s.nq; ++i)
    for (int j = 0; j < s.nq; ++j)
      if (op.allowed[(size_t)i * s.nq + j]) {
        op.off[(size_t)i * s.nq + j] = dev;
        dev += align_up((int64_t)s.dims[i] * pad_ld(s.dims[j]), BLK_ALIGN);
        packed += (int64_t)s.dims[i] * s.dims[j];
      }
  op.dev_size = dev;
  op.packed_size = packed;
}

// an operator or its Transposeview (BaseOperator.h:208-243) as TensorMultiply sees it
struct View {
  const Side* side;
  const OpRec* op;
  bool t;
  bool allowed(int i, int j) const { return t ? op->allowed[(size_t)j * side->nq + i] : op->allowed[(size_t)i * side->nq + j]; }
  int spin() const { return op->dq[1]; }
  bool fermion() const { return op->fermion; }
  // stored block that backs view element (i,j): (j,i) for a Transposeview
  int64_t stored_off(int i, int j) const { return t ? op->off[(size_t)j * side->nq + i] : op->off[(size_t)i * side->nq + j]; }
  int stored_ld(int i, int j) const { return pad_ld(t ? side->dims[i] : side->dims[j]); }
  double scaling(AngMom& am, int i, int j) const {
    return t ? am.transpose_scaling(op->dq[1], side->quantum(i)[1], side->quantum(j)[1]) : 1.0;
  }
  // view element (i,j) as scaled sub-blocks: one full-size entry for a materialised operator, the factor list otherwise
  template <class F>
  void for_each_sub(int i, int j, F&& f) const {
    if (!op->factorised) {
      SubBlock s;
      s.r0 = 0; s.c0 = 0; s.m = side->dims[i]; s.n = side->dims[j];
      s.a = op->dev + stored_off(i, j); s.lda = stored_ld(i, j); s.t = t; s.alpha = 1.0;
      f(s);
      return;
    }
    const size_t b = t ? (size_t)j * side->nq + i : (size_t)i * side->nq + j;
    for (int32_t k = op->sub_begin[b]; k < op->sub_begin[b + 1]; ++k) {
      SubBlock s = op->subs[k];
      if (t) { std::swap(s.r0, s.c0); std::swap(s.m, s.n); s.t = !s.t; }
      f(s);
    }
  }
};

struct PsiLayout {
  int nl = 0, nr = 0;
  int dq[3] = {0, 0, 0};
  std::vector<int> blk;            // nl x nr -> block index or -1
  std::vector<int> bl, br, rows, cols, ld;
  std::vector<int64_t> ref_off, dev_off;
  int64_t W = 0, Wp = 0;
  int nblocks() const { return (int)bl.size(); }
  bool allowed(int l, int r) const { return blk[(size_t)l * nr + r] >= 0; }
  void build(const Side& L, const Side& R, const int* target) {
    nl = L.nq; nr = R.nq;
    std::memcpy(dq, target, sizeof(dq));
    blk.assign((size_t)nl * nr, -1);
    bl.clear(); br.clear(); rows.clear(); cols.clear(); ld.clear(); ref_off.clear(); dev_off.clear();
    W = Wp = 0;
    for (int l = 0; l < nl; ++l)
      for (int r = 0; r < nr; ++r)
        if (qn_allow(target, L.quantum(l), R.quantum(r))) {
          blk[(size_t)l * nr + r] = (int)bl.size();
          bl.push_back(l); br.push_back(r);
          rows.push_back(L.dims[l]); cols.push_back(R.dims[r]); ld.push_back(pad_ld(R.dims[r]));
          ref_off.push_back(W); dev_off.push_back(Wp);
          W += (int64_t)L.dims[l] * R.dims[r];
          Wp += align_up((int64_t)L.dims[l] * pad_ld(R.dims[r]), BLK_ALIGN);
        }
  }
};

struct Term {
  int lop, rop;     // operator ids on the left / right child
  bool lt, rt;      // Transposeview flags
  double scale;
  int owner;        // rank that executes it
  int kind;         // TERM_CORE / TERM_HAM_LEFT / TERM_HAM_RIGHT / TERM_PAIR
};
enum { TERM_CORE = 0, TERM_HAM_LEFT = 1, TERM_HAM_RIGHT = 2, TERM_PAIR = 3 };

inline int tristore_2d(int i) { return i * (i + 1) / 2; }
// para_array.h:360-383
inline int trimap_2d(int i, int j, int length) {
  if (i < j) std::swap(i, j);
  int halflen = length / 2;
  if (i >= halflen && j >= halflen) return tristore_2d(length - j - 1) + length - i - 1;
  if (i < halflen && j < halflen) return tristore_2d(length - halflen - 1) + length - halflen + tristore_2d(i) + j;
  int base = tristore_2d(length - halflen - 1) + length - halflen + tristore_2d(halflen);
  return base + (i - halflen) * halflen + j;
}

// operator arrays of one type in storage order: (orbs) -> components, like Op_component::get_local_element
struct OpArray {
  std::vector<std::vector<int>> comps;   // per element: op ids by component index
  std::vector<const int*> orbs;
};
inline OpArray op_array(const Side& s, int optype) {
  OpArray a;
  std::map<std::pair<int, int>, int> index;
  for (size_t m = 0; m < s.ops.size(); ++m) {
    const OpRec& o = s.ops[m];
    if (o.optype != optype) continue;
    std::pair<int, int> key(o.orbs[0], o.orbs[1]);
    auto it = index.find(key);
    if (it == index.end()) {
      it = index.emplace(key, (int)a.comps.size()).first;
      a.comps.emplace_back();
      a.orbs.push_back(o.orbs);
    }
    std::vector<int>& c = a.comps[it->second];
    if ((int)c.size() <= o.comp) c.resize(o.comp + 1, -1);
    c[o.comp] = (int)m;
  }
  return a;
}

// The TensorMultiply calls of one multiplyH (energy sweep: implicit transposes, no DES/.. arrays), every rank's.
inline std::vector<Term> enumerate_terms(const Side& L, const Side& R, double core_energy, bool hubbard, int norbs, int nranks, AngMom& am) {
  std::vector<Term> terms;
  const int hq[3] = {0, 0, 0};
  auto neg = [](const int* q, int* out) { out[0] = -q[0]; out[1] = q[1]; out[2] = q[2]; };
  auto pair = [&](bool other_is_left, int op_other, bool t_other, int op_loop, bool t_loop, double scale, int owner) {
    Term t;
    if (other_is_left) { t.lop = op_other; t.lt = t_other; t.rop = op_loop; t.rt = t_loop; }
    else { t.lop = op_loop; t.lt = t_loop; t.rop = op_other; t.rt = t_other; }
    t.scale = scale; t.owner = owner; t.kind = TERM_PAIR;
    terms.push_back(t);
  };
  int none[2] = {-1, -1};
  int ovl_l = L.find(OP_OVERLAP, none, 0, 0), ovl_r = R.find(OP_OVERLAP, none, 0, 0);
  int ham_l = L.find(OP_HAM, none, 0, 0), ham_r = R.find(OP_HAM, none, 0, 0);
  if (ovl_l < 0 || ovl_r < 0 || ham_l < 0 || ham_r < 0) throw std::runtime_error("plan: both children need HAM and OVERLAP operators");
  if (std::fabs(core_energy) > 1e-20) terms.push_back(Term{ovl_l, ovl_r, false, false, core_energy, 0, TERM_CORE});   // spinblock.C:735-740 (rank 0)
  terms.push_back(Term{ham_l, ovl_r, false, false, 1.0, 0, TERM_HAM_LEFT});                                              // :742-744
  terms.push_back(Term{ovl_l, ham_r, false, false, 1.0, 0, TERM_HAM_RIGHT});                                              // :745-747

  // c x ccd_comp, both directions (:757-763 -> opxop.C:232-285)
  for (int dir = 0; dir < 2; ++dir) {
    const Side& other = dir == 0 ? L : R;
    const Side& loopb = dir == 0 ? R : L;
    bool other_is_left = dir == 0;
    OpArray arr = op_array(loopb, OP_CRE);
    for (size_t e = 0; e < arr.comps.size(); ++e)
      for (size_t k = 0; k < arr.comps[e].size(); ++k) {
        int i1 = arr.comps[e][k];
        if (i1 < 0) continue;
        const OpRec& op1 = loopb.ops[i1];
        int i2 = other.find(OP_CRE_CRE_DESCOMP, op1.orbs, 1, (int)k);
        if (i2 < 0) break;                                            // has_local_index false (opxop.C:240)
        const OpRec& op2 = other.ops[i2];
        int owner = nranks > 1 ? op1.orbs[0] % nranks : 0;            // processorindex(i)
        int nq1[3], nq2[3];
        neg(op1.dq, nq1); neg(op2.dq, nq2);
        double par = !other_is_left ? am.commute_parity(nq1, op2.dq, hq) : 1.0;   // opxop.C:273-274
        pair(other_is_left, i2, false, i1, true, par, owner);                       // :276
        par = other_is_left ? am.commute_parity(op1.dq, nq2, hq) : 1.0;            // :278-279
        pair(other_is_left, i2, true, i1, false, par, owner);                       // :282
      }
  }
  if (!hubbard) {                                                     // spinblock.C:771
    const Side& loopb = L.loop ? L : R;
    const Side& other = L.loop ? R : L;
    bool other_is_left = !L.loop;
    OpArray arr = op_array(loopb, OP_CRE_DES);                        // cdxcdcomp opxop.C:155-185
    for (size_t e = 0; e < arr.comps.size(); ++e)
      for (size_t k = 0; k < arr.comps[e].size(); ++k) {
        int i1 = arr.comps[e][k];
        if (i1 < 0) continue;
        const OpRec& op1 = loopb.ops[i1];
        int i2 = other.find(OP_CRE_DESCOMP, op1.orbs, 2, (int)k);
        if (i2 < 0) break;
        int owner = nranks > 1 ? trimap_2d(op1.orbs[0], op1.orbs[1], norbs) % nranks : 0;
        pair(other_is_left, i2, false, i1, false, 1.0, owner);
        if (op1.orbs[0] != op1.orbs[1]) pair(other_is_left, i2, true, i1, true, 1.0, owner);
      }
    arr = op_array(loopb, OP_CRE_CRE);                                // ddxcccomp opxop.C:187-228
    for (size_t e = 0; e < arr.comps.size(); ++e)
      for (size_t k = 0; k < arr.comps[e].size(); ++k) {
        int i1 = arr.comps[e][k];
        if (i1 < 0) continue;
        const OpRec& op1 = loopb.ops[i1];
        int i2 = other.find(OP_DES_DESCOMP, op1.orbs, 2, (int)k);
        if (i2 < 0) break;
        const OpRec& op2 = other.ops[i2];
        int owner = nranks > 1 ? trimap_2d(op1.orbs[0], op1.orbs[1], norbs) % nranks : 0;
        double factor = op1.orbs[0] == op1.orbs[1] ? 1.0 : 2.0;
        double par = other_is_left ? am.commute_parity(op1.dq, op2.dq, hq) : 1.0;
        pair(other_is_left, i2, false, i1, false, factor * par, owner);
        par *= AngMom::transpose_factor_dd(op1.dq[1]) * AngMom::transpose_factor_dd(op2.dq[1]);
        pair(other_is_left, i2, true, i1, true, factor * par, owner);
      }
  }
  return terms;
}

// Cost-weighted ownership ("allow a cost-weighted reassignment as long as the sum is unchanged"): longest-processing-time
// greedy over the terms' executed flops - sort by cost (ties: term order), give each term to the least loaded rank (ties: lowest rank).
// Deterministic, so every rank computes the same assignment without communication.
inline void balance_owners(std::vector<Term>& terms, const std::vector<double>& cost, int nranks) {
  std::vector<int> order(terms.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  std::vector<double> load(nranks, 0.0);
  for (int i : order) {
    int best = 0;
    for (int r = 1; r < nranks; ++r) if (load[r] < load[best]) best = r;
    terms[i].owner = best;
    load[best] += cost[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// grouped-contraction schedule
// ---------------------------------------------------------------------------------------------------------------
struct GemmBatch {
  std::vector<GSeg> segs;
  std::vector<GGroup> groups;
  std::vector<GTile> tiles[B2D_NUM_TILE_CLASSES];
  double flops = 0.0;
  double class_flops[B2D_NUM_TILE_CLASSES] = {};    // useful flops (2 m n k) executed by each tile class
  double class_padded[B2D_NUM_TILE_CLASSES] = {};   // flops the tiles issue (tile area x pipeline iterations x 16)
  bool unit_alpha = false;                           // every segment has alpha == 1
  bool empty() const { return groups.empty(); }
};

// Cover a dimension of `d` rows (or columns) with bands: 128-wide bands, then ONE narrower band (or 64 + 32) for the
// remainder, so that ragged quantum-number sectors do not feed the tensor pipe padding.  Returns (origin, size index)
// pairs, size index 0/1/2 = 128/64/32.  forced >= 0 tiles everything with that size.
inline void make_bands(int d, int forced, std::vector<std::pair<int, int>>& out) {
  out.clear();
  if (forced >= 0) {
    int b = 128 >> forced;
    for (int o = 0; o < d; o += b) out.emplace_back(o, forced);
    return;
  }
  int o = 0;
  while (d - o > 96) { out.emplace_back(o, 0); o += 128; }
  int r = d - o;
  if (r <= 0) return;
  if (r <= 32) out.emplace_back(o, 2);
  else if (r <= 64) out.emplace_back(o, 1);
  else { out.emplace_back(o, 1); out.emplace_back(o + 64, 2); }   // 65..96
}

// fills batch.tiles from batch.groups; tiles of a class are ordered by descending cost so the hardware's in-order
// CTA dispatch approximates longest-processing-time scheduling over the 148 SMs
inline void make_tiles(GemmBatch& b, int forced_class) {
  for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) { b.tiles[c].clear(); b.class_flops[c] = b.class_padded[c] = 0.0; }
  b.unit_alpha = true;
  for (const GSeg& s : b.segs) if (s.alpha != 1.0) { b.unit_alpha = false; break; }
  std::vector<std::pair<int, int>> mb, nb;
  for (size_t g = 0; g < b.groups.size(); ++g) {
    GGroup& G = b.groups[g];
    int64_t ktot = 0;
    int kiters = 0;
    for (int s = G.seg_begin; s < G.seg_end; ++s) { ktot += b.segs[s].k; kiters += (b.segs[s].k + 15) / 16; }
    G.kiters = kiters;
    // tiny sectors (both dimensions <= B2D_TINY_DIM): one warp per block, no shared-memory pipeline (forced_class 3 = only these)
    if ((forced_class == -1 || forced_class == 3) && G.m <= B2D_TINY_DIM && G.n <= B2D_TINY_DIM) {
      b.class_flops[B2D_TINY_CLASS] += 2.0 * G.m * G.n * (double)ktot;
      b.class_padded[B2D_TINY_CLASS] += 2.0 * G.m * G.n * (double)ktot;   // DFMA path: no padded tensor-pipe work is issued
      GTile t;
      t.group = (int)g; t.m0 = 0; t.n0 = 0;
      t.cost = (int)std::min<int64_t>(ktot + 8 * (G.seg_end - G.seg_begin), 0x7fffffff);
      b.tiles[B2D_TINY_CLASS].push_back(t);
      continue;
    }
    // forced_class: -1 auto; 3 tiny where eligible, auto elsewhere; 0/1/2 square 128/64/32; 10*(r+1)+c forces (128>>r) x (128>>c) tiles (experiments)
    const int fc = forced_class == 3 ? -1 : forced_class;
    make_bands(G.m, fc >= 10 ? fc / 10 - 1 : fc, mb);
    make_bands(G.n, fc >= 10 ? fc % 10 : fc, nb);
    for (const auto& bm : mb)
      for (const auto& bn : nb) {
        const int c = 3 * bm.second + bn.second;
        if ((G.pad0 == 1 && c != 0) || (G.pad0 == 2 && c == 0)) continue;   // split-K families of one sigma block (build_schedule): 128 x 128 tiles / the others
        const int tm = 128 >> bm.second, tn = 128 >> bn.second;
        const int um = std::min(tm, G.m - bm.first), un = std::min(tn, G.n - bn.first);
        b.class_flops[c] += 2.0 * um * un * (double)ktot;
        b.class_padded[c] += 2.0 * tm * tn * 16.0 * kiters;
        GTile t;
        t.group = (int)g; t.m0 = bm.first; t.n0 = bn.first;
        t.cost = (int)std::min<int64_t>(((int64_t)kiters * 16 + 8 * (G.seg_end - G.seg_begin)) * (tm / 32) * (tn / 32), 0x7fffffff);
        b.tiles[c].push_back(t);
      }
  }
  for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c)
    std::stable_sort(b.tiles[c].begin(), b.tiles[c].end(), [](const GTile& x, const GTile& y) { return x.cost > y.cost; });
}

struct Chunk {
  GemmBatch step1, step2;
  int64_t work = 0;   // doubles of T workspace
  int nterms = 0;
  bool zero_work = false;   // a factorised left operator writes only the row pieces it has factors for: T must start from zero
};

struct Schedule {
  std::vector<Chunk> chunks;
  double flops_alg = 0.0;    // the reference's dgemm flops
  double flops_exec = 0.0;   // what the kernels execute (T blocks nobody consumes are skipped)
  int64_t work_max = 0;
  int64_t n_step1 = 0, n_step2 = 0, n_tiles = 0;
  int nslices = 1;           // split-K copies of the destination used by step 2 (slice 0 is the destination itself)
  std::vector<double> term_flops;   // executed flops of each term, in the order of the term list (cost-weighted ownership)
};

constexpr int SPLITK_MAX = 8;   // K slices of one sigma block computed by different CTAs into private partial copies
constexpr int SPLITK_NARROW_MAX = 32;   // ... for the narrow tiles of a sigma block (option slice_iters_narrow)

// Build the two-step schedule for a list of operator pairs:  dst[lQ,rQ] += F * (s A_L[lQ,lQ'] src[lQ',rQ']) A_R[rQ,rQ']^T
inline Schedule build_schedule(const Side& L, const Side& R, const PsiLayout& P, const std::vector<Term>& terms, int opq_spin,
                               int64_t work_budget, int forced_class, AngMom& am, int slice_iters = 256, int slice_iters_narrow = 0) {
  Schedule S;
  const int S_psi = P.dq[1];
  Chunk cur;
  std::map<std::pair<int, int>, int> group_of;   // (psi block, first column of the output piece) -> group index inside cur.step2
  std::vector<std::vector<GSeg>> pending;   // per group: segments (merged into contiguous ranges at chunk close)
  auto open_chunk = [&]() {
    cur = Chunk();
    group_of.clear();
    pending.clear();
  };
  auto close_chunk = [&]() {
    if (cur.nterms == 0) return;
    // Split-K: a sigma block receives hundreds of segments per chunk but there are only a few hundred sigma tiles in
    // total, i.e. ~3 waves over 148 SMs with a long serial K loop each.  Cut the segment list of a group into up to
    // SPLITK_MAX slices of ~slice_iters pipeline iterations; slice 0 accumulates into the destination, slice s >= 1
    // into partial copy s - 1 (base AUX, stride P.Wp) which the caller sums in a fixed order afterwards: deterministic.
    const size_t ngroups = cur.step2.groups.size();
    std::vector<GGroup> sliced;
    for (size_t g = 0; g < ngroups; ++g) {
      const GGroup G0 = cur.step2.groups[g];
      int64_t iters = 0;
      for (const GSeg& sg : pending[g]) iters += (sg.k + 15) / 16;
      int ns = (int)std::min<int64_t>(SPLITK_MAX, std::max<int64_t>(1, (iters + slice_iters / 2) / std::max(slice_iters, 1)));
      if (slice_iters <= 0) ns = 1;
      // The narrow tiles of a sigma block (remainder bands of a ragged sector: few warps per CTA, latency bound per pipeline iteration) run
      // the same K loop as its 128 x 128 tiles and finish long after them.  Option slice_iters_narrow: they get their OWN, finer slicing
      // (a second family of slice groups that make_tiles restricts to the narrow classes), i.e. more and shorter CTAs.
      int ns_narrow = 0;
      bool has_big = false, has_narrow = false;
      if (slice_iters_narrow > 0 && slice_iters > 0 && !((forced_class == -1 || forced_class == 3) && G0.m <= B2D_TINY_DIM && G0.n <= B2D_TINY_DIM)) {
        const int fc = forced_class == 3 ? -1 : forced_class;
        std::vector<std::pair<int, int>> mb, nb;
        make_bands(G0.m, fc >= 10 ? fc / 10 - 1 : fc, mb);
        make_bands(G0.n, fc >= 10 ? fc % 10 : fc, nb);
        for (const auto& bm : mb)
          for (const auto& bn : nb) { if (bm.second == 0 && bn.second == 0) has_big = true; else has_narrow = true; }
        ns_narrow = (int)std::min<int64_t>(SPLITK_NARROW_MAX, std::max<int64_t>(1, (iters + slice_iters_narrow / 2) / slice_iters_narrow));
        if (!has_narrow || ns_narrow <= ns) ns_narrow = 0;
      }
      auto emit_family = [&](int nsl, uint8_t filter) {
        S.nslices = std::max(S.nslices, nsl);
        size_t pos = 0;
        int64_t done = 0;
        for (int sl = 0; sl < nsl; ++sl) {
          GGroup G = G0;
          G.pad0 = filter;
          if (sl > 0) { G.c_base = B2D_BASE_AUX; G.c = G0.c + (int64_t)(sl - 1) * P.Wp; }
          G.seg_begin = (int)cur.step2.segs.size();
          const int64_t target = iters * (sl + 1) / nsl;
          while (pos < pending[g].size() && (done < target || sl == nsl - 1)) {
            done += (pending[g][pos].k + 15) / 16;
            cur.step2.segs.push_back(pending[g][pos++]);
          }
          G.seg_end = (int)cur.step2.segs.size();
          if (G.seg_end > G.seg_begin) sliced.push_back(G);
        }
      };
      if (ns_narrow > 0 && has_big) { emit_family(ns, 1); emit_family(ns_narrow, 2); }
      else emit_family(ns_narrow > 0 ? ns_narrow : ns, 0);
    }
    cur.step2.groups.swap(sliced);
    make_tiles(cur.step1, forced_class);
    make_tiles(cur.step2, forced_class);
    S.work_max = std::max(S.work_max, cur.work);
    S.n_step1 += (int64_t)cur.step1.groups.size();
    S.n_step2 += (int64_t)cur.step2.segs.size();
    for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) S.n_tiles += (int64_t)(cur.step1.tiles[c].size() + cur.step2.tiles[c].size());
    S.chunks.push_back(std::move(cur));
  };
  open_chunk();

  std::vector<std::vector<int>> psi_row(L.nq);   // lQ' -> allowed rQ'
  for (int l = 0; l < L.nq; ++l)
    for (int r = 0; r < R.nq; ++r)
      if (P.allowed(l, r)) psi_row[l].push_back(r);

  for (const Term& t : terms) {
    View lop{&L, &L.ops[t.lop], t.lt}, rop{&R, &R.ops[t.rop], t.rt};
    // size this term's T blocks first so a chunk never splits a term
    std::vector<std::vector<int>> rcol(R.nq);   // rQ' -> rQ with rop.allowed(rQ, rQ')
    for (int rq = 0; rq < R.nq; ++rq)
      for (int rqp = 0; rqp < R.nq; ++rqp)
        if (rop.allowed(rq, rqp)) rcol[rqp].push_back(rq);
    struct TBlock { int lQ, lQp, rQp; };
    std::vector<TBlock> tb;
    int64_t need = 0;
    for (int lQ = 0; lQ < L.nq; ++lQ)
      for (int lQp = 0; lQp < L.nq; ++lQp) {
        if (!lop.allowed(lQ, lQp)) continue;
        for (int rQp : psi_row[lQp]) {
          double f1 = 2.0 * L.dims[lQ] * L.dims[lQp] * R.dims[rQp];
          S.flops_alg += f1;
          bool used = false;
          for (int rQ : rcol[rQp])
            if (P.allowed(lQ, rQ)) { used = true; S.flops_alg += 2.0 * L.dims[lQ] * R.dims[rQp] * R.dims[rQ]; }
          if (!used) continue;
          tb.push_back(TBlock{lQ, lQp, rQp});
          need += align_up((int64_t)L.dims[lQ] * pad_ld(R.dims[rQp]), BLK_ALIGN);
        }
      }
    if (cur.nterms > 0 && cur.work + need > work_budget) { close_chunk(); open_chunk(); }
    cur.nterms++;
    const double flops_before = cur.step1.flops + cur.step2.flops;
    // a factorised left operator writes only the row pieces it has factors for: the others are written as ZERO by groups without
    // segments (the kernel stores its cleared accumulators) - cheaper than clearing the whole workspace per chunk
    std::vector<SubBlock> lsubs;
    for (const TBlock& b : tb) {
      const int dl = L.dims[b.lQ], drp = R.dims[b.rQp];
      const int ldt = pad_ld(drp);
      const int64_t toff = cur.work;
      cur.work += align_up((int64_t)dl * ldt, BLK_ALIGN);
      // step 1:  T = s A_L^(c)[lQ,lQ'] src[lQ',rQ']                                  (operatorfunctions.C:512-516)
      // A materialised operator block is ONE product; a factorised one is a product per (row piece, column piece) factor: the row
      // pieces are the output groups, the column pieces the K segments, each reading its rows of the source block
      const int pb = P.blk[(size_t)b.lQp * P.nr + b.rQp];
      lsubs.clear();
      lop.for_each_sub(b.lQ, b.lQp, [&](const SubBlock& sb) { lsubs.push_back(sb); });
      std::stable_sort(lsubs.begin(), lsubs.end(), [](const SubBlock& x, const SubBlock& y) { return x.r0 < y.r0; });
      int covered = 0;   // rows [0, covered) of this T block are written by the groups emitted so far
      auto zero_rows = [&](int r_begin, int r_end) {
        if (r_end <= r_begin) return;
        GGroup gz;
        std::memset(&gz, 0, sizeof(gz));
        gz.c = toff + (int64_t)r_begin * ldt; gz.c_base = B2D_BASE_WORK; gz.ldc = ldt; gz.m = r_end - r_begin; gz.n = drp; gz.accumulate = 0;
        gz.seg_begin = gz.seg_end = (int)cur.step1.segs.size();
        cur.step1.groups.push_back(gz);
      };
      for (size_t k0 = 0; k0 < lsubs.size();) {
        size_t k1 = k0;
        while (k1 < lsubs.size() && lsubs[k1].r0 == lsubs[k0].r0) ++k1;
        zero_rows(covered, lsubs[k0].r0);
        covered = lsubs[k0].r0 + lsubs[k0].m;
        GGroup g1;
        std::memset(&g1, 0, sizeof(g1));
        g1.c = toff + (int64_t)lsubs[k0].r0 * ldt; g1.c_base = B2D_BASE_WORK; g1.ldc = ldt; g1.m = lsubs[k0].m; g1.n = drp; g1.accumulate = 0;
        g1.seg_begin = (int)cur.step1.segs.size();
        for (size_t k = k0; k < k1; ++k) {
          const SubBlock& sb = lsubs[k];
          GSeg s1;
          std::memset(&s1, 0, sizeof(s1));
          s1.a = (int64_t)(intptr_t)sb.a;                                           // absolute byte address
          s1.a_base = B2D_BASE_ABS;
          s1.a_trans = sb.t ? 1 : 0;                 // stored block is k x m
          s1.lda = sb.lda;
          s1.b = P.dev_off[pb] + (int64_t)sb.c0 * P.ld[pb]; s1.b_base = B2D_BASE_SRC; s1.b_kmajor = 0; s1.ldb = P.ld[pb];
          s1.k = sb.n;
          s1.alpha = sb.alpha;                       // 1 for a materialised operator; the left scaling (:514) is folded into the step-2 factor below
          cur.step1.segs.push_back(s1);
          cur.step1.flops += 2.0 * sb.m * sb.n * drp;
        }
        g1.seg_end = (int)cur.step1.segs.size();
        cur.step1.groups.push_back(g1);
        k0 = k1;
      }
      zero_rows(covered, dl);
      const double left_scaling = lop.scaling(am, b.lQ, b.lQp);
      // step 2:  dst[lQ,rQ] += F T (A_R^(c)[rQ,rQ'])^T                                 (operatorfunctions.C:517-531)
      for (int rQ : rcol[b.rQp]) {
        if (!P.allowed(b.lQ, rQ)) continue;
        double F = t.scale * am.ninej(L.quantum(b.lQp)[1], R.quantum(b.rQp)[1], S_psi, lop.spin(), rop.spin(), opq_spin,
                                      L.quantum(b.lQ)[1], R.quantum(rQ)[1], S_psi);            // :522-524
        if (rop.fermion() && (L.quantum(b.lQp)[0] & 1)) F = -F;                                 // :528
        F *= rop.scaling(am, rQ, b.rQp);                                                        // :529
        F *= left_scaling;                                                                      // :514
        const int db = P.blk[(size_t)b.lQ * P.nr + rQ];
        rop.for_each_sub(rQ, b.rQp, [&](const SubBlock& sb) {
          // the rows [r0, r0 + m) of A_R[rQ,rQ'] are the COLUMNS [r0, r0 + m) of the destination block, its columns [c0, c0 + n) the
          // columns of T this factor contracts
          auto key = std::make_pair(db, (int)sb.r0);
          auto it = group_of.find(key);
          int g;
          if (it == group_of.end()) {
            g = (int)cur.step2.groups.size();
            group_of.emplace(key, g);
            GGroup G;
            std::memset(&G, 0, sizeof(G));
            G.c = P.dev_off[db] + sb.r0; G.c_base = B2D_BASE_DST; G.ldc = P.ld[db]; G.m = dl; G.n = sb.m; G.accumulate = 1;
            cur.step2.groups.push_back(G);
            pending.emplace_back();
          } else g = it->second;
          cur.step2.flops += 2.0 * dl * sb.n * sb.m;
          const double Fa = F * sb.alpha;
          if (Fa == 0.0) return;                     // a vanishing recoupling coefficient contributes nothing
          GSeg s2;
          std::memset(&s2, 0, sizeof(s2));
          s2.a = toff + sb.c0; s2.a_base = B2D_BASE_WORK; s2.a_trans = 0; s2.lda = ldt;
          s2.b = (int64_t)(intptr_t)sb.a;
          s2.b_base = B2D_BASE_ABS;
          s2.b_kmajor = sb.t ? 0 : 1;                // plain: stored block is n x k; transposed: stored k x n
          s2.ldb = sb.lda;
          s2.k = sb.n;
          s2.alpha = Fa;
          pending[g].push_back(s2);
        });
      }
    }
    S.term_flops.push_back(cur.step1.flops + cur.step2.flops - flops_before);
  }
  close_chunk();
  for (const Chunk& c : S.chunks) S.flops_exec += c.step1.flops + c.step2.flops;
  return S;
}

// ---------------------------------------------------------------------------------------------------------------
// one-operator products (operatorfunctions::TensorMultiply(ablock, a, cblock, c, v, dQ, scale), operatorfunctions.C:331-404)
// and wavefunction outer products (MultiplyProduct, operatorfunctions.C:630-650) as single-step grouped contractions
// ---------------------------------------------------------------------------------------------------------------
// dst (layout Pd, vector at dst_base + dst_off) += scale (a x 1) src   [side 0]   or   scale (1 x a) src   [side 1];
// src has layout Ps and lives at src_base + src_off.  One group per destination block, one segment per contributing
// source block.  Returns the dgemm flops.
inline double add_one_op_groups(const Side& L, const Side& R, const PsiLayout& Ps, const PsiLayout& Pd, int side, const OpRec& op, bool transposed,
                                double scale, uint8_t src_base, int64_t src_off, uint8_t dst_base, int64_t dst_off, AngMom& am, GemmBatch& batch) {
  const Side& S = side == 0 ? L : R;
  View a{&S, &op, transposed};
  const int Sc = Ps.dq[1], Sv = Pd.dq[1];
  double flops = 0.0;
  struct Item { SubBlock sb; double fac; int ps; };
  std::vector<Item> items;
  for (int p = 0; p < Pd.nblocks(); ++p) {
    const int lQ = Pd.bl[p], rQ = Pd.br[p];
    items.clear();
    if (side == 0) {
      for (int lQp = 0; lQp < L.nq; ++lQp) {                                              // :347-366
        if (!a.allowed(lQ, lQp) || !Ps.allowed(lQp, rQ)) continue;
        double fac = scale * am.ninej(L.quantum(lQp)[1], R.quantum(rQ)[1], Sc, a.spin(), 0, a.spin(), L.quantum(lQ)[1], R.quantum(rQ)[1], Sv);
        fac *= a.scaling(am, lQ, lQp);
        flops += 2.0 * L.dims[lQ] * L.dims[lQp] * R.dims[rQ];
        if (fac == 0.0) continue;
        const int ps = Ps.blk[(size_t)lQp * Ps.nr + rQ];
        a.for_each_sub(lQ, lQp, [&](const SubBlock& sb) { items.push_back(Item{sb, fac, ps}); });
      }
    } else {
      for (int rQp = 0; rQp < R.nq; ++rQp) {                                              // :378-397
        if (!a.allowed(rQ, rQp) || !Ps.allowed(lQ, rQp)) continue;
        double fac = scale * am.ninej(L.quantum(lQ)[1], R.quantum(rQp)[1], Sc, 0, a.spin(), a.spin(), L.quantum(lQ)[1], R.quantum(rQ)[1], Sv);
        fac *= a.scaling(am, rQ, rQp);
        if (a.fermion() && (L.quantum(lQ)[0] & 1)) fac = -fac;
        flops += 2.0 * L.dims[lQ] * R.dims[rQp] * R.dims[rQ];
        if (fac == 0.0) continue;
        const int ps = Ps.blk[(size_t)lQ * Ps.nr + rQp];
        a.for_each_sub(rQ, rQp, [&](const SubBlock& sb) { items.push_back(Item{sb, fac, ps}); });
      }
    }
    // one group per output piece (rows r0.. of the destination block for a left operator, columns r0.. for a right operator);
    // a materialised operator has the single piece r0 = 0
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.sb.r0 < y.sb.r0; });
    for (size_t k0 = 0; k0 < items.size();) {
      size_t k1 = k0;
      while (k1 < items.size() && items[k1].sb.r0 == items[k0].sb.r0) ++k1;
      const SubBlock& first = items[k0].sb;
      GGroup G;
      std::memset(&G, 0, sizeof(G));
      G.c_base = dst_base; G.ldc = Pd.ld[p]; G.accumulate = 1;
      if (side == 0) { G.c = dst_off + Pd.dev_off[p] + (int64_t)first.r0 * Pd.ld[p]; G.m = first.m; G.n = Pd.cols[p]; }
      else { G.c = dst_off + Pd.dev_off[p] + first.r0; G.m = Pd.rows[p]; G.n = first.m; }
      G.seg_begin = (int)batch.segs.size();
      for (size_t k = k0; k < k1; ++k) {
        const SubBlock& sb = items[k].sb;
        const int ps = items[k].ps;
        const double f = items[k].fac * sb.alpha;
        if (f == 0.0) continue;
        GSeg sg;
        std::memset(&sg, 0, sizeof(sg));
        if (side == 0) {
          sg.a = (int64_t)(intptr_t)sb.a; sg.a_base = B2D_BASE_ABS; sg.a_trans = sb.t ? 1 : 0; sg.lda = sb.lda;
          sg.b = src_off + Ps.dev_off[ps] + (int64_t)sb.c0 * Ps.ld[ps]; sg.b_base = src_base; sg.b_kmajor = 0; sg.ldb = Ps.ld[ps];
        } else {
          sg.a = src_off + Ps.dev_off[ps] + sb.c0; sg.a_base = src_base; sg.a_trans = 0; sg.lda = Ps.ld[ps];
          sg.b = (int64_t)(intptr_t)sb.a; sg.b_base = B2D_BASE_ABS; sg.b_kmajor = sb.t ? 0 : 1; sg.ldb = sb.lda;
        }
        sg.k = sb.n; sg.alpha = f;
        batch.segs.push_back(sg);
      }
      G.seg_end = (int)batch.segs.size();
      if (G.seg_end > G.seg_begin) batch.groups.push_back(G);
      k0 = k1;
    }
  }
  return flops;
}

// per left sector lQ: segments of  rho[lQ] += alpha * w[lQ,rQ] w[lQ,rQ]^T  for the wavefunction (layout P) at base + off
inline void add_density_segments(const PsiLayout& P, uint8_t base, int64_t off, double alpha, std::vector<std::vector<GSeg>>& per_sector) {
  for (int p = 0; p < P.nblocks(); ++p) {
    GSeg s;
    std::memset(&s, 0, sizeof(s));
    s.a = s.b = off + P.dev_off[p];
    s.a_base = s.b_base = base;
    s.a_trans = 0; s.b_kmajor = 1;
    s.lda = s.ldb = P.ld[p];
    s.k = P.cols[p];
    s.alpha = alpha;
    per_sector[P.bl[p]].push_back(s);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// diag(H)
// ---------------------------------------------------------------------------------------------------------------
// `regions` receives one BlockDesc per (psi block, row piece, column piece) - the whole block when neither child is a product of
// factorised operators -, `region_begin` the task range of each.  dev_off / ld address the region inside the padded wavefunction.
inline void build_diag_tasks(const Side& L, const Side& R, const PsiLayout& P, const std::vector<Term>& terms, double core_energy,
                             bool hubbard, AngMom& am, std::vector<DiagTask>& tasks, std::vector<int>& region_begin, std::vector<BlockDesc>& regions) {
  // per psi block: list of (f, diagA, diagB).  The term list of diagonalH mirrors multiplyH's with the *_d functors
  // (opxop.C:295-365): same operator pairs, scale 1 for c x ccd, `factor` (no parity) for cc x dd; H and e_core by trace.
  (void)hubbard;
  const int S_psi = P.dq[1];
  auto pieces_of = [](const Side& s, int q) {
    if ((int)s.pieces.size() == s.nq && !s.pieces[q].empty()) return s.pieces[q];
    return std::vector<std::pair<int, int>>(1, std::make_pair(0, s.dims[q]));
  };
  regions.clear();
  std::vector<int> first_region(P.nblocks() + 1, 0);
  std::vector<std::pair<int, int>> rp_of, cp_of;   // per region: row piece, column piece (offset, size)
  for (int p = 0; p < P.nblocks(); ++p) {
    first_region[p] = (int)regions.size();
    for (const auto& rp : pieces_of(L, P.bl[p]))
      for (const auto& cp : pieces_of(R, P.br[p])) {
        BlockDesc d;
        d.ref_off = 0; d.dev_off = P.dev_off[p] + (int64_t)rp.first * P.ld[p] + cp.first; d.rows = rp.second; d.cols = cp.second; d.ld = P.ld[p]; d.pad = 0;
        regions.push_back(d);
        rp_of.push_back(rp); cp_of.push_back(cp);
      }
  }
  first_region[P.nblocks()] = (int)regions.size();
  std::vector<std::vector<DiagTask>> per(regions.size());
  struct DiagRef { int64_t addr; int stride; double alpha; };
  // diagonal of view block (q,q) restricted to the piece [o, o + n): the factors whose sub-block lies on the block diagonal and covers it
  auto diag_refs = [&](const View& v, int q, const std::pair<int, int>& piece, std::vector<DiagRef>& out) {
    out.clear();
    v.for_each_sub(q, q, [&](const SubBlock& sb) {
      if (sb.r0 != sb.c0 || sb.m != sb.n) return;
      if (piece.first < sb.r0 || piece.first + piece.second > sb.r0 + sb.m) return;
      out.push_back(DiagRef{(int64_t)(intptr_t)sb.a + 8 * (int64_t)(piece.first - sb.r0) * (sb.lda + 1), sb.lda + 1, sb.alpha});
    });
  };
  std::vector<DiagRef> da, db;
  auto add = [&](const Term& t, double scale, bool left_only, bool right_only) {
    View a{&L, &L.ops[t.lop], t.lt}, b{&R, &R.ops[t.rop], t.rt};
    for (int p = 0; p < P.nblocks(); ++p) {
      int l = P.bl[p], r = P.br[p];
      if (!left_only && !right_only && !(a.allowed(l, l) && b.allowed(r, r))) continue;
      if (left_only && !a.allowed(l, l)) continue;
      if (right_only && !b.allowed(r, r)) continue;
      int sa = right_only ? 0 : a.spin(), sb = left_only ? 0 : b.spin();
      double f = scale * am.ninej(L.quantum(l)[1], R.quantum(r)[1], S_psi, sa, sb, 0, L.quantum(l)[1], R.quantum(r)[1], S_psi);
      if (!left_only && b.fermion() && (L.quantum(l)[0] & 1)) f = -f;
      for (int g = first_region[p]; g < first_region[p + 1]; ++g) {
        if (!right_only) diag_refs(a, l, rp_of[g], da); else da.assign(1, DiagRef{0, 0, 1.0});
        if (!left_only) diag_refs(b, r, cp_of[g], db); else db.assign(1, DiagRef{0, 0, 1.0});
        for (const DiagRef& x : da)
          for (const DiagRef& y : db) {
            DiagTask d;
            std::memset(&d, 0, sizeof(d));
            d.f = f * x.alpha * y.alpha;
            d.a = x.addr; d.sa = x.stride;
            d.b = y.addr; d.sb = y.stride;
            if (d.f != 0.0) per[g].push_back(d);
          }
      }
    }
  };
  // `terms` is (this rank's share of) enumerate_terms() output; under a term partition every rank adds its own
  // contributions and the caller all-reduces the result
  bool have_core = false;
  for (const Term& t : terms) {
    if (t.kind == TERM_CORE) { have_core = true; continue; }    // handled as a constant below
    if (t.kind == TERM_HAM_LEFT) { add(t, 1.0, true, false); continue; }    // TensorTrace(H_L)   spinblock.C:864
    if (t.kind == TERM_HAM_RIGHT) { add(t, 1.0, false, true); continue; }   // TensorTrace(H_R)   :867
    const OpRec& lo = L.ops[t.lop];
    const OpRec& ro = R.ops[t.rop];
    double scale = 1.0;
    int ty = lo.optype == OP_CRE_CRE || ro.optype == OP_CRE_CRE ? OP_CRE_CRE : 0;
    if (ty == OP_CRE_CRE) {
      const OpRec& cc = lo.optype == OP_CRE_CRE ? lo : ro;
      scale = cc.orbs[0] == cc.orbs[1] ? 1.0 : 2.0;      // opxop.C:314-339 (no commute parity in the diagonal form)
    }
    add(t, scale, false, false);
  }
  tasks.clear();
  region_begin.assign(regions.size() + 1, 0);
  for (size_t g = 0; g < regions.size(); ++g) {
    region_begin[g] = (int)tasks.size();
    if (have_core && core_energy != 0.0) {
      DiagTask d;
      std::memset(&d, 0, sizeof(d));
      d.f = core_energy;
      tasks.push_back(d);
    }
    tasks.insert(tasks.end(), per[g].begin(), per[g].end());
  }
  region_begin[regions.size()] = (int)tasks.size();
}

// ---------------------------------------------------------------------------------------------------------------
// truncation: sort_weights rotationmat.C:313-346 + assign_matrix_by_dm :149-256 (keptqstates = 0)
// ---------------------------------------------------------------------------------------------------------------
inline double select_states(const std::vector<std::vector<double>>& evals, int keep, std::vector<std::vector<int>>& kept) {
  struct E { double w; int64_t ord; int q, s; };
  std::vector<E> all;
  for (size_t q = 0; q < evals.size(); ++q)
    for (size_t s = 0; s < evals[q].size(); ++s) all.push_back(E{evals[q][s], (int64_t)all.size(), (int)q, (int)s});
  // multimap reverse iteration: descending key, equal keys in reverse insertion order
  std::sort(all.begin(), all.end(), [](const E& a, const E& b) { return a.w != b.w ? a.w > b.w : a.ord > b.ord; });
  size_t total = std::min<size_t>(all.size(), (size_t)std::max(keep, 0));
  kept.assign(evals.size(), std::vector<int>());
  double norm_kept = 0.0, norm = 0.0;
  for (size_t i = 0; i < total; ++i)
    if (all[i].w > 1e-13) { kept[all[i].q].push_back(all[i].s); norm_kept += all[i].w; }
  for (size_t q = 0; q < evals.size(); ++q) {
    double s = 0.0;
    for (double w : evals[q]) s += w;
    norm += s;
  }
  return norm - norm_kept;
}

}  // namespace b2d
