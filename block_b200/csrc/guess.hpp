// Host planner of the guess-wavefunction transform of a two-dot step.
//
// Replaces GuessWave::transform_previous_wavefunction, two-dot branch (guess_wavefunction.C:524-636):
//   stage 1  TransformLeftBlock (:17-31)        T1[a, b]   = L[olda]^T . old[olda, b]            grouped GEMM
//   stage 2  onedot_shufflesysdot (:434-485, :200-256)   T2[lq, c][r0.., :] += f . T1[a, b][:, c0..]   HBM-bound scatter
//            f = getCommuteParity(E, dot, E.dot) . sixj(A, B, AB, C, J, CB) sqrt((AB+1)(CB+1)) (-1)^((A+B+J+C)/2) . spatial_sixj
//   stage 3  TransformRightBlock (:33-50)       trial[lq, tb] = T2[lq, c] . R[tb]^T              grouped GEMM
// One-dot branch (GuessWave::onedot_transform_wavefunction, guess_wavefunction.C:832-936), mode 1 / 2 of b2d_guess_desc:
//   stage 0  T0[olda, tc] = old[olda, c] . R[tc]^T   (:870-890)      stage 1  rows rotated as above (:892-911)
//   mode 1 (dot on the system side): stage 2 moves the dot from the environment to the system and writes the trial vector (:916-935);
//   mode 2 (dot on the environment side): stage 1 writes the trial vector, there is no shuffle.
// Everything here is integer / scalar work on the StateInfo tables of the reference; the plan is executed on the device by
// b2d_guess_transform (ctx.cpp) with the grouped contraction kernel and kron_scatter_kernel.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/block_b200.h"
#include "kernels.h"
#include "plan.hpp"

namespace b2d {

struct GuessPlan {
  bool valid = false;
  int mode = 0;
  int dq[3] = {0, 0, 0};
  // input image: old wavefunction blocks, left and right rotation matrices, packed one after the other in ONE device buffer
  std::vector<BlockDesc> in_old, in_lrot, in_rrot;    // ref_off: offset in the caller's flat array, dev_off: offset in the image
  int64_t old_size = 0, lrot_size = 0, rrot_size = 0; // doubles in the caller's arrays
  int64_t image_size = 0;                              // doubles on the device
  // grouped contractions in execution order: gemm_a, gemm_b (reads what gemm_a wrote), the shuffle rounds, gemm_c
  //   mode 0: a = rows -> S' (T1),  b = -,              c = columns -> un-truncated right basis (trial)
  //   mode 1: a = columns (T0),     b = rows (T1),      c = -        (the shuffle writes the trial vector)
  //   mode 2: a = columns (T0),     b = rows (trial),   c = -        (no shuffle)
  GemmBatch gemm_a, gemm_b, gemm_c;
  std::vector<std::vector<KronTask>> rounds;           // a / dst hold OFFSETS (doubles) until execution: a into WORK, dst into WORK or - pad = 1 -
                                                       // into the trial vector, pad bit 1: a is an offset into the image; tasks of one round never overlap (a destination's j-th
                                                       // source goes to round j)
  int64_t t1_size = 0, t2_off = 0, work_size = 0;      // WORK: [0, t1_size) what the shuffle reads (and T0 before it), [t2_off, work_size) T2
  Side left, right;                                    // sector tables of the big block's children (trial layout)
  PsiLayout trial;
  double flops = 0.0;
  int64_t shuffle_bytes = 0;                           // algorithmic bytes of the shuffle: 8 x (read + read-modify-write) elements
};

inline void guess_check(bool ok, const char* what) {
  if (!ok) throw std::runtime_error(std::string("b2d_guess_plan: ") + what);
}

// mode 3: GuessWave::transpose_previous_wavefunction (guess_wavefunction.C:55-84), two-dot to two-dot, first block iteration of a sweep:
//   trial(i, j) = getCommuteParity(q_oldcol[i], q_oldleft[j], dq) . old(j, i)^T      - one scatter task per block, no contraction
inline GuessPlan plan_guess_transpose(const b2d_guess_desc& d, AngMom& am) {
  GuessPlan P;
  P.mode = 3;
  std::memcpy(P.dq, d.dq, sizeof(P.dq));
  const b2d_stateinfo &left = d.left, &right = d.right, &oldleft = d.oldleft, &oldcol = d.oldcol;
  guess_check(left.nq > 0 && right.nq > 0 && oldleft.nq == right.nq && oldcol.nq == left.nq && d.old_allowed,
              "transpose: the previous wavefunction must live on (right sectors) x (left sectors)");
  for (int i = 0; i < left.nq; ++i) guess_check(left.dims[i] == oldcol.dims[i], "transpose: column sectors of the previous wavefunction are not the new left sectors");
  for (int j = 0; j < right.nq; ++j) guess_check(right.dims[j] == oldleft.dims[j], "transpose: row sectors of the previous wavefunction are not the new right sectors");
  std::vector<int64_t> old_off((size_t)oldleft.nq * oldcol.nq, -1);
  int64_t dev = 0, ref = 0;
  for (int a = 0; a < oldleft.nq; ++a)
    for (int b = 0; b < oldcol.nq; ++b)
      if (d.old_allowed[(size_t)a * oldcol.nq + b]) {
        BlockDesc bd; bd.ref_off = ref; bd.dev_off = dev; bd.rows = oldleft.dims[a]; bd.cols = oldcol.dims[b]; bd.ld = pad_ld(bd.cols); bd.pad = 0;
        old_off[(size_t)a * oldcol.nq + b] = dev;
        ref += (int64_t)bd.rows * bd.cols; dev += align_up((int64_t)bd.rows * bd.ld, BLK_ALIGN);
        P.in_old.push_back(bd);
      }
  P.old_size = ref;
  P.image_size = dev;
  P.left.nq = left.nq; P.left.q.assign(left.q, left.q + 3 * left.nq); P.left.dims.assign(left.dims, left.dims + left.nq);
  P.right.nq = right.nq; P.right.q.assign(right.q, right.q + 3 * right.nq); P.right.dims.assign(right.dims, right.dims + right.nq);
  P.trial.build(P.left, P.right, d.dq);
  P.rounds.resize(1);
  for (int p = 0; p < P.trial.nblocks(); ++p) {
    const int i = P.trial.bl[p], j = P.trial.br[p];
    const int64_t src = old_off[(size_t)j * oldcol.nq + i];
    guess_check(src >= 0, "transpose: a block of the trial vector has no counterpart in the previous wavefunction");   // the reference asserts it (:72)
    KronTask t; std::memset(&t, 0, sizeof(t));
    t.a = src; t.b = 0; t.dst = P.trial.dev_off[p];
    t.coef = am.commute_parity(&oldcol.q[3 * i], &oldleft.q[3 * j], d.dq);
    t.a_rows = left.dims[i]; t.a_cols = right.dims[j]; t.lda = pad_ld(oldcol.dims[i]); t.a_t = 1;     // stored a_cols x a_rows
    t.b_rows = 1; t.b_cols = 1; t.ldb = 1; t.b_t = 0;
    t.row0 = 0; t.col0 = 0; t.ldd = P.trial.ld[p];
    t.pad = 3;                                                                                          // source: the image, destination: the trial vector
    P.rounds[0].push_back(t);
    P.shuffle_bytes += 8ll * 3 * t.a_rows * t.a_cols;
  }
  P.valid = true;
  return P;
}

// mode 4: first block iteration of a ONE-DOT sweep, GuessWave::onedot_transpose_wavefunction (guess_wavefunction.C:140-198 with :402-432 and
// :200-256): [S.d][E] -> [E.d][S].  Every un-collected row piece u = (S sector a, dot sector b, quantum AB) of a block of the previous
// wavefunction is transposed with the sign getCommuteParity(S, d, AB) . getCommuteParity(AB, E, dq) and re-coupled into the row pieces
// v = (E sector c, dot sector b, quantum EB) of the new left block with sixj(E, d, EB, S, J, AB) sqrt((EB+1)(AB+1)) (-1)^((E+d+J+S)/2):
// scatter tasks only (source: the input image, stored transposed; destination: the trial vector).
inline GuessPlan plan_guess_onedot_transpose(const b2d_guess_desc& d, AngMom& am) {
  GuessPlan P;
  P.mode = 4;
  std::memcpy(P.dq, d.dq, sizeof(P.dq));
  const b2d_stateinfo &left = d.left, &sys = d.sys, &dot = d.dot, &right = d.right, &oldleft = d.oldleft, &oldcol = d.oldcol;
  guess_check(left.nq > 0 && right.nq > 0 && sys.nq > 0 && dot.nq > 0 && oldleft.nq > 0 && oldcol.nq == sys.nq && d.old_allowed, "one-dot transpose: bad StateInfos");
  guess_check(left.nunc > 0 && left.unc_q && left.unc_dims && left.unc_left && left.unc_right && left.old_to_new_begin && left.old_to_new, "left needs its un-collected tables");
  guess_check(oldleft.nunc > 0 && oldleft.unc_q && oldleft.unc_dims && oldleft.unc_left && oldleft.unc_right && oldleft.old_to_new_begin && oldleft.old_to_new,
              "oldleft needs its un-collected tables");
  for (int b = 0; b < dot.nq; ++b) guess_check(dot.dims[b] == 1, "dot sectors must hold one state (spin-adapted single site)");
  for (int c = 0; c < sys.nq; ++c) guess_check(sys.dims[c] == oldcol.dims[c], "one-dot transpose: the new system block is not the previous wavefunction's column block");
  std::vector<int64_t> old_off((size_t)oldleft.nq * oldcol.nq, -1);
  int64_t dev = 0, ref = 0;
  for (int a = 0; a < oldleft.nq; ++a)
    for (int b = 0; b < oldcol.nq; ++b)
      if (d.old_allowed[(size_t)a * oldcol.nq + b]) {
        BlockDesc bd; bd.ref_off = ref; bd.dev_off = dev; bd.rows = oldleft.dims[a]; bd.cols = oldcol.dims[b]; bd.ld = pad_ld(bd.cols); bd.pad = 0;
        old_off[(size_t)a * oldcol.nq + b] = dev;
        ref += (int64_t)bd.rows * bd.cols; dev += align_up((int64_t)bd.rows * bd.ld, BLK_ALIGN);
        P.in_old.push_back(bd);
      }
  P.old_size = ref;
  P.image_size = dev;
  P.left.nq = left.nq; P.left.q.assign(left.q, left.q + 3 * left.nq); P.left.dims.assign(left.dims, left.dims + left.nq);
  P.right.nq = right.nq; P.right.q.assign(right.q, right.q + 3 * right.nq); P.right.dims.assign(right.dims, right.dims + right.nq);
  P.trial.build(P.left, P.right, d.dq);
  std::vector<int> ou_parent(oldleft.nunc, -1), ou_first(oldleft.nunc, 0), lu_parent(left.nunc, -1), lu_first(left.nunc, 0);
  for (int q = 0; q < oldleft.nq; ++q) {
    int first = 0;
    for (int k = oldleft.old_to_new_begin[q]; k < oldleft.old_to_new_begin[q + 1]; ++k) {
      const int u = oldleft.old_to_new[k];
      guess_check(u >= 0 && u < oldleft.nunc, "oldleft.oldToNewState out of range");
      ou_parent[u] = q; ou_first[u] = first; first += oldleft.unc_dims[u];
    }
    guess_check(first == oldleft.dims[q], "oldleft: un-collected pieces do not add up to the collected sector");
  }
  for (int q = 0; q < left.nq; ++q) {
    int first = 0;
    for (int k = left.old_to_new_begin[q]; k < left.old_to_new_begin[q + 1]; ++k) {
      const int v = left.old_to_new[k];
      guess_check(v >= 0 && v < left.nunc, "left.oldToNewState out of range");
      lu_parent[v] = q; lu_first[v] = first; first += left.unc_dims[v];
    }
    guess_check(first == left.dims[q], "left: un-collected pieces do not add up to the collected sector");
  }
  std::vector<int> hits((size_t)left.nunc * right.nq, 0);
  const int J = d.dq[1];
  for (int u = 0; u < oldleft.nunc; ++u) {
    const int a = oldleft.unc_left[u], b = oldleft.unc_right[u], oq = ou_parent[u];
    guess_check(a >= 0 && a < right.nq && b >= 0 && b < dot.nq, "oldleft un-collected maps out of range");
    if (oq < 0) continue;
    guess_check(oldleft.unc_dims[u] == right.dims[a] * dot.dims[b], "one-dot transpose: the new right block is not the previous wavefunction's system block");
    for (int c = 0; c < oldcol.nq; ++c) {
      const int64_t src = old_off[(size_t)oq * oldcol.nq + c];
      if (src < 0 || !qn_allow(d.dq, &oldleft.unc_q[3 * u], &oldcol.q[3 * c])) continue;
      const double parity = am.commute_parity(&right.q[3 * a], &dot.q[3 * b], &oldleft.unc_q[3 * u]) * am.commute_parity(&oldleft.unc_q[3 * u], &oldcol.q[3 * c], d.dq);
      for (int v = 0; v < left.nunc; ++v) {
        if (left.unc_left[v] != c || left.unc_right[v] != b || lu_parent[v] < 0) continue;
        if (!qn_allow(d.dq, &left.unc_q[3 * v], &right.q[3 * a])) continue;
        const int A = sys.q[3 * c + 1], B = dot.q[3 * b + 1], AB = left.unc_q[3 * v + 1], C = right.q[3 * a + 1], CB = oldleft.unc_q[3 * u + 1];
        double f = parity * am.six_j(A, B, AB, C, J, CB) * std::sqrt((AB + 1.0) * (CB + 1.0)) * ((((A + B + J + C) / 2) & 1) ? -1.0 : 1.0);
        const int Al = sys.q[3 * c + 2], Bl = dot.q[3 * b + 2], ABl = left.unc_q[3 * v + 2], Cl = right.q[3 * a + 2], CBl = oldleft.unc_q[3 * u + 2];
        if (ABl != (Al ^ Bl) || CBl != (Bl ^ Cl) || d.dq[2] != (ABl ^ Cl)) f = 0.0;
        if (f == 0.0) continue;
        const int p = P.trial.blk[(size_t)lu_parent[v] * right.nq + a];
        guess_check(p >= 0, "one-dot transpose: destination outside the target quantum number");
        KronTask t; std::memset(&t, 0, sizeof(t));
        t.a = src + (int64_t)ou_first[u] * pad_ld(oldcol.dims[c]); t.b = 0; t.dst = P.trial.dev_off[p]; t.coef = f;
        t.a_rows = sys.dims[c]; t.a_cols = right.dims[a]; t.lda = pad_ld(oldcol.dims[c]); t.a_t = 1;      // stored (S piece rows) x (E columns)
        t.b_rows = 1; t.b_cols = 1; t.ldb = 1; t.b_t = 0;
        t.row0 = lu_first[v]; t.col0 = 0; t.ldd = P.trial.ld[p];
        t.pad = 3;
        guess_check(left.unc_dims[v] == sys.dims[c] * dot.dims[b], "left: un-collected sector size is not the product of its factors");
        int& h = hits[(size_t)v * right.nq + a];
        if ((int)P.rounds.size() <= h) P.rounds.resize(h + 1);
        P.rounds[h].push_back(t);
        ++h;
        P.shuffle_bytes += 8ll * 3 * t.a_rows * t.a_cols;
      }
    }
  }
  P.valid = true;
  return P;
}

inline GuessPlan plan_guess_transform(const b2d_guess_desc& d, AngMom& am, int forced_class) {
  GuessPlan P;
  P.mode = d.mode;
  std::memcpy(P.dq, d.dq, sizeof(P.dq));
  guess_check(d.mode >= 0 && d.mode <= 4, "mode must be 0 (two-dot), 1 (one-dot, dot moved to the system), 2 (one-dot, rotation only), 3 (transpose) or 4 (one-dot transpose)");
  if (d.mode == 3) return plan_guess_transpose(d, am);
  if (d.mode == 4) return plan_guess_onedot_transpose(d, am);
  const bool shuffle = d.mode != 2, onedot = d.mode != 0;
  const b2d_stateinfo &sys = d.sys, &dot = d.dot, &left = d.left, &right = d.right, &oldleft = d.oldleft, &oldright = d.oldright, &oldcol = d.oldcol;
  const b2d_stateinfo& env = d.mode == 0 ? d.env : d.right;      // the environment sectors the shuffle keeps as columns
  const b2d_stateinfo& rows = d.mode == 2 ? d.left : d.sys;      // the renormalised row space the left rotation produces
  const b2d_stateinfo& wcols = onedot ? oldcol : oldright;       // column space of the previous wavefunction
  const b2d_stateinfo& rbasis = d.mode == 1 ? oldright : right;  // un-truncated basis the right rotation matrix is indexed by
  guess_check(left.nq > 0 && right.nq > 0 && oldleft.nq > 0 && rows.nq > 0 && wcols.nq > 0, "empty StateInfo");
  guess_check(rows.new_quanta_map != nullptr, "the renormalised row StateInfo needs newQuantaMap");
  guess_check(d.old_allowed && d.lrot_cols && d.rrot_cols, "null tables");
  if (shuffle) {
    guess_check(sys.nq > 0 && dot.nq > 0 && oldright.nq > 0 && env.nq > 0, "empty StateInfo");
    guess_check(left.nunc > 0 && left.unc_q && left.unc_dims && left.unc_left && left.unc_right && left.old_to_new_begin && left.old_to_new, "left needs its un-collected tables");
    guess_check(oldright.nunc > 0 && oldright.unc_q && oldright.unc_dims && oldright.unc_left && oldright.unc_right && oldright.old_to_new_begin && oldright.old_to_new,
                "oldright needs its un-collected tables");
    for (int b = 0; b < dot.nq; ++b) guess_check(dot.dims[b] == 1, "dot sectors must hold one state (spin-adapted single site)");
  }
  if (d.mode == 0) guess_check(d.env.nq > 0 && d.env.new_quanta_map, "env needs newQuantaMap");
  if (onedot) guess_check(oldcol.new_quanta_map != nullptr, "oldcol needs newQuantaMap");

  // ---- input image ------------------------------------------------------------------------------------------------------
  std::vector<int64_t> old_off((size_t)oldleft.nq * wcols.nq, -1), lrot_off(oldleft.nq, -1), rrot_off(rbasis.nq, -1);
  int64_t dev = 0, ref = 0;
  for (int i = 0; i < oldleft.nq; ++i)
    for (int j = 0; j < wcols.nq; ++j)
      if (d.old_allowed[(size_t)i * wcols.nq + j]) {
        BlockDesc bd; bd.ref_off = ref; bd.dev_off = dev; bd.rows = oldleft.dims[i]; bd.cols = wcols.dims[j]; bd.ld = pad_ld(bd.cols); bd.pad = 0;
        old_off[(size_t)i * wcols.nq + j] = dev;
        ref += (int64_t)bd.rows * bd.cols; dev += align_up((int64_t)bd.rows * bd.ld, BLK_ALIGN);
        P.in_old.push_back(bd);
      }
  P.old_size = ref; ref = 0;
  for (int q = 0; q < oldleft.nq; ++q)
    if (d.lrot_cols[q] > 0) {
      BlockDesc bd; bd.ref_off = ref; bd.dev_off = dev; bd.rows = oldleft.dims[q]; bd.cols = d.lrot_cols[q]; bd.ld = pad_ld(bd.cols); bd.pad = 0;
      lrot_off[q] = dev;
      ref += (int64_t)bd.rows * bd.cols; dev += align_up((int64_t)bd.rows * bd.ld, BLK_ALIGN);
      P.in_lrot.push_back(bd);
    }
  P.lrot_size = ref; ref = 0;
  for (int q = 0; q < rbasis.nq; ++q)
    if (d.rrot_cols[q] > 0) {
      BlockDesc bd; bd.ref_off = ref; bd.dev_off = dev; bd.rows = rbasis.dims[q]; bd.cols = d.rrot_cols[q]; bd.ld = pad_ld(bd.cols); bd.pad = 0;
      rrot_off[q] = dev;
      ref += (int64_t)bd.rows * bd.cols; dev += align_up((int64_t)bd.rows * bd.ld, BLK_ALIGN);
      P.in_rrot.push_back(bd);
    }
  P.rrot_size = ref;
  P.image_size = dev;

  // trial layout
  P.left.nq = left.nq; P.left.q.assign(left.q, left.q + 3 * left.nq); P.left.dims.assign(left.dims, left.dims + left.nq);
  P.right.nq = right.nq; P.right.q.assign(right.q, right.q + 3 * right.nq); P.right.dims.assign(right.dims, right.dims + right.nq);
  P.trial.build(P.left, P.right, d.dq);

  auto add_gemm = [&](GemmBatch& B, int m, int n, int k, int64_t a, uint8_t a_base, uint8_t a_trans, int lda, int64_t b, uint8_t b_base, uint8_t b_kmajor, int ldb,
                      int64_t c, uint8_t c_base, int ldc) {
    GSeg s; std::memset(&s, 0, sizeof(s));
    s.a = a; s.a_base = a_base; s.a_trans = a_trans; s.lda = lda; s.b = b; s.b_base = b_base; s.b_kmajor = b_kmajor; s.ldb = ldb; s.k = k; s.alpha = 1.0;
    GGroup g; std::memset(&g, 0, sizeof(g));
    g.c = c; g.c_base = c_base; g.ldc = ldc; g.m = m; g.n = n; g.accumulate = 0;
    g.seg_begin = (int)B.segs.size(); g.seg_end = g.seg_begin + 1;
    B.segs.push_back(s); B.groups.push_back(g);
    P.flops += 2.0 * m * n * k;
  };

  int64_t work = 0;
  // ---- one-dot stage 0: T0[olda, tc] = old[olda, c] R[tc]^T   (guess_wavefunction.C:870-890) -----------------------------------
  const int ncol1 = onedot ? rbasis.nq : oldright.nq;      // column sectors of what stage 1 reads
  std::vector<int64_t> src_off((size_t)oldleft.nq * ncol1, -1);
  std::vector<int> src_ld(ncol1, 0), col_dim(ncol1, 0);
  if (onedot) {
    std::vector<char> seen(rbasis.nq, 0);
    for (int c = 0; c < oldcol.nq; ++c) {
      const int tc = oldcol.new_quanta_map[c];
      guess_check(tc >= 0 && tc < rbasis.nq && !seen[tc], "oldcol.newQuantaMap is not an injective map into the right rotation's sectors");
      seen[tc] = 1;
      bool used = false;
      for (int a = 0; a < oldleft.nq; ++a) used = used || old_off[(size_t)a * oldcol.nq + c] >= 0;
      if (!used) continue;
      guess_check(rrot_off[tc] >= 0 && d.rrot_cols[tc] == oldcol.dims[c], "right rotation matrix does not match the column space of the previous wavefunction");
      src_ld[tc] = pad_ld(rbasis.dims[tc]); col_dim[tc] = rbasis.dims[tc];
      for (int a = 0; a < oldleft.nq; ++a) {
        if (old_off[(size_t)a * oldcol.nq + c] < 0) continue;
        const int m = oldleft.dims[a], n = rbasis.dims[tc], k = oldcol.dims[c];
        add_gemm(P.gemm_a, m, n, k, old_off[(size_t)a * oldcol.nq + c], B2D_BASE_AUX, 0, pad_ld(k), rrot_off[tc], B2D_BASE_AUX, 1, pad_ld(k), work, B2D_BASE_WORK, src_ld[tc]);
        src_off[(size_t)a * ncol1 + tc] = work;
        work += align_up((int64_t)m * src_ld[tc], BLK_ALIGN);
      }
    }
  } else {
    for (int b = 0; b < oldright.nq; ++b) { src_ld[b] = pad_ld(oldright.dims[b]); col_dim[b] = oldright.dims[b]; }
    for (int a = 0; a < oldleft.nq; ++a)
      for (int b = 0; b < oldright.nq; ++b) src_off[(size_t)a * ncol1 + b] = old_off[(size_t)a * oldright.nq + b];
  }
  const uint8_t src_base = onedot ? B2D_BASE_WORK : B2D_BASE_AUX;

  // ---- stage 1: rows -> the renormalised block:  out[a, col] = L[olda]^T src[olda, col]   (:17-31 / :892-911) ------------------------
  GemmBatch& G1 = onedot ? P.gemm_b : P.gemm_a;
  std::vector<int64_t> t1_off((size_t)rows.nq * ncol1, -1);
  for (int a = 0; a < rows.nq; ++a) {
    const int olda = rows.new_quanta_map[a];
    guess_check(olda >= 0 && olda < oldleft.nq, "newQuantaMap of the renormalised row space out of range");
    for (int b = 0; b < ncol1; ++b) {
      if (src_off[(size_t)olda * ncol1 + b] < 0) continue;
      guess_check(lrot_off[olda] >= 0 && d.lrot_cols[olda] == rows.dims[a], "left rotation matrix does not match the renormalised row space");
      const int m = rows.dims[a], n = col_dim[b], k = oldleft.dims[olda];
      if (d.mode == 2) {       // straight into the trial vector
        guess_check(P.trial.allowed(a, b), "trial block outside the target quantum number");
        const int p = P.trial.blk[(size_t)a * right.nq + b];
        add_gemm(G1, m, n, k, lrot_off[olda], B2D_BASE_AUX, 1, pad_ld(m), src_off[(size_t)olda * ncol1 + b], src_base, 0, src_ld[b], P.trial.dev_off[p], B2D_BASE_DST, P.trial.ld[p]);
      } else {
        guess_check(qn_allow(d.dq, &rows.q[3 * a], &(onedot ? rbasis : oldright).q[3 * b]), "previous wavefunction block outside the target quantum number");
        add_gemm(G1, m, n, k, lrot_off[olda], B2D_BASE_AUX, 1, pad_ld(m), src_off[(size_t)olda * ncol1 + b], src_base, 0, src_ld[b], work, B2D_BASE_WORK, src_ld[b]);
        t1_off[(size_t)a * ncol1 + b] = work;
        work += align_up((int64_t)m * src_ld[b], BLK_ALIGN);
      }
    }
  }
  P.t1_size = work;
  P.t2_off = work;

  if (shuffle) {
    // ---- destination of the shuffle: T2[lq, c] in WORK (mode 0) or the trial vector itself (mode 1), allowed by dq ----------------
    std::vector<int64_t> t2_off((size_t)left.nq * env.nq, -1);
    std::vector<int> t2_ld((size_t)left.nq * env.nq, 0);
    for (int lq = 0; lq < left.nq; ++lq)
      for (int c = 0; c < env.nq; ++c)
        if (qn_allow(d.dq, &left.q[3 * lq], &env.q[3 * c])) {
          if (d.mode == 1) {
            const int p = P.trial.blk[(size_t)lq * right.nq + c];
            guess_check(p >= 0, "trial block outside the target quantum number");
            t2_off[(size_t)lq * env.nq + c] = P.trial.dev_off[p]; t2_ld[(size_t)lq * env.nq + c] = P.trial.ld[p];
          } else {
            t2_off[(size_t)lq * env.nq + c] = work; t2_ld[(size_t)lq * env.nq + c] = pad_ld(env.dims[c]);
            work += align_up((int64_t)left.dims[lq] * pad_ld(env.dims[c]), BLK_ALIGN);
          }
        }
    // ---- stage 2: the shuffle [S'][E.d] -> [S'.d][E]   (:434-485, :200-256) ----------------------------------------------------------
    // where each un-collected piece sits inside its collected sector (Un/CollectQuanta: oldToNewState order)
    std::vector<int> oru_parent(oldright.nunc, -1), oru_first(oldright.nunc, 0), lu_parent(left.nunc, -1), lu_first(left.nunc, 0);
    for (int b = 0; b < oldright.nq; ++b) {
      int first = 0;
      for (int k = oldright.old_to_new_begin[b]; k < oldright.old_to_new_begin[b + 1]; ++k) {
        const int u = oldright.old_to_new[k];
        guess_check(u >= 0 && u < oldright.nunc, "oldright.oldToNewState out of range");
        oru_parent[u] = b; oru_first[u] = first; first += oldright.unc_dims[u];
      }
      guess_check(first == oldright.dims[b], "oldright: un-collected pieces do not add up to the collected sector");
    }
    for (int lq = 0; lq < left.nq; ++lq) {
      int first = 0;
      for (int k = left.old_to_new_begin[lq]; k < left.old_to_new_begin[lq + 1]; ++k) {
        const int u = left.old_to_new[k];
        guess_check(u >= 0 && u < left.nunc, "left.oldToNewState out of range");
        lu_parent[u] = lq; lu_first[u] = first; first += left.unc_dims[u];
      }
      guess_check(first == left.dims[lq], "left: un-collected pieces do not add up to the collected sector");
    }
    std::vector<int> hits((size_t)left.nunc * env.nq, 0);   // sources already planned for destination (ab, c)
    const int J = d.dq[1];
    for (int ab = 0; ab < left.nunc; ++ab) {
      const int a = left.unc_left[ab], b = left.unc_right[ab], lq = lu_parent[ab];
      guess_check(a >= 0 && a < sys.nq && b >= 0 && b < dot.nq, "left un-collected maps out of range");
      if (lq < 0) continue;
      guess_check(left.unc_dims[ab] == sys.dims[a] * dot.dims[b], "left: un-collected sector size is not the product of its factors");
      for (int cb = 0; cb < oldright.nunc; ++cb) {     // prevUnCollectedSI.quantaMap(c, b): increasing un-collected index
        if (oldright.unc_right[cb] != b) continue;
        const int c = oldright.unc_left[cb], bcol = oru_parent[cb];
        guess_check(c >= 0 && c < env.nq, "oldright un-collected maps out of range");
        if (bcol < 0) continue;
        if (!qn_allow(d.dq, &left.unc_q[3 * ab], &env.q[3 * c])) continue;                 // twowavefunction.allowed(ab, c)
        if (!qn_allow(d.dq, &sys.q[3 * a], &oldright.unc_q[3 * cb])) continue;             // the (a, cb) piece of the un-collected wave
        const int64_t src = t1_off[(size_t)a * ncol1 + bcol];
        if (src < 0) continue;                                                                // block absent from the previous wavefunction
        const int64_t dst = t2_off[(size_t)lq * env.nq + c];
        guess_check(dst >= 0, "shuffle destination outside the target quantum number");
        const int A = sys.q[3 * a + 1], B = dot.q[3 * b + 1], AB = left.unc_q[3 * ab + 1], C = env.q[3 * c + 1], CB = oldright.unc_q[3 * cb + 1];
        double f = am.commute_parity(&env.q[3 * c], &dot.q[3 * b], &oldright.unc_q[3 * cb]);
        f *= am.six_j(A, B, AB, C, J, CB) * std::sqrt((AB + 1.0) * (CB + 1.0)) * ((((A + B + J + C) / 2) & 1) ? -1.0 : 1.0);
        const int Al = sys.q[3 * a + 2], Bl = dot.q[3 * b + 2], ABl = left.unc_q[3 * ab + 2], Cl = env.q[3 * c + 2], CBl = oldright.unc_q[3 * cb + 2];
        if (ABl != (Al ^ Bl) || CBl != (Bl ^ Cl) || d.dq[2] != (ABl ^ Cl)) f = 0.0;          // Symmetry::spatial_sixj, abelian (Symmetry.C:520-526)
        if (f == 0.0) continue;
        KronTask t; std::memset(&t, 0, sizeof(t));
        t.a = src + oru_first[cb]; t.b = 0; t.dst = dst; t.coef = f;
        t.a_rows = sys.dims[a]; t.a_cols = oldright.unc_dims[cb]; t.lda = src_ld[bcol]; t.a_t = 0;
        t.b_rows = 1; t.b_cols = 1; t.ldb = 1; t.b_t = 0;
        t.row0 = lu_first[ab]; t.col0 = 0; t.ldd = t2_ld[(size_t)lq * env.nq + c];
        t.pad = d.mode == 1 ? 1 : 0;                                                          // destination: the trial vector / WORK
        guess_check(t.a_cols == env.dims[c] * dot.dims[b], "oldright: un-collected sector size is not the product of its factors");
        int& h = hits[(size_t)ab * env.nq + c];
        if ((int)P.rounds.size() <= h) P.rounds.resize(h + 1);
        P.rounds[h].push_back(t);
        ++h;
        P.shuffle_bytes += 8ll * 3 * t.a_rows * t.a_cols;
      }
    }
    // ---- stage 3 (two-dot only): trial[lq, tb] = T2[lq, c] R[tb]^T   (:33-50) --------------------------------------------------------
    if (d.mode == 0) {
      std::vector<char> seen(right.nq, 0);
      for (int c = 0; c < env.nq; ++c) {
        const int tb = env.new_quanta_map[c];
        guess_check(tb >= 0 && tb < right.nq && !seen[tb], "env.newQuantaMap is not an injective map into the right sectors");
        seen[tb] = 1;
        guess_check(rrot_off[tb] >= 0 && d.rrot_cols[tb] == env.dims[c], "right rotation matrix does not match the renormalised environment block");
        for (int lq = 0; lq < left.nq; ++lq) {
          if (t2_off[(size_t)lq * env.nq + c] < 0) continue;
          guess_check(P.trial.allowed(lq, tb), "trial block outside the target quantum number");
          const int p = P.trial.blk[(size_t)lq * right.nq + tb];
          const int k = env.dims[c];
          add_gemm(P.gemm_c, left.dims[lq], right.dims[tb], k, t2_off[(size_t)lq * env.nq + c], B2D_BASE_WORK, 0, pad_ld(k), rrot_off[tb], B2D_BASE_AUX, 1, pad_ld(k),
                   P.trial.dev_off[p], B2D_BASE_DST, P.trial.ld[p]);
        }
      }
    }
  }
  P.work_size = work;
  make_tiles(P.gemm_a, forced_class);
  make_tiles(P.gemm_b, forced_class);
  make_tiles(P.gemm_c, forced_class);
  P.valid = true;
  return P;
}

}  // namespace b2d
