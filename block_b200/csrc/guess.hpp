// Host planner of the guess-wavefunction transform of a two-dot step.
//
// Replaces GuessWave::transform_previous_wavefunction, two-dot branch (guess_wavefunction.C:524-636):
//   stage 1  TransformLeftBlock (:17-31)        T1[a, b]   = L[olda]^T . old[olda, b]            grouped GEMM
//   stage 2  onedot_shufflesysdot (:434-485, :200-256)   T2[lq, c][r0.., :] += f . T1[a, b][:, c0..]   HBM-bound scatter
//            f = getCommuteParity(E, dot, E.dot) . sixj(A, B, AB, C, J, CB) sqrt((AB+1)(CB+1)) (-1)^((A+B+J+C)/2) . spatial_sixj
//   stage 3  TransformRightBlock (:33-50)       trial[lq, tb] = T2[lq, c] . R[tb]^T              grouped GEMM
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
  int dq[3] = {0, 0, 0};
  // input image: old wavefunction blocks, left and right rotation matrices, packed one after the other in ONE device buffer
  std::vector<BlockDesc> in_old, in_lrot, in_rrot;    // ref_off: offset in the caller's flat array, dev_off: offset in the image
  int64_t old_size = 0, lrot_size = 0, rrot_size = 0; // doubles in the caller's arrays
  int64_t image_size = 0;                              // doubles on the device
  GemmBatch stage1, stage3;                            // stage 1 writes WORK[0, t1_size), stage 3 reads WORK[t2_off, ...) and writes DST
  std::vector<std::vector<KronTask>> rounds;           // stage 2; a / dst hold OFFSETS (doubles) into WORK until execution; tasks of
                                                       // one round never overlap (a destination's j-th source goes to round j)
  int64_t t1_size = 0, t2_off = 0, work_size = 0;
  Side left, right;                                    // sector tables of the big block's children (trial layout)
  PsiLayout trial;
  double flops = 0.0;
  int64_t shuffle_bytes = 0;                           // algorithmic bytes of stage 2: 8 x (read + read-modify-write) elements
};

inline void guess_check(bool ok, const char* what) {
  if (!ok) throw std::runtime_error(std::string("b2d_guess_plan: ") + what);
}

inline GuessPlan plan_guess_transform(const b2d_guess_desc& d, AngMom& am, int forced_class) {
  GuessPlan P;
  std::memcpy(P.dq, d.dq, sizeof(P.dq));
  const b2d_stateinfo &sys = d.sys, &dot = d.dot, &left = d.left, &right = d.right, &oldleft = d.oldleft, &oldright = d.oldright, &env = d.env;
  guess_check(sys.nq > 0 && dot.nq > 0 && left.nq > 0 && right.nq > 0 && oldleft.nq > 0 && oldright.nq > 0 && env.nq > 0, "empty StateInfo");
  guess_check(sys.new_quanta_map && env.new_quanta_map, "sys / env need newQuantaMap");
  guess_check(left.nunc > 0 && left.unc_q && left.unc_dims && left.unc_left && left.unc_right && left.old_to_new_begin && left.old_to_new, "left needs its un-collected tables");
  guess_check(oldright.nunc > 0 && oldright.unc_q && oldright.unc_dims && oldright.unc_left && oldright.unc_right && oldright.old_to_new_begin && oldright.old_to_new,
              "oldright needs its un-collected tables");
  guess_check(d.old_allowed && d.lrot_cols && d.rrot_cols, "null tables");
  for (int b = 0; b < dot.nq; ++b) guess_check(dot.dims[b] == 1, "dot sectors must hold one state (spin-adapted single site)");

  // ---- input image ------------------------------------------------------------------------------------------------------
  std::vector<int64_t> old_off((size_t)oldleft.nq * oldright.nq, -1), lrot_off(oldleft.nq, -1), rrot_off(right.nq, -1);
  int64_t dev = 0, ref = 0;
  for (int i = 0; i < oldleft.nq; ++i)
    for (int j = 0; j < oldright.nq; ++j)
      if (d.old_allowed[(size_t)i * oldright.nq + j]) {
        BlockDesc bd; bd.ref_off = ref; bd.dev_off = dev; bd.rows = oldleft.dims[i]; bd.cols = oldright.dims[j]; bd.ld = pad_ld(bd.cols); bd.pad = 0;
        old_off[(size_t)i * oldright.nq + j] = dev;
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
  for (int q = 0; q < right.nq; ++q)
    if (d.rrot_cols[q] > 0) {
      BlockDesc bd; bd.ref_off = ref; bd.dev_off = dev; bd.rows = right.dims[q]; bd.cols = d.rrot_cols[q]; bd.ld = pad_ld(bd.cols); bd.pad = 0;
      rrot_off[q] = dev;
      ref += (int64_t)bd.rows * bd.cols; dev += align_up((int64_t)bd.rows * bd.ld, BLK_ALIGN);
      P.in_rrot.push_back(bd);
    }
  P.rrot_size = ref;
  P.image_size = dev;

  // ---- stage 1: T1[a, b] = L[olda]^T old[olda, b]   (tempoldWave, allowed by dq like Wavefunction::AllowQuantaFor) -----------
  std::vector<int64_t> t1_off((size_t)sys.nq * oldright.nq, -1);
  int64_t work = 0;
  for (int a = 0; a < sys.nq; ++a) {
    const int olda = sys.new_quanta_map[a];
    guess_check(olda >= 0 && olda < oldleft.nq, "sys.newQuantaMap out of range");
    for (int b = 0; b < oldright.nq; ++b) {
      if (old_off[(size_t)olda * oldright.nq + b] < 0) continue;
      guess_check(qn_allow(d.dq, &sys.q[3 * a], &oldright.q[3 * b]), "previous wavefunction block outside the target quantum number");
      guess_check(lrot_off[olda] >= 0 && d.lrot_cols[olda] == sys.dims[a], "left rotation matrix does not match the renormalised system block");
      const int m = sys.dims[a], n = oldright.dims[b], k = oldleft.dims[olda];
      GSeg s; std::memset(&s, 0, sizeof(s));
      s.a = lrot_off[olda]; s.a_base = B2D_BASE_AUX; s.a_trans = 1; s.lda = pad_ld(m);          // stored k x m
      s.b = old_off[(size_t)olda * oldright.nq + b]; s.b_base = B2D_BASE_AUX; s.b_kmajor = 0; s.ldb = pad_ld(n);
      s.k = k; s.alpha = 1.0;
      GGroup g; std::memset(&g, 0, sizeof(g));
      g.c = work; g.c_base = B2D_BASE_WORK; g.ldc = pad_ld(n); g.m = m; g.n = n; g.accumulate = 0;
      g.seg_begin = (int)P.stage1.segs.size(); g.seg_end = g.seg_begin + 1;
      t1_off[(size_t)a * oldright.nq + b] = work;
      work += align_up((int64_t)m * g.ldc, BLK_ALIGN);
      P.stage1.segs.push_back(s); P.stage1.groups.push_back(g);
      P.flops += 2.0 * m * n * k;
    }
  }
  P.t1_size = work;
  P.t2_off = work;

  // ---- T2 layout: tempnewWave[lq, c], allowed by dq ---------------------------------------------------------------------------
  std::vector<int64_t> t2_off((size_t)left.nq * env.nq, -1);
  for (int lq = 0; lq < left.nq; ++lq)
    for (int c = 0; c < env.nq; ++c)
      if (qn_allow(d.dq, &left.q[3 * lq], &env.q[3 * c])) {
        t2_off[(size_t)lq * env.nq + c] = work;
        work += align_up((int64_t)left.dims[lq] * pad_ld(env.dims[c]), BLK_ALIGN);
      }
  P.work_size = work;

  // ---- stage 2: the shuffle ------------------------------------------------------------------------------------------------------
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
      const int64_t src = t1_off[(size_t)a * oldright.nq + bcol];
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
      t.a_rows = sys.dims[a]; t.a_cols = oldright.unc_dims[cb]; t.lda = pad_ld(oldright.dims[bcol]); t.a_t = 0;
      t.b_rows = 1; t.b_cols = 1; t.ldb = 1; t.b_t = 0;
      t.row0 = lu_first[ab]; t.col0 = 0; t.ldd = pad_ld(env.dims[c]);
      guess_check(t.a_cols == env.dims[c] * dot.dims[b], "oldright: un-collected sector size is not the product of its factors");
      int& h = hits[(size_t)ab * env.nq + c];
      if ((int)P.rounds.size() <= h) P.rounds.resize(h + 1);
      P.rounds[h].push_back(t);
      ++h;
      P.shuffle_bytes += 8ll * 3 * t.a_rows * t.a_cols;
    }
  }

  // ---- stage 3: trial[lq, tb] = T2[lq, c] R[tb]^T ---------------------------------------------------------------------------------
  P.left.nq = left.nq; P.left.q.assign(left.q, left.q + 3 * left.nq); P.left.dims.assign(left.dims, left.dims + left.nq);
  P.right.nq = right.nq; P.right.q.assign(right.q, right.q + 3 * right.nq); P.right.dims.assign(right.dims, right.dims + right.nq);
  P.trial.build(P.left, P.right, d.dq);
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
      const int m = left.dims[lq], n = right.dims[tb], k = env.dims[c];
      GSeg s; std::memset(&s, 0, sizeof(s));
      s.a = t2_off[(size_t)lq * env.nq + c]; s.a_base = B2D_BASE_WORK; s.a_trans = 0; s.lda = pad_ld(k);
      s.b = rrot_off[tb]; s.b_base = B2D_BASE_AUX; s.b_kmajor = 1; s.ldb = pad_ld(k);      // stored n x k
      s.k = k; s.alpha = 1.0;
      GGroup g; std::memset(&g, 0, sizeof(g));
      g.c = P.trial.dev_off[p]; g.c_base = B2D_BASE_DST; g.ldc = P.trial.ld[p]; g.m = m; g.n = n; g.accumulate = 0;
      g.seg_begin = (int)P.stage3.segs.size(); g.seg_end = g.seg_begin + 1;
      P.stage3.segs.push_back(s); P.stage3.groups.push_back(g);
      P.flops += 2.0 * m * n * k;
    }
  }
  make_tiles(P.stage1, forced_class);
  make_tiles(P.stage3, forced_class);
  P.valid = true;
  return P;
}

}  // namespace b2d
