// Reference-side marshalling of the guess-wavefunction transform: from the reference's own objects (the big block of the
// block iteration, its scratch files) to b2d_guess_desc + the three flat arrays of b2d_guess_transform.  This is the binding
// INTEGRATION.md section 5b shows; it is shared by the drop-in binary (tests/dropin/block_gpu_hooks.cpp) and by the CPU-only checker
// (tests/dropin/guess_plan_cpu_check.cpp), which runs it on every guess of whole reference sweeps.  TEST INFRASTRUCTURE like the rest of
// tests/dropin: compiled together with the unmodified reference headers; include after them and after block_b200.h.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

namespace b2d_binding {
using namespace SpinAdapted;

struct GuessBinding {
  b2d_guess_desc d;
  std::vector<std::vector<int32_t> > keep;     // the integer tables the descriptor points into
  std::vector<uint8_t> allowed;
  std::vector<double> old, lrot, rrot;         // previous wavefunction (allowed blocks, row-major), rotation matrices (kept sectors)
  std::vector<int32_t> lcols, rcols;
  StateInfo oldSI, newenv;
  const char* why = "";                        // reason when the guess is left to the reference
  ~GuessBinding() { if (oldSI.hasAllocatedMemory) oldSI.Free(); }     // LoadWavefunctionInfo allocated the tree (StateInfo.C:379-399)
};

inline void fill_stateinfo(b2d_stateinfo& o, const StateInfo& s, std::vector<std::vector<int32_t> >& keep) {
  memset(&o, 0, sizeof(o));
  auto hold = [&](const std::vector<int>& v) -> const int32_t* {
    keep.push_back(std::vector<int32_t>(v.begin(), v.end()));
    if (keep.back().empty()) keep.back().push_back(0);
    return keep.back().data();
  };
  auto quanta = [&](const StateInfo& t) -> const int32_t* {
    std::vector<int> q;
    for (size_t i = 0; i < t.quanta.size(); ++i) { q.push_back(t.quanta[i].get_n()); q.push_back(t.quanta[i].get_s().getirrep()); q.push_back(t.quanta[i].get_symm().getirrep()); }
    return hold(q);
  };
  o.nq = (int32_t)s.quanta.size();
  o.q = quanta(s);
  o.dims = hold(s.quantaStates);
  o.new_quanta_map = s.newQuantaMap.size() == s.quanta.size() && !s.newQuantaMap.empty() ? hold(s.newQuantaMap) : 0;
  if (s.hasCollectedQuanta && s.unCollectedStateInfo) {
    const StateInfo& u = *s.unCollectedStateInfo;
    o.nunc = (int32_t)u.quanta.size();
    o.unc_q = quanta(u);
    o.unc_dims = hold(u.quantaStates);
    o.unc_left = hold(u.leftUnMapQuanta);
    o.unc_right = hold(u.rightUnMapQuanta);
    std::vector<int> flat, begin(1, 0);
    for (size_t q = 0; q < s.oldToNewState.size(); ++q) { flat.insert(flat.end(), s.oldToNewState[q].begin(), s.oldToNewState[q].end()); begin.push_back((int)flat.size()); }
    o.old_to_new_begin = hold(begin);
    o.old_to_new = hold(flat);
  }
}

inline void pack_rotation(const std::vector<Matrix>& rot, std::vector<int32_t>& cols, std::vector<double>& data) {
  for (size_t q = 0; q < rot.size(); ++q) {
    cols.push_back(rot[q].Ncols());
    if (rot[q].Ncols()) data.insert(data.end(), rot[q].Store(), rot[q].Store() + rot[q].Storage());
  }
}

// Loads what the reference's guess function would load for root `state` of this block iteration (same files, same calls) and describes
// it for b2d_guess_plan.  false: this guess is not one of the five covered forms (B.why says which) - the caller uses the reference's own.
inline bool make_guess_binding(GuessBinding& B, const SpinBlock& big, guessWaveTypes gw, bool onedot, bool transpose_guess_wave, int state) {
  if (gw != TRANSFORM && gw != TRANSPOSE) { B.why = "BASIC guess"; return false; }
  if (!dmrginp.spinAdapted() || dmrginp.hamiltonian() == BCS || dmrginp.transition_diff_irrep()) { B.why = "run type"; return false; }
  if (!big.get_leftBlock() || !big.get_rightBlock() || !big.get_leftBlock()->get_leftBlock() || !big.get_leftBlock()->get_rightBlock()) { B.why = "left child is not a product block"; return false; }
  const StateInfo& bs = big.get_stateInfo();
  Wavefunction oldWave;
  std::vector<Matrix> lrot, rrot;
  B.keep.reserve(160);
  memset(&B.d, 0, sizeof(B.d));
  b2d_guess_desc& d = B.d;
  const std::vector<int>& sys_sites = big.get_leftBlock()->get_leftBlock()->get_sites();
  std::vector<int> right_plus_dot = big.get_rightBlock()->get_sites();
  right_plus_dot.insert(right_plus_dot.end(), big.get_leftBlock()->get_rightBlock()->get_sites().begin(), big.get_leftBlock()->get_rightBlock()->get_sites().end());
  std::sort(right_plus_dot.begin(), right_plus_dot.end());
  if (gw == TRANSFORM && !onedot) {                      // transform_previous_wavefunction, guess_wavefunction.C:524-636
    d.mode = 0;
    oldWave.LoadWavefunctionInfo(B.oldSI, sys_sites, state);                                   // :537
    LoadRotationMatrix(sys_sites, lrot, state);                                                // :538
    LoadRotationMatrix(big.get_rightBlock()->get_sites(), rrot, state);                        // :613
  } else if (gw == TRANSFORM && transpose_guess_wave) {  // one-dot, dot on the system side: :537-538, :617-622, :832-936
    d.mode = 1;
    oldWave.LoadWavefunctionInfo(B.oldSI, sys_sites, state);
    LoadRotationMatrix(sys_sites, lrot, state);
    LoadRotationMatrix(right_plus_dot, rrot, state);
    TensorProduct(*(bs.rightStateInfo), *(bs.leftStateInfo->rightStateInfo), B.newenv, NO_PARTICLE_SPIN_NUMBER_CONSTRAINT);   // :853-856, the reference's own integer bookkeeping
    B.newenv.CollectQuanta();
  } else if (gw == TRANSFORM) {                          // one-dot, dot on the environment side: :541-542, :625
    d.mode = 2;
    oldWave.LoadWavefunctionInfo(B.oldSI, big.get_leftBlock()->get_sites(), state);
    LoadRotationMatrix(big.get_leftBlock()->get_sites(), lrot, state);
    LoadRotationMatrix(big.get_rightBlock()->get_sites(), rrot, state);
  } else if (!onedot) {                                  // transpose_previous_wavefunction, :55-84
    d.mode = 3;
    oldWave.LoadWavefunctionInfo(B.oldSI, big.get_rightBlock()->get_sites(), state);          // :62
    if (oldWave.get_onedot()) { B.why = "one-dot -> two-dot switch"; return false; }
  } else {                                               // onedot_transpose_wavefunction, :100-112, :140-198
    d.mode = 4;
    oldWave.LoadWavefunctionInfo(B.oldSI, right_plus_dot, state);
  }
  if (oldWave.get_deltaQuantum_size() != 1) { B.why = "wavefunction with several target quanta"; return false; }
  SpinQuantum dq = oldWave.get_deltaQuantum(0);
  d.dq[0] = dq.get_n(); d.dq[1] = dq.get_s().getirrep(); d.dq[2] = dq.get_symm().getirrep();
  fill_stateinfo(d.left, *bs.leftStateInfo, B.keep);
  fill_stateinfo(d.right, *bs.rightStateInfo, B.keep);
  fill_stateinfo(d.oldleft, *B.oldSI.leftStateInfo, B.keep);
  if (d.mode == 0 || d.mode == 1 || d.mode == 4) {
    fill_stateinfo(d.sys, *bs.leftStateInfo->leftStateInfo, B.keep);
    fill_stateinfo(d.dot, *bs.leftStateInfo->rightStateInfo, B.keep);
  }
  if (d.mode == 0) {
    fill_stateinfo(d.oldright, *B.oldSI.rightStateInfo, B.keep);
    fill_stateinfo(d.env, *B.oldSI.rightStateInfo->leftStateInfo, B.keep);
  } else {
    fill_stateinfo(d.oldcol, *B.oldSI.rightStateInfo, B.keep);
    if (d.mode == 1) fill_stateinfo(d.oldright, B.newenv, B.keep);
  }
  for (int a = 0; a < oldWave.nrows(); ++a)
    for (int b = 0; b < oldWave.ncols(); ++b) {
      B.allowed.push_back(oldWave.allowed(a, b) ? 1 : 0);
      if (oldWave.allowed(a, b)) { const Matrix& m = oldWave.operator_element(a, b); B.old.insert(B.old.end(), m.Store(), m.Store() + m.Storage()); }
    }
  d.old_allowed = B.allowed.data();
  if (d.mode <= 2) {
    pack_rotation(lrot, B.lcols, B.lrot);
    pack_rotation(rrot, B.rcols, B.rrot);
    if ((int)B.lcols.size() != d.oldleft.nq || (int)B.rcols.size() != (d.mode == 1 ? d.oldright.nq : d.right.nq)) { B.why = "rotation matrices do not match the StateInfo of their blocks"; return false; }
    d.lrot_cols = B.lcols.data(); d.rrot_cols = B.rcols.data();
  }
  if (B.old.empty()) B.old.push_back(0);
  if (B.lrot.empty()) B.lrot.push_back(0);
  if (B.rrot.empty()) B.rrot.push_back(0);
  return true;
}

}  // namespace b2d_binding
