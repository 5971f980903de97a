// block_guesscheck: CPU-only check of the guess-wavefunction binding + planner on EVERY guess of whole reference sweeps.
//
// The UNMODIFIED reference sweep with one link-time wrap, on GuessWave::guess_wavefunctions (solver.C:77).  After the reference's own
// function has produced its trial vectors, the same marshalling the drop-in binary uses (tests/dropin/guess_binding.hpp) describes the
// guess for b2d_guess_plan on a planning-only context (no GPU), the plan's descriptors are exported through the C ABI
// (b2d_guess_plan_export) and executed here with plain loops - grouped-GEMM segments, scatter tasks, padded layouts exactly as the
// device kernels read them - and the result is compared with the reference's trial vector.  One line per guess on stderr:
//   B2D_GUESSCHECK call=<n> mode=<m> root=<i> W=<len> max_abs_diff=<d>        (or ... skipped=<reason> for the forms left to the reference)
// TEST INFRASTRUCTURE (tests/test_guess_binding_cpu.py); built by `make -C oracle guesscheck` into oracle/_ref/block_guesscheck.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "wrap_syms_guess.h"

#include "spinblock.h"
#include "wavefunction.h"
#include "rotationmat.h"
#include "global.h"
#include "input.h"
#include "guess_wavefunction.h"
#include "sweep_params.h"

#include "block_b200.h"
#include "../../block_b200/csrc/gemm_desc.h"
#include "guess_binding.hpp"

using namespace SpinAdapted;
using std::vector;

namespace {

// KronTask of block_b200/csrc/kernels.h (that header needs the CUDA runtime; the layout is pinned by b2d_guess_plan_export(9))
struct KronTaskPod {
  int64_t a, b, dst;
  double coef;
  int32_t a_rows, a_cols, lda, a_t, b_rows, b_cols, ldb, b_t, row0, col0, ldd, pad;
};

b2d_ctx* g_ctx = 0;
int g_call = -1;

template <class T> vector<T> fetch(int what) {
  int64_t n = b2d_guess_plan_export(g_ctx, what, 0, 0);
  vector<T> v((size_t)(n < 0 ? 0 : n) / sizeof(T));
  if (n > 0) b2d_guess_plan_export(g_ctx, what, v.data(), n);
  return v;
}

void run_groups(const vector<GSeg>& segs, const vector<GGroup>& groups, double* const bases[B2D_NUM_BASES]) {
  for (const GGroup& g : groups) {
    vector<double> acc((size_t)g.m * g.n, 0.0);
    for (int s = g.seg_begin; s < g.seg_end; ++s) {
      const GSeg& sg = segs[s];
      const double* A = bases[sg.a_base] + sg.a;
      const double* Bm = bases[sg.b_base] + sg.b;
      for (int i = 0; i < g.m; ++i)
        for (int j = 0; j < g.n; ++j) {
          double x = 0;
          for (int k = 0; k < sg.k; ++k) {
            const double av = sg.a_trans ? A[(size_t)k * sg.lda + i] : A[(size_t)i * sg.lda + k];
            const double bv = sg.b_kmajor ? Bm[(size_t)j * sg.ldb + k] : Bm[(size_t)k * sg.ldb + j];
            x += av * bv;
          }
          acc[(size_t)i * g.n + j] += sg.alpha * x;
        }
    }
    double* C = bases[g.c_base] + g.c;
    for (int i = 0; i < g.m; ++i)
      for (int j = 0; j < g.n; ++j) C[(size_t)i * g.ldc + j] = (g.accumulate ? C[(size_t)i * g.ldc + j] : 0.0) + acc[(size_t)i * g.n + j];
  }
}

// executes the exported plan; returns the trial vector in FlattenInto order
vector<double> execute_plan(const b2d_binding::GuessBinding& B) {
  vector<int32_t> sz = fetch<int32_t>(9);
  if (sz.size() != 4 || sz[0] != (int)sizeof(GSeg) || sz[1] != (int)sizeof(GGroup) || sz[2] != (int)sizeof(KronTaskPod) || sz[3] != (int)sizeof(BlockDesc)) {
    fprintf(stderr, "B2D_GUESSCHECK descriptor sizes differ from the library's\n"); abort();
  }
  vector<char> head = fetch<char>(7);
  const int32_t* cnt = (const int32_t*)head.data();
  const int64_t* sizes = (const int64_t*)(head.data() + 16);
  const int64_t image_size = sizes[0], work_size = sizes[2];
  vector<BlockDesc> in = fetch<BlockDesc>(6), tb = fetch<BlockDesc>(8);
  vector<double> image((size_t)std::max<int64_t>(image_size, 1), 0.0), work((size_t)std::max<int64_t>(work_size, 1), 0.0);
  for (size_t k = 0; k < in.size(); ++k) {
    const double* src = (int)k < cnt[0] ? B.old.data() : ((int)k < cnt[0] + cnt[1] ? B.lrot.data() : B.rrot.data());
    for (int r = 0; r < in[k].rows; ++r)
      for (int c = 0; c < in[k].cols; ++c) image[(size_t)in[k].dev_off + (size_t)r * in[k].ld + c] = src[(size_t)in[k].ref_off + (size_t)r * in[k].cols + c];
  }
  int64_t Wp = 1, W = 0;
  for (const BlockDesc& b : tb) { Wp = std::max<int64_t>(Wp, b.dev_off + (int64_t)b.rows * b.ld); W = std::max<int64_t>(W, b.ref_off + (int64_t)b.rows * b.cols); }
  vector<double> dst((size_t)Wp, 0.0);
  double* bases[B2D_NUM_BASES] = {0, 0, work.data(), dst.data(), image.data()};
  run_groups(fetch<GSeg>(0), fetch<GGroup>(1), bases);
  run_groups(fetch<GSeg>(10), fetch<GGroup>(11), bases);
  vector<KronTaskPod> tasks = fetch<KronTaskPod>(2);
  for (const KronTaskPod& t : tasks) {      // rounds are concatenated in execution order
    const double* A = ((t.pad & 2) ? image.data() : work.data()) + t.a;
    double* D = ((t.pad & 1) ? dst.data() : work.data()) + t.dst;
    for (int i = 0; i < t.a_rows; ++i)
      for (int j = 0; j < t.a_cols; ++j) {
        const double v = t.a_t ? A[(size_t)j * t.lda + i] : A[(size_t)i * t.lda + j];
        D[(size_t)(t.row0 + i) * t.ldd + t.col0 + j] += t.coef * v;
      }
  }
  run_groups(fetch<GSeg>(4), fetch<GGroup>(5), bases);
  vector<double> flat((size_t)W, 0.0);
  for (const BlockDesc& b : tb)
    for (int r = 0; r < b.rows; ++r)
      for (int c = 0; c < b.cols; ++c) flat[(size_t)b.ref_off + (size_t)r * b.cols + c] = dst[(size_t)b.dev_off + (size_t)r * b.ld + c];
  return flat;
}

void flatten(const SparseMatrix& w, vector<double>& out) {
  out.clear();
  for (int l = 0; l < w.nrows(); ++l)
    for (int r = 0; r < w.ncols(); ++r)
      if (w.allowed(l, r)) { const Matrix& m = w.operator_element(l, r); out.insert(out.end(), m.Store(), m.Store() + m.Storage()); }
}

}  // namespace

namespace SpinAdapted {

void real_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) asm("__real_" SYM_guess_wavefunctions);
void wrap_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) asm("__wrap_" SYM_guess_wavefunctions);
void wrap_guess(vector<Wavefunction>& solution, DiagonalMatrix& e, const SpinBlock& big, const guessWaveTypes& gw, const bool& onedot,
                const bool& transpose_guess_wave, double additional_noise, int currentState) {
  ++g_call;
  real_guess(solution, e, big, gw, onedot, transpose_guess_wave, additional_noise, currentState);
  if (gw == BASIC) return;
  if (!g_ctx && b2d_create(-1, &g_ctx)) { fprintf(stderr, "B2D_GUESSCHECK b2d_create(-1) failed: %s\n", b2d_last_error(0)); abort(); }
  for (size_t i = 0; i < solution.size(); ++i) {
    const int state = (dmrginp.setStateSpecific() || dmrginp.calc_type() == COMPRESS || dmrginp.calc_type() == MPS_NEVPT) ? currentState : (int)i;
    b2d_binding::GuessBinding B;
    if (!b2d_binding::make_guess_binding(B, big, gw, onedot, transpose_guess_wave, state)) {
      fprintf(stderr, "B2D_GUESSCHECK call=%d root=%d skipped=%s\n", g_call, (int)i, B.why);
      continue;
    }
    double info[8];
    if (b2d_guess_plan(g_ctx, &B.d, info, 8)) { fprintf(stderr, "B2D_GUESSCHECK call=%d mode=%d root=%d plan_error=%s\n", g_call, (int)B.d.mode, (int)i, b2d_last_error(g_ctx)); continue; }
    vector<double> got = execute_plan(B), ref;
    flatten(solution[i], ref);
    double worst = got.size() == ref.size() ? 0.0 : 1e300, scale = 0;
    for (size_t k = 0; k < ref.size() && k < got.size(); ++k) { worst = std::max(worst, fabs(ref[k] - got[k])); scale = std::max(scale, fabs(ref[k])); }
    fprintf(stderr, "B2D_GUESSCHECK call=%d mode=%d root=%d W=%d max_abs_diff=%.3e max_abs=%.3e tasks=%d rounds=%d\n", g_call, (int)B.d.mode, (int)i, (int)ref.size(), worst, scale,
            (int)info[6], (int)info[7]);
  }
}

}  // namespace SpinAdapted
