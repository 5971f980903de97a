// CPU-baseline timer: the UNMODIFIED reference's operatorfunctions::TensorMultiply (operatorfunctions.C:485-537) run on
// the same synthetic big block bench.py gives the GPU.  TEST / MEASUREMENT INFRASTRUCTURE ONLY.
//
// This translation unit is OUR code, linked against the reference objects (oracle/_ref/libblockref.a); no reference
// source is edited or copied.  It builds the reference's own StateInfo / SpinBlock / SparseMatrix / Wavefunction objects
// for the sector tables it is handed, fills the operator blocks with a cheap counter-based stream (timing does not
// depend on the values) and calls the reference's TensorMultiply for a sample of multiplyH's operator terms.
// Parallelism follows the reference: an OpenMP loop over operator terms with one sigma accumulator per thread
// (operatorloops.h:87-97, distribute.h:128-170) and single-threaded BLAS inside (run with OPENBLAS_NUM_THREADS=1).
//
// usage: ref_bench <spec file> <repetitions> [threads] [psi.bin sigma_out.bin]
// spec (text):  nL / nL x "N 2S dim" / nR / nR x "N 2S dim" / "N 2S" of psi / nterms /
//               nterms x "l_dN l_2S l_fermion l_transposed  r_dN r_2S r_fermion r_transposed  scale  l_seed l_amp r_seed r_amp"
// A non-zero seed fills that operator with the SAME counter-based stream the CUDA library's b2d_fill_op_random uses
// (splitmix64 of seed and the packed element index), so that with psi.bin (flat doubles, Wavefunction::CollectFrom order)
// the reference's sum of the sampled TensorMultiply terms lands in sigma_out.bin (FlattenInto order) and can be compared
// with the GPU's at full benchmark size.
// output (stdout): REFBENCH seconds=<s> flops=<f> threads=<t> terms=<n> reps=<r>
#include <omp.h>
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "BaseOperator.h"
#include "StateInfo.h"
#include "Symmetry.h"
#include "global.h"
#include "input.h"
#include "operatorfunctions.h"
#include "spinblock.h"
#include "wavefunction.h"

using namespace SpinAdapted;

namespace {

double now_s() { timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }

// a concrete SparseMatrix: storage only (the hot path never calls the virtuals below)
struct SynthOp : public SparseMatrix {
  boost::shared_ptr<SparseMatrix> getworkingrepresentation(const SpinBlock*) { return boost::shared_ptr<SparseMatrix>(this, boostutils::null_deleter()); }
  void build(const SpinBlock&) {}
  double redMatrixElement(Csf, std::vector<Csf>&, const SpinBlock*) { return 0.0; }
};

void fill(SparseMatrix& m, unsigned long long seed) {
  unsigned long long x = seed * 0x9E3779B97F4A7C15ull + 1;
  for (int i = 0; i < m.nrows(); ++i)
    for (int j = 0; j < m.ncols(); ++j)
      if (m.allowed(i, j)) {
        Matrix& a = m.operator_element(i, j);
        double* p = a.Store();
        for (int k = 0; k < a.Storage(); ++k) {
          x ^= x << 13; x ^= x >> 7; x ^= x << 17;
          p[k] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        }
      }
}

// the generator of b2d_fill_op_random (block_b200/csrc/kernels.cu counter_uniform): value = 2 amp u(seed, packed index)
double counter_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}
void fill_counter(SparseMatrix& m, unsigned long long seed, double amp) {
  unsigned long long off = 0;
  for (int i = 0; i < m.nrows(); ++i)
    for (int j = 0; j < m.ncols(); ++j)
      if (m.allowed(i, j)) {
        Matrix& a = m.operator_element(i, j);
        double* p = a.Store();
        const long n = a.Storage();
#pragma omp parallel for
        for (long k = 0; k < n; ++k) p[k] = 2.0 * amp * counter_uniform(seed, off + k);
        off += n;
      }
}

struct TermSpec { int ldn, ls, lf, lt, rdn, rs, rf, rt; double scale; unsigned long long lseed, rseed; double lamp, ramp; };

StateInfo read_stateinfo(std::ifstream& in) {
  int n; in >> n;
  std::vector<SpinQuantum> q(n);
  std::vector<int> d(n);
  for (int i = 0; i < n; ++i) {
    int N, S; in >> N >> S >> d[i];
    q[i] = SpinQuantum(N, SpinSpace(S), IrrepSpace(0));
  }
  return StateInfo(n, &q[0], &d[0]);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_bench <spec> <reps> [threads]\n"); return 2; }
  const int reps = atoi(argv[2]);
  int threads = argc > 3 ? atoi(argv[3]) : omp_get_max_threads();
  omp_set_num_threads(threads);
  // global state the path reads: the reference's own defaults (spin-adapted, 9j table built: input.C:61-186),
  // abelian C1 symmetry
  dmrginp.initialize_defaults();
  sym = "c1";
  NonabelianSym = false;
  Symmetry::InitialiseTable(sym);

  std::ifstream in(argv[1]);
  if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  SpinBlock lb, rb, big;
  lb.braStateInfo = lb.ketStateInfo = read_stateinfo(in);
  rb.braStateInfo = rb.ketStateInfo = read_stateinfo(in);
  big.leftBlock = &lb; big.rightBlock = &rb;
  big.braStateInfo.leftStateInfo = big.ketStateInfo.leftStateInfo = &lb.ketStateInfo;
  big.braStateInfo.rightStateInfo = big.ketStateInfo.rightStateInfo = &rb.ketStateInfo;
  int pn, ps; in >> pn >> ps;
  const SpinQuantum target(pn, SpinSpace(ps), IrrepSpace(0));
  int nterms; in >> nterms;
  std::vector<TermSpec> spec(nterms);
  for (TermSpec& t : spec) in >> t.ldn >> t.ls >> t.lf >> t.lt >> t.rdn >> t.rs >> t.rf >> t.rt >> t.scale >> t.lseed >> t.lamp >> t.rseed >> t.ramp;
  if (!in) { fprintf(stderr, "short spec file\n"); return 2; }

  Wavefunction c;
  c.initialise(target, &big, false);
  fill(c, 7);
  const char* psi_path = argc > 5 ? argv[4] : 0;
  const char* out_path = argc > 5 ? argv[5] : 0;
  if (psi_path) {
    FILE* f = fopen(psi_path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", psi_path); return 2; }
    for (int l = 0; l < c.nrows(); ++l)
      for (int r = 0; r < c.ncols(); ++r)
        if (c.allowed(l, r)) {
          Matrix& m = c.operator_element(l, r);
          if (fread(m.Store(), 8, m.Storage(), f) != (size_t)m.Storage()) { fprintf(stderr, "short psi file\n"); return 2; }
        }
    fclose(f);
  }
  std::vector<Wavefunction> v(threads);
  for (Wavefunction& w : v) w.initialise(target, &big, false);

  std::vector<SynthOp> lops(nterms), rops(nterms);
  double flops = 0.0;
  for (int k = 0; k < nterms; ++k) {
    const TermSpec& t = spec[k];
    lops[k].set_deltaQuantum(1, SpinQuantum(t.ldn, SpinSpace(t.ls), IrrepSpace(0)));
    lops[k].set_fermion() = t.lf != 0;
    lops[k].allocate(lb.ketStateInfo);
    lops[k].set_initialised() = true;
    if (t.lseed) fill_counter(lops[k], t.lseed, t.lamp); else fill(lops[k], 100 + 2 * k);
    rops[k].set_deltaQuantum(1, SpinQuantum(t.rdn, SpinSpace(t.rs), IrrepSpace(0)));
    rops[k].set_fermion() = t.rf != 0;
    rops[k].allocate(rb.ketStateInfo);
    rops[k].set_initialised() = true;
    if (t.rseed) fill_counter(rops[k], t.rseed, t.ramp); else fill(rops[k], 101 + 2 * k);
    // the dgemm flops TensorMultiply issues for this term (operatorfunctions.C:515,530)
    const StateInfo& sl = lb.ketStateInfo;
    const StateInfo& sr = rb.ketStateInfo;
    const int nl = sl.quanta.size(), nr = sr.quanta.size();
    for (int lQ = 0; lQ < nl; ++lQ)
      for (int rQp = 0; rQp < nr; ++rQp)
        for (int lQp = 0; lQp < nl; ++lQp) {
          const bool la = t.lt ? lops[k].allowed(lQp, lQ) : lops[k].allowed(lQ, lQp);
          if (!la || !c.allowed(lQp, rQp)) continue;
          flops += 2.0 * sl.quantaStates[lQ] * sl.quantaStates[lQp] * sr.quantaStates[rQp];
          for (int rQ = 0; rQ < nr; ++rQ) {
            const bool ra = t.rt ? rops[k].allowed(rQp, rQ) : rops[k].allowed(rQ, rQp);
            if (c.allowed(lQ, rQ) && ra) flops += 2.0 * sl.quantaStates[lQ] * sr.quantaStates[rQp] * sr.quantaStates[rQ];
          }
        }
  }

  const SpinQuantum opQ(0, SpinSpace(0), IrrepSpace(0));
  double secs = 0.0;
  for (int r = 0; r < reps; ++r) {
    const double t0 = now_s();
#pragma omp parallel for schedule(guided)
    for (int k = 0; k < nterms; ++k) {
      const TermSpec& t = spec[k];
      Wavefunction& out = v[omp_get_thread_num()];
      Transposeview lt(lops[k]), rt(rops[k]);
      const Baseoperator<Matrix>& a = t.lt ? (const Baseoperator<Matrix>&)lt : (const Baseoperator<Matrix>&)lops[k];
      const Baseoperator<Matrix>& b = t.rt ? (const Baseoperator<Matrix>&)rt : (const Baseoperator<Matrix>&)rops[k];
      operatorfunctions::TensorMultiply(&lb, a, b, &big, c, out, opQ, t.scale);
    }
    secs += now_s() - t0;
  }
  double check = 0.0;
  for (Wavefunction& w : v) check += DotProduct(w, w);
  if (out_path) {   // sum of the per-thread accumulators (accumulateMultiThread, distribute.h:172-224), flattened
    for (size_t i = 1; i < v.size(); ++i) ScaleAdd(1.0, v[i], v[0]);
    FILE* f = fopen(out_path, "wb");
    for (int l = 0; l < v[0].nrows(); ++l)
      for (int r = 0; r < v[0].ncols(); ++r)
        if (v[0].allowed(l, r)) {
          Matrix& m = v[0].operator_element(l, r);
          fwrite(m.Store(), 8, m.Storage(), f);
        }
    fclose(f);
  }
  printf("REFBENCH seconds=%.6f flops=%.6e threads=%d terms=%d reps=%d check=%.6e\n", secs, flops * reps, threads, nterms, reps, check);
  return 0;
}
