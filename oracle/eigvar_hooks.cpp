// Reference-vs-reference conditioning experiment.  TEST INFRASTRUCTURE ONLY - nothing here is shipped or measured as product.
//
// The unmodified reference sweep with ONE link-time interposition (GNU ld --wrap, no reference source edited): the per-sector
// eigen-decomposition of the reduced density matrix, diagonalise_dm (rotationmat.C:258-279; the reference calls dsyev_ 'V','U'
// through diagonalise, MatrixBLAS.C:381-414).  ORACLE_EIGVAR selects the solver:
//
//   dsyev     the reference's own call, re-issued here (control: must reproduce block.spin_adapted digit for digit)
//   dsyevd    LAPACK divide & conquer from the SAME OpenBLAS
//   dsyevr    LAPACK MRRR from the same OpenBLAS
//   ulp       dsyev_ on rho with every element multiplied by (1 + s * 2^-52), s in {-1, 0, +1} from a fixed counter-based hash:
//             a perturbation the size of ONE rounding error of the density matrix build
//   jacobi    cyclic one-sided (Hestenes) Jacobi in plain C++ (the algorithm family of the device solver)
//
// Everything else - wavefunction, density matrix, the 1e-14 clamp, sort_weights / assign_matrix_by_dm, operator rotation - is the
// reference's own code.  The per-sweep energies of these runs measure how far the REFERENCE's sweeps move when only the basis
// returned inside (near-)degenerate eigenspaces of rho changes: tests/golden/make_eigvar_golden.py -> tests/golden/eigvar_spread.npz.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "wrap_syms_eigvar.h"

#include "BaseOperator.h"
#include "rotationmat.h"
#include "global.h"

using namespace SpinAdapted;
using std::vector;

extern "C" {
void x_dsyev(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w, double* work, const int* lwork, int* info) asm("dsyev_");
void x_dsyevd(const char* jobz, const char* uplo, const int* n, double* a, const int* lda, double* w, double* work, const int* lwork, int* iwork,
             const int* liwork, int* info) asm("dsyevd_");
void x_dsyevr(const char* jobz, const char* range, const char* uplo, const int* n, double* a, const int* lda, const double* vl, const double* vu,
             const int* il, const int* iu, const double* abstol, int* m, double* w, double* z, const int* ldz, int* isuppz, double* work,
             const int* lwork, int* iwork, const int* liwork, int* info) asm("dsyevr_");
}

namespace {

uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull; x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; return x ^ (x >> 31);
}

// a: n x n symmetric, row-major == column-major.  On return w ascending, a holds eigenvectors as LAPACK columns (a[j*n + i] = v_j[i]).
void solve(const std::string& how, int n, vector<double>& a, vector<double>& w, uint64_t salt) {
  int info = 0;
  w.assign(n, 0.0);
  if (n == 0) return;
  if (how == "dsyev" || how == "ulp") {
    if (how == "ulp")
      for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
          int s = (int)(mix(salt * 1315423911ull + (uint64_t)i * n + j) % 3) - 1;
          double v = a[(size_t)i * n + j] * (1.0 + s * 2.220446049250313e-16);
          a[(size_t)i * n + j] = a[(size_t)j * n + i] = v;
        }
    double q; int lw = -1;
    x_dsyev("V", "L", &n, a.data(), &n, w.data(), &q, &lw, &info);           // MatrixBLAS.C:397 (query with 'L', solve with 'U')
    lw = (int)q; vector<double> work(lw);
    x_dsyev("V", "U", &n, a.data(), &n, w.data(), work.data(), &lw, &info);
  } else if (how == "dsyevd") {
    double q; int iq, lw = -1, liw = -1;
    x_dsyevd("V", "U", &n, a.data(), &n, w.data(), &q, &lw, &iq, &liw, &info);
    lw = (int)q; liw = iq; vector<double> work(lw); vector<int> iwork(liw);
    x_dsyevd("V", "U", &n, a.data(), &n, w.data(), work.data(), &lw, iwork.data(), &liw, &info);
  } else if (how == "dsyevr") {
    vector<double> z((size_t)n * n); vector<int> isuppz(2 * n);
    double q, vl = 0, vu = 0, abstol = 0; int iq, lw = -1, liw = -1, il = 0, iu = 0, m = 0;
    x_dsyevr("V", "A", "U", &n, a.data(), &n, &vl, &vu, &il, &iu, &abstol, &m, w.data(), z.data(), &n, isuppz.data(), &q, &lw, &iq, &liw, &info);
    lw = (int)q; liw = iq; vector<double> work(lw); vector<int> iwork(liw);
    x_dsyevr("V", "A", "U", &n, a.data(), &n, &vl, &vu, &il, &iu, &abstol, &m, w.data(), z.data(), &n, isuppz.data(), work.data(), &lw, iwork.data(),
            &liw, &info);
    a = z;
  } else if (how == "jacobi") {
    // one-sided Jacobi on the rows of G = rho (G <- G J; V accumulates): at convergence the rows of G are mutually orthogonal,
    // |g_i| = |lambda_i| and V's rows are the eigenvectors; eigenvalues as Rayleigh quotients with the original matrix.
    vector<double> g = a, v((size_t)n * n, 0.0), a0 = a;
    for (int i = 0; i < n; ++i) v[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
      double off = 0.0;
      for (int p = 0; p < n - 1; ++p)
        for (int q2 = p + 1; q2 < n; ++q2) {
          double app = 0, aqq = 0, apq = 0;
          const double* gp = &g[(size_t)p * n]; const double* gq = &g[(size_t)q2 * n];
          for (int k = 0; k < n; ++k) { app += gp[k] * gp[k]; aqq += gq[k] * gq[k]; apq += gp[k] * gq[k]; }
          if (std::fabs(apq) <= 1e-17 * std::sqrt(app * aqq) || apq == 0.0) continue;
          off = std::max(off, std::fabs(apq) / std::sqrt(app * aqq));
          double zeta = (aqq - app) / (2.0 * apq);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
          double* hp = &g[(size_t)p * n]; double* hq = &g[(size_t)q2 * n];
          double* vp = &v[(size_t)p * n]; double* vq = &v[(size_t)q2 * n];
          for (int k = 0; k < n; ++k) {
            double x = hp[k], y = hq[k]; hp[k] = c * x - s * y; hq[k] = s * x + c * y;
            x = vp[k]; y = vq[k]; vp[k] = c * x - s * y; vq[k] = s * x + c * y;
          }
        }
      if (off < 1e-15) break;
    }
    vector<std::pair<double, int>> ord(n);
    for (int i = 0; i < n; ++i) {
      double r = 0.0;
      for (int k = 0; k < n; ++k) { double t = 0.0; for (int l = 0; l < n; ++l) t += a0[(size_t)k * n + l] * v[(size_t)i * n + l]; r += v[(size_t)i * n + k] * t; }
      ord[i] = std::make_pair(r, i);
    }
    std::sort(ord.begin(), ord.end());
    for (int j = 0; j < n; ++j) { w[j] = ord[j].first; for (int k = 0; k < n; ++k) a[(size_t)j * n + k] = v[(size_t)ord[j].second * n + k]; }
  } else {
    fprintf(stderr, "ORACLE_EIGVAR=%s unknown\n", how.c_str());
    abort();
  }
  if (info != 0) { fprintf(stderr, "eigvar: %s info=%d\n", how.c_str(), info); abort(); }
}

uint64_t g_calls = 0;

}  // namespace

void wrap_diagdm(SparseMatrix& traced, SparseMatrix& transform, vector<DiagonalMatrix>& eigs) asm("__wrap_" SYM_diagonalise_dm);
void wrap_diagdm(SparseMatrix& traced, SparseMatrix& transform, vector<DiagonalMatrix>& eigs) {
  const char* e = getenv("ORACLE_EIGVAR");
  const std::string how = e ? e : "dsyev";
  const int nq = traced.nrows();
  eigs.resize(nq);
  for (int q = 0; q < nq; ++q) {
    Matrix& rho = traced.operator_element(q, q);
    const int n = rho.Nrows();
    vector<double> a(rho.Store(), rho.Store() + (size_t)n * n), w;
    solve(how, n, a, w, ++g_calls);
    Matrix& vec = transform.operator_element(q, q);
    vec.ReSize(n, n);
    for (int i = 0; i < n; ++i)                       // MatrixBLAS.C:411-413: eigenvector i becomes COLUMN i of the row-major matrix
      for (int j = 0; j < n; ++j) vec(j + 1, i + 1) = a[(size_t)i * n + j];
    DiagonalMatrix weights(n);
    for (int i = 0; i < n; ++i) weights.element(i, i) = w[i] < 1.e-14 ? 0.0 : w[i];   // rotationmat.C:274-276
    eigs[q] = weights;
  }
}
