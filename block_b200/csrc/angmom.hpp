// Host-side SU(2) recoupling scalars of the sigma path.  All spin arguments are the integers 2S.
//
// What the reference computes (file:line under the reference root):
//   Clebsch-Gordan   new_anglib.C:135-206   (Racah's closed form)
//   6j               new_anglib.C:76-116
//   9j               new_anglib.C:19-72     (sum over k of three 6j)
//   ninej            couplingCoeffs.h:88-97 / couplingCoeffs.C:54-68  sqrt((jg+1)(jh+1)(jc+1)(jf+1)) {9j}
//   getCommuteParity BaseOperator.C:20-53
//   Transposeview::get_scaling  BaseOperator.C:56-91
//   TensorOp::getTransposeFactorDD  tensor_operator.h:164-190
// Abelian point groups only: every spatial factor is 1 (Symmetry.C:534-538).
//
// These are closed-form sums of factorial ratios; this file evaluates the same closed forms from a factorial table,
// memoised per argument tuple because the schedule builder asks for the same few hundred tuples millions of times.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <unordered_map>

namespace b2d {

struct AngMom {
  double fact[171];
  std::unordered_map<uint64_t, double> memo9, memo6, memocg, memots;
  AngMom() {
    fact[0] = 1.0;
    for (int i = 1; i <= 170; ++i) fact[i] = fact[i - 1] * i;
  }
  double f(int n) const {
    if (n < 0 || n > 170) throw std::runtime_error("angmom: factorial argument out of range");
    return fact[n];
  }
  static uint64_t key(std::initializer_list<int> v) {
    uint64_t k = 1469598103934665603ull;
    for (int x : v) { k ^= (uint64_t)(x + 1024); k *= 1099511628211ull; }
    return k;
  }

  // <j1 m1 j2 m2 | j3 m3>
  double clebsch(int j1, int m1, int j2, int m2, int j3, int m3) {
    if (j1 < 0 || j2 < 0 || j3 < 0 || std::abs(m1) > j1 || std::abs(m2) > j2 || std::abs(m3) > j3) return 0.0;
    if (j1 + j2 < j3 || std::abs(j1 - j2) > j3 || m1 + m2 != m3) return 0.0;
    if (((j1 + m1) & 1) || ((j2 + m2) & 1) || ((j3 + m3) & 1) || ((j1 + j2 + j3) & 1)) return 0.0;
    uint64_t k = key({j1, m1, j2, m2, j3, m3});
    auto it = memocg.find(k);
    if (it != memocg.end()) return it->second;
    auto h = [](int x) { return x / 2; };   // arguments are even by the checks above
    double pref = std::sqrt((j3 + 1) * f(h(j1 + j2 - j3)) * f(h(j1 - j2 + j3)) * f(h(-j1 + j2 + j3)) / f(h(j1 + j2 + j3) + 1));
    pref *= std::sqrt(f(h(j1 + m1)) * f(h(j1 - m1)) * f(h(j2 + m2)) * f(h(j2 - m2)) * f(h(j3 + m3)) * f(h(j3 - m3)));
    int kmin = std::max(0, std::max(h(j2 - j3 - m1), h(j1 - j3 + m2)));
    int kmax = std::min(h(j1 + j2 - j3), std::min(h(j1 - m1), h(j2 + m2)));
    double s = 0.0;
    for (int q = kmin; q <= kmax; ++q) {
      double term = 1.0 / (f(q) * f(h(j1 + j2 - j3) - q) * f(h(j1 - m1) - q) * f(h(j2 + m2) - q) * f(h(j3 - j2 + m1) + q) * f(h(j3 - j1 - m2) + q));
      s += (q & 1) ? -term : term;
    }
    double r = pref * s;
    memocg.emplace(k, r);
    return r;
  }

  double tri(int a, int b, int c) const { return f((a + b - c) / 2) * f((a - b + c) / 2) * f((-a + b + c) / 2) / f((a + b + c) / 2 + 1); }
  static bool bad_triad(int x, int y, int z) { return ((x + y + z) & 1) || x + y < z || std::abs(x - y) > z; }

  double six_j(int a, int b, int c, int d, int e, int ff) {
    if (bad_triad(a, b, c) || bad_triad(c, d, e) || bad_triad(a, e, ff) || bad_triad(b, d, ff)) return 0.0;
    uint64_t k = key({a, b, c, d, e, ff});
    auto it = memo6.find(k);
    if (it != memo6.end()) return it->second;
    double pref = std::sqrt(tri(a, b, c) * tri(c, d, e) * tri(a, e, ff) * tri(b, d, ff));
    int t1 = (a + b + c) / 2, t2 = (c + d + e) / 2, t3 = (a + e + ff) / 2, t4 = (b + d + ff) / 2;
    int u1 = (a + b + d + e) / 2, u2 = (a + c + d + ff) / 2, u3 = (b + c + e + ff) / 2;
    int tmin = std::max(std::max(t1, t2), std::max(t3, t4));
    int tmax = std::min(u1, std::min(u2, u3));
    double s = 0.0;
    for (int t = tmin; t <= tmax; ++t) {
      double term = f(t + 1) / (f(t - t1) * f(t - t2) * f(t - t3) * f(t - t4) * f(u1 - t) * f(u2 - t) * f(u3 - t));
      s += (t & 1) ? -term : term;
    }
    double r = pref * s;
    memo6.emplace(k, r);
    return r;
  }

  double nine_j(int a, int b, int c, int d, int e, int ff, int g, int hh, int i) {
    auto tri_bad = [](int x, int y, int z) { return x + y < z || std::abs(x - y) > z; };
    if (tri_bad(a, b, c) || tri_bad(d, e, ff) || tri_bad(g, hh, i) || tri_bad(a, d, g) || tri_bad(b, e, hh) || tri_bad(c, ff, i)) return 0.0;
    int kmin = std::max(std::abs(hh - d), std::max(std::abs(b - ff), std::abs(a - i)));
    int kmax = std::min(hh + d, std::min(b + ff, a + i));
    double s = 0.0;
    for (int k = kmin; k <= kmax; ++k) {
      double term = (k + 1) * six_j(a, b, c, ff, i, k) * six_j(d, e, ff, b, k, hh) * six_j(g, hh, i, k, a, d);
      s += (k & 1) ? -term : term;
    }
    return s;
  }

  // ninejCoeffs::operator() / Ninej: the normalised recoupling coefficient TensorMultiply multiplies in
  double ninej(int ja, int jb, int jc, int jd, int je, int jf, int jg, int jh, int ji) {
    uint64_t k = key({ja, jb, jc, jd, je, jf, jg, jh, ji});
    auto it = memo9.find(k);
    if (it != memo9.end()) return it->second;
    double r = std::sqrt((double)(jg + 1) * (jh + 1) * (jc + 1) * (jf + 1)) * nine_j(ja, jb, jc, jd, je, jf, jg, jh, ji);
    memo9.emplace(k, r);
    return r;
  }

  static constexpr double NUMERICAL_ZERO = 1e-15;   // dmrg.C:85

  // getCommuteParity(a, b, c): a, b, c are (N, 2S, irrep)
  double commute_parity(const int* a, const int* b, const int* c) {
    double parity = ((a[0] & 1) && (b[0] & 1)) ? -1.0 : 1.0;
    for (int asz = -a[1]; asz <= a[1]; asz += 2)
      for (int bsz = -b[1]; bsz <= b[1]; bsz += 2) {
        double cleb = clebsch(a[1], asz, b[1], bsz, c[1], c[1]);
        if (std::fabs(cleb) <= NUMERICAL_ZERO) continue;
        return parity * cleb / clebsch(b[1], bsz, a[1], asz, c[1], c[1]);
      }
    throw std::runtime_error("getCommuteParity: inappropriate operator quanta");
  }

  // Transposeview::get_scaling(leftq, rightq) for an operator of spin cs (conjugacy 't')
  double transpose_scaling(int cs, int ls, int rs) {
    const uint64_t k = key({cs, ls, rs});
    auto it = memots.find(k);
    if (it != memots.end()) return it->second;
    for (int lsz = -ls; lsz <= ls; lsz += 2)
      for (int rsz = -rs; rsz <= rs; rsz += 2) {
        double cleb = clebsch(ls, lsz, cs, -cs, rs, rsz);
        if (std::fabs(cleb) <= NUMERICAL_ZERO) continue;
        const double r = ((cs & 1) ? -1.0 : 1.0) * cleb / clebsch(rs, rsz, cs, cs, ls, lsz);
        memots.emplace(k, r);
        return r;
      }
    throw std::runtime_error("Transposeview::get_scaling: inappropriate sector quanta");
  }

  static double transpose_factor_dd(int pspin) { return pspin == 0 ? -1.0 : 1.0; }
};

// SpinQuantum::allow (SpinQuantum.C:99-107): q in q1 (+) q2
inline bool qn_allow(const int* q, const int* q1, const int* q2) {
  if (q[0] != q1[0] + q2[0] || q[2] != (q1[2] ^ q2[2])) return false;
  return std::abs(q1[1] - q2[1]) <= q[1] && q[1] <= q1[1] + q2[1] && ((q1[1] + q2[1] - q[1]) & 1) == 0;
}

}  // namespace b2d
