// Block one-sided Jacobi eigen-decomposition of the LARGE reduced-density-matrix sectors (replaces dsyev_ at rotationmat.C:268 /
// MatrixBLAS.C:381-414 for sectors beyond the single-CTA kernel, and the cusolverDnDsyevd call round 1 used there).
//
// State per sector (same as sector_eig_kernel): G (d x d, starts as rho_q) and V (starts as the identity), both row-major with leading
// dimension ld; the iteration applies the same orthogonal row transformations to both, G <- J G, V <- J V, until the rows of G are mutually
// orthogonal.  Then row i of V is an eigenvector and v_i . g_i its eigenvalue (G = V rho throughout).
//
// The rows are cut into blocks of BJ_B = 32.  One CTA owns one PAIR of blocks (64 rows) for one step:
//   1. Gram matrix A = X X^T of its 64 rows of G (streamed once through shared memory in 64-column chunks, FP64 FMA register tiles);
//   2. two-sided cyclic Jacobi on the 64 x 64 matrix A in shared memory, accumulating the rotations in W (parallel round-robin ordering:
//      32 disjoint rotations per round) - for a positive semi-definite A = D B D this is accurate relative to the row norms D (Demmel &
//      Veselic), which is what keeps the near-null space of rho at its own scale;
//   3. X <- W^T X for the 64 rows of G and of V (second streaming pass, in place: the pair owns its rows).
// A step pairs every block with one partner (round-robin tournament over the blocks of a sector), all sectors and all pairs of a step in ONE
// launch; nblocks - 1 steps make a sweep, after which every row pair of the sector has met once.  A sector has converged when a whole
// sweep applied no rotation.  Traffic per sweep: 5 x 8 d^2 bytes x (nblocks - 1) / ... = 32x less than the row-pair form (each row is
// re-read once per BLOCK it meets, not once per row), and the flops are dense 64 x 64 x d products.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2d {

constexpr int BJ_B = 32;             // rows per block
constexpr int BJ_R = 2 * BJ_B;       // rows per pair
constexpr int BJ_THREADS = 256;
constexpr int BJ_CH = 64;            // columns per streamed chunk
constexpr int BJ_INNER_SWEEPS = 2;
constexpr int BJ_LDX = BJ_CH + 1;    // shared-memory leading dimensions (odd: conflict-free column access)
constexpr int BJ_LDA = BJ_R + 1;
constexpr int BJ_LDW = BJ_R + 2;     // even: W rows are read as double2
constexpr size_t BJ_SMEM = (size_t)(BJ_R * BJ_LDX + BJ_R * BJ_LDA + BJ_R * BJ_LDW) * 8 + 32 * 4 * 8 + 64;

struct BJPair {
  int64_t off;        // offset (doubles) of the sector's G / V block
  int32_t d, ld;      // sector size, leading dimension
  int32_t i0, ni;     // first block: rows [i0, i0 + ni)
  int32_t j0, nj;     // second block: rows [j0, j0 + nj)   (nj = 0: the block sits out this step - odd block count - or the sector has one block)
  int32_t sector, pad;
};

#ifdef B2D_EIG_KERNELS   // kernels.cu only; ctx.cpp includes this file for the descriptors
// flags[sector] != 0 on entry: the sector is still active; rotated[sector] is set when this step applied a rotation
__global__ void __launch_bounds__(BJ_THREADS) block_jacobi_step_kernel(const BJPair* __restrict__ pairs, double* __restrict__ g, double* __restrict__ vt,
                                                                        const int* __restrict__ active, int* __restrict__ rotated, double tol) {
  extern __shared__ __align__(16) double bj_smem[];
  const BJPair pr = pairs[blockIdx.x];
  if (!active[pr.sector]) return;
  double* Xs = bj_smem;                              // [BJ_R][BJ_LDX]
  double* As = Xs + BJ_R * BJ_LDX;                   // [BJ_R][BJ_LDA]
  double* Ws = As + BJ_R * BJ_LDA;                   // [BJ_R][BJ_LDW]
  double* cs = Ws + BJ_R * BJ_LDW;                   // [32][4]: c, s, p, q of the round's rotations (p < 0: none)
  int* any_flag = reinterpret_cast<int*>(cs + 32 * 4);
  int* sig_flag = any_flag + 1;                      // a rotation well above the noise threshold happened: the sector has not converged
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = pr.d, ld = pr.ld;
  // the Gram entries are d-term dot products: their rounding noise grows like sqrt(d), and a threshold below it would rotate noise for ever
  tol *= fmax(1.0, 0.5 * sqrt((double)d));
  // Rows of G whose norm is below 1e-16 belong to eigenvalues below 1e-16 (rho has trace <= 1): two orders under the reference's clamp
  // (1e-14 -> 0, rotationmat.C:274) and three under its keep threshold (1e-13, :161) - they are never retained.  Their direction is pure
  // rounding noise of the density-matrix build, so "orthogonalising" them never converges; pairs that involve such a row are left alone
  // (V stays orthogonal whatever is skipped; the retained vectors pick up at most a 1e-16 / lambda rotation into the null space).
  const double floor2 = 1e-32;
  double* G = g + pr.off;
  double* V = vt + pr.off;
  auto row_of = [&](int r) -> int {   // local row -> row of the sector, or -1
    if (r < BJ_B) return r < pr.ni ? pr.i0 + r : -1;
    return (r - BJ_B) < pr.nj ? pr.j0 + (r - BJ_B) : -1;
  };
  auto load_chunk = [&](const double* M, int c0) {
    for (int e = tid; e < BJ_R * BJ_CH; e += BJ_THREADS) {
      const int r = e / BJ_CH, c = e % BJ_CH;
      const int gr = row_of(r);
      Xs[r * BJ_LDX + c] = (gr >= 0 && c0 + c < d) ? M[(int64_t)gr * ld + c0 + c] : 0.0;
    }
  };

  // ---- 1. Gram matrix of the 64 rows of G ----
  {
    const int ty = tid / 16, tx = tid % 16;          // 4 x 4 register tile of A per thread
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int c0 = 0; c0 < d; c0 += BJ_CH) {
      __syncthreads();
      load_chunk(G, c0);
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < BJ_CH; ++k) {
        double xa[4], xb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) xa[a] = Xs[(ty * 4 + a) * BJ_LDX + k];
#pragma unroll
        for (int b = 0; b < 4; ++b) xb[b] = Xs[(tx * 4 + b) * BJ_LDX + k];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fma(xa[a], xb[b], acc[a][b]);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) As[(ty * 4 + a) * BJ_LDA + tx * 4 + b] = acc[a][b];
  }
  for (int e = tid; e < BJ_R * BJ_R; e += BJ_THREADS) Ws[(e / BJ_R) * BJ_LDW + e % BJ_R] = (e / BJ_R == e % BJ_R) ? 1.0 : 0.0;
  if (tid == 0) { *any_flag = 0; *sig_flag = 0; }
  __syncthreads();

  // ---- 2. two-sided cyclic Jacobi on A (64 x 64), rotations accumulated in W ----
  int applied_total = 0;
  // at most BJ_INNER_SWEEPS sweeps over the 64 x 64 pivot block per step: the step is latency-bound (three CTA barriers per round, 63
  // rounds per sweep), and diagonalising the pivot block to convergence buys no outer sweep - the pairs meet again next sweep anyway
  for (int sweep = 0; sweep < BJ_INNER_SWEEPS; ++sweep) {
    int applied_sweep = 0;
    for (int round = 0; round < BJ_R - 1; ++round) {
      if (tid < 32) {
        int a = (round + tid) % (BJ_R - 1);
        int b = tid == 0 ? BJ_R - 1 : (round - tid + (BJ_R - 1)) % (BJ_R - 1);
        if (a > b) { int t = a; a = b; b = t; }
        const double app = As[a * BJ_LDA + a], aqq = As[b * BJ_LDA + b], apq = As[a * BJ_LDA + b];
        double c = 1.0, s = 0.0;
        int p = -1;
        if (app > floor2 && aqq > floor2 && fabs(apq) > tol * sqrt(app * aqq)) {
          const double zeta = (aqq - app) / (2.0 * apq);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          c = 1.0 / sqrt(1.0 + t * t);
          s = c * t;
          p = a;
        }
        cs[tid * 4 + 0] = c; cs[tid * 4 + 1] = s; cs[tid * 4 + 2] = (double)p; cs[tid * 4 + 3] = (double)b;
        const unsigned m = __ballot_sync(0xffffffffu, p >= 0);
        // rotations between tol and 4 tol are applied but do not keep the sector alive: at that level the fresh Gram entries of the next
        // sweep are rounding noise of the d-term dot products, and a whole sweep without ANY rotation would never happen
        // ... and only pairs of rows that are BOTH above the reference's keep threshold (|g| = eigenvalue > 1e-13, rotationmat.C:161) decide
        // convergence: rows below it can never be retained, their directions are dominated by the rounding noise of rho (relative error
        // 1e-16 / eigenvalue) and never settle; they are still rotated against everything every sweep
        const unsigned ms = __ballot_sync(0xffffffffu, p >= 0 && app > 1e-26 && aqq > 1e-26 && fabs(apq) > 4.0 * tol * sqrt(app * aqq));
        if (tid == 0) { *any_flag = m != 0; if (ms) *sig_flag = 1; }
      }
      __syncthreads();
      const int any = *any_flag;
      if (any) {
        // columns: A <- A J, W <- W J
        for (int e = tid; e < BJ_R * 32; e += BJ_THREADS) {
          const int i = e >> 5, k = e & 31;
          const int p = (int)cs[k * 4 + 2];
          if (p < 0) continue;
          const int q = (int)cs[k * 4 + 3];
          const double c = cs[k * 4], s = cs[k * 4 + 1];
          double x = As[i * BJ_LDA + p], y = As[i * BJ_LDA + q];
          As[i * BJ_LDA + p] = c * x - s * y; As[i * BJ_LDA + q] = s * x + c * y;
          x = Ws[i * BJ_LDW + p]; y = Ws[i * BJ_LDW + q];
          Ws[i * BJ_LDW + p] = c * x - s * y; Ws[i * BJ_LDW + q] = s * x + c * y;
        }
        __syncthreads();
        // rows: A <- J^T A
        for (int e = tid; e < BJ_R * 32; e += BJ_THREADS) {
          const int j = e & 63, k = e >> 6;
          const int p = (int)cs[k * 4 + 2];
          if (p < 0) continue;
          const int q = (int)cs[k * 4 + 3];
          const double c = cs[k * 4], s = cs[k * 4 + 1];
          const double x = As[p * BJ_LDA + j], y = As[q * BJ_LDA + j];
          As[p * BJ_LDA + j] = c * x - s * y; As[q * BJ_LDA + j] = s * x + c * y;
        }
        applied_sweep = 1;
      }
      __syncthreads();
    }
    if (!applied_sweep) break;
    applied_total = 1;
  }
  if (!applied_total) return;                        // the 64 rows were already mutually orthogonal: nothing to write
  if (tid == 0 && *sig_flag) rotated[pr.sector] = 1;

  // ---- 3. X <- W^T X for the rows of G and of V (new row r = sum_i W[i][r] x_i), in place ----
  for (int which = 0; which < 2; ++which) {
    double* M = which == 0 ? G : V;
    for (int c0 = 0; c0 < d; c0 += BJ_CH) {
      __syncthreads();
      load_chunk(M, c0);
      __syncthreads();
      double acc[8][2];
#pragma unroll
      for (int a = 0; a < 8; ++a) acc[a][0] = acc[a][1] = 0.0;
#pragma unroll 4
      for (int i = 0; i < BJ_R; ++i) {
        const double x0 = Xs[i * BJ_LDX + lane], x1 = Xs[i * BJ_LDX + lane + 32];
        const double2* wrow = reinterpret_cast<const double2*>(Ws + i * BJ_LDW + warp * 8);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const double2 w = wrow[a];
          acc[2 * a][0] = fma(w.x, x0, acc[2 * a][0]); acc[2 * a][1] = fma(w.x, x1, acc[2 * a][1]);
          acc[2 * a + 1][0] = fma(w.y, x0, acc[2 * a + 1][0]); acc[2 * a + 1][1] = fma(w.y, x1, acc[2 * a + 1][1]);
        }
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int gr = row_of(warp * 8 + a);
        if (gr < 0) continue;
        if (c0 + lane < d) M[(int64_t)gr * ld + c0 + lane] = acc[a][0];
        if (c0 + lane + 32 < d) M[(int64_t)gr * ld + c0 + lane + 32] = acc[a][1];
      }
    }
  }
}

// V = identity for the listed sectors (G is a copy of rho made by the caller)
__global__ void block_jacobi_init_kernel(const BJPair* __restrict__ sectors, double* __restrict__ vt) {
  const BJPair s = sectors[blockIdx.y];
  double* V = vt + s.off;
  const int64_t n = (int64_t)s.d * s.ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) V[e] = (e / s.ld == e % s.ld) ? 1.0 : 0.0;
}

#endif  // B2D_EIG_KERNELS

}  // namespace b2d
