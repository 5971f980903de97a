// Grouped, variable-size FP64 contraction kernel for sm_100a.
//
// Every dense step of the DMRG hot path is an instance of
//        C_g (m_g x n_g)  (+)=  sum_{s in segs(g)}  alpha_s * op(A_s) (m_g x k_s) * op(B_s) (k_s x n_g)
// with thousands of ragged (m, n, k) per launch:
//   sigma step 1   T = s A_L psi                         (one segment per group)          operatorfunctions.C:512-516
//   sigma step 2   sigma[lQ,rQ] += sum F T A_R^T          (hundreds of segments per group)  operatorfunctions.C:517-531
//   density        rho[q] = sum w psi[q,r] psi[q,r]^T                                      operatorfunctions.C:630-650
//   rotation       O' = U^T (O U)                                                          MatrixBLAS.C:553-572
// which on the CPU are one dgemm_ call each (MatrixBLAS.C:438-498).
//
// Design (B200):
//   * FP64 has no tcgen05 path: the tensor-pipe instruction for doubles is the warp-level DMMA
//     mma.sync.aligned.m16n8k8.f64, accumulators in registers.
//   * one CTA per output tile (tile lists are cost-sorted on the host: in-order CTA dispatch ~ LPT over the SMs);
//     three tile classes 128x128 / 64x64 / 32x32 so that tiny quantum-number sectors do not pay for big tiles.
//   * the K loop runs over ALL segments of the group back to back through one multi-stage cp.async (LDGSTS.128)
//     pipeline: a 5-row segment and a 900-row segment cost what their K says, no per-segment pipeline drain.
//   * operands may be stored K-major or M/N-major (Transposeview operators are never materialised); shared-memory
//     tiles keep the global orientation and are padded so the DMMA fragment reads are bank-conflict free in both.
//   * ragged edges are zero-filled by cp.async's src-size operand: no scalar tail loops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_desc.h"

namespace b2d {

struct Bases {
  double* p[B2D_NUM_BASES];
};

constexpr int GEMM_BK = 16;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(16x8) += A(16x8) * B(8x8), FP64 tensor pipe.  Fragment layout (PTX ISA, mma.m16n8k8 .f64), g = lane/4, t = lane%4:
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);   b0 (k=t, n=g)  b1 (k=t+4, n=g);
//   c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
#ifndef B2D_MMA_884
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
#else
  // the same product out of four m8n8k4 DMMAs (sm_80 shape); kept as a cross-check of the fragment mapping
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a[2]), "d"(b[1]));
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[2]), "+d"(c[3]) : "d"(a[1]), "d"(b[0]));
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[2]), "+d"(c[3]) : "d"(a[3]), "d"(b[1]));
#endif
}

template <int BM, int BN>
struct TileSmem {
  static constexpr int A_DOUBLES = (BM * (GEMM_BK + 4) > GEMM_BK * (BM + 4)) ? BM * (GEMM_BK + 4) : GEMM_BK * (BM + 4);
  static constexpr int B_DOUBLES = (BN * (GEMM_BK + 4) > GEMM_BK * (BN + 4)) ? BN * (GEMM_BK + 4) : GEMM_BK * (BN + 4);
  static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
};

struct StageMeta {
  double alpha;
  int a_trans, b_kmajor;
};

// Stage an (R x GEMM_BK) operand slab.  kmajor = true: global is [R][K] (K contiguous) -> smem [R][BK+4];
// kmajor = false: global is [K][R] (R contiguous) -> smem [BK][R+4].  r_valid / k_valid bound the ragged edge.
template <int R, int THREADS>
__device__ __forceinline__ void stage_operand(double* smem, const double* g, int ld, bool kmajor, int r0, int k0, int r_total, int k_total) {
  const int tid = threadIdx.x;
  if (kmajor) {
    constexpr int CPR = GEMM_BK / 2;           // 16-byte chunks per row
    constexpr int CHUNKS = R * CPR;
#pragma unroll
    for (int c = tid; c < CHUNKS; c += THREADS) {
      int row = c / CPR, kc = (c % CPR) * 2;
      int gr = r0 + row, gk = k0 + kc;
      int nv = (gr < r_total) ? min(max(k_total - gk, 0), 2) : 0;
      const double* src = nv > 0 ? g + (int64_t)gr * ld + gk : g;
      cp_async16(smem + row * (GEMM_BK + 4) + kc, src, nv * 8);
    }
  } else {
    constexpr int CPR = R / 2;
    constexpr int CHUNKS = GEMM_BK * CPR;
#pragma unroll
    for (int c = tid; c < CHUNKS; c += THREADS) {
      int krow = c / CPR, rc = (c % CPR) * 2;
      int gk = k0 + krow, gr = r0 + rc;
      int nv = (gk < k_total) ? min(max(r_total - gr, 0), 2) : 0;
      const double* src = nv > 0 ? g + (int64_t)gk * ld + gr : g;
      cp_async16(smem + krow * (R + 4) + rc, src, nv * 8);
    }
  }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32)
    grouped_gemm_kernel(const GSeg* __restrict__ segs, const GGroup* __restrict__ groups, const GTile* __restrict__ tiles, Bases bases) {
  constexpr int THREADS = WARPS_M * WARPS_N * 32;
  constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  constexpr int MT = WM / 16, NT = WN / 8;
  using SM = TileSmem<BM, BN>;
  extern __shared__ __align__(16) double smem[];
  __shared__ StageMeta meta[STAGES];

  const GTile tile = tiles[blockIdx.x];
  const GGroup grp = groups[tile.group];
  const int m0 = tile.m0, n0 = tile.n0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;

  double acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0;

  // producer cursor over (segment, k0)
  int ps = grp.seg_begin, pk = 0;
  auto issue = [&](int stage) {
    if (ps < grp.seg_end) {
      const GSeg sg = segs[ps];
      const double* A = sg.a_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.a) : bases.p[sg.a_base] + sg.a;
      const double* B = sg.b_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.b) : bases.p[sg.b_base] + sg.b;
      double* sA = smem + stage * SM::STAGE_DOUBLES;
      double* sB = sA + SM::A_DOUBLES;
      stage_operand<BM, THREADS>(sA, A, sg.lda, sg.a_trans == 0, m0, pk, grp.m, sg.k);
      stage_operand<BN, THREADS>(sB, B, sg.ldb, sg.b_kmajor != 0, n0, pk, grp.n, sg.k);
      if (threadIdx.x == 0) {
        meta[stage].alpha = sg.alpha;
        meta[stage].a_trans = sg.a_trans;
        meta[stage].b_kmajor = sg.b_kmajor;
      }
      pk += GEMM_BK;
      if (pk >= sg.k) { pk = 0; ++ps; }
    }
    cp_async_commit();
  };

  const int total = grp.kiters;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  for (int it = 0; it < total; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue((it + STAGES - 1) % STAGES);
    const int stage = it % STAGES;
    const double* sA = smem + stage * SM::STAGE_DOUBLES;
    const double* sB = sA + SM::A_DOUBLES;
    const StageMeta mt = meta[stage];
    const int sAr = mt.a_trans ? 1 : (GEMM_BK + 4), sAk = mt.a_trans ? (BM + 4) : 1;
    const int sBn = mt.b_kmajor ? (GEMM_BK + 4) : 1, sBk = mt.b_kmajor ? 1 : (BN + 4);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; kk += 8) {
      double bf[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const double* pb = sB + (wn * WN + j * 8 + g) * sBn + (kk + t) * sBk;
        bf[j][0] = pb[0] * mt.alpha;
        bf[j][1] = pb[4 * sBk] * mt.alpha;
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        double af[4];
        const double* pa = sA + (wm * WM + i * 16 + g) * sAr + (kk + t) * sAk;
        af[0] = pa[0];
        af[1] = pa[8 * sAr];
        af[2] = pa[4 * sAk];
        af[3] = pa[8 * sAr + 4 * sAk];
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma_16x8x8(acc[i][j], af, bf[j]);
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: registers -> global (c0,c1 are adjacent columns: one 16-byte store per row pair)
  double* C = bases.p[grp.c_base] + grp.c;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int row = m0 + wm * WM + i * 16 + g + h * 8;
        int col = n0 + wn * WN + j * 8 + 2 * t;
        if (row >= grp.m || col >= grp.n) continue;
        double* dst = C + (int64_t)row * grp.ldc + col;
        double v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
        if (col + 1 < grp.n) {
          double2 o;
          if (grp.accumulate) { o = *reinterpret_cast<double2*>(dst); o.x += v0; o.y += v1; }
          else { o.x = v0; o.y = v1; }
          *reinterpret_cast<double2*>(dst) = o;
        } else {
          dst[0] = grp.accumulate ? dst[0] + v0 : v0;
        }
      }
}

}  // namespace b2d
