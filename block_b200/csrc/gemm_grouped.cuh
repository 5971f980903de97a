// Grouped, variable-size FP64 contraction kernel for sm_100a.
//
// Every dense step of the DMRG hot path is an instance of
//        C_g (m_g x n_g)  (+)=  sum_{s in segs(g)}  alpha_s * op(A_s) (m_g x k_s) * op(B_s) (k_s x n_g)
// with thousands of ragged (m, n, k) per launch:
//   sigma step 1   T = A_L psi                            (one segment per group)          operatorfunctions.C:512-516
//   sigma step 2   sigma[lQ,rQ] += sum F T A_R^T          (hundreds of segments per group)  operatorfunctions.C:517-531
//   density        rho[q] = sum w psi[q,r] psi[q,r]^T                                      operatorfunctions.C:630-650
//   rotation       O' = U^T (O U)                                                          MatrixBLAS.C:553-572
// which on the CPU are one dgemm_ call each (MatrixBLAS.C:438-498).
//
// Design (B200):
//   * FP64 has no tcgen05 path: the tensor-pipe instruction for doubles is the warp-level DMMA
//     (mma.sync.aligned.m16n8k8.f64, four DMMA.8x8x4 in SASS), accumulators in registers.  Measured on B200: the
//     pipe retires one DMMA.8x8x4 per 16 clocks per SM sub-partition (37.1 TFLOP/s) and one warp per sub-partition
//     with >= 2 independent accumulators already saturates it - so the kernel is organised to keep FOUR warps per
//     sub-partition issuing DMMAs (16 warps on a 128x128 tile, 32x32 per warp, 64 accumulator registers) so that
//     fragment loads, address arithmetic and barriers of one warp hide under the DMMAs of the others.
//   * one CTA per output tile; nine tile classes {128,64,32} x {128,64,32}: a ragged sector is covered by 128-wide
//     bands plus ONE narrower band for the remainder, so the tensor pipe is not fed padding (tile lists are cost-sorted
//     on the host: in-order CTA dispatch ~ longest-processing-time-first over the 148 SMs).
//   * the K loop runs over ALL segments of the group back to back through one multi-stage cp.async (LDGSTS.128)
//     pipeline: a 5-row segment and a 900-row segment cost what their K says, no per-segment pipeline drain.  The
//     stages are handed over through mbarriers (cp.async.mbarrier.arrive for "data landed", one arrive per warp for
//     "slot free"), not __syncthreads: warps drift up to a stage apart instead of draining the pipe at a barrier.
//   * operands may be stored K-major or M/N-major (Transposeview operators are never materialised); shared-memory
//     tiles keep the global orientation, padded so the DMMA fragment reads are bank-conflict free in both; the
//     orientation is resolved ONCE per pipeline stage into one of four fully unrolled code paths with compile-time
//     strides (no per-fragment address multiplies).
//   * ragged edges are zero-filled by cp.async's src-size operand: no scalar tail loops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_desc.h"

namespace b2d {

struct Bases {
  double* p[B2D_NUM_BASES];
  unsigned long long* trace;   // diagnostic: {min CTA start, max CTA end} of this launch in %globaltimer ns, or nullptr
  int* tile_counter;           // persistent variant: next unclaimed tile of the launch (zeroed before the launch)
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_PAD = 4;
#ifndef B2D_SMALL_STAGES
#define B2D_SMALL_STAGES 3
#endif
#ifndef B2D_SMALL_PREFETCH
#define B2D_SMALL_PREFETCH 1
#endif

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
// 8-byte variant for operands that start on an odd element (a column piece of a wavefunction / T block at an odd offset: the
// factorised operators of an enlarged block address sub-blocks of psi, T and sigma): same zero-fill semantics, twice the copies
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
// mbarrier primitives (shared::cta).  full[stage]: data of a pipeline stage has landed (every thread's cp.async group
// arrives asynchronously); empty[stage]: every warp has finished reading the stage.
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
// arrive on b once all cp.async operations this thread has issued so far have completed (counts as an expected arrival)
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint64_t* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, int parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(b);
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  } while (!ok);
}

// D(16x8) += A(16x8) * B(8x8), FP64 tensor pipe.  Fragment layout (PTX ISA, mma.m16n8k8 .f64), g = lane/4, t = lane%4:
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);   b0 (k=t, n=g)  b1 (k=t+4, n=g);
//   c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int BM, int BN>
struct TileCfg {
  static constexpr int WARPS_M = BM / 32, WARPS_N = BN / 32;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int A_DOUBLES = (BM * (GEMM_BK + GEMM_PAD) > GEMM_BK * (BM + GEMM_PAD)) ? BM * (GEMM_BK + GEMM_PAD) : GEMM_BK * (BM + GEMM_PAD);
  static constexpr int B_DOUBLES = (BN * (GEMM_BK + GEMM_PAD) > GEMM_BK * (BN + GEMM_PAD)) ? BN * (GEMM_BK + GEMM_PAD) : GEMM_BK * (BN + GEMM_PAD);
  static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
  // 128x128: one 16-warp CTA per SM, 4 x 40 KB stages.  Narrower classes: fewer stages so that 2-3 CTAs share an SM and
  // one CTA's prologue / epilogue hides under the others' K loops.
  static constexpr bool BIG = (BM == 128 && BN == 128);
  static constexpr int STAGES = BIG ? 4 : B2D_SMALL_STAGES;
  static constexpr int PREFETCH = BIG ? 2 : B2D_SMALL_PREFETCH;   // stages in flight; STAGES - PREFETCH - 1 stages of slack between warps
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_DOUBLES * 8;
};

struct StageMeta {
  double alpha;
  int layout;   // bit0: A stored k x m (Transposeview / U^T), bit1: B stored n x k (K contiguous)
};

// Stage an (R x GEMM_BK) operand slab.  kmajor = true: global is [R][K] (K contiguous) -> smem [R][BK+PAD];
// kmajor = false: global is [K][R] (R contiguous) -> smem [BK][R+PAD].  r_total / k_total bound the ragged edge.
template <int R, int THREADS>
__device__ __forceinline__ void stage_operand(double* smem, const double* g, int ld, bool kmajor, int r0, int k0, int r_total, int k_total) {
  const int tid = threadIdx.x;
  const bool al16 = (reinterpret_cast<uintptr_t>(g) & 15) == 0;   // leading dimensions are even: the base address decides for every chunk
  if (kmajor) {
    constexpr int CPR = GEMM_BK / 2;           // 16-byte chunks per row
    constexpr int CHUNKS = R * CPR;
#pragma unroll
    for (int c = tid; c < CHUNKS; c += THREADS) {
      int row = c / CPR, kc = (c % CPR) * 2;
      int gr = r0 + row, gk = k0 + kc;
      int nv = (gr < r_total) ? min(max(k_total - gk, 0), 2) : 0;
      const double* src = nv > 0 ? g + (int64_t)gr * ld + gk : g;
      double* dst = smem + row * (GEMM_BK + GEMM_PAD) + kc;
      if (al16) cp_async16(dst, src, nv * 8);
      else { cp_async8(dst, src, nv > 0 ? 8 : 0); cp_async8(dst + 1, nv > 1 ? src + 1 : g, nv > 1 ? 8 : 0); }
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
      double* dst = smem + krow * (R + GEMM_PAD) + rc;
      if (al16) cp_async16(dst, src, nv * 8);
      else { cp_async8(dst, src, nv > 0 ? 8 : 0); cp_async8(dst + 1, nv > 1 ? src + 1 : g, nv > 1 ? 8 : 0); }
    }
  }
}

// One pipeline stage (GEMM_BK deep) of a warp's 32 x 32 sub-tile; orientation fixed at compile time.
template <int BM, int BN, bool AT, bool BKM, bool ALPHA>
__device__ __forceinline__ void mma_stage(const double* __restrict__ sA, const double* __restrict__ sB, double alpha, double (&acc)[2][4][4]) {
  constexpr int sAr = AT ? 1 : (GEMM_BK + GEMM_PAD), sAk = AT ? (BM + GEMM_PAD) : 1;
  constexpr int sBn = BKM ? (GEMM_BK + GEMM_PAD) : 1, sBk = BKM ? 1 : (BN + GEMM_PAD);
#pragma unroll
  for (int kk = 0; kk < GEMM_BK; kk += 8) {
    double bf[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bf[j][0] = sB[j * 8 * sBn + kk * sBk];
      bf[j][1] = sB[j * 8 * sBn + (kk + 4) * sBk];
      if (ALPHA) { bf[j][0] *= alpha; bf[j][1] *= alpha; }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double af[4];
      af[0] = sA[(i * 16) * sAr + kk * sAk];
      af[1] = sA[(i * 16 + 8) * sAr + kk * sAk];
      af[2] = sA[(i * 16) * sAr + (kk + 4) * sAk];
      af[3] = sA[(i * 16 + 8) * sAr + (kk + 4) * sAk];
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma_16x8x8(acc[i][j], af, bf[j]);
    }
  }
}

template <int BM, int BN, bool ALPHA>
__global__ void __launch_bounds__(TileCfg<BM, BN>::THREADS)
    grouped_gemm_kernel(const GSeg* __restrict__ segs, const GGroup* __restrict__ groups, const GTile* __restrict__ tiles, Bases bases) {
  using Cfg = TileCfg<BM, BN>;
  constexpr int THREADS = Cfg::THREADS;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(16) double smem[];
  __shared__ StageMeta meta[STAGES];

  if (bases.trace && threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    atomicMin(bases.trace, t0);
  }
  const GTile tile = tiles[blockIdx.x];
  const GGroup grp = groups[tile.group];
  const int m0 = tile.m0, n0 = tile.n0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;

  double acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0;

  // Pipeline: a ring of STAGES slots, loads run PREFETCH iterations ahead.  There is no CTA-wide barrier in the loop:
  // a warp waits on full[slot] for the data of its iteration and releases the slot through empty[slot]; the slot being
  // refilled at iteration `it` was consumed at iteration it + PREFETCH - STAGES, so a fast warp may run that far ahead of
  // the slowest one and its DMMAs cover the other warps' fragment loads, address arithmetic and load issue.
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], THREADS + 1); mbar_init(&empty_bar[s], THREADS / 32); }
  }
  __syncthreads();

  // producer cursor over (segment, k0); the segment descriptor stays in registers until the segment is exhausted
  int ps = grp.seg_begin, pk = 0;
  GSeg sg = segs[grp.seg_begin < grp.seg_end ? ps : 0];   // a group without segments (kiters == 0) stores zeros: rows of T no factor writes
  auto produce = [&](int stage) {
    const double* A = sg.a_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.a) : bases.p[sg.a_base] + sg.a;
    const double* B = sg.b_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.b) : bases.p[sg.b_base] + sg.b;
    double* sA = smem + stage * Cfg::STAGE_DOUBLES;
    double* sB = sA + Cfg::A_DOUBLES;
    stage_operand<BM, THREADS>(sA, A, sg.lda, sg.a_trans == 0, m0, pk, grp.m, sg.k);
    stage_operand<BN, THREADS>(sB, B, sg.ldb, sg.b_kmajor != 0, n0, pk, grp.n, sg.k);
    mbar_arrive_on_cp_async(&full_bar[stage]);
    if (threadIdx.x == 0) {
      meta[stage].alpha = sg.alpha;
      meta[stage].layout = (sg.a_trans ? 1 : 0) | (sg.b_kmajor ? 2 : 0);
      mbar_arrive(&full_bar[stage]);
    }
    pk += GEMM_BK;
    if (pk >= sg.k) {
      pk = 0;
      if (++ps < grp.seg_end) sg = segs[ps];
    }
  };

  const int total = grp.kiters;
#pragma unroll
  for (int p = 0; p < Cfg::PREFETCH; ++p)
    if (p < total) produce(p);

  // this lane's fragment origin inside a stage, per orientation
  const int ar = wm * 32 + g, br = wn * 32 + g;
  const int a_off_n = ar * (GEMM_BK + GEMM_PAD) + t, a_off_t = ar + t * (BM + GEMM_PAD);
  const int b_off_n = br + t * (BN + GEMM_PAD), b_off_k = br * (GEMM_BK + GEMM_PAD) + t;

  for (int it = 0; it < total; ++it) {
    const int p = it + Cfg::PREFETCH;
    if (p < total) {
      const int slot = p % STAGES;
      if (p >= STAGES) mbar_wait(&empty_bar[slot], ((p / STAGES) & 1) ^ 1);
      produce(slot);
    }
    const int stage = it % STAGES;
    mbar_wait(&full_bar[stage], (it / STAGES) & 1);
    const double* sA = smem + stage * Cfg::STAGE_DOUBLES;
    const double* sB = sA + Cfg::A_DOUBLES;
    const StageMeta mt = meta[stage];
    switch (mt.layout) {
      case 0: mma_stage<BM, BN, false, false, ALPHA>(sA + a_off_n, sB + b_off_n, mt.alpha, acc); break;
      case 1: mma_stage<BM, BN, true, false, ALPHA>(sA + a_off_t, sB + b_off_n, mt.alpha, acc); break;
      case 2: mma_stage<BM, BN, false, true, ALPHA>(sA + a_off_n, sB + b_off_k, mt.alpha, acc); break;
      default: mma_stage<BM, BN, true, true, ALPHA>(sA + a_off_t, sB + b_off_k, mt.alpha, acc); break;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
  }

  // epilogue: registers -> global (c0,c1 are adjacent columns: one 16-byte store per row pair)
  double* C = bases.p[grp.c_base] + grp.c;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int row = m0 + wm * 32 + i * 16 + g + h * 8;
        int col = n0 + wn * 32 + j * 8 + 2 * t;
        if (row >= grp.m || col >= grp.n) continue;
        double* dst = C + (int64_t)row * grp.ldc + col;
        double v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
        if (col + 1 < grp.n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          double2 o;
          if (grp.accumulate) { o = *reinterpret_cast<double2*>(dst); o.x += v0; o.y += v1; }
          else { o.x = v0; o.y = v1; }
          *reinterpret_cast<double2*>(dst) = o;
        } else {
          dst[0] = grp.accumulate ? dst[0] + v0 : v0;
          if (col + 1 < grp.n) dst[1] = grp.accumulate ? dst[1] + v1 : v1;   // output block at an odd column offset
        }
      }
  if (bases.trace && threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    atomicMax(bases.trace + 1, t1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent variant (used for the 128 x 128 class, one 16-warp CTA per SM): a CTA claims tiles from an atomic counter in the
// host's cost-sorted order and runs ONE pipeline across tile boundaries - the loads of the next tile's first stages are
// already in flight while the warps store the current tile's accumulators, so the per-tile prologue (first loads) and epilogue
// (stores) no longer leave the DMMA pipe idle.  Tile ids travel from thread 0 to the other threads through an 8-slot ring
// guarded by mbarriers (no CTA-wide barrier anywhere in the loop).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TILE_RING = 8;

template <int BM, int BN, bool ALPHA>
__global__ void __launch_bounds__(TileCfg<BM, BN>::THREADS)
    grouped_gemm_persistent_kernel(const GSeg* __restrict__ segs, const GGroup* __restrict__ groups, const GTile* __restrict__ tiles, int ntiles, Bases bases) {
  using Cfg = TileCfg<BM, BN>;
  constexpr int THREADS = Cfg::THREADS;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(16) double smem[];
  __shared__ StageMeta meta[STAGES];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tile_bar[TILE_RING];
  __shared__ int tile_ring[TILE_RING];

  if (bases.trace && threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    atomicMin(bases.trace, t0);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp / Cfg::WARPS_N, wn = warp % Cfg::WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], THREADS + 1); mbar_init(&empty_bar[s], THREADS / 32); }
#pragma unroll
    for (int s = 0; s < TILE_RING; ++s) mbar_init(&tile_bar[s], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) { tile_ring[0] = atomicAdd(bases.tile_counter, 1); mbar_arrive(&tile_bar[0]); }

  // ---- producer cursor: (tile, segment, k0); every thread stages its share of each operand slab ----
  int p_seq = 0;            // sequence number of the tile being produced
  bool p_valid = false;
  int p_m0 = 0, p_n0 = 0, p_m = 0, p_n = 0, p_left = 0, p_send = 0, ps = 0, pk = 0;
  GSeg sg;
  auto enter_tile = [&](int seq) {
    mbar_wait(&tile_bar[seq % TILE_RING], (seq / TILE_RING) & 1);
    const int id = tile_ring[seq % TILE_RING];
    p_valid = id < ntiles;
    if (threadIdx.x == 0 && p_valid) {   // publish the following tile now: its id is needed PREFETCH stages before this one ends
      tile_ring[(seq + 1) % TILE_RING] = atomicAdd(bases.tile_counter, 1);
      mbar_arrive(&tile_bar[(seq + 1) % TILE_RING]);
    }
    if (!p_valid) return;
    const GTile tl = tiles[id];
    const GGroup gr = groups[tl.group];
    p_m0 = tl.m0; p_n0 = tl.n0; p_m = gr.m; p_n = gr.n; p_left = gr.kiters; ps = gr.seg_begin; p_send = gr.seg_end; pk = 0;
    sg = segs[ps];
  };
  enter_tile(0);
  while (p_valid && p_left == 0) enter_tile(++p_seq);   // tiles of groups without segments need no stages (they only store zeros)
  int pg = 0;               // stages produced so far (over all tiles of this CTA)
  auto produce_one = [&]() {
    const int slot = pg % STAGES;
    if (pg >= STAGES) mbar_wait(&empty_bar[slot], ((pg / STAGES) & 1) ^ 1);
    const double* A = sg.a_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.a) : bases.p[sg.a_base] + sg.a;
    const double* B = sg.b_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.b) : bases.p[sg.b_base] + sg.b;
    double* sA = smem + slot * Cfg::STAGE_DOUBLES;
    double* sB = sA + Cfg::A_DOUBLES;
    stage_operand<BM, THREADS>(sA, A, sg.lda, sg.a_trans == 0, p_m0, pk, p_m, sg.k);
    stage_operand<BN, THREADS>(sB, B, sg.ldb, sg.b_kmajor != 0, p_n0, pk, p_n, sg.k);
    mbar_arrive_on_cp_async(&full_bar[slot]);
    if (threadIdx.x == 0) {
      meta[slot].alpha = sg.alpha;
      meta[slot].layout = (sg.a_trans ? 1 : 0) | (sg.b_kmajor ? 2 : 0);
      mbar_arrive(&full_bar[slot]);
    }
    ++pg;
    pk += GEMM_BK;
    if (pk >= sg.k) {
      pk = 0;
      if (++ps < p_send) sg = segs[ps];
    }
    if (--p_left == 0) {
      enter_tile(++p_seq);
      while (p_valid && p_left == 0) enter_tile(++p_seq);
    }
  };

  const int ar = wm * 32 + g, br = wn * 32 + g;
  const int a_off_n = ar * (GEMM_BK + GEMM_PAD) + t, a_off_t = ar + t * (BM + GEMM_PAD);
  const int b_off_n = br + t * (BN + GEMM_PAD), b_off_k = br * (GEMM_BK + GEMM_PAD) + t;

  int cg = 0;               // stages consumed so far
  for (int seq = 0;; ++seq) {
    mbar_wait(&tile_bar[seq % TILE_RING], (seq / TILE_RING) & 1);
    const int id = tile_ring[seq % TILE_RING];
    if (id >= ntiles) break;
    const GTile tile = tiles[id];
    const GGroup grp = groups[tile.group];
    const int m0 = tile.m0, n0 = tile.n0;
    double acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0;
    const int total = grp.kiters;
    for (int it = 0; it < total; ++it) {
      while (p_valid && pg <= cg + Cfg::PREFETCH) produce_one();   // keeps running into the NEXT tile near the end of this one
      const int stage = cg % STAGES;
      mbar_wait(&full_bar[stage], (cg / STAGES) & 1);
      const double* sA = smem + stage * Cfg::STAGE_DOUBLES;
      const double* sB = sA + Cfg::A_DOUBLES;
      const StageMeta mt = meta[stage];
      switch (mt.layout) {
        case 0: mma_stage<BM, BN, false, false, ALPHA>(sA + a_off_n, sB + b_off_n, mt.alpha, acc); break;
        case 1: mma_stage<BM, BN, true, false, ALPHA>(sA + a_off_t, sB + b_off_n, mt.alpha, acc); break;
        case 2: mma_stage<BM, BN, false, true, ALPHA>(sA + a_off_n, sB + b_off_k, mt.alpha, acc); break;
        default: mma_stage<BM, BN, true, true, ALPHA>(sA + a_off_t, sB + b_off_k, mt.alpha, acc); break;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      ++cg;
    }
    while (p_valid && pg <= cg + Cfg::PREFETCH) produce_one();     // the next tile's first stages are in flight during the stores
    double* C = bases.p[grp.c_base] + grp.c;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int row = m0 + wm * 32 + i * 16 + g + h * 8;
          int col = n0 + wn * 32 + j * 8 + 2 * t;
          if (row >= grp.m || col >= grp.n) continue;
          double* dst = C + (int64_t)row * grp.ldc + col;
          double v0 = acc[i][j][2 * h], v1 = acc[i][j][2 * h + 1];
          if (col + 1 < grp.n && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            double2 o;
            if (grp.accumulate) { o = *reinterpret_cast<double2*>(dst); o.x += v0; o.y += v1; }
            else { o.x = v0; o.y = v1; }
            *reinterpret_cast<double2*>(dst) = o;
          } else {
            dst[0] = grp.accumulate ? dst[0] + v0 : v0;
            if (col + 1 < grp.n) dst[1] = grp.accumulate ? dst[1] + v1 : v1;
          }
        }
  }
  if (bases.trace && threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    atomicMax(bases.trace + 1, t1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tiny sectors: output blocks of at most 8 x 8 (the one- to eight-state quantum-number sectors at the edges of the
// particle-number / spin distribution).  A 32 x 32 DMMA tile would be > 94 % padding and pay a shared-memory pipeline per
// segment; here ONE WARP owns the block: lanes split the K index of every segment (coalesced along K for K-contiguous
// operands), accumulate the full m x n product in registers with DFMA, and the 32 partial sums of each output element are
// combined by a fixed-order butterfly of warp shuffles (deterministic).  Four warps per CTA, one block each.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TINY_WARPS = 4;

__global__ void __launch_bounds__(TINY_WARPS * 32)
    tiny_gemm_kernel(const GSeg* __restrict__ segs, const GGroup* __restrict__ groups, const GTile* __restrict__ tiles, int ntiles, Bases bases) {
  const int w = blockIdx.x * TINY_WARPS + (threadIdx.x >> 5);
  if (bases.trace && threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    atomicMin(bases.trace, t0);
  }
  if (w < ntiles) {
    const int lane = threadIdx.x & 31;
    const GGroup grp = groups[tiles[w].group];
    const int m = grp.m, n = grp.n;
    double acc[B2D_TINY_DIM][B2D_TINY_DIM];
#pragma unroll
    for (int i = 0; i < B2D_TINY_DIM; ++i)
#pragma unroll
      for (int j = 0; j < B2D_TINY_DIM; ++j) acc[i][j] = 0.0;
    for (int s = grp.seg_begin; s < grp.seg_end; ++s) {
      const GSeg sg = segs[s];
      const double* A = sg.a_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.a) : bases.p[sg.a_base] + sg.a;
      const double* B = sg.b_base == B2D_BASE_ABS ? reinterpret_cast<const double*>(sg.b) : bases.p[sg.b_base] + sg.b;
      const int64_t a_rs = sg.a_trans ? 1 : sg.lda, a_ks = sg.a_trans ? sg.lda : 1;     // op(A)(i, k) = A[i * a_rs + k * a_ks]
      const int64_t b_ns = sg.b_kmajor ? sg.ldb : 1, b_ks = sg.b_kmajor ? 1 : sg.ldb;   // op(B)(k, j) = B[j * b_ns + k * b_ks]
      for (int k = lane; k < sg.k; k += 32) {
        double a[B2D_TINY_DIM], b[B2D_TINY_DIM];
#pragma unroll
        for (int i = 0; i < B2D_TINY_DIM; ++i) a[i] = i < m ? sg.alpha * A[i * a_rs + k * a_ks] : 0.0;
#pragma unroll
        for (int j = 0; j < B2D_TINY_DIM; ++j) b[j] = j < n ? B[j * b_ns + k * b_ks] : 0.0;
#pragma unroll
        for (int i = 0; i < B2D_TINY_DIM; ++i)
          if (i < m) {
#pragma unroll
            for (int j = 0; j < B2D_TINY_DIM; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
          }
      }
    }
    double* C = bases.p[grp.c_base] + grp.c;
#pragma unroll
    for (int i = 0; i < B2D_TINY_DIM; ++i)
#pragma unroll
      for (int j = 0; j < B2D_TINY_DIM; ++j) {
        if (i < m && j < n) {   // warp-uniform
          double v = acc[i][j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == ((i * B2D_TINY_DIM + j) & 31)) {
            double* dst = C + (int64_t)i * grp.ldc + j;
            *dst = grp.accumulate ? *dst + v : v;
          }
        }
      }
  }
  if (bases.trace && threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    atomicMax(bases.trace + 1, t1);
  }
}

}  // namespace b2d
