// POD descriptors shared by the host planner and the device kernels.
#pragma once
#include <stdint.h>

#define B2D_NUM_TILE_CLASSES 10  // {128,64,32} x {128,64,32}: class = 3 * (row size index) + (column size index); class 9 = tiny
#define B2D_TINY_CLASS 9         // output blocks of at most B2D_TINY_DIM x B2D_TINY_DIM: one WARP per block, lanes split K, shuffle reduction
#define B2D_TINY_DIM 8
#define B2D_BASE_ABS 0    // pointer field is an absolute device address (operator arenas)
#define B2D_BASE_SRC 1    // offset (doubles) into the source wavefunction
#define B2D_BASE_WORK 2   // offset into the T workspace
#define B2D_BASE_DST 3    // offset into the destination wavefunction
#define B2D_BASE_AUX 4    // offset into an auxiliary buffer (density matrix, rotation matrices, ...)
#define B2D_NUM_BASES 5

#if defined(__CUDACC__)
#define B2D_HD __host__ __device__
#else
#define B2D_HD
#endif

B2D_HD inline int b2d_tile_m(int cls) { return cls == B2D_TINY_CLASS ? B2D_TINY_DIM : 128 >> (cls / 3); }
B2D_HD inline int b2d_tile_n(int cls) { return cls == B2D_TINY_CLASS ? B2D_TINY_DIM : 128 >> (cls % 3); }

// One K-segment of a grouped contraction:  C += alpha * op(A) * op(B),  op(A) is m x k, op(B) is k x n.
//   a_trans  = 0: A stored m x k row-major (K contiguous);  1: stored k x m row-major (M contiguous)
//   b_kmajor = 1: B stored n x k row-major (K contiguous);  0: stored k x n row-major (N contiguous)
struct GSeg {
  int64_t a, b;        // absolute byte address (base ABS) or offset in doubles from the base
  double alpha;
  int32_t lda, ldb;    // leading dimensions of the STORED blocks (doubles, even)
  int32_t k;
  uint8_t a_base, b_base, a_trans, b_kmajor;
};

// One output block C (m x n, row-major, leading dimension ldc) and its K-segments [seg_begin, seg_end)
struct GGroup {
  int64_t c;
  int32_t ldc, m, n;
  int32_t seg_begin, seg_end;
  int32_t kiters;      // sum over segments of ceil(k / 16): pipeline iterations of one tile
  uint8_t c_base, accumulate, pad0, pad1;
};

// One CTA work item: a tile of one group's output
struct GTile {
  int32_t group, m0, n0, cost;
};

// diag(H) contribution to one psi block: e[i*dr + j] += f * a[i*sa] * b[j*sb]   (a == 0 / b == 0: factor 1)
struct DiagTask {
  int64_t a, b;   // absolute byte addresses of the first diagonal element, or 0
  double f;
  int32_t sa, sb; // diagonal strides (ld + 1)
};

// per-block table used by pack / unpack / diag / density kernels
struct BlockDesc {
  int64_t ref_off, dev_off;
  int32_t rows, cols, ld, pad;
};
