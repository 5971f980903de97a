// Host-callable launch wrappers around the sm_100a kernels (kernels.cu).  Everything is asynchronous on `stream`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_desc.h"

namespace b2d {

struct Bases;   // gemm_grouped.cuh

struct DevBatch {
  const GSeg* segs = nullptr;
  const GGroup* groups = nullptr;
  const GTile* tiles[B2D_NUM_TILE_CLASSES] = {};
  int ntiles[B2D_NUM_TILE_CLASSES] = {};
  bool unit_alpha = false;   // every segment has alpha == 1: the kernel variant without the operand scaling is used
};

constexpr int L1_MAX_VECS = 32;      // vectors per multi-vector level-1 kernel
constexpr int L1_MAX_BLOCKS = 1184;  // 148 SMs x 8 CTAs: upper bound of the level-1 grids (partials buffer rows)

struct VecList {
  double* p[L1_MAX_VECS];
};

cudaError_t gemm_init();   // opt in to > 48 KB dynamic shared memory
// base pointers: p[B2D_BASE_SRC], p[B2D_BASE_WORK], p[B2D_BASE_DST], p[B2D_BASE_AUX]
cudaError_t launch_gemm_batch(const DevBatch& b, double* const* bases, cudaStream_t stream, int64_t* launches);
// one tile class of the batch only (profiling: events around each class)
// tile_counter != nullptr (one int of device memory, private to the stream): the 128 x 128 class runs as a persistent kernel
cudaError_t launch_gemm_class(const DevBatch& b, int cls, double* const* bases, cudaStream_t stream, int64_t* launches,
                              unsigned long long* trace_slot = nullptr, int* tile_counter = nullptr);

// flat reference layout <-> padded device layout (Wavefunction::CollectFrom / FlattenInto; operator upload)
cudaError_t launch_pack(const BlockDesc* blocks, int nblocks, const double* flat, double* dev, cudaStream_t s, int64_t* launches);
cudaError_t launch_unpack(const BlockDesc* blocks, int nblocks, const double* dev, double* flat, cudaStream_t s, int64_t* launches);

// level-1 (Davidson) kernels; `partials` is scratch of L1_MAX_BLOCKS * L1_MAX_VECS doubles; results land in `out` on the device
cudaError_t launch_multi_dot(int K, const VecList& x, const double* y, int64_t n, double* partials, double* out, cudaStream_t s, int64_t* launches);
// in place: x_j <- sum_i alpha[i * lda + j] x_i   (j < n_out), alpha on the device
cudaError_t launch_rotate(int n_in, int n_out, const VecList& x, const double* alpha, int lda, int64_t n, cudaStream_t s, int64_t* launches);
// r = sigma - theta[0] b ; out[0] = r.r
cudaError_t launch_residual(const double* sigma, const double* b, const double* theta, double* r, int64_t n, double* partials, double* out, cudaStream_t s, int64_t* launches);
// Olsen preconditioner (linear.C:27-60): r <- P(theta)( r - <P c0|r>/<c0|P c0> c0 ),  P = 1/(theta - diag) where |theta - diag| > 1e-12
cudaError_t launch_olsen(double* r, const double* c0, const double* diag, const double* theta, int64_t n, double* partials, double* scratch2, cudaStream_t s, int64_t* launches);
// one modified Gram-Schmidt step (linear.C:358-366): r <- r/|r| ; r <- r - <r|b> b
cudaError_t launch_mgs_step(double* r, const double* b, int64_t n, double* partials, double* scratch2, cudaStream_t s, int64_t* launches);
// r <- r / |r|
cudaError_t launch_normalise(double* r, int64_t n, double* partials, double* scratch1, cudaStream_t s, int64_t* launches);
// r <- r / |r| if |r|^2 > tiny, else r <- 0  (noise wavefunctions O.psi, density.C:220-224: a vanishing one is skipped)
cudaError_t launch_normalise_guarded(double* r, int64_t n, double tiny, double* partials, double* scratch1, cudaStream_t s, int64_t* launches);
// out[0] = sum over sectors of the trace of the d x d block at buf + dev_off (leading dimension ld)
cudaError_t launch_trace(const BlockDesc* sectors, int nsectors, const double* buf, double* out, cudaStream_t s, int64_t* launches);
// y += mult * coef[0] * x  (coef on the device; coef == nullptr means 1)
cudaError_t launch_axpy(double* y, const double* x, const double* coef, double mult, int64_t n, cudaStream_t s, int64_t* launches);
cudaError_t launch_scale(double* x, double a, int64_t n, cudaStream_t s, int64_t* launches);
// dst += parts[0] + parts[1] + ... (nparts copies, `stride` doubles apart), summed in that fixed order (split-K epilogue)
cudaError_t launch_sum_parts(double* dst, const double* parts, int nparts, int64_t stride, int64_t n, cudaStream_t s, int64_t* launches);

// symmetric eigenproblem of the Davidson subspace matrix (n <= 32): two-sided cyclic Jacobi in one warp.
// G: n x ldg (upper triangle G[j][i], i >= j, valid; mirrored inside); theta[n] ascending; alpha[i*ldg + j] = component i of eigenvector j
cudaError_t launch_subspace_eig(const double* G, int n, int ldg, double* theta, double* alpha, cudaStream_t s, int64_t* launches);

// enlarged-block operator construction (operatorfunctions::TensorProduct / TensorTrace, operatorfunctions.C:19-254 -> MatrixTensorProduct
// MatrixBLAS.C:125-200): dst[row0 + ia * b_rows + ib][col0 + ja * b_cols + jb] += coef * opA(ia, ja) * opB(ib, jb), one task per
// (destination block, uncollected row piece, uncollected column piece).  HBM-bound scatter: A and B sub-blocks read once, dst written once.
struct KronTask {
  int64_t a, b;        // absolute byte addresses of the STORED blocks; 0 = identity of size a_rows / b_rows (TensorTrace)
  int64_t dst;         // absolute byte address of the destination block
  double coef;
  int32_t a_rows, a_cols, lda, a_t;   // op(A) is a_rows x a_cols; a_t: stored transposed (a_cols x a_rows, leading dimension lda)
  int32_t b_rows, b_cols, ldb, b_t;
  int32_t row0, col0, ldd, pad;
};
// work item of the scatter kernel: a band of KRON_BAND destination rows of task `task` (index into the task array of the launch)
constexpr int KRON_BAND = 32;
struct KronTile { int32_t task, band; };
cudaError_t launch_kron_scatter(const KronTask* tasks, const KronTile* tiles, int ntiles, cudaStream_t s, int64_t* launches);
// diagonals of operator sector blocks gathered into a compact pool (stride ld + 1 -> 1) before diag(H) reads them thousands of times
struct DiagGather {
  int64_t src;      // absolute byte address of the first diagonal element
  int64_t dst;      // offset (doubles) in the pool
  int32_t n, stride;
};
cudaError_t launch_gather_diag(const DiagGather* items, int nitems, double* pool, cudaStream_t s, int64_t* launches);
// diag(H): e[block p][i, j] += sum_tasks f a_i b_j
cudaError_t launch_diag(const BlockDesc* blocks, int nblocks, const DiagTask* tasks, const int* block_begin, double* e, cudaStream_t s, int64_t* launches);

// per-sector eigen-decomposition of the density matrix: one-sided (Hestenes) Jacobi on the rows, one CTA per sector.
// sectors[q] = {rows=cols=d_q, ld, dev_off = offset of G_q / Vt_q in g / vt, ref_off = offset of the eigenvalues}
// On exit: rows of vt are the eigenvectors, evals[ref_off + i] the eigenvalue of row i (unsorted), sweeps[q] the sweep count.
cudaError_t launch_sector_eig(const BlockDesc* sectors, int nsectors, double* g, double* vt, double* evals, int* sweeps, cudaStream_t s, int64_t* launches);
// LARGE sectors: block one-sided Jacobi (eig_block_jacobi.cuh).  init: V = identity for the listed sectors; step: one launch = one step of
// the round-robin tournament over 32-row blocks, one CTA per block pair, all sectors together.
struct BJPair;
cudaError_t block_jacobi_setup();
cudaError_t launch_block_jacobi_init(const BJPair* sectors, int nsectors, double* vt, cudaStream_t s, int64_t* launches);
cudaError_t launch_block_jacobi_step(const BJPair* pairs, int npairs, double* g, double* vt, const int* active, int* rotated, double tol, cudaStream_t s,
                                     int64_t* launches);
// evals[ref_off + i] = vt_i . g_i (Rayleigh quotients of the rows of vt, g = vt * rho): eigenvalues of library eigenvectors to the
// absolute accuracy of one FP64 matrix product
cudaError_t launch_rayleigh(const BlockDesc* sectors, int nsectors, const double* g, const double* vt, double* evals, cudaStream_t s, int64_t* launches);
// U_q[:, c] = vt_q[src_row[c], :]   (gather the kept eigenvectors as columns, selection order)
struct GatherDesc {
  int64_t vt_off, u_off;
  int32_t d, ld_vt, ncols, ld_u, row_begin, pad;
};
cudaError_t launch_gather_rotation(const GatherDesc* desc, int nsectors, const int* src_rows, const double* vt, double* u, cudaStream_t s, int64_t* launches);

cudaError_t launch_fill_random(double* dst, const BlockDesc* blocks, int nblocks, uint64_t seed, double amplitude, cudaStream_t s, int64_t* launches);
// make an operator self-adjoint in the reduced sense: for block pairs (ij, ji): A_ij <- (A_ij + f_ij * A_ji^T) / 2 ...
struct SymPair {
  int64_t off_a, off_b;   // (i,j) block and (j,i) block (doubles from the operator base); off_a == off_b for diagonal blocks
  int32_t rows, cols, ld_a, ld_b;
  double f;               // A_ji = f * A_ij^T
};
cudaError_t launch_symmetrise(double* base, const SymPair* pairs, int npairs, cudaStream_t s, int64_t* launches);

// FP64 yardsticks: register-resident DMMA / DFMA loops, returns FLOP counts and fills ms
cudaError_t measure_fp64(cudaStream_t s, double* dmma_tflops, double* dfma_tflops);

}  // namespace b2d
