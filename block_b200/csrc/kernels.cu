// sm_100a kernels of the DMRG hot path other than the grouped contraction itself (gemm_grouped.cuh), plus the launch
// wrappers declared in kernels.h.  Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
#include "kernels.h"

#include <math.h>

#include <algorithm>
#include <vector>

#include "gemm_grouped.cuh"
#define B2D_EIG_KERNELS
#include "eig_block_jacobi.cuh"

namespace b2d {

#define B2D_LAUNCH_CHECK()                 \
  do {                                     \
    cudaError_t e_ = cudaGetLastError();   \
    if (e_ != cudaSuccess) return e_;      \
    if (launches) ++*launches;             \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// grouped contraction launches
// ---------------------------------------------------------------------------------------------------------------
template <int BM, int BN, bool ALPHA>
static cudaError_t init_one() {
  return cudaFuncSetAttribute(grouped_gemm_kernel<BM, BN, ALPHA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<BM, BN>::SMEM_BYTES);
}
template <int BM, int BN>
static cudaError_t init_pair() {
  cudaError_t e = init_one<BM, BN, true>();
  return e != cudaSuccess ? e : init_one<BM, BN, false>();
}
static int g_num_sms = 0;
static cudaError_t init_persistent() {
  cudaError_t e = cudaFuncSetAttribute(grouped_gemm_persistent_kernel<128, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<128, 128>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(grouped_gemm_persistent_kernel<128, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileCfg<128, 128>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  int dev = 0;
  e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  return cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
}

cudaError_t gemm_init() {   // opt in to > 48 KB dynamic shared memory
  cudaError_t e;
#define B2D_INIT(BM, BN) if ((e = init_pair<BM, BN>()) != cudaSuccess) return e;
  B2D_INIT(128, 128) B2D_INIT(128, 64) B2D_INIT(128, 32)
  B2D_INIT(64, 128) B2D_INIT(64, 64) B2D_INIT(64, 32)
  B2D_INIT(32, 128) B2D_INIT(32, 64) B2D_INIT(32, 32)
#undef B2D_INIT
  if ((e = block_jacobi_setup()) != cudaSuccess) return e;
  return init_persistent();
}

template <int BM, int BN>
static void launch_one(const DevBatch& b, int cls, const Bases& B, cudaStream_t stream) {
  using Cfg = TileCfg<BM, BN>;
  if (b.unit_alpha) grouped_gemm_kernel<BM, BN, false><<<b.ntiles[cls], Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(b.segs, b.groups, b.tiles[cls], B);
  else grouped_gemm_kernel<BM, BN, true><<<b.ntiles[cls], Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(b.segs, b.groups, b.tiles[cls], B);
}

cudaError_t launch_gemm_class(const DevBatch& b, int cls, double* const* bases, cudaStream_t stream, int64_t* launches,
                              unsigned long long* trace_slot, int* tile_counter) {
  if (cls < 0 || cls >= B2D_NUM_TILE_CLASSES || b.ntiles[cls] <= 0) return cudaSuccess;
  Bases B;
  for (int i = 0; i < B2D_NUM_BASES; ++i) B.p[i] = bases[i];
  B.trace = trace_slot;
  B.tile_counter = tile_counter;
  if (cls == 0 && tile_counter && g_num_sms > 0 && b.ntiles[0] > g_num_sms) {
    // persistent 128 x 128: one CTA per SM claims tiles from the counter; the pipeline runs across tile boundaries
    cudaError_t e = cudaMemsetAsync(tile_counter, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    using Cfg = TileCfg<128, 128>;
    if (b.unit_alpha) grouped_gemm_persistent_kernel<128, 128, false><<<g_num_sms, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(b.segs, b.groups, b.tiles[0], b.ntiles[0], B);
    else grouped_gemm_persistent_kernel<128, 128, true><<<g_num_sms, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(b.segs, b.groups, b.tiles[0], b.ntiles[0], B);
    B2D_LAUNCH_CHECK();
    return cudaSuccess;
  }
  switch (cls) {
    case 0: launch_one<128, 128>(b, cls, B, stream); break;
    case 1: launch_one<128, 64>(b, cls, B, stream); break;
    case 2: launch_one<128, 32>(b, cls, B, stream); break;
    case 3: launch_one<64, 128>(b, cls, B, stream); break;
    case 4: launch_one<64, 64>(b, cls, B, stream); break;
    case 5: launch_one<64, 32>(b, cls, B, stream); break;
    case 6: launch_one<32, 128>(b, cls, B, stream); break;
    case 7: launch_one<32, 64>(b, cls, B, stream); break;
    case 8: launch_one<32, 32>(b, cls, B, stream); break;
    default: tiny_gemm_kernel<<<(b.ntiles[cls] + TINY_WARPS - 1) / TINY_WARPS, TINY_WARPS * 32, 0, stream>>>(b.segs, b.groups, b.tiles[cls], b.ntiles[cls], B); break;
  }
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_gemm_batch(const DevBatch& b, double* const* bases, cudaStream_t stream, int64_t* launches) {
  for (int c = 0; c < B2D_NUM_TILE_CLASSES; ++c) {
    cudaError_t e = launch_gemm_class(b, c, bases, stream, launches, nullptr, nullptr);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// layout conversion
// ---------------------------------------------------------------------------------------------------------------
__global__ void pack_kernel(const BlockDesc* __restrict__ blocks, int nblocks, const double* __restrict__ flat, double* __restrict__ dev, int to_dev) {
  // one CTA strides over blocks; threads stride over elements (column index fastest: coalesced on both sides)
  for (int b = blockIdx.y; b < nblocks; b += gridDim.y) {
    const BlockDesc d = blocks[b];
    const int64_t n = (int64_t)d.rows * d.cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
      int r = (int)(e / d.cols), c = (int)(e % d.cols);
      if (to_dev) dev[d.dev_off + (int64_t)r * d.ld + c] = flat[d.ref_off + e];
      else ((double*)flat)[d.ref_off + e] = dev[d.dev_off + (int64_t)r * d.ld + c];
    }
  }
}

static dim3 pack_grid(int nblocks) { return dim3(8, (unsigned)min(nblocks, 4096), 1); }

cudaError_t launch_pack(const BlockDesc* blocks, int nblocks, const double* flat, double* dev, cudaStream_t s, int64_t* launches) {
  if (nblocks == 0) return cudaSuccess;
  pack_kernel<<<pack_grid(nblocks), 256, 0, s>>>(blocks, nblocks, flat, dev, 1);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t launch_unpack(const BlockDesc* blocks, int nblocks, const double* dev, double* flat, cudaStream_t s, int64_t* launches) {
  if (nblocks == 0) return cudaSuccess;
  pack_kernel<<<pack_grid(nblocks), 256, 0, s>>>(blocks, nblocks, flat, (double*)dev, 0);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// level-1 kernels.  Reductions are two-stage and ordered (per-CTA partials, then one CTA sums them in a fixed
// order), so results are bit-reproducible run to run and identical on every rank.
// ---------------------------------------------------------------------------------------------------------------
constexpr int L1_THREADS = 256;

static int l1_grid(int64_t n) {
  int64_t want = (n + L1_THREADS * 4 - 1) / (L1_THREADS * 4);
  if (want < 1) want = 1;
  if (want > L1_MAX_BLOCKS) want = L1_MAX_BLOCKS;
  return (int)want;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// sums `v` over the CTA; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sh /* 32 doubles */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

template <int K>
__global__ void __launch_bounds__(L1_THREADS) multi_dot_kernel(VecList x, const double* __restrict__ y, int64_t n, double* __restrict__ partials) {
  __shared__ double sh[32];
  double acc[K];
#pragma unroll
  for (int j = 0; j < K; ++j) acc[j] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double yv = y[i];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] += x.p[j][i] * yv;
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    double v = block_sum(acc[j], sh);
    if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * L1_MAX_VECS + j] = v;
  }
}

// out[j] = sum over blocks of partials[b][j], fixed order
__global__ void finish_kernel(const double* __restrict__ partials, int nblocks, int K, double* __restrict__ out) {
  __shared__ double sh[32];
  for (int j = 0; j < K; ++j) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v += partials[(int64_t)b * L1_MAX_VECS + j];
    v = block_sum(v, sh);
    if (threadIdx.x == 0) out[j] = v;
  }
}

cudaError_t launch_multi_dot(int K, const VecList& x, const double* y, int64_t n, double* partials, double* out, cudaStream_t s, int64_t* launches) {
  int done = 0;
  const int grid = l1_grid(n);
  while (done < K) {
    int k = K - done >= 8 ? 8 : (K - done >= 4 ? 4 : (K - done >= 2 ? 2 : 1));
    VecList sub;
    for (int j = 0; j < k; ++j) sub.p[j] = x.p[done + j];
    switch (k) {
      case 8: multi_dot_kernel<8><<<grid, L1_THREADS, 0, s>>>(sub, y, n, partials); break;
      case 4: multi_dot_kernel<4><<<grid, L1_THREADS, 0, s>>>(sub, y, n, partials); break;
      case 2: multi_dot_kernel<2><<<grid, L1_THREADS, 0, s>>>(sub, y, n, partials); break;
      default: multi_dot_kernel<1><<<grid, L1_THREADS, 0, s>>>(sub, y, n, partials); break;
    }
    B2D_LAUNCH_CHECK();
    finish_kernel<<<1, 256, 0, s>>>(partials, grid, k, out + done);
    B2D_LAUNCH_CHECK();
    done += k;
  }
  return cudaSuccess;
}

__global__ void __launch_bounds__(L1_THREADS) rotate_kernel(int n_in, int n_out, VecList x, const double* __restrict__ alpha, int lda, int64_t n) {
  __shared__ double a[L1_MAX_VECS * L1_MAX_VECS];
  for (int i = threadIdx.x; i < n_in * n_out; i += blockDim.x) a[i] = alpha[(i / n_out) * lda + (i % n_out)];
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    double v[L1_MAX_VECS];
#pragma unroll
    for (int i = 0; i < L1_MAX_VECS; ++i) v[i] = i < n_in ? x.p[i][e] : 0.0;
    for (int j = 0; j < n_out; ++j) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < L1_MAX_VECS; ++i)
        if (i < n_in) s += a[i * n_out + j] * v[i];
      x.p[j][e] = s;
    }
  }
}

// Square in-place rotation with the vector count as a template parameter: every thread owns TWO consecutive elements of all N
// vectors (16-byte loads and stores), the N x N coefficients sit in shared memory and each broadcast read feeds two DFMAs.
// HBM bound: N reads + N writes of the vector per launch.
template <int N>
__global__ void __launch_bounds__(L1_THREADS) rotate_square_kernel(VecList x, const double* __restrict__ alpha, int lda, int64_t n2) {
  __shared__ double a[N * N];        // a[j * N + i] = alpha(i, j): the coefficients of output j are contiguous
  __shared__ double2* ptr[N];
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) a[(i % N) * N + (i / N)] = alpha[(i / N) * lda + (i % N)];
  if (threadIdx.x < N) ptr[threadIdx.x] = reinterpret_cast<double2*>(x.p[threadIdx.x]);
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += (int64_t)gridDim.x * blockDim.x) {
    double2 v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = ptr[i][e];
#pragma unroll 1
    for (int j = 0; j < N; ++j) {     // not unrolled: N x N x 2 DFMAs would not fit an unrolled body for N = 32
      const double* aj = a + j * N;
      double2 s = make_double2(0.0, 0.0);
#pragma unroll
      for (int i = 0; i < N; ++i) { const double c = aj[i]; s.x = fma(c, v[i].x, s.x); s.y = fma(c, v[i].y, s.y); }
      ptr[j][e] = s;
    }
  }
}

template <int N>
static void launch_rotate_square(const VecList& x, const double* alpha, int lda, int64_t n, cudaStream_t s) {
  rotate_square_kernel<N><<<l1_grid(n / 2), L1_THREADS, 0, s>>>(x, alpha, lda, n / 2);
}

cudaError_t launch_rotate(int n_in, int n_out, const VecList& x, const double* alpha, int lda, int64_t n, cudaStream_t s, int64_t* launches) {
  bool aligned = (n % 2) == 0;
  for (int i = 0; i < n_in && aligned; ++i) aligned = (reinterpret_cast<uintptr_t>(x.p[i]) % 16) == 0;
  if (n_in == n_out && aligned && n_in >= 1 && n_in <= L1_MAX_VECS) {
    switch (n_in) {
#define B2D_ROT(N) case N: launch_rotate_square<N>(x, alpha, lda, n, s); break;
      B2D_ROT(1) B2D_ROT(2) B2D_ROT(3) B2D_ROT(4) B2D_ROT(5) B2D_ROT(6) B2D_ROT(7) B2D_ROT(8) B2D_ROT(9) B2D_ROT(10) B2D_ROT(11)
      B2D_ROT(12) B2D_ROT(13) B2D_ROT(14) B2D_ROT(15) B2D_ROT(16) B2D_ROT(17) B2D_ROT(18) B2D_ROT(19) B2D_ROT(20) B2D_ROT(21)
      B2D_ROT(22) B2D_ROT(23) B2D_ROT(24) B2D_ROT(25) B2D_ROT(26) B2D_ROT(27) B2D_ROT(28) B2D_ROT(29) B2D_ROT(30) B2D_ROT(31)
      B2D_ROT(32)
#undef B2D_ROT
    }
  } else {
    rotate_kernel<<<l1_grid(n), L1_THREADS, 0, s>>>(n_in, n_out, x, alpha, lda, n);
  }
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

__global__ void __launch_bounds__(L1_THREADS) residual_kernel(const double* __restrict__ sigma, const double* __restrict__ b, const double* __restrict__ theta,
                                                              double* __restrict__ r, int64_t n, double* __restrict__ partials) {
  __shared__ double sh[32];
  const double th = theta[0];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = sigma[i] - th * b[i];
    r[i] = v;
    acc += v * v;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * L1_MAX_VECS] = acc;
}

cudaError_t launch_residual(const double* sigma, const double* b, const double* theta, double* r, int64_t n, double* partials, double* out, cudaStream_t s, int64_t* launches) {
  const int grid = l1_grid(n);
  residual_kernel<<<grid, L1_THREADS, 0, s>>>(sigma, b, theta, r, n, partials);
  B2D_LAUNCH_CHECK();
  finish_kernel<<<1, 256, 0, s>>>(partials, grid, 1, out);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

__device__ __forceinline__ double precond(double v, double den) { return fabs(den) > 1e-12 ? v / den : v; }   // linear.C:35-40

__global__ void __launch_bounds__(L1_THREADS) olsen_dots_kernel(const double* __restrict__ r, const double* __restrict__ c0, const double* __restrict__ diag,
                                                                const double* __restrict__ theta, int64_t n, double* __restrict__ partials) {
  __shared__ double sh[32];
  const double th = theta[0];
  double d1 = 0.0, d2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double c = c0[i], cp = precond(c, th - diag[i]);
    d1 += cp * r[i];
    d2 += c * cp;
  }
  d1 = block_sum(d1, sh);
  if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * L1_MAX_VECS] = d1;
  d2 = block_sum(d2, sh);
  if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * L1_MAX_VECS + 1] = d2;
}

__global__ void __launch_bounds__(L1_THREADS) olsen_apply_kernel(double* __restrict__ r, const double* __restrict__ c0, const double* __restrict__ diag,
                                                                 const double* __restrict__ theta, const double* __restrict__ dots, int64_t n) {
  const double th = theta[0];
  const double ratio = dots[0] / dots[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    r[i] = precond(r[i] - ratio * c0[i], th - diag[i]);
}

cudaError_t launch_olsen(double* r, const double* c0, const double* diag, const double* theta, int64_t n, double* partials, double* scratch2, cudaStream_t s, int64_t* launches) {
  const int grid = l1_grid(n);
  olsen_dots_kernel<<<grid, L1_THREADS, 0, s>>>(r, c0, diag, theta, n, partials);
  B2D_LAUNCH_CHECK();
  finish_kernel<<<1, 256, 0, s>>>(partials, grid, 2, scratch2);
  B2D_LAUNCH_CHECK();
  olsen_apply_kernel<<<grid, L1_THREADS, 0, s>>>(r, c0, diag, theta, scratch2, n);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

__global__ void __launch_bounds__(L1_THREADS) mgs_dots_kernel(const double* __restrict__ r, const double* __restrict__ b, int64_t n, double* __restrict__ partials) {
  __shared__ double sh[32];
  double rr = 0.0, rb = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = r[i];
    rr += v * v;
    rb += v * (b ? b[i] : 0.0);
  }
  rr = block_sum(rr, sh);
  if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * L1_MAX_VECS] = rr;
  rb = block_sum(rb, sh);
  if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * L1_MAX_VECS + 1] = rb;
}

__global__ void __launch_bounds__(L1_THREADS) mgs_apply_kernel(double* __restrict__ r, const double* __restrict__ b, const double* __restrict__ dots, int64_t n) {
  const double inv = 1.0 / sqrt(dots[0]);
  const double coef = dots[1] * inv;   // <r/|r| | b>
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    r[i] = b ? r[i] * inv - coef * b[i] : r[i] * inv;
}

cudaError_t launch_mgs_step(double* r, const double* b, int64_t n, double* partials, double* scratch2, cudaStream_t s, int64_t* launches) {
  const int grid = l1_grid(n);
  mgs_dots_kernel<<<grid, L1_THREADS, 0, s>>>(r, b, n, partials);
  B2D_LAUNCH_CHECK();
  finish_kernel<<<1, 256, 0, s>>>(partials, grid, 2, scratch2);
  B2D_LAUNCH_CHECK();
  mgs_apply_kernel<<<grid, L1_THREADS, 0, s>>>(r, b, scratch2, n);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_normalise(double* r, int64_t n, double* partials, double* scratch1, cudaStream_t s, int64_t* launches) {
  return launch_mgs_step(r, nullptr, n, partials, scratch1, s, launches);
}

__global__ void __launch_bounds__(L1_THREADS) guarded_scale_kernel(double* __restrict__ r, const double* __restrict__ dots, double tiny, int64_t n) {
  const double nrm = dots[0];
  const double inv = fabs(nrm) > tiny ? 1.0 / sqrt(nrm) : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) r[i] *= inv;
}
cudaError_t launch_normalise_guarded(double* r, int64_t n, double tiny, double* partials, double* scratch1, cudaStream_t s, int64_t* launches) {
  const int grid = l1_grid(n);
  mgs_dots_kernel<<<grid, L1_THREADS, 0, s>>>(r, nullptr, n, partials);
  B2D_LAUNCH_CHECK();
  finish_kernel<<<1, 256, 0, s>>>(partials, grid, 1, scratch1);
  B2D_LAUNCH_CHECK();
  guarded_scale_kernel<<<grid, L1_THREADS, 0, s>>>(r, scratch1, tiny, n);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

__global__ void __launch_bounds__(256) trace_kernel(const BlockDesc* __restrict__ sectors, int nsectors, const double* __restrict__ buf, double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int q = 0; q < nsectors; ++q) {
    const BlockDesc d = sectors[q];
    for (int i = threadIdx.x; i < d.rows; i += blockDim.x) acc += buf[d.dev_off + (int64_t)i * d.ld + i];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = acc;
}
cudaError_t launch_trace(const BlockDesc* sectors, int nsectors, const double* buf, double* out, cudaStream_t s, int64_t* launches) {
  trace_kernel<<<1, 256, 0, s>>>(sectors, nsectors, buf, out);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

__global__ void __launch_bounds__(L1_THREADS) axpy_kernel(double* __restrict__ y, const double* __restrict__ x, const double* __restrict__ coef, double mult, int64_t n) {
  const double a = coef ? mult * coef[0] : mult;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}
cudaError_t launch_axpy(double* y, const double* x, const double* coef, double mult, int64_t n, cudaStream_t s, int64_t* launches) {
  axpy_kernel<<<l1_grid(n), L1_THREADS, 0, s>>>(y, x, coef, mult, n);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}
__global__ void __launch_bounds__(L1_THREADS) scale_kernel(double* __restrict__ x, double a, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= a;
}
cudaError_t launch_scale(double* x, double a, int64_t n, cudaStream_t s, int64_t* launches) {
  scale_kernel<<<l1_grid(n), L1_THREADS, 0, s>>>(x, a, n);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

__global__ void __launch_bounds__(L1_THREADS) sum_parts_kernel(double* __restrict__ dst, const double* __restrict__ parts, int nparts, int64_t stride, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = dst[i];
    for (int p = 0; p < nparts; ++p) v += parts[(int64_t)p * stride + i];
    dst[i] = v;
  }
}
cudaError_t launch_sum_parts(double* dst, const double* parts, int nparts, int64_t stride, int64_t n, cudaStream_t s, int64_t* launches) {
  if (nparts <= 0) return cudaSuccess;
  sum_parts_kernel<<<l1_grid(n), L1_THREADS, 0, s>>>(dst, parts, nparts, stride, n);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// Davidson subspace eigenproblem (replaces dsyev_ at linear.C:273): two-sided cyclic Jacobi, one warp, n <= 32
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) subspace_eig_kernel(const double* __restrict__ G, int n, int ldg, double* __restrict__ theta, double* __restrict__ alpha) {
  __shared__ double A[32][33], V[32][33];
  __shared__ int order[32];
  const int lane = threadIdx.x;
  for (int i = 0; i < n; ++i)
    if (lane < n) {
      A[i][lane] = lane >= i ? G[i * ldg + lane] : G[lane * ldg + i];   // G[j][i] = <b_i|sigma_j>, i >= j, is authoritative (linear.C:266-270)
      V[i][lane] = i == lane ? 1.0 : 0.0;
    }
  __syncwarp();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dg = 0.0;
    if (lane < n)
      for (int j = 0; j < n; ++j) {
        double v = A[lane][j];
        if (j == lane) dg += v * v; else off += v * v;
      }
    off = warp_sum(off); dg = warp_sum(dg);
    off = __shfl_sync(0xffffffffu, off, 0); dg = __shfl_sync(0xffffffffu, dg, 0);
    if (off <= 1e-32 * dg || off == 0.0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p][q];
        if (fabs(apq) > 1e-300) {
          const double app = A[p][p], aqq = A[q][q];
          const double zeta = (aqq - app) / (2.0 * apq);
          const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
          __syncwarp();
          if (lane < n) {   // columns p, q of A and V
            double akp = A[lane][p], akq = A[lane][q];
            A[lane][p] = c * akp - s * akq;
            A[lane][q] = s * akp + c * akq;
            double vkp = V[lane][p], vkq = V[lane][q];
            V[lane][p] = c * vkp - s * vkq;
            V[lane][q] = s * vkp + c * vkq;
          }
          __syncwarp();
          if (lane < n) {   // rows p, q of A
            double apk = A[p][lane], aqk = A[q][lane];
            A[p][lane] = c * apk - s * aqk;
            A[q][lane] = s * apk + c * aqk;
          }
          __syncwarp();
        }
      }
  }
  __syncwarp();
  if (lane == 0) {   // ascending order (dsyev convention), stable
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 1; i < n; ++i) {
      int o = order[i], j = i - 1;
      while (j >= 0 && A[order[j]][order[j]] > A[o][o]) { order[j + 1] = order[j]; --j; }
      order[j + 1] = o;
    }
  }
  __syncwarp();
  if (lane < n) {
    theta[lane] = A[order[lane]][order[lane]];
    for (int i = 0; i < n; ++i) alpha[i * ldg + lane] = V[i][order[lane]];
  }
}

cudaError_t launch_subspace_eig(const double* G, int n, int ldg, double* theta, double* alpha, cudaStream_t s, int64_t* launches) {
  if (n > 32) return cudaErrorInvalidValue;
  subspace_eig_kernel<<<1, 32, 0, s>>>(G, n, ldg, theta, alpha);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// diag(H)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) diag_kernel(const BlockDesc* __restrict__ blocks, const DiagTask* __restrict__ tasks, const int* __restrict__ block_begin,
                                                   double* __restrict__ e) {
  const BlockDesc d = blocks[blockIdx.y];
  const int t0 = block_begin[blockIdx.y], t1 = block_begin[blockIdx.y + 1];
  const int64_t n = (int64_t)d.rows * d.cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / d.cols), j = (int)(idx % d.cols);
    double acc = 0.0;
    for (int t = t0; t < t1; ++t) {
      const DiagTask k = tasks[t];
      double v = k.f;
      if (k.a) v *= reinterpret_cast<const double*>(k.a)[(int64_t)i * k.sa];
      if (k.b) v *= reinterpret_cast<const double*>(k.b)[(int64_t)j * k.sb];
      acc += v;
    }
    e[d.dev_off + (int64_t)i * d.ld + j] = acc;
  }
}

// c += coef (a x b) on sub-blocks (HBM-bound scatter).  Work item = a band of KRON_BAND destination rows of one task (`tiles`, built on the
// host: big and small pieces load-balance over the SMs); 8 warps, 4 rows each, lanes run along the destination columns (coalesced stores).
// The common case - B is the 1 x 1 block of a one-site dot or the identity on it - needs no index arithmetic per element: the
// destination is a scaled (possibly transposed) copy of the A block; a transposed A goes through a 32 x 33 shared-memory tile so that
// both the reads and the writes are coalesced.  k.pad bit 0: the destination piece is known to be zero (first contribution to freshly
// zero-filled storage): store instead of read-modify-write.  Tasks of one launch never overlap.

__global__ void __launch_bounds__(256) kron_scatter_kernel(const KronTask* __restrict__ tasks, const KronTile* __restrict__ tiles, int ntiles) {
  __shared__ double tile[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = blockIdx.x; w < ntiles; w += gridDim.x) {
    const KronTask k = tasks[tiles[w].task];
    const double* A = reinterpret_cast<const double*>(k.a);
    const double* B = reinterpret_cast<const double*>(k.b);
    double* D = reinterpret_cast<double*>(k.dst);
    const int rows = k.a_rows * k.b_rows, cols = k.a_cols * k.b_cols;
    const int r0 = tiles[w].band * KRON_BAND, r1 = min(r0 + KRON_BAND, rows);
    const bool store = (k.pad & 1) != 0;
    if (k.b_rows == 1 && k.b_cols == 1) {
      const double f = k.coef * (B ? B[0] : 1.0);
      if (!A) {                                   // identity (x) scalar: only the diagonal of the band
        for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x)
          if (r < cols) { double* d = D + (int64_t)(k.row0 + r) * k.ldd + k.col0 + r; *d = store ? f : *d + f; }
        // a stored piece must be fully defined: the off-diagonal part stays as it is (zero-filled storage) in both modes
      } else if (!k.a_t) {
        // each warp owns 4 rows of the band and walks the columns two at a time: 4 independent 16-byte loads (and, when the destination is
        // accumulated, 4 more) are in flight per lane before the first store - the kernel is latency-bound otherwise
        const int rb = r0 + warp * 4;
        const bool dst16 = ((k.col0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0) && ((k.ldd & 1) == 0);
        const bool src16 = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((k.lda & 1) == 0);
        const int cpair = cols & ~1;
        if (dst16 && src16) {
          for (int c = 2 * lane; c < cpair; c += 64) {
            double2 v[4], o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (rb + j < r1) v[j] = *reinterpret_cast<const double2*>(A + (int64_t)(rb + j) * k.lda + c);
            if (!store) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (rb + j < r1) o[j] = *reinterpret_cast<const double2*>(D + (int64_t)(k.row0 + rb + j) * k.ldd + k.col0 + c);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (rb + j < r1) {
                double2 w;
                w.x = store ? f * v[j].x : o[j].x + f * v[j].x;
                w.y = store ? f * v[j].y : o[j].y + f * v[j].y;
                *reinterpret_cast<double2*>(D + (int64_t)(k.row0 + rb + j) * k.ldd + k.col0 + c) = w;
              }
          }
          if ((cols & 1) && lane == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (rb + j < r1) {
                double* d = D + (int64_t)(k.row0 + rb + j) * k.ldd + k.col0 + cpair;
                const double v = f * A[(int64_t)(rb + j) * k.lda + cpair];
                *d = store ? v : *d + v;
              }
          }
        } else {
          for (int c = lane; c < cols; c += 32) {
            double v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (rb + j < r1) v[j] = A[(int64_t)(rb + j) * k.lda + c];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (rb + j < r1) {
                double* d = D + (int64_t)(k.row0 + rb + j) * k.ldd + k.col0 + c;
                *d = store ? f * v[j] : *d + f * v[j];
              }
          }
        }
      } else {                                    // op(A)(r, c) = A[c][r]: 32 x 32 tiles through shared memory
        for (int c0 = 0; c0 < cols; c0 += 32) {
          __syncthreads();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cc = c0 + warp + 8 * j, rr = r0 + lane;
            tile[warp + 8 * j][lane] = (cc < cols && rr < r1) ? A[(int64_t)cc * k.lda + rr] : 0.0;
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = r0 + warp + 8 * j, cc = c0 + lane;
            if (rr < r1 && cc < cols) {
              double* d = D + (int64_t)(k.row0 + rr) * k.ldd + k.col0 + cc;
              const double v = f * tile[lane][warp + 8 * j];
              *d = store ? v : *d + v;
            }
          }
        }
      }
      continue;
    }
    // general Kronecker product (a dot of several sites: warm-up blocks): row indices are split once per row
    for (int r = r0 + warp; r < r1; r += 8) {
      const int ia = r / k.b_rows, ib = r % k.b_rows;
      double* dst = D + (int64_t)(k.row0 + r) * k.ldd + k.col0;
      for (int c = lane; c < cols; c += 32) {
        const int ja = c / k.b_cols, jb = c % k.b_cols;
        double va, vb;
        if (A) va = k.a_t ? A[(int64_t)ja * k.lda + ia] : A[(int64_t)ia * k.lda + ja]; else va = ia == ja ? 1.0 : 0.0;
        if (B) vb = k.b_t ? B[(int64_t)jb * k.ldb + ib] : B[(int64_t)ib * k.ldb + jb]; else vb = ib == jb ? 1.0 : 0.0;
        const double v = k.coef * va * vb;
        dst[c] = store ? v : dst[c] + v;
      }
    }
  }
}

cudaError_t launch_kron_scatter(const KronTask* tasks, const KronTile* tiles, int ntiles, cudaStream_t s, int64_t* launches) {
  if (ntiles == 0) return cudaSuccess;
  kron_scatter_kernel<<<min(ntiles, 148 * 16), 256, 0, s>>>(tasks, tiles, ntiles);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

// gather the diagonals of operator blocks into one compact pool: pool[dst + i] = *(src + i * stride)
__global__ void gather_diag_kernel(const DiagGather* __restrict__ items, int nitems, double* __restrict__ pool) {
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const DiagGather g = items[it];
    const double* src = reinterpret_cast<const double*>(g.src);
    for (int i = threadIdx.x; i < g.n; i += blockDim.x) pool[g.dst + i] = src[(int64_t)i * g.stride];
  }
}
cudaError_t launch_gather_diag(const DiagGather* items, int nitems, double* pool, cudaStream_t s, int64_t* launches) {
  if (nitems == 0) return cudaSuccess;
  gather_diag_kernel<<<min(nitems, 8192), 128, 0, s>>>(items, nitems, pool);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_diag(const BlockDesc* blocks, int nblocks, const DiagTask* tasks, const int* block_begin, double* e, cudaStream_t s, int64_t* launches) {
  if (nblocks == 0) return cudaSuccess;
  for (int b0 = 0; b0 < nblocks; b0 += 32768) {
    int nb = min(32768, nblocks - b0);
    diag_kernel<<<dim3(8, nb, 1), 256, 0, s>>>(blocks + b0, tasks, block_begin + b0, e);
    B2D_LAUNCH_CHECK();
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// density-matrix eigen-decomposition (replaces dsyev_ at rotationmat.C:268): one-sided Jacobi on rows
// ---------------------------------------------------------------------------------------------------------------
constexpr int EIG_THREADS = 512;

__global__ void __launch_bounds__(EIG_THREADS) sector_eig_kernel(const BlockDesc* __restrict__ sectors, double* __restrict__ g, double* __restrict__ vt,
                                                                 double* __restrict__ evals, int* __restrict__ sweeps) {
  const BlockDesc sd = sectors[blockIdx.x];
  const int d = sd.rows, ld = sd.ld;
  double* G = g + sd.dev_off;
  double* V = vt + sd.dev_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  __shared__ int rotated;
  // V = identity
  for (int64_t e = threadIdx.x; e < (int64_t)d * ld; e += blockDim.x) V[e] = (e / ld == e % ld) ? 1.0 : 0.0;
  __syncthreads();
  const int D = (d + 1) & ~1;   // round-robin tournament over an even number of players
  int sweep = 0;
  for (; sweep < 60 && d > 1; ++sweep) {
    if (threadIdx.x == 0) rotated = 0;
    __syncthreads();
    for (int round = 0; round < D - 1; ++round) {
      for (int k = warp; k < D / 2; k += nwarps) {
        int a = (round + k) % (D - 1);
        int b = k == 0 ? D - 1 : (round - k + (D - 1)) % (D - 1);
        if (a >= d || b >= d) continue;
        if (a > b) { int tmp = a; a = b; b = tmp; }
        double* ga = G + (int64_t)a * ld;
        double* gb = G + (int64_t)b * ld;
        double aa = 0.0, bb = 0.0, ab = 0.0;
        for (int j = lane; j < d; j += 32) {
          double x = ga[j], y = gb[j];
          aa += x * x; bb += y * y; ab += x * y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          aa += __shfl_xor_sync(0xffffffffu, aa, o);
          bb += __shfl_xor_sync(0xffffffffu, bb, o);
          ab += __shfl_xor_sync(0xffffffffu, ab, o);
        }
        if (fabs(ab) <= 1e-15 * sqrt(aa * bb) || aa < 1e-40 || bb < 1e-40) continue;
        const double zeta = (bb - aa) / (2.0 * ab);
        const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
        double* va = V + (int64_t)a * ld;
        double* vb = V + (int64_t)b * ld;
        for (int j = lane; j < d; j += 32) {
          double x = ga[j], y = gb[j];
          ga[j] = c * x - s * y;
          gb[j] = s * x + c * y;
          x = va[j]; y = vb[j];
          va[j] = c * x - s * y;
          vb[j] = s * x + c * y;
        }
        if (lane == 0) rotated = 1;
      }
      __syncthreads();
    }
    const int any = rotated;
    __syncthreads();
    if (!any) break;
  }
  // eigenvalue of row i: Rayleigh quotient v_i . (A v_i) = v_i . g_i
  for (int i = warp; i < d; i += nwarps) {
    double acc = 0.0;
    for (int j = lane; j < d; j += 32) acc += V[(int64_t)i * ld + j] * G[(int64_t)i * ld + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) evals[sd.ref_off + i] = acc;
  }
  if (threadIdx.x == 0) sweeps[blockIdx.x] = sweep;
}

// evals[ref_off + i] = v_i . g_i for the rows of V and G = V rho (Rayleigh quotients of given eigenvectors): one CTA per sector
__global__ void __launch_bounds__(EIG_THREADS) rayleigh_kernel(const BlockDesc* __restrict__ sectors, const double* __restrict__ g, const double* __restrict__ vt,
                                                               double* __restrict__ evals) {
  const BlockDesc sd = sectors[blockIdx.x];
  const int d = sd.rows, ld = sd.ld;
  const double* G = g + sd.dev_off;
  const double* V = vt + sd.dev_off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = warp; i < d; i += nwarps) {
    double acc = 0.0;
    for (int j = lane; j < d; j += 32) acc += V[(int64_t)i * ld + j] * G[(int64_t)i * ld + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) evals[sd.ref_off + i] = acc;
  }
}
cudaError_t launch_rayleigh(const BlockDesc* sectors, int nsectors, const double* g, const double* vt, double* evals, cudaStream_t s, int64_t* launches) {
  if (nsectors == 0) return cudaSuccess;
  rayleigh_kernel<<<nsectors, EIG_THREADS, 0, s>>>(sectors, g, vt, evals);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t block_jacobi_setup() {
  return cudaFuncSetAttribute(block_jacobi_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BJ_SMEM);
}
cudaError_t launch_block_jacobi_init(const BJPair* sectors, int nsectors, double* vt, cudaStream_t s, int64_t* launches) {
  if (nsectors == 0) return cudaSuccess;
  block_jacobi_init_kernel<<<dim3(64, nsectors), 256, 0, s>>>(sectors, vt);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}
cudaError_t launch_block_jacobi_step(const BJPair* pairs, int npairs, double* g, double* vt, const int* active, int* rotated, double tol, cudaStream_t s,
                                     int64_t* launches) {
  if (npairs == 0) return cudaSuccess;
  block_jacobi_step_kernel<<<npairs, BJ_THREADS, BJ_SMEM, s>>>(pairs, g, vt, active, rotated, tol);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t launch_sector_eig(const BlockDesc* sectors, int nsectors, double* g, double* vt, double* evals, int* sweeps, cudaStream_t s, int64_t* launches) {
  if (nsectors == 0) return cudaSuccess;
  sector_eig_kernel<<<nsectors, EIG_THREADS, 0, s>>>(sectors, g, vt, evals, sweeps);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

// U_q[j, c] = Vt_q[row(c), j]: kept eigenvectors (rows of Vt) become the columns of the rotation matrix.  32 x 32 tiles through shared
// memory: reads run along j (an eigenvector row), writes along c (a row of U) - both coalesced; grid.y = sector, grid.x strides over
// the tiles of that sector.
__global__ void __launch_bounds__(256) gather_rotation_kernel(const GatherDesc* __restrict__ desc, const int* __restrict__ src_rows, const double* __restrict__ vt, double* __restrict__ u) {
  __shared__ double tile[32][33];
  const GatherDesc g = desc[blockIdx.y];
  const int tj = (g.d + 31) / 32, tc = (g.ncols + 31) / 32;
  for (int t = blockIdx.x; t < tj * tc; t += gridDim.x) {
    const int j0 = (t % tj) * 32, c0 = (t / tj) * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int c = c0 + r, j = j0 + threadIdx.x;
      if (c < g.ncols && j < g.d) tile[r][threadIdx.x] = vt[g.vt_off + (int64_t)src_rows[g.row_begin + c] * g.ld_vt + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int j = j0 + r, c = c0 + threadIdx.x;
      if (c < g.ncols && j < g.d) u[g.u_off + (int64_t)j * g.ld_u + c] = tile[threadIdx.x][r];
    }
    __syncthreads();
  }
}
cudaError_t launch_gather_rotation(const GatherDesc* desc, int nsectors, const int* src_rows, const double* vt, double* u, cudaStream_t s, int64_t* launches) {
  if (nsectors == 0) return cudaSuccess;
  for (int s0 = 0; s0 < nsectors; s0 += 65535) {   // grid.y limit
    const int ns = std::min(nsectors - s0, 65535);
    gather_rotation_kernel<<<dim3(48, ns), dim3(32, 8), 0, s>>>(desc + s0, src_rows, vt, u);
    B2D_LAUNCH_CHECK();
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// synthetic operators for the benchmark
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double counter_uniform(uint64_t seed, uint64_t idx) {   // splitmix64
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

__global__ void fill_random_kernel(double* __restrict__ dst, const BlockDesc* __restrict__ blocks, int nblocks, uint64_t seed, double amplitude) {
  for (int b = blockIdx.y; b < nblocks; b += gridDim.y) {
    const BlockDesc d = blocks[b];
    const int64_t n = (int64_t)d.rows * d.cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
      int r = (int)(e / d.cols), c = (int)(e % d.cols);
      dst[d.dev_off + (int64_t)r * d.ld + c] = 2.0 * amplitude * counter_uniform(seed, (uint64_t)(d.ref_off + e));
    }
  }
}
cudaError_t launch_fill_random(double* dst, const BlockDesc* blocks, int nblocks, uint64_t seed, double amplitude, cudaStream_t s, int64_t* launches) {
  if (nblocks == 0) return cudaSuccess;
  fill_random_kernel<<<pack_grid(nblocks), 256, 0, s>>>(dst, blocks, nblocks, seed, amplitude);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

__global__ void symmetrise_kernel(double* __restrict__ base, const SymPair* __restrict__ pairs, int npairs) {
  for (int p = blockIdx.y; p < npairs; p += gridDim.y) {
    const SymPair sp = pairs[p];
    const int64_t n = (int64_t)sp.rows * sp.cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
      int r = (int)(e / sp.cols), c = (int)(e % sp.cols);
      double* pa = base + sp.off_a + (int64_t)r * sp.ld_a + c;
      double* pb = base + sp.off_b + (int64_t)c * sp.ld_b + r;
      if (sp.off_a == sp.off_b) {
        if (c >= r) { double v = 0.5 * (*pa + *pb); *pa = v; *pb = v; }   // diagonal block, f = 1
      } else {
        double v = *pa;
        *pb = sp.f * v;
      }
    }
  }
}
cudaError_t launch_symmetrise(double* base, const SymPair* pairs, int npairs, cudaStream_t s, int64_t* launches) {
  if (npairs == 0) return cudaSuccess;
  symmetrise_kernel<<<pack_grid(npairs), 256, 0, s>>>(base, pairs, npairs);
  B2D_LAUNCH_CHECK();
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 yardsticks
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double acc[8][4];
  double a[4], b[2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.0;
  a[0] = threadIdx.x * 1e-3; a[1] = a[0] + 1; a[2] = a[0] + 2; a[3] = a[0] + 3;
  b[0] = 1e-6 * threadIdx.x; b[1] = b[0] + 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma_16x8x8(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) s += acc[i][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i;
  const double a = 1.0000001, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t measure_fp64(cudaStream_t s, double* dmma_tflops, double* dfma_tflops) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 4, threads = 256, iters = 20000;
  double* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(double) * grid * threads);
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0.f;
  dmma_peak_kernel<<<grid, threads, 0, s>>>(out, 100);
  cudaEventRecord(e0, s);
  dmma_peak_kernel<<<grid, threads, 0, s>>>(out, iters);
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  // per warp per iteration: 8 DMMA x (16*8*8) FMA x 2 flops
  *dmma_tflops = (double)grid * (threads / 32) * iters * 8.0 * 16 * 8 * 8 * 2 / (ms * 1e-3) / 1e12;
  dfma_peak_kernel<<<grid, threads, 0, s>>>(out, 100);
  cudaEventRecord(e0, s);
  dfma_peak_kernel<<<grid, threads, 0, s>>>(out, iters);
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  *dfma_tflops = (double)grid * threads * iters * 16.0 * 2 / (ms * 1e-3) / 1e12;
  e = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return e;
}

}  // namespace b2d
