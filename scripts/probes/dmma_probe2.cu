// Where do the 30% go?  The warp-tile inner loop of the grouped kernel (2 x 4 m16n8k8 per k8 step) in isolation:
//   v0: operands fixed in registers     v1: fragments re-loaded from shared memory every step (LDS.64)
//   v2: v1 + the alpha DMULs            v3: v1 + __syncthreads every 2 steps (one pipeline stage)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
template <int V>
__global__ void __launch_bounds__(512) k(double* out, int iters, double alpha) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 2 * 256 * 20; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp / 4, wn = warp % 4;
  const double* sA = sm + (wm * 32 + g) * 20 + t;
  const double* sB = sm + 128 * 20 + (wn * 32 + g) * 20 + t;
  double acc[2][4][4] = {};
  double af[2][4], bf[4][2];
  for (int i = 0; i < 2; ++i) for (int e = 0; e < 4; ++e) af[i][e] = 1e-3 * (threadIdx.x + i + e);
  for (int j = 0; j < 4; ++j) for (int e = 0; e < 2; ++e) bf[j][e] = 1e-4 * (threadIdx.x + j + e);
  for (int it = 0; it < iters; ++it) {
    const int kk = (it & 1) * 8;
    if (V >= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { bf[j][0] = sB[j * 8 * 20 + kk]; bf[j][1] = sB[j * 8 * 20 + kk + 4]; if (V == 2) { bf[j][0] *= alpha; bf[j][1] *= alpha; } }
#pragma unroll
      for (int i = 0; i < 2; ++i) { af[i][0] = sA[i * 16 * 20 + kk]; af[i][1] = sA[(i * 16 + 8) * 20 + kk]; af[i][2] = sA[i * 16 * 20 + kk + 4]; af[i][3] = sA[(i * 16 + 8) * 20 + kk + 4]; }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma_16x8x8(acc[i][j], af[i], bf[j]);
    if (V == 3 && (it & 1)) __syncthreads();
  }
  double s = 0;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += acc[i][j][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V> void run(int sms, int threads, double* out) {
  const int iters = 20000;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 20 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<sms, threads, 2 * 256 * 20 * 8>>>(out, 100, 1.0000001);
  cudaEventRecord(e0); k<V><<<sms, threads, 2 * 256 * 20 * 8>>>(out, iters, 1.0000001); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("variant %d threads %3d : %6.2f TFLOP/s  (%s)\n", V, threads, (double)sms * (threads / 32) * iters * 8 * 2048.0 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8 * sms * 512);
  for (int threads : {128, 256, 512}) { run<0>(sms, threads, out); run<1>(sms, threads, out); run<2>(sms, threads, out); run<3>(sms, threads, out); }
  return 0;
}
