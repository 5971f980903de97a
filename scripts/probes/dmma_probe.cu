// FP64 DMMA issue/latency probe for sm_100a: throughput vs warps per SM and independent accumulators per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k16(double* out, int iters) {
  double acc[ILP][4], a[4], b[2];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.0;
  a[0] = threadIdx.x * 1e-3; a[1] = a[0] + 1; a[2] = a[0] + 2; a[3] = a[0] + 3;
  b[0] = 1e-6 * threadIdx.x; b[1] = b[0] + 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma_16x8x8(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) s += acc[i][e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k8(double* out, int iters) {
  double acc[ILP][2], a, b;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = 0.0;
  a = threadIdx.x * 1e-3; b = 1e-6 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma_8x8x4(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
double run(F launch, double flops_per_launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return flops_per_launch / (ms * 1e-3) / 1e12;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, 8 * sms * 1024 * 4);
  const int iters = 4000;
  printf("SMs %d\n", sms);
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32;   // one CTA per SM
#define R16(ILP) printf("m16n8k8 warps/SM %2d ILP %2d : %6.2f TFLOP/s\n", warps, ILP, run([&] { k16<ILP><<<sms, threads>>>(out, iters); }, (double)sms * warps * iters * ILP * 2048.0));
    R16(1) R16(2) R16(4) R16(8) R16(16)
#define R8(ILP) printf("m8n8k4  warps/SM %2d ILP %2d : %6.2f TFLOP/s\n", warps, ILP, run([&] { k8<ILP><<<sms, threads>>>(out, iters); }, (double)sms * warps * iters * ILP * 512.0));
    R8(1) R8(4) R8(16) R8(32)
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
