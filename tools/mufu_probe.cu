// Hardware probe (developer tool): issue rate of MUFU.TANH (fp32) vs MUFU.TANH.F16 (one per half of tanh.approx.f16x2) vs MUFU.EX2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mufu_probe tools/mufu_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)
template <int MODE> __global__ void k(float* out, int iters) {
  float a[8]; unsigned h[8];
  for (int j = 0; j < 8; ++j) { a[j] = 0.001f * (threadIdx.x + j); h[j] = 0x3c003800u + threadIdx.x + j; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[j]));
      else if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[j]));
      else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[j]));
    }
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += a[j] + __uint_as_float(h[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; CK(cudaMalloc(&d, 148 * 8 * 1024 * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  const char* names[3] = {"tanh.approx.f32 (MUFU.TANH)", "tanh.approx.f16x2 (2 x MUFU.TANH.F16)", "ex2.approx.ftz.f32 (MUFU.EX2)"};
  for (int m = 0; m < 3; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<148 * 2, 1024>>>(d, iters); else if (m == 1) k<1><<<148 * 2, 1024>>>(d, iters); else k<2><<<148 * 2, 1024>>>(d, iters);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 2 * 1024 * iters * 8 * (m == 1 ? 2 : 1);           // scalar function evaluations
    printf("%-42s %8.3f ms  %7.2f evaluations / clk / SM at 1.9 GHz\n", names[m], ms, ops / (ms * 1e-3) / 148 / 1.9e9);
  }
  return 0;
}
