// Hardware probe (developer tool): register layout of tcgen05.ld.sync.aligned.16x256b.x4.b32 (16 TMEM lanes x 32 columns
// spread over the 32 threads of a warp), with the lane base at a quadrant's first and second 16 lanes.
// TMEM is filled with value = lane * 1000 + column through tcgen05.st.32x32b.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ld16_probe tools/ld16_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128, 1) probe(uint32_t lane_off, float* out) {
  __shared__ uint32_t tslot;
  if (threadIdx.x < 32) { tmem_alloc(&tslot, 64); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tslot;
  const uint32_t q = threadIdx.x >> 5;
  for (int c = 0; c < 64; ++c) {
    const uint32_t v = __float_as_uint(float(threadIdx.x * 1000 + c));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + ((q * 32u) << 16) + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  uint32_t r[16];
  const uint32_t taddr = tm + ((q * 32u + lane_off) << 16);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  tmem_ld_wait();
  for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = __uint_as_float(r[i]);
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 64);
}

int main() {
  float* d; CK(cudaMalloc(&d, 128 * 16 * 4));
  float h[128 * 16];
  for (uint32_t off = 0; off <= 16; off += 16) {
    probe<<<1, 128>>>(off, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lane offset %u: %s\n", off, cudaGetErrorString(e)); return 0; }
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("16x256b.x4, lane base = 32 * quadrant + %u: registers r0..r15 of threads 0-7, 31 of warp 0 and thread 0 of warp 1 (value = lane * 1000 + column)\n", off);
    const int ts[] = {0, 1, 2, 3, 4, 5, 6, 7, 31, 32};
    for (int t : ts) {
      printf("  t%3d:", t);
      for (int i = 0; i < 16; ++i) printf(" %6.0f", h[t * 16 + i]);
      printf("\n");
    }
  }
  return 0;
}
