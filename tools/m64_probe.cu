// Hardware probe (developer tool, not part of the product): where does tcgen05.mma.cta_group::1.kind::f16 with M = 64 put the
// rows of D in TMEM, and may the D address carry a lane offset?  (Needed to decide whether two M = 64 accumulators can share
// TMEM columns in different lanes for the TF-GridNet LSTM sweep.)
// A[i][0] = i + 1 (64 x 16, K-major SW128), B[j][0] = 1 -> D[i][j] = i + 1.  All 128 lanes x 32 columns are dumped.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o m64_probe tools/m64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200/csrc/tc05.cuh"
using namespace tc05;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__host__ __device__ inline int sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(128, 1) probe(int M, uint32_t d_lane, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t done;
  __shared__ uint32_t tslot;
  uint8_t* sA = smem;              // [128 rows][128 B]
  uint8_t* sB = smem + 16384;      // [64 rows][128 B]
  for (int i = threadIdx.x; i < (16384 + 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (threadIdx.x < 128) *reinterpret_cast<__half*>(sA + sw128_off(threadIdx.x, 0)) = __float2half(float(threadIdx.x + 1));
  if (threadIdx.x < 64) *reinterpret_cast<__half*>(sB + sw128_off(threadIdx.x, 0)) = __float2half(1.0f);
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tslot, 128); tmem_relinquish(); }
  asm volatile("fence.proxy.async;" ::: "memory");
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tslot;
  // clear the 128 lanes x 64 columns first (tcgen05.st), so untouched lanes read as -1
  {
    const uint32_t lane_sel = ((threadIdx.x >> 5) * 32u) << 16;
    for (int c = 0; c < 64; ++c) {
      const uint32_t v = __float_as_uint(-1.0f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + lane_sel + c), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(M, 64, 0);
      mma_f16(tm + (d_lane << 16), make_desc_sw128(smem_u32(sA), 1024), make_desc_sw128(smem_u32(sB), 1024), idesc, 0u);
      mma_commit(&done);
    }
    __syncwarp();
  }
  mbar_wait(&done, 0);
  fence_after_sync();
  uint32_t v[32];
  tmem_ld_32x32(tm + (((threadIdx.x >> 5) * 32u) << 16), v);
  tmem_ld_wait();
  for (int c = 0; c < 32; ++c) out[threadIdx.x * 32 + c] = __uint_as_float(v[c]);
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 128);
}

int main() {
  float* d; CK(cudaMalloc(&d, 128 * 32 * 4));
  float h[128 * 32];
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  const int Ms[] = {128, 64, 64, 64, 64};
  const uint32_t lanes[] = {0, 0, 16, 32, 64};
  for (int t = 0; t < 5; ++t) {
    CK(cudaMemset(d, 0, sizeof(h)));
    probe<<<1, 128, 32768>>>(Ms[t], lanes[t], d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d d_lane=%u: %s\n", Ms[t], lanes[t], cudaGetErrorString(e)); return 0; }
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("M=%d D lane offset %u: value of column 0 (and 1, 31) per TMEM lane (-1 = untouched)\n", Ms[t], lanes[t]);
    for (int l = 0; l < 128; ++l) {
      printf("%4.0f/%.0f/%.0f", h[l * 32], h[l * 32 + 1], h[l * 32 + 31]);
      if (l % 16 == 15) printf("\n");
    }
  }
  return 0;
}
