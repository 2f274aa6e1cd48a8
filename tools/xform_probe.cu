// Hardware probe (developer tool): cost of the convolution's operand transform (GroupNorm affine in fp32 + SiLU as
// h + h tanh(h) in packed 16-bit, conv_igemm.cu "operand transform") on one 18 x 10 pixel x 64 channel tile in shared
// memory, as a function of how many warps per SM sub-partition run it and how many 16-byte slots are in flight per pass.
// Answers: is the transform bound by the MUFU / conversion pipes (then more warps do not help) or by latency?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xform_probe tools/xform_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 2; } } while (0)

constexpr int TILE_PX = 180, TILE_BYTES = 23552;

__device__ __forceinline__ float2 h22f2(__half2 h) { return __half22float2(h); }

// MODE 0: tanh.approx.f16x2 (the shipped arithmetic); 1: tanh.approx.f32 per element; 2: affine only (no SiLU); 3: all-16-bit (HFMA2 affine)
template <int MODE>
__device__ __forceinline__ void xform(uint4& rawv, const float2 (&sc)[4], const float2 (&sh)[4]) {
  __half2* h2 = reinterpret_cast<__half2*>(&rawv);
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    if (MODE == 3) {
      const __half2 s = __floats2half2_rn(sc[w].x, sc[w].y), b = __floats2half2_rn(sh[w].x, sh[w].y);
      const __half2 h = __hfma2(h2[w], s, b);
      uint32_t hi = *reinterpret_cast<const uint32_t*>(&h), t;
      asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hi));
      h2[w] = __hfma2(h, *reinterpret_cast<__half2*>(&t), h);
      continue;
    }
    const float2 y = __ffma2_rn(h22f2(h2[w]), sc[w], sh[w]);
    if (MODE == 0) {
      const __half2 h = __floats2half2_rn(y.x, y.y);
      uint32_t hi = *reinterpret_cast<const uint32_t*>(&h), t;
      asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hi));
      h2[w] = __hfma2(h, *reinterpret_cast<__half2*>(&t), h);
    } else if (MODE == 1) {
      float tx, ty;
      asm("tanh.approx.f32 %0, %1;" : "=f"(tx) : "f"(y.x));
      asm("tanh.approx.f32 %0, %1;" : "=f"(ty) : "f"(y.y));
      const float2 o = __ffma2_rn(y, make_float2(tx, ty), y);
      h2[w] = __floats2half2_rn(o.x, o.y);
    } else {
      h2[w] = __floats2half2_rn(y.x, y.y);
    }
  }
}

// NT threads transform `tiles` tiles one after the other; thread -> 8-channel group g and pixel lane p_lane (NT / 8 lanes)
template <int MODE, int NT, int XS>
__global__ void __launch_bounds__(NT) k(float* out, int tiles, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < TILE_BYTES / 4; i += NT) reinterpret_cast<uint32_t*>(smem)[i] = 0x38003400u + (i & 255);
  __syncthreads();
  const int g = threadIdx.x & 7, p_lane = threadIdx.x >> 3;
  constexpr int LANES = NT / 8;                       // 16 (128 threads) or 32 (256 threads)
  constexpr int SLOTS = (TILE_PX + LANES - 1) / LANES;   // 12 or 6
  float2 sc[4], sh[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) { sc[u] = make_float2(0.5f + 0.01f * g, 0.4f + 0.01f * u); sh[u] = make_float2(0.01f * u, -0.02f * g); }
  uint8_t* base = smem + p_lane * 128 + ((g ^ (p_lane & 7)) << 4);
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
#pragma unroll
    for (int pass = 0; pass < SLOTS / XS; ++pass) {
      uint4 raw[XS];
#pragma unroll
      for (int u = 0; u < XS; ++u) {
        const int i = pass * XS + u;
        if (i * LANES + p_lane < TILE_PX) raw[u] = *reinterpret_cast<uint4*>(base + i * LANES * 128);
      }
#pragma unroll
      for (int u = 0; u < XS; ++u) xform<MODE>(raw[u], sc, sh);
#pragma unroll
      for (int u = 0; u < XS; ++u) {
        const int i = pass * XS + u;
        if (i * LANES + p_lane < TILE_PX) *reinterpret_cast<uint4*>(base + i * LANES * 128) = raw[u];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * NT + threadIdx.x] = __half2float(reinterpret_cast<__half*>(smem)[threadIdx.x]);
}

template <int MODE, int NT, int XS>
int run(const char* name, float* d, long long* dc) {
  const int tiles = 2000;
  CK(cudaFuncSetAttribute(k<MODE, NT, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int rep = 0; rep < 2; ++rep) k<MODE, NT, XS><<<148, NT, 200 * 1024>>>(d, tiles, dc);
  CK(cudaDeviceSynchronize());
  long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
  printf("%-58s %7.0f cycles per tile (%5.1f per 16-byte slot and thread-slot lane)\n", name, double(c) / tiles, double(c) / tiles / (TILE_PX * 8.0 / NT));
  return 0;
}

int main() {
  float* d; long long* dc;
  CK(cudaMalloc(&d, 148 * 256 * 4)); CK(cudaMalloc(&dc, 8));
  run<0, 128, 4>("f16x2 tanh, 1 warp / sub-partition, 4 slots in flight", d, dc);
  run<0, 128, 6>("f16x2 tanh, 1 warp / sub-partition, 6 slots in flight", d, dc);
  run<0, 128, 12>("f16x2 tanh, 1 warp / sub-partition, 12 slots in flight", d, dc);
  run<0, 256, 3>("f16x2 tanh, 2 warps / sub-partition, 3 slots in flight", d, dc);
  run<0, 256, 6>("f16x2 tanh, 2 warps / sub-partition, 6 slots in flight", d, dc);
  run<1, 128, 4>("f32 tanh, 1 warp / sub-partition, 4 slots in flight", d, dc);
  run<1, 256, 3>("f32 tanh, 2 warps / sub-partition, 3 slots in flight", d, dc);
  run<2, 128, 4>("affine only, 1 warp / sub-partition, 4 slots in flight", d, dc);
  run<2, 256, 3>("affine only, 2 warps / sub-partition, 3 slots in flight", d, dc);
  run<3, 128, 4>("all 16-bit, 1 warp / sub-partition, 4 slots in flight", d, dc);
  run<3, 256, 3>("all 16-bit, 2 warps / sub-partition, 3 slots in flight", d, dc);
  return 0;
}
