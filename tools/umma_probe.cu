// Hardware probe (developer tool, not part of the product): validates the UMMA shared-memory
// descriptor conventions the implicit-GEMM convolution relies on, on a real sm_100a part:
//   T0  plain 128xNx128 GEMM, TMA SW128 tiles, aligned descriptors
//   T1  A descriptor start shifted by whole 8-row groups (1024 B)           (3x3 tap, row shift)
//   T2  A descriptor start shifted by s rows (s*128 B, not 1024-aligned), base_offset = 0
//   T3  same, base_offset = (addr >> 7) & 7
//   T4  A rows with 10-pixel pitch (SBO = 1280 B) + shifted start               (single-halo tile)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "../rethinking-flow-and-diffusion-bridge-models-for-speech-enhancement_b200/csrc/tc05.cuh"

using namespace tc05;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int N_ = 128;
constexpr int KB = 2;          // k-blocks of 64
constexpr int A_ROWS_MAX = 176;

struct Params {
  uint32_t a_box_bytes;    // bytes per A k-block landed by TMA
  uint32_t a_start_off;    // byte offset added to the A descriptor start
  uint32_t a_sbo;          // stride between 8-row groups
  uint32_t base_offset;    // descriptor base_offset field
  int a_dims;              // 2 or 3 (tensor-map rank for A)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, Params p,
             float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t A_STRIDE = A_ROWS_MAX * 128;            // 22528 = 22 * 1024
  uint8_t* sA = smem;
  uint8_t* sB = smem + KB * A_STRIDE;
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x / 32;

  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 128);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, KB * (p.a_box_bytes + N_ * 128));
    for (int kb = 0; kb < KB; ++kb) {
      if (p.a_dims == 2) tma_load_2d(sA + kb * A_STRIDE, &map_a, &full_bar, kb * 64, 0);
      else tma_load_3d(sA + kb * A_STRIDE, &map_a, &full_bar, kb * 64, 0, 0);
      tma_load_2d(sB + kb * N_ * 128, &map_b, &full_bar, kb * 64, 0);
    }
    mbar_wait(&full_bar, 0);
    fence_after_sync();
    const uint32_t idesc = make_idesc_f16(128, N_, 1);
    for (int kb = 0; kb < KB; ++kb) {
      for (int k = 0; k < 4; ++k) {
        uint64_t da = make_desc_sw128(smem_u32(sA + kb * A_STRIDE) + p.a_start_off + k * 32, p.a_sbo, p.base_offset);
        uint64_t db = make_desc_sw128(smem_u32(sB + kb * N_ * 128) + k * 32, 1024, 0);
        mma_f16(tmem, da, db, idesc, (kb | k) != 0);
      }
    }
    mma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  fence_after_sync();
  for (int c = 0; c < N_ / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((warp * 32u) << 16) + c * 32, r);
    tmem_ld_wait();
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int j = 0; j < 32; ++j) D[row * N_ + c * 32 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return reinterpret_cast<EncodeFn>(fn);
}

static float bf(__nv_bfloat16 v) { return __bfloat162float(v); }

int main() {
  EncodeFn encode = get_encode();
  const int K = KB * 64;
  std::vector<__nv_bfloat16> hA(A_ROWS_MAX * K), hB(N_ * K);
  srand(1);
  for (auto& v : hA) v = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (auto& v : hB) v = __float2bfloat16((rand() % 13 - 6) / 8.0f);
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, 128 * N_ * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));

  CUtensorMap mapB, mapA2, mapA3;
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N_};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)N_};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", r); return 2; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)A_ROWS_MAX};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)A_ROWS_MAX};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode(&mapA2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A2 failed %d\n", r); return 2; }
  }
  {
    // A viewed as an image patch [17 rows][10 px][K] (170 <= 176 rows of the same buffer)
    cuuint64_t dims[3] = {(cuuint64_t)K, 10, 17};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)K * 2 * 10};
    cuuint32_t box[3] = {64, 10, 17};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&mapA3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A3 failed %d\n", r); return 2; }
  }

  const size_t smem_bytes = 1024 + KB * A_ROWS_MAX * 128 + KB * N_ * 128;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));

  struct Case { const char* name; Params p; int mode; int shift; };
  std::vector<Case> cases;
  cases.push_back({"T0 aligned", {A_ROWS_MAX * 128, 0, 1024, 0, 2}, 0, 0});
  cases.push_back({"T1 start+8rows", {A_ROWS_MAX * 128, 1024, 1024, 0, 2}, 0, 8});
  cases.push_back({"T1 start+16rows", {A_ROWS_MAX * 128, 2048, 1024, 0, 2}, 0, 16});
  for (int s = 1; s <= 3; ++s) {
    cases.push_back({"T2 start+s rows, base_offset=0", {A_ROWS_MAX * 128, (uint32_t)s * 128, 1024, 0, 2}, 0, s});
    cases.push_back({"T3 start+s rows, base_offset=s", {A_ROWS_MAX * 128, (uint32_t)s * 128, 1024, (uint32_t)s, 2}, 0, s});
  }
  cases.push_back({"T2 start+9 rows, base_offset=0", {A_ROWS_MAX * 128, 9 * 128, 1024, 0, 2}, 0, 9});
  for (int s = 0; s <= 2; ++s) {
    cases.push_back({"T4 pitch10 SBO=1280 shift s, bo=0", {170 * 128, (uint32_t)s * 128, 1280, 0, 3}, 1, s});
    cases.push_back({"T4 pitch10 SBO=1280 shift s+row, bo=0", {170 * 128, (uint32_t)(s + 10) * 128, 1280, 0, 3}, 1, s + 10});
  }

  std::vector<float> hD(128 * N_);
  int n_fail = 0;
  for (auto& c : cases) {
    CK(cudaMemset(dD, 0xff, 128 * N_ * 4));
    probe_kernel<<<1, 128, smem_bytes>>>(c.p.a_dims == 2 ? mapA2 : mapA3, mapB, c.p, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s shift=%2d  KERNEL ERROR %s\n", c.name, c.shift, cudaGetErrorString(e)); return 3; }
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0;
    for (int m = 0; m < 128; ++m) {
      int arow;
      if (c.mode == 0) arow = m + c.shift;                       // consecutive rows
      else arow = (m / 8) * 10 + (m % 8) + c.shift;              // 8 px per image row, pitch 10
      for (int n = 0; n < N_; ++n) {
        float acc = 0;
        for (int k = 0; k < K; ++k) acc += bf(hA[arow * K + k]) * bf(hB[n * K + k]);
        double d = fabs(acc - hD[m * N_ + n]);
        if (!(d <= max_err)) max_err = d;
      }
    }
    bool ok = max_err < 1e-3;
    if (!ok) n_fail++;
    printf("%-44s shift=%2d  max_err=%-12g %s\n", c.name, c.shift, max_err, ok ? "PASS" : "FAIL");
  }
  printf("probe done, %d failing variants\n", n_fail);
  return 0;
}
