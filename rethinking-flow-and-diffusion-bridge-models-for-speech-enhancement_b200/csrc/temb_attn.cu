// Time embedding (layerspp.py:32-41, ncsnpp_v2.py:108-113,252-270), the 49 per-block Dense_0
// projections (layerspp.py:263) as one batched GEMV, and the attention core of AttnBlockpp
// (layerspp.py:82-86).  All tiny next to the convolutions (< 0.1 % of the FLOPs at 4 s).
#include "common.cuh"

namespace fdbm {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// One block per batch item.  emb = [sin, cos](log(t) * W * 2 * pi)  -> Linear -> SiLU -> Linear -> SiLU.
// The result is SiLU(temb): every consumer applies the activation first (layerspp.py:263).
__global__ void __launch_bounds__(512)
temb_kernel(const float* __restrict__ t, const float* __restrict__ fw, int nf, const float* __restrict__ w1,
            const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
            int t_stride, float* __restrict__ out) {
  extern __shared__ float sm[];          // emb[2nf] | h1[4nf]
  float* emb = sm;
  float* h1 = sm + 2 * nf;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // fp32 op order of the reference: ((log(t) * W) * 2) * pi, pi rounded to fp32
  const float lt = static_cast<float>(log(static_cast<double>(t[b * t_stride])));
  for (int j = threadIdx.x; j < nf; j += blockDim.x) {
    const float proj = __fmul_rn(__fmul_rn(__fmul_rn(lt, fw[j]), 2.0f), 3.14159274101257324f);
    emb[j] = sinf(proj);
    emb[nf + j] = cosf(proj);
  }
  __syncthreads();
  const int D = 4 * nf;
  for (int r = warp; r < D; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < 2 * nf; k += 32) acc = fmaf(w1[r * 2 * nf + k], emb[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) h1[r] = silu_f(acc + b1[r]);
  }
  __syncthreads();
  for (int r = warp; r < D; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(w2[r * D + k], h1[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[static_cast<int64_t>(b) * D + r] = silu_f(acc + b2[r]);
  }
}

// out[b][r] = sum_k act[b][k] * w[r][k] + bias[r];  warp per row, up to 8 batch items per pass.
__global__ void __launch_bounds__(256)
dense_all_kernel(const float* __restrict__ act, const float* __restrict__ w, const float* __restrict__ bias, int B,
                 int K, int rows, float* __restrict__ out) {
  extern __shared__ float sa[];          // [8][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = min(8, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * K; i += 256) sa[i] = act[static_cast<int64_t>(b0) * K + i];
    __syncthreads();
    for (int r = blockIdx.x * 8 + warp; r < rows; r += 8 * gridDim.x) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float wv = w[static_cast<int64_t>(r) * K + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nb) acc[j] = fmaf(wv, sa[j * K + k], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nb) {
          const float v = warp_sum(acc[j]);
          if (lane == 0) out[static_cast<int64_t>(b0 + j) * rows + r] = v + bias[r];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Attention: one warp per query.  Pass 1 scores all keys (lane = key), pass 2 soft-max, pass 3
// o = P V (lane = 8 channels).  K/V stay L1/L2 resident (L*C*2 bytes = 128 KB at L=256, C=256).
// ------------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8;

__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_kernel(const op_t* __restrict__ q, const op_t* __restrict__ k,
                 const op_t* __restrict__ v, int ld, int L, int C, float scale,
                 op_t* __restrict__ o, int ldo) {
  extern __shared__ float sm[];                       // per warp: q[C] | p[L]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int qi = blockIdx.x * ATT_WARPS + warp;
  if (qi >= L) return;
  float* sq = sm + warp * (C + L);
  float* sp = sq + C;
  const int64_t base = static_cast<int64_t>(b) * L;
  for (int c = lane; c < C; c += 32) sq[c] = op2f(q[(base + qi) * ld + c]) * scale;
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(k + (base + j) * ld);
    float acc = 0.f;
    for (int c8 = 0; c8 < C / 8; ++c8) {
      const uint4 raw = kr[c8];
      const op2_t* h2 = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f2 = op22f2(h2[u]);
        acc = fmaf(f2.x, sq[c8 * 8 + 2 * u], acc);
        acc = fmaf(f2.y, sq[c8 * 8 + 2 * u + 1], acc);
      }
    }
    sp[j] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = __expf(sp[j] - mx);
    sp[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __syncwarp();
  // lane owns channels [lane*8, lane*8+8) (+256 per repeat)
  for (int c0 = lane * 8; c0 < C; c0 += 256) {
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    for (int j = 0; j < L; ++j) {
      const float pj = sp[j];
      const uint4 raw = *reinterpret_cast<const uint4*>(v + (base + j) * ld + c0);
      const op2_t* h2 = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f2 = op22f2(h2[u]);
        acc[2 * u] = fmaf(pj, f2.x, acc[2 * u]);
        acc[2 * u + 1] = fmaf(pj, f2.y, acc[2 * u + 1]);
      }
    }
    op2_t ov[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ov[u] = f2op2(acc[2 * u] * inv, acc[2 * u + 1] * inv);
    *reinterpret_cast<uint4*>(o + (base + qi) * ldo + c0) = *reinterpret_cast<uint4*>(ov);
  }
}

}  // namespace

int launch_temb(const float* t, const float* fourier_w, int nf, const float* w1, const float* b1, const float* w2,
                const float* b2, int B, int t_stride, float* temb_act, cudaStream_t s) {
  temb_kernel<<<B, 512, sizeof(float) * 6 * nf, s>>>(t, fourier_w, nf, w1, b1, w2, b2, t_stride, temb_act);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_dense_all(const float* temb_act, const float* w, const float* bias, int B, int K, int rows, float* out,
                     cudaStream_t s) {
  const int grid = std::min(ceil_div(rows, 8), num_sms() * 8);
  dense_all_kernel<<<grid, 256, sizeof(float) * 8 * K, s>>>(temb_act, w, bias, B, K, rows, out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_attention(const op_t* q, const op_t* k, const op_t* v, int ld, int B, int L,
                     int C, op_t* o, int ldo, cudaStream_t s) {
  FDBM_REQUIRE(C % 8 == 0 && ld % 8 == 0 && ldo % 8 == 0, "attention: channels / strides must be multiples of 8");
  const size_t smem = sizeof(float) * ATT_WARPS * (C + L);
  FDBM_REQUIRE(smem <= 200 * 1024, "attention: sequence length %d too long for the shared-memory score buffer", L);
  static size_t attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    FDBM_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_smem = 200 * 1024;
  }
  dim3 grid(ceil_div(L, ATT_WARPS), B);
  attention_kernel<<<grid, ATT_WARPS * 32, smem, s>>>(q, k, v, ld, L, C, 1.0f / sqrtf(static_cast<float>(C)), o, ldo);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

}  // namespace fdbm

using namespace fdbm;

extern "C" int fdbm_attention(const void* q, const void* k, const void* v, int batch, int L, int C, void* o,
                              void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(q && k && v && o && batch > 0 && L > 0, "fdbm_attention: bad arguments");
  return launch_attention(reinterpret_cast<const op_t*>(q), reinterpret_cast<const op_t*>(k),
                          reinterpret_cast<const op_t*>(v), C, batch, L, C,
                          reinterpret_cast<op_t*>(o), C, as_stream(stream));
}
