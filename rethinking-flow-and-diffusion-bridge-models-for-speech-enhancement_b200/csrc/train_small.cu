// Small backward passes of the training step around the 4-channel pyramids, the Combine layers and the time
// embedding (backward).  All are tiny next to the convolutions.
#include "common.cuh"

namespace fdbm {
namespace {

__device__ __forceinline__ float silu_grad(float y) {
  const float sg = 1.0f / (1.0f + __expf(-y));
  return sg * (1.0f + y * (1.0f - sg));
}

// ------------------------------------------------------------------------------------------------
// output_layer backward (ncsnpp_v2.py:392-399): g_out complex [B,257,T] -> g_pyr fp32 [B,T,F,Cp];
// dW[o][k] += inv * sum g_o * pyr_k, db[o] += inv * sum g_o  (fp32 atomics, 2*Cp+2 values)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
output_layer_bwd_kernel(const float2* __restrict__ g_out, const float* __restrict__ pyr, int Cp, const float* __restrict__ w,
                        int T, int F, int F_out, float inv, float* __restrict__ g_pyr, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float2 tile[32][33];
  __shared__ float red[256][10];
  const int b = blockIdx.z;
  const int f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {            // r: frequency, tx: frame
    const int f = f0 + r, t = t0 + tx;
    tile[r][tx] = (t < T && f < F) ? g_out[(static_cast<int64_t>(b) * F_out + f) * T + t] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  float acc[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) acc[j] = 0.f;
  for (int r = ty; r < 32; r += 8) {            // r: frame, tx: frequency
    const int t = t0 + r, f = f0 + tx;
    if (t < T && f < F) {
      const float2 g = tile[tx][r];
      const int64_t o = ((static_cast<int64_t>(b) * T + t) * F + f) * Cp;
      for (int k = 0; k < Cp; ++k) {
        g_pyr[o + k] = w[k] * g.x + w[Cp + k] * g.y;
        const float pk = pyr[o + k];
        acc[k] = fmaf(g.x, pk, acc[k]);
        acc[4 + k] = fmaf(g.y, pk, acc[4 + k]);
      }
      acc[8] += g.x; acc[9] += g.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 10; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 10) {
    float s = 0.f;
    for (int i = 0; i < 256; ++i) s += red[i][threadIdx.x];
    const int j = threadIdx.x;
    if (j < 8) { if ((j & 3) < Cp) atomicAdd(dw + (j >> 2) * Cp + (j & 3), inv * s); }
    else atomicAdd(db + (j - 8), inv * s);
  }
}

// sums of a small-channel fp32 tensor [N, Cp] (+ optional products with the channels of a second tensor):
//   db[c] += inv * sum_px g[px,c]                                   (pyramid conv bias)
__global__ void __launch_bounds__(256)
small_col_sums_kernel(const float* __restrict__ g, int64_t n_px, int Cp, float inv, float* __restrict__ db) {
  __shared__ float red[256][4];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t p = blockIdx.x * 256ll + threadIdx.x; p < n_px; p += 256ll * gridDim.x)
    for (int k = 0; k < Cp; ++k) acc[k] += g[p * Cp + k];
#pragma unroll
  for (int k = 0; k < 4; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < Cp) {
    float s = 0.f;
    for (int i = 0; i < 256; ++i) s += red[i][threadIdx.x];
    atomicAdd(db + threadIdx.x, inv * s);
  }
}

// Combine backward (layerspp.py:52-59): dW[c][k] += inv * sum_px g[px,c] * pyr[px,k],  db[c] += inv * sum_px g[px,c]
// block = 64 channels x 4 pixel lanes over a chunk of pixels; fp32 atomics at the end.
__global__ void __launch_bounds__(256)
combine_bwd_kernel(const float* __restrict__ g, const float* __restrict__ pyr, int Cp, int64_t n_px, int C, int64_t px_per_block,
                   float inv, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float red[256][5];
  const int cl = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  const int64_t p0 = blockIdx.x * px_per_block, p1 = min(n_px, p0 + px_per_block);
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t p = p0 + pl; p < p1; p += 4) {
    const float gv = g[p * C + c];
    for (int k = 0; k < Cp; ++k) acc[k] = fmaf(gv, pyr[p * Cp + k], acc[k]);
    acc[4] += gv;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  if (pl == 0) {
    for (int k = 0; k < 5; ++k) {
      const float s = red[cl][k] + red[64 + cl][k] + red[128 + cl][k] + red[192 + cl][k];
      if (k < Cp) atomicAdd(dw + c * Cp + k, inv * s);
      else if (k == 4) atomicAdd(db + c, inv * s);
    }
  }
}

// packed dgrad weights of the pyramid conv C -> Cp (3x3) as ONE K-block of 64 over the im2col'd pyramid gradient:
//   out[c][k] = W[o][c][2-kf'][2-kt']  with k = tap' * Cp + o, tap' = kf'*3 + kt'   (zero for k >= 9*Cp)
__global__ void pack_pyr_dgrad_kernel(const float* __restrict__ w, int C, int Cp, op_t* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= C * 64) return;
  const int k = i % 64, c = i / 64;
  float v = 0.f;
  if (k < 9 * Cp) {
    const int tap = k / Cp, o = k % Cp;
    v = w[(static_cast<int64_t>(o) * C + c) * 9 + (8 - tap)];
  }
  out[i] = f2op(v);
}

// ------------------------------------------------------------------------------------------------
// Dense_0 (FiLM) backward: d[b][r] = gradient w.r.t. the projection outputs, act[b][k] = SiLU(temb)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_bwd_w_kernel(const float* __restrict__ d, const float* __restrict__ act, int B, int K, int rows, float* __restrict__ dw,
                   float* __restrict__ db) {
  const int64_t total = static_cast<int64_t>(rows) * K;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int k = static_cast<int>(i % K), r = static_cast<int>(i / K);
    float s = 0.f, sb = 0.f;
    for (int b = 0; b < B; ++b) { const float dv = d[static_cast<int64_t>(b) * rows + r]; s = fmaf(dv, act[b * K + k], s); sb += dv; }
    dw[i] += s;
    if (k == 0) db[r] += sb;
  }
}
// g_act[b][k] = sum_r d[b][r] * w[r][k]: block = (256 k-values, a run of rows); the d values of up to 8 utterances for the
// block's rows sit in shared memory, so every weight is loaded once per 8 utterances (8 FMAs per load, 8 independent
// accumulators); fp32 atomics into a zeroed buffer combine the row splits.
constexpr int DBA_ROWS = 128;
__global__ void __launch_bounds__(256)
dense_bwd_act_kernel(const float* __restrict__ d, const float* __restrict__ w, int B, int K, int rows, float* __restrict__ g_act) {
  __shared__ float sd[8][DBA_ROWS];
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * DBA_ROWS, nr = min(DBA_ROWS, rows - r0);
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = min(8, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * DBA_ROWS; i += 256) {
      const int j = i / DBA_ROWS, r = i % DBA_ROWS;
      sd[j][r] = (j < nb && r < nr) ? d[static_cast<int64_t>(b0 + j) * rows + r0 + r] : 0.f;
    }
    __syncthreads();
    if (k < K) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      const float* wp = w + static_cast<int64_t>(r0) * K + k;
#pragma unroll 4
      for (int r = 0; r < nr; ++r) {
        const float wv = __ldg(wp + static_cast<int64_t>(r) * K);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(sd[j][r], wv, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nb) atomicAdd(g_act + static_cast<int64_t>(b0 + j) * K + k, acc[j]);
    }
  }
}

// time-embedding MLP backward (layerspp.py:32-41, ncsnpp_v2.py:108-113,252-270): one block per utterance recomputes the
// forward (emb -> z1 -> h1 = SiLU -> z2 -> act = SiLU) and accumulates dW1, db1, dW2, db2 with fp32 atomics.
__global__ void __launch_bounds__(512)
temb_bwd_kernel(const float* __restrict__ t, const float* __restrict__ fw, int nf, const float* __restrict__ w1,
                const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, int t_stride,
                const float* __restrict__ g_act, float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                float* __restrict__ db2) {
  extern __shared__ float sm[];          // emb[2nf] | z1[4nf] | h1[4nf] | gz2[4nf] | gz1[4nf]
  const int D = 4 * nf, E = 2 * nf;
  float* emb = sm; float* z1 = emb + E; float* h1 = z1 + D; float* gz2 = h1 + D; float* gz1 = gz2 + D;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float lt = static_cast<float>(log(static_cast<double>(t[b * t_stride])));
  for (int j = threadIdx.x; j < nf; j += blockDim.x) {
    const float proj = __fmul_rn(__fmul_rn(__fmul_rn(lt, fw[j]), 2.0f), 3.14159274101257324f);
    emb[j] = sinf(proj);
    emb[nf + j] = cosf(proj);
  }
  __syncthreads();
  for (int r = warp; r < D; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < E; k += 32) acc = fmaf(w1[r * E + k], emb[k], acc);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) { z1[r] = acc + b1[r]; h1[r] = silu_f(acc + b1[r]); }
  }
  __syncthreads();
  for (int r = warp; r < D; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(w2[r * D + k], h1[k], acc);
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) gz2[r] = g_act[static_cast<int64_t>(b) * D + r] * silu_grad(acc + b2[r]);
  }
  __syncthreads();
  // dW2[r][k] += gz2[r] * h1[k], db2[r] += gz2[r];  g_h1[k] = sum_r w2[r][k] gz2[r]
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) atomicAdd(dw2 + i, gz2[i / D] * h1[i % D]);
  for (int r = threadIdx.x; r < D; r += blockDim.x) atomicAdd(db2 + r, gz2[r]);
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < D; ++r) s = fmaf(w2[r * D + k], gz2[r], s);
    gz1[k] = s * silu_grad(z1[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D * E; i += blockDim.x) atomicAdd(dw1 + i, gz1[i / E] * emb[i % E]);
  for (int r = threadIdx.x; r < D; r += blockDim.x) atomicAdd(db1 + r, gz1[r]);
}

int grid_for(int64_t n) { return static_cast<int>(std::min<int64_t>(ceil_div64(n, 256), static_cast<int64_t>(num_sms()) * 16)); }

}  // namespace

int launch_output_layer_bwd(const float* g_out, const float* pyr, int Cp, const float* w, int B, int T, int F, int F_out, float inv,
                            float* g_pyr, float* dw, float* db, cudaStream_t s) {
  FDBM_REQUIRE(Cp <= 4, "output_layer_bwd: Cp must be <= 4");
  dim3 grid(ceil_div(T, 32), ceil_div(F, 32), B);
  output_layer_bwd_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(g_out), pyr, Cp, w, T, F, F_out, inv, g_pyr, dw, db);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_small_col_sums(const float* g, int64_t n_px, int Cp, float inv, float* db, cudaStream_t s) {
  small_col_sums_kernel<<<std::min(grid_for(n_px), 512), 256, 0, s>>>(g, n_px, Cp, inv, db);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_combine_bwd(const float* g, const float* pyr, int Cp, int64_t n_px, int C, float inv, float* dw, float* db, cudaStream_t s) {
  FDBM_REQUIRE(C % 64 == 0 && Cp <= 4, "combine_bwd: unsupported channels");
  const int64_t bx = std::max<int64_t>(1, std::min<int64_t>(n_px / 64, (static_cast<int64_t>(num_sms()) * 8) / (C / 64)));
  const int64_t ppb = ceil_div64(n_px, bx);
  dim3 grid(static_cast<unsigned>(ceil_div64(n_px, ppb)), C / 64);
  combine_bwd_kernel<<<grid, 256, 0, s>>>(g, pyr, Cp, n_px, C, ppb, inv, dw, db);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_pack_pyr_dgrad(const float* w, int C, int Cp, op_t* out, cudaStream_t s) {
  pack_pyr_dgrad_kernel<<<ceil_div(C * 64, 256), 256, 0, s>>>(w, C, Cp, out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

// all Dense_0 layers at once: dW[rows][K] += d^T act, db[rows] += sum_b d, g_act = d W; then the embedding MLP
int launch_dense_temb_bwd(const float* d, const float* act, const float* w, int B, int K, int rows, float* dw, float* db, float* g_act,
                          const float* t, const float* fw, int nf, const float* w1, const float* b1, const float* w2, const float* b2,
                          int t_stride, float* dw1, float* db1, float* dw2, float* db2, cudaStream_t s) {
  dense_bwd_w_kernel<<<grid_for(static_cast<int64_t>(rows) * K), 256, 0, s>>>(d, act, B, K, rows, dw, db);
  FDBM_LAUNCH_CHECK();
  FDBM_CUDA(cudaMemsetAsync(g_act, 0, sizeof(float) * B * K, s));
  dense_bwd_act_kernel<<<dim3(ceil_div(K, 256), ceil_div(rows, DBA_ROWS)), 256, 0, s>>>(d, w, B, K, rows, g_act);
  FDBM_LAUNCH_CHECK();
  temb_bwd_kernel<<<B, 512, sizeof(float) * (2 * nf + 16 * nf), s>>>(t, fw, nf, w1, b1, w2, b2, t_stride, g_act, dw1, db1, dw2, db2);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

}  // namespace fdbm
