// "Skinny" layers of NCSN++ whose GEMM K or N is <= 4: HBM-bound, CUDA-core kernels.
//   pack_input     ncsnpp_v2.py:247-250   complex [B,1,257,T] x,y -> fp32 [B,T,256,4]
//   im2col_input   ncsnpp_v2.py:278       conv3x3 4 -> nf as one 64-wide K-block for the tensor-core kernel
//   combine        layerspp.py:52-59      h += conv1x1(4 -> C)(input pyramid)
//   (the C -> 4 pyramid convolutions run in conv_igemm with MMA N = 16 and a progressive-output epilogue)
//   output_layer   ncsnpp_v2.py:392-399   conv1x1(4 -> 2) -> complex [B,1,257,T], Nyquist row = 0
// Reference conv weights are OIHW with H = frequency, W = frames; activations here are [B,T,F,C].
#include "common.cuh"

namespace fdbm {
namespace {

// ------------------------------------------------------------------------------------------------
// pack_input: transpose through a 32(f) x 32(t) shared-memory tile.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_input_kernel(const float2* __restrict__ x, const float2* __restrict__ y, int T, int F_in, int F, int Cin,
                  float* __restrict__ out) {
  __shared__ float4 tile[32][33];
  const int b = blockIdx.z;
  const int f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {            // r: frequency row, tx: frame
    const int f = f0 + r, t = t0 + tx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f < F && t < T) {
      const int64_t idx = (static_cast<int64_t>(b) * F_in + f) * T + t;
      const float2 xv = x[idx];
      v.x = xv.x; v.y = xv.y;
      if (y) { const float2 yv = y[idx]; v.z = yv.x; v.w = yv.y; }
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {            // r: frame, tx: frequency
    const int t = t0 + r, f = f0 + tx;
    if (f < F && t < T) {
      const float4 v = tile[tx][r];
      float* o = out + ((static_cast<int64_t>(b) * T + t) * F + f) * Cin;
      if (Cin == 4) *reinterpret_cast<float4*>(o) = v;
      else *reinterpret_cast<float2*>(o) = make_float2(v.x, v.y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// im2col of the 4 (or 2) channel input for the first 3x3 convolution: [B,T,F,Cin] fp32 -> [B,T,F,64] h16
// with k = (kf*3 + kt)*Cin + ci (zero border, zero for k >= 9*Cin).  The convolution itself then runs on
// the tensor cores as a K=64 GEMM with the fused-statistics epilogue (conv_igemm, ksize 1).
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256)
im2col_input_kernel(const float* __restrict__ in, int B, int T, int F, uint4* __restrict__ out) {
  constexpr int TPG = 8 / CIN;                                         // taps per group of 8 k-values
  const int total = B * T * F * 8;                                     // 8 groups of 8 k-values per pixel (host checks < 2^31)
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += 256 * gridDim.x) {
    const int g = i & 7;
    const int p = i >> 3;
    const int f = p % F;
    const int bt = p / F;
    const int t = bt % T;
    float v[8];
#pragma unroll
    for (int u = 0; u < TPG; ++u) {
      const int tap = g * TPG + u;
      const int ff = f + tap / 3 - 1, tt = t + tap % 3 - 1;
      const bool ok = tap < 9 && ff >= 0 && ff < F && tt >= 0 && tt < T;
      const float* src = in + (static_cast<int64_t>(bt - t + tt) * F + ff) * CIN;
      if (CIN == 4) {
        const float4 x = ok ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * u] = x.x; v[4 * u + 1] = x.y; v[4 * u + 2] = x.z; v[4 * u + 3] = x.w;
      } else {
        const float2 x = ok ? __ldg(reinterpret_cast<const float2*>(src)) : make_float2(0.f, 0.f);
        v[2 * u] = x.x; v[2 * u + 1] = x.y;
      }
    }
    out[i] = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
  }
}

// ------------------------------------------------------------------------------------------------
// combine: h[b,p,c] += sum_k w[c][k] * pyr[b,p,k] + bias[c]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
combine_kernel(float4* __restrict__ h, const float* __restrict__ pyr, int Cp, const float* __restrict__ w,
               const float* __restrict__ bias, int64_t n_px, int C) {
  const int cg = C / 4;
  const int64_t total = n_px * cg;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int g = static_cast<int>(i % cg);
    const int64_t p = i / cg;
    float4 v = h[i];
    float o[4] = {v.x, v.y, v.z, v.w};
    float pv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < Cp; ++k) pv[k] = pyr[p * Cp + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = g * 4 + j;
      float acc = bias[c];
      for (int k = 0; k < Cp; ++k) acc = fmaf(w[c * Cp + k], pv[k], acc);
      o[j] += acc;
    }
    h[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// output_layer: per pixel 2 outputs from Cp inputs, transposed to the reference layout.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
output_layer_kernel(const float* __restrict__ pyr, int Cp, const float* __restrict__ w, const float* __restrict__ bias,
                    int T, int F, int F_out, float2* __restrict__ out) {
  __shared__ float2 tile[32][33];
  const int b = blockIdx.z;
  const int f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {            // r: frame, tx: frequency
    const int t = t0 + r, f = f0 + tx;
    float2 v = make_float2(0.f, 0.f);
    if (t < T && f < F) {
      const float* p = pyr + ((static_cast<int64_t>(b) * T + t) * F + f) * Cp;
      float re = bias[0], im = bias[1];
      for (int k = 0; k < Cp; ++k) { re = fmaf(w[k], p[k], re); im = fmaf(w[Cp + k], p[k], im); }
      v = make_float2(re, im);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {            // r: frequency, tx: frame
    const int f = f0 + r, t = t0 + tx;
    if (t < T && f < F_out) out[(static_cast<int64_t>(b) * F_out + f) * T + t] = f < F ? tile[tx][r] : make_float2(0.f, 0.f);
  }
}

}  // namespace

int launch_pack_input(const float* x, const float* y, int B, int T, int F_in, int F, int Cin, float* out,
                      cudaStream_t s) {
  FDBM_REQUIRE(Cin == 4 || Cin == 2, "pack_input: Cin must be 2 or 4");
  dim3 grid(ceil_div(T, 32), ceil_div(F, 32), B);
  pack_input_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(y), T, F_in,
                                         F, Cin, out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_im2col_input(const float* in, int Cin, int B, int T, int F, op_t* out, cudaStream_t s) {
  FDBM_REQUIRE(Cin == 4 || Cin == 2, "im2col_input: Cin must be 2 or 4");
  const int64_t total = static_cast<int64_t>(B) * T * F * 8;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(total, 256), static_cast<int64_t>(num_sms()) * 16));
  FDBM_REQUIRE(total < (1ll << 31), "im2col_input: tensor too large");
  if (Cin == 4) im2col_input_kernel<4><<<grid, 256, 0, s>>>(in, B, T, F, reinterpret_cast<uint4*>(out));
  else im2col_input_kernel<2><<<grid, 256, 0, s>>>(in, B, T, F, reinterpret_cast<uint4*>(out));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_combine(float* h, const float* pyr, int Cp, const float* w, const float* bias, int B, int T, int F, int C,
                   cudaStream_t s) {
  FDBM_REQUIRE(Cp <= 4 && C % 4 == 0, "combine: unsupported channels");
  const int64_t n_px = static_cast<int64_t>(B) * T * F;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(n_px * (C / 4), 256), static_cast<int64_t>(num_sms()) * 16));
  combine_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(h), pyr, Cp, w, bias, n_px, C);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_output_layer(const float* pyr, int Cp, const float* w, const float* bias, int B, int T, int F, int F_out,
                        float* out, cudaStream_t s) {
  dim3 grid(ceil_div(T, 32), ceil_div(F_out, 32), B);
  output_layer_kernel<<<grid, 256, 0, s>>>(pyr, Cp, w, bias, T, F, F_out, reinterpret_cast<float2*>(out));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

}  // namespace fdbm

using namespace fdbm;

// C ABI of the skinny layers (kernel-level parity tests and other hosts; the backbone plan calls the launchers directly)
extern "C" int fdbm_pack_input(const float* x, const float* y, int batch, int n_frames, int f_in, int f_used, int c_in, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x && out && batch > 0 && n_frames > 0 && f_used > 0 && f_used <= f_in && (c_in == 2 || (c_in == 4 && y)), "fdbm_pack_input: bad arguments");
  return launch_pack_input(x, c_in == 4 ? y : nullptr, batch, n_frames, f_in, f_used, c_in, out, as_stream(stream));
}
extern "C" int fdbm_im2col_input(const float* in, int c_in, int batch, int n_frames, int n_freq, void* out_h16, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(in && out_h16 && batch > 0 && n_frames > 0 && n_freq > 0, "fdbm_im2col_input: bad arguments");
  return launch_im2col_input(in, c_in, batch, n_frames, n_freq, reinterpret_cast<op_t*>(out_h16), as_stream(stream));
}
extern "C" int fdbm_combine(float* h, const float* pyramid, int c_pyr, const float* weight, const float* bias, int batch, int n_frames, int n_freq,
                            int channels, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(h && pyramid && weight && bias && batch > 0 && n_frames > 0 && n_freq > 0 && c_pyr >= 1, "fdbm_combine: bad arguments");
  return launch_combine(h, pyramid, c_pyr, weight, bias, batch, n_frames, n_freq, channels, as_stream(stream));
}
extern "C" int fdbm_output_layer(const float* pyramid, int c_pyr, const float* weight, const float* bias, int batch, int n_frames, int n_freq,
                                 int f_out, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(pyramid && weight && bias && out && batch > 0 && n_frames > 0 && n_freq > 0 && f_out >= n_freq && c_pyr >= 1 && c_pyr <= 4,
               "fdbm_output_layer: bad arguments");
  return launch_output_layer(pyramid, c_pyr, weight, bias, batch, n_frames, n_freq, f_out, out, as_stream(stream));
}
