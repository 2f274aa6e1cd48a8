// "Skinny" layers of NCSN++ whose GEMM K or N is <= 4: HBM-bound, CUDA-core kernels.
//   pack_input     ncsnpp_v2.py:247-250   complex [B,1,257,T] x,y -> fp32 [B,T,256,4]
//   conv_in        ncsnpp_v2.py:278       conv3x3 4 -> nf
//   combine        layerspp.py:52-59      h += conv1x1(4 -> C)(input pyramid)
//   pyramid_conv   ncsnpp_v2.py:338-359   pyramid = FIR-up(pyramid) + conv3x3(C -> 4)(act)
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
// conv_in: a warp computes one pixel; lane l owns output channels 4l..4l+3 (+128 per repeat).
// Weights sit in shared memory as [tap*Cin + ci][Cout].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ in, int Cin, const float* __restrict__ w, const float* __restrict__ bias,
               int B, int T, int F, int Cout, float* __restrict__ out) {
  extern __shared__ float ws[];                  // [9*Cin][Cout]
  for (int i = threadIdx.x; i < 9 * Cin * Cout; i += 256) {
    const int co = i % Cout, k = i / Cout;
    const int ci = k % Cin, tap = k / Cin;
    const int kf = tap / 3, kt = tap % 3;         // OIHW: H = frequency, W = frames
    ws[i] = w[((co * Cin + ci) * 3 + kf) * 3 + kt];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_px = static_cast<int64_t>(B) * T * F;
  for (int64_t p = blockIdx.x * 8ll + warp; p < n_px; p += 8ll * gridDim.x) {
    const int f = static_cast<int>(p % F);
    const int t = static_cast<int>((p / F) % T);
    const int b = static_cast<int>(p / (static_cast<int64_t>(F) * T));
    float xin[36];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ff = f + tap / 3 - 1, tt = t + tap % 3 - 1;
      const bool ok = ff >= 0 && ff < F && tt >= 0 && tt < T;
      const float* src = in + ((static_cast<int64_t>(b) * T + tt) * F + ff) * Cin;
      for (int ci = 0; ci < 4; ++ci) xin[tap * 4 + ci] = (ok && ci < Cin) ? src[ci] : 0.f;
    }
    for (int c0 = lane * 4; c0 < Cout; c0 += 128) {
      float4 acc = *reinterpret_cast<const float4*>(bias + c0);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        for (int ci = 0; ci < Cin; ++ci) {
          const float4 wv = *reinterpret_cast<const float4*>(ws + (tap * Cin + ci) * Cout + c0);
          const float xv = xin[tap * 4 + ci];
          acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y);
          acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
        }
      }
      *reinterpret_cast<float4*>(out + p * Cout + c0) = acc;
    }
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
// pyramid_conv: a warp computes one pixel; lane l owns VEC = C/32 consecutive input channels.
// Shared weights are laid out [tap][j][lane][Cp] so that a warp reads consecutive float4s.
// ------------------------------------------------------------------------------------------------
template <int VEC, int CP>
__global__ void __launch_bounds__(256)
pyramid_conv_kernel(const op_t* __restrict__ act, const float* __restrict__ w, const float* __restrict__ bias,
                    const float* __restrict__ prev, int B, int T, int F, float* __restrict__ out) {
  constexpr int C = VEC * 32;
  extern __shared__ float ws[];                  // [9][VEC][32][CP]
  for (int i = threadIdx.x; i < 9 * C * CP; i += 256) {
    const int o = i % CP;
    const int l = (i / CP) % 32;
    const int j = (i / (CP * 32)) % VEC;
    const int tap = i / (CP * 32 * VEC);
    const int c = l * VEC + j;
    const int kf = tap / 3, kt = tap % 3;
    ws[i] = w[((o * C + c) * 3 + kf) * 3 + kt];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_px = static_cast<int64_t>(B) * T * F;
  for (int64_t p = blockIdx.x * 8ll + warp; p < n_px; p += 8ll * gridDim.x) {
    const int f = static_cast<int>(p % F);
    const int t = static_cast<int>((p / F) % T);
    const int b = static_cast<int>(p / (static_cast<int64_t>(F) * T));
    float acc[CP];
#pragma unroll
    for (int o = 0; o < CP; ++o) acc[o] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ff = f + tap / 3 - 1, tt = t + tap % 3 - 1;
      if (ff < 0 || ff >= F || tt < 0 || tt >= T) continue;
      const op_t* src = act + ((static_cast<int64_t>(b) * T + tt) * F + ff) * C + lane * VEC;
      float xv[VEC];
      if (VEC == 8) {
        const uint4 raw = *reinterpret_cast<const uint4*>(src);
        const op2_t* h2 = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f2 = op22f2(h2[j]); xv[2 * j] = f2.x; xv[2 * j + 1] = f2.y; }
      } else {
        const uint2 raw = *reinterpret_cast<const uint2*>(src);
        const op2_t* h2 = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
        for (int j = 0; j < 2; ++j) { const float2 f2 = op22f2(h2[j]); xv[2 * j] = f2.x; xv[2 * j + 1] = f2.y; }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float* wp = ws + ((tap * VEC + j) * 32 + lane) * CP;
#pragma unroll
        for (int o = 0; o < CP; ++o) acc[o] = fmaf(xv[j], wp[o], acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < CP; ++o) {
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], sft);
    }
    if (lane < CP) {
      float v = bias[lane];
#pragma unroll
      for (int o = 0; o < CP; ++o) if (o == lane) v += acc[o];
      if (prev) {                                  // + FIR-up x2 of the coarser pyramid level at (t, f)
        const int Tp = T / 2, Fp = F / 2;
        const int it = t >> 1, jf = f >> 1;
        const int t_a = (t & 1) ? it : it - 1, t_b = t_a + 1;
        const float wt_a = (t & 1) ? 0.75f : 0.25f, wt_b = 1.0f - wt_a;
        const int f_a = (f & 1) ? jf : jf - 1, f_b = f_a + 1;
        const float wf_a = (f & 1) ? 0.75f : 0.25f, wf_b = 1.0f - wf_a;
        const float* pb = prev + static_cast<int64_t>(b) * Tp * Fp * CP + lane;
        float up = 0.f;
        if (t_a >= 0 && t_a < Tp) {
          if (f_a >= 0 && f_a < Fp) up += wt_a * wf_a * pb[(static_cast<int64_t>(t_a) * Fp + f_a) * CP];
          if (f_b >= 0 && f_b < Fp) up += wt_a * wf_b * pb[(static_cast<int64_t>(t_a) * Fp + f_b) * CP];
        }
        if (t_b >= 0 && t_b < Tp) {
          if (f_a >= 0 && f_a < Fp) up += wt_b * wf_a * pb[(static_cast<int64_t>(t_b) * Fp + f_a) * CP];
          if (f_b >= 0 && f_b < Fp) up += wt_b * wf_b * pb[(static_cast<int64_t>(t_b) * Fp + f_b) * CP];
        }
        v += up;
      }
      out[p * CP + lane] = v;
    }
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

int launch_conv_in(const float* in, int Cin, const float* w, const float* bias, int B, int T, int F, int Cout,
                   float* out, cudaStream_t s) {
  FDBM_REQUIRE(Cin <= 4 && Cout % 128 == 0, "conv_in: unsupported channels %d -> %d", Cin, Cout);
  const size_t smem = sizeof(float) * 9 * Cin * Cout;
  const int64_t n_px = static_cast<int64_t>(B) * T * F;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(n_px, 8), static_cast<int64_t>(num_sms()) * 8));
  conv_in_kernel<<<grid, 256, smem, s>>>(in, Cin, w, bias, B, T, F, Cout, out);
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

template <int VEC, int CP>
static int launch_pyr(const op_t* act, const float* w, const float* bias, const float* prev, int B, int T,
                      int F, float* out, cudaStream_t s) {
  const size_t smem = sizeof(float) * 9 * VEC * 32 * CP;
  static PerDeviceOnce attr_once;
  if (attr_once.first(current_device()))
    FDBM_CUDA(cudaFuncSetAttribute(pyramid_conv_kernel<VEC, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int64_t n_px = static_cast<int64_t>(B) * T * F;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(n_px, 8), static_cast<int64_t>(num_sms()) * 6));
  pyramid_conv_kernel<VEC, CP><<<grid, 256, smem, s>>>(act, w, bias, prev, B, T, F, out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_pyramid_conv(const op_t* act, int C, const float* w, const float* bias, const float* prev, int Cp,
                        int B, int T, int F, float* out, cudaStream_t s) {
  FDBM_REQUIRE((C == 128 || C == 256) && (Cp == 4 || Cp == 2), "pyramid_conv: unsupported channels %d -> %d", C, Cp);
  FDBM_REQUIRE(prev == nullptr || (T % 2 == 0 && F % 2 == 0), "pyramid_conv: odd size with a coarser level");
  if (C == 128 && Cp == 4) return launch_pyr<4, 4>(act, w, bias, prev, B, T, F, out, s);
  if (C == 256 && Cp == 4) return launch_pyr<8, 4>(act, w, bias, prev, B, T, F, out, s);
  if (C == 128 && Cp == 2) return launch_pyr<4, 2>(act, w, bias, prev, B, T, F, out, s);
  return launch_pyr<8, 2>(act, w, bias, prev, B, T, F, out, s);
}

int launch_output_layer(const float* pyr, int Cp, const float* w, const float* bias, int B, int T, int F, int F_out,
                        float* out, cudaStream_t s) {
  dim3 grid(ceil_div(T, 32), ceil_div(F_out, 32), B);
  output_layer_kernel<<<grid, 256, 0, s>>>(pyr, Cp, w, bias, T, F, F_out, reinterpret_cast<float2*>(out));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

}  // namespace fdbm
