// TF-GridNet backbones (fdbm/backbones/tfgridnet.py:126-229 TFGridNet, :236-427 GridNetV3Block, :430-484 the two
// normalisation layers; tfgridnet_predictive.py) -- the backbones config.yaml / config_predictive.yaml select.
//
// 97 % of the arithmetic is the ten bidirectional LSTM sweeps of a forward (intra: one sequence of 260 unfolded positions
// along frequency per (utterance, frame); inter: one sequence along time per (utterance, bin)); a 4 s utterance is
// ~250 GFLOP.  The sweep is the tcgen05 cluster kernel in tfg_lstm_tc.cu; this file holds everything around it: the fused
// pad + time-embedding add + LayerNorm pass, the pass after a sweep (bias + both directions + residual + next LayerNorm / crop),
// the 3x3 input conv + GroupNorm, the time embedding, the attention (1x1 convs / PReLU-LayerNorms, two batched tensor-core
// GEMMs for Q K^T and P V, projection) and the output ConvTranspose2d -- all HBM / latency bound.
#include <mma.h>
#include "common.cuh"

namespace fdbm {
namespace {

constexpr int TC = 32;            // emb_dim
constexpr int OLP = 3;            // emb_ks - emb_hs
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// ---- per-position passes ---------------------------------------------------------------------------------------------
// LayerNorm over the 32 channels of one position: one warp per position, lane = channel
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Position-wise kernels below: 8 lanes own one position (4 channels each, 16-byte accesses), a warp four positions per
// iteration, so a warp instruction moves 512 contiguous bytes and LayerNorm(32) is a 3-step shuffle inside the 8-lane group.
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float4 ln32_q(float4 v, float4 ga, float4 be, float eps) {
  const float mu = group8_sum(v.x + v.y + v.z + v.w) * (1.0f / 32.0f);
  const float4 d = make_float4(v.x - mu, v.y - mu, v.z - mu, v.w - mu);
  const float r = rsqrtf(group8_sum(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w) * (1.0f / 32.0f) + eps);
  return make_float4(d.x * r * ga.x + be.x, d.y * r * ga.y + be.y, d.z * r * ga.z + be.z, d.w * r * ga.w + be.w);
}
__device__ __forceinline__ uint2 pack_h4(float4 v) { return make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w)); }
__device__ __forceinline__ float4 unpack_h4(uint2 u) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// xp[b, t', q', c] = (inside ? h[b, t, q, c] + emb[b, c] : 0);  xn = fp16 LayerNorm(xp)   (tfgridnet.py:206, :329-334)
__global__ void __launch_bounds__(256)
pad_add_norm_kernel(const float* __restrict__ hcur, const float* __restrict__ emb, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int B, int T, int Q, float eps, float* __restrict__ xp, __half* __restrict__ xn) {
  const int Tp = T + 2 * OLP, Qp = Q + 2 * OLP;
  const int64_t n_pos = static_cast<int64_t>(B) * Tp * Qp;
  const int lane = threadIdx.x & 31, cq = lane & 7;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cq), be = __ldg(reinterpret_cast<const float4*>(beta) + cq);
  for (int64_t p = (blockIdx.x * 8ll + (threadIdx.x >> 5)) * 4 + (lane >> 3); p < n_pos; p += 32ll * gridDim.x) {
    const int qp = static_cast<int>(p % Qp), tp = static_cast<int>((p / Qp) % Tp), b = static_cast<int>(p / (static_cast<int64_t>(Qp) * Tp));
    const int t = tp - OLP, q = qp - OLP;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T && q >= 0 && q < Q) {
      v = __ldg(reinterpret_cast<const float4*>(hcur + ((static_cast<int64_t>(b) * T + t) * Q + q) * TC) + cq);
      if (emb) { const float4 e = __ldg(reinterpret_cast<const float4*>(emb + b * TC) + cq); v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w; }
    }
    reinterpret_cast<float4*>(xp + p * TC)[cq] = v;
    reinterpret_cast<uint2*>(xn + p * TC)[cq] = pack_h4(ln32_q(v, ga, be, eps));
  }
}

// After a sweep: out[b, t', q', c] = bias[c] + yf + yb + resid, where yf / yb are the sweep's per-direction ConvTranspose1d
// outputs [seq][pos][32] (taps already overlap-added; tfgridnet.py:346-350 / :371-375).  intra (mode 0, seq = (b, t'), pos = q'):
// writes the full padded tensor and its LayerNorm for the inter sweep.  inter (mode 1, seq = (b, q'), pos = t'): writes only the
// un-padded crop [B,T,Q,C] (:381).
__global__ void __launch_bounds__(256)
sweep_post_kernel(const __half* __restrict__ yf, const __half* __restrict__ yb, const float* __restrict__ lin_bias, const float* __restrict__ resid,
                  int B, int T, int Q, int mode, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                  float* __restrict__ out_full, __half* __restrict__ xn, float* __restrict__ out_crop, int xn_transposed) {
  const int Tp = T + 2 * OLP, Qp = Q + 2 * OLP;
  const int lane = threadIdx.x & 31, cq = lane & 7;
  const float4 lb = __ldg(reinterpret_cast<const float4*>(lin_bias) + cq);
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 ga = gamma ? __ldg(reinterpret_cast<const float4*>(gamma) + cq) : one, be = beta ? __ldg(reinterpret_cast<const float4*>(beta) + cq) : zero;
  const int64_t n_pos = mode == 0 ? static_cast<int64_t>(B) * Tp * Qp : static_cast<int64_t>(B) * T * Q;
  for (int64_t p = (blockIdx.x * 8ll + (threadIdx.x >> 5)) * 4 + (lane >> 3); p < n_pos; p += 32ll * gridDim.x) {
    int b, tp, qp;
    if (mode == 0) { qp = static_cast<int>(p % Qp); tp = static_cast<int>((p / Qp) % Tp); b = static_cast<int>(p / (static_cast<int64_t>(Qp) * Tp)); }
    else { qp = static_cast<int>(p % Q) + OLP; tp = static_cast<int>((p / Q) % T) + OLP; b = static_cast<int>(p / (static_cast<int64_t>(Q) * T)); }
    const int64_t pfull = (static_cast<int64_t>(b) * Tp + tp) * Qp + qp;
    const int64_t ptr = (static_cast<int64_t>(b) * Qp + qp) * Tp + tp;
    const int64_t py = mode == 0 ? pfull : ptr;
    const float4 f = unpack_h4(__ldg(reinterpret_cast<const uint2*>(yf + py * TC) + cq)), r = unpack_h4(__ldg(reinterpret_cast<const uint2*>(yb + py * TC) + cq));
    const float4 x = __ldg(reinterpret_cast<const float4*>(resid + pfull * TC) + cq);
    const float4 acc = make_float4(lb.x + f.x + r.x + x.x, lb.y + f.y + r.y + x.y, lb.z + f.z + r.z + x.z, lb.w + f.w + r.w + x.w);
    if (mode == 0) {
      reinterpret_cast<float4*>(out_full + pfull * TC)[cq] = acc;
      // xn_transposed: [B, Q', T', C], the inter sweep's sequences (one per padded bin) contiguous along time
      reinterpret_cast<uint2*>(xn + (xn_transposed ? ptr : pfull) * TC)[cq] = pack_h4(ln32_q(acc, ga, be, eps));
    } else {
      reinterpret_cast<float4*>(out_crop + p * TC)[cq] = acc;
    }
  }
}

// input conv 3x3 (Cin = 4: x.re, x.im, y.re, y.im; or 2) over the [T, F] plane of complex [B,1,F,T] inputs (tfgridnet.py:152,201-214)
// + the sums for GroupNorm(1, C) (a LayerNorm over the whole (C, T, F) of an utterance).
// A block owns 32 frames x 8 bins of one utterance: the input patch (frames contiguous in memory) is staged in shared memory,
// a thread computes the 32 channels of one position, and the outputs leave through shared memory so that a warp store is one
// position's 128 contiguous bytes.  (First version: one warp per position, 784 us at B = 16.)
constexpr int IC_TT = 32, IC_TQ = 8;
__global__ void __launch_bounds__(256)
tfg_input_conv_kernel(const float2* __restrict__ x, const float2* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bias,
                      int B, int T, int Q, int Cin, float* __restrict__ out, double* __restrict__ sums) {
  __shared__ __align__(16) float sw[4 * 9 * TC];                 // [ci * 9 + tap][c]
  __shared__ float2 sx[2][IC_TQ + 2][IC_TT + 2];                  // [x | y][bin][frame]
  __shared__ float so[256][TC + 1];
  __shared__ double red[2][8];
  // conv weight [C, Cin, kh = time, kw = freq] on input [B, Cin, T, F]
  for (int i = threadIdx.x; i < TC * Cin * 9; i += 256) { const int c = i / (Cin * 9), r = i % (Cin * 9); sw[r * TC + c] = w[i]; }
  const int b = blockIdx.z, t0 = blockIdx.y * IC_TT, q0 = blockIdx.x * IC_TQ;
  const int n_src = Cin == 4 ? 2 : 1;
  for (int i = threadIdx.x; i < n_src * (IC_TQ + 2) * (IC_TT + 2); i += 256) {
    const int tl = i % (IC_TT + 2), ql = (i / (IC_TT + 2)) % (IC_TQ + 2), sidx = i / ((IC_TT + 2) * (IC_TQ + 2));
    const int t = t0 + tl - 1, q = q0 + ql - 1;
    float2 v = make_float2(0.f, 0.f);
    if (t >= 0 && t < T && q >= 0 && q < Q) v = __ldg((sidx ? y : x) + (static_cast<int64_t>(b) * Q + q) * T + t);
    sx[sidx][ql][tl] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int ql = threadIdx.x & 7, tl = threadIdx.x >> 3;          // position of this thread: bins fastest (the output's order)
  float acc[TC];
#pragma unroll
  for (int c = 0; c < TC; ++c) acc[c] = __ldg(bias + c);
#pragma unroll
  for (int dt = 0; dt < 3; ++dt)
#pragma unroll
    for (int dq = 0; dq < 3; ++dq) {
      float in[4];
      const float2 xv = sx[0][ql + dq][tl + dt];
      in[0] = xv.x; in[1] = xv.y;
      if (Cin == 4) { const float2 yv = sx[1][ql + dq][tl + dt]; in[2] = yv.x; in[3] = yv.y; }
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (ci >= Cin) break;
        const float4* wr = reinterpret_cast<const float4*>(sw + (ci * 9 + dt * 3 + dq) * TC);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 wv = wr[c4];
          acc[4 * c4] = fmaf(in[ci], wv.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(in[ci], wv.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(in[ci], wv.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(in[ci], wv.w, acc[4 * c4 + 3]);
        }
      }
    }
  const bool ok = t0 + tl < T && q0 + ql < Q;
  float f1 = 0.f, f2 = 0.f;
#pragma unroll
  for (int c = 0; c < TC; ++c) { so[threadIdx.x][c] = acc[c]; if (ok) { f1 += acc[c]; f2 = fmaf(acc[c], acc[c], f2); } }
  __syncthreads();
  // warp w stores positions 32 w .. 32 w + 31, one position (128 contiguous bytes) per instruction
  for (int r = 0; r < 32; ++r) {
    const int pos = wrp * 32 + r, pq = q0 + (pos & 7), pt = t0 + (pos >> 3);
    if (pt < T && pq < Q) out[((static_cast<int64_t>(b) * T + pt) * Q + pq) * TC + lane] = so[pos][lane];
  }
  double s1 = f1, s2 = f2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if (lane == 0) { red[0][wrp] = s1; red[1][wrp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; c += red[1][i]; }
    atomicAdd(sums + 2 * b, a); atomicAdd(sums + 2 * b + 1, c);
  }
}
__global__ void __launch_bounds__(256)
tfg_groupnorm1_kernel(float* __restrict__ h, const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                      int64_t per_utt, float eps) {
  const int b = blockIdx.y;
  const double n = static_cast<double>(per_utt);
  const double mu = sums[2 * b] / n, var = fmax(sums[2 * b + 1] / n - mu * mu, 0.0);
  const float m = static_cast<float>(mu), r = static_cast<float>(1.0 / sqrt(var + eps));
  float* p = h + static_cast<int64_t>(b) * per_utt;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < per_utt; i += 256ll * gridDim.x) {
    const int c = static_cast<int>(i & (TC - 1));
    p[i] = (p[i] - m) * r * __ldg(gamma + c) + __ldg(beta + c);
  }
}

// time embedding (tfgridnet.py:177-192, 203-219): Fourier features of log t -> Linear, SiLU, Linear, SiLU -> one Linear per block
__global__ void __launch_bounds__(128)
tfg_temb_kernel(const float* __restrict__ t, int t_stride, const float* __restrict__ fw, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ wb, const float* __restrict__ bb, int n_layers,
                float* __restrict__ emb /* [n_layers][B][32] */, int B) {
  __shared__ float f[64], h1[128], h2[128];
  const int b = blockIdx.x, i = threadIdx.x;
  const float lt = static_cast<float>(log(static_cast<double>(t[b * t_stride])));
  if (i < 32) {
    const float proj = __fmul_rn(__fmul_rn(__fmul_rn(lt, fw[i]), 2.0f), 3.14159274101257324f);
    f[i] = sinf(proj); f[32 + i] = cosf(proj);
  }
  __syncthreads();
  float acc = b1[i];
  for (int k = 0; k < 64; ++k) acc = fmaf(w1[i * 64 + k], f[k], acc);
  h1[i] = silu_f(acc);
  __syncthreads();
  acc = b2[i];
  for (int k = 0; k < 128; ++k) acc = fmaf(w2[i * 128 + k], h1[k], acc);
  h2[i] = silu_f(acc);
  __syncthreads();
  for (int o = i; o < n_layers * 32; o += 128) {
    const int l = o >> 5, c = o & 31;
    float s = bb[l * 32 + c];
    for (int k = 0; k < 128; ++k) s = fmaf(wb[(l * 32 + c) * 128 + k], h2[k], s);
    emb[(static_cast<int64_t>(l) * B + b) * 32 + c] = s;
  }
}

// attention front: 1x1 convs Q (8), K (8), V (32) + per-head PReLU + normalisation over the head's E channels + affine
// (tfgridnet.py:383-385, 458-484).  Outputs are the fp16 operands of the two batched GEMMs:
//   Qh, Kh [B*4][T][E*F] with feature e * F + f  (tfgridnet.py:392-396);  Vt [B*4][8*F][T] (feature f * 8 + c8, frames contiguous:
//   the order of V's features only has to match the un-flattening of P V, and this one makes those stores contiguous)
// A block owns 16 frames x 32 bins of one utterance; a thread owns one bin of two frames, so the whole per-position chain
// (48 dot products, the head statistics) is register-local, Q / K stores are contiguous along bins across the warp, and V goes
// through a shared-memory transpose so that Vt is written in 32-byte runs along frames.  (The first version, one warp per
// position with 2-byte scattered stores, took 3.2 ms per call at B = 16.)
constexpr int QKV_TT = 16, QKV_TQ = 32, QKV_PITCH = 18;
__global__ void __launch_bounds__(256)
tfg_qkv_kernel(const float* __restrict__ z, const float* __restrict__ wq, const float* __restrict__ bq, const float* __restrict__ wk,
               const float* __restrict__ bk, const float* __restrict__ wv, const float* __restrict__ bv, const float* __restrict__ aq,
               const float* __restrict__ ak, const float* __restrict__ av, const float* __restrict__ gq, const float* __restrict__ betq,
               const float* __restrict__ gk, const float* __restrict__ betk, const float* __restrict__ gv, const float* __restrict__ betv,
               int B, int T, int Q, float eps, int ldf, int ldt, __half* __restrict__ Qh, __half* __restrict__ Kh, __half* __restrict__ Vt) {
  __shared__ __align__(16) float sw[48 * 32];
  __shared__ float sb[48], s_slope[12], s_ga[48], s_be[48];
  __shared__ __align__(16) __half sv[32 * QKV_TQ * QKV_PITCH];
  for (int i = threadIdx.x; i < 48 * 32; i += 256) sw[i] = i < 256 ? wq[i] : (i < 512 ? wk[i - 256] : wv[i - 512]);
  if (threadIdx.x < 48) {
    const int i = threadIdx.x;
    sb[i] = i < 8 ? bq[i] : (i < 16 ? bk[i - 8] : bv[i - 16]);
    s_ga[i] = i < 8 ? gq[i] : (i < 16 ? gk[i - 8] : gv[i - 16]);
    s_be[i] = i < 8 ? betq[i] : (i < 16 ? betk[i - 8] : betv[i - 16]);
    if (i < 12) s_slope[i] = i < 4 ? aq[i] : (i < 8 ? ak[i - 4] : av[i - 8]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int b = blockIdx.z, t0 = blockIdx.y * QKV_TT, q = blockIdx.x * QKV_TQ + lane;
  const int tl[2] = {wrp, wrp + 8};
  bool ok[2];
  float zv[2][32];
  // a warp's 32 positions (one frame, 32 consecutive bins) are 4 KB contiguous: read them as fully used 16-byte pieces into a
  // per-warp [32][33] tile (aliased onto the V transpose buffer, which is not written before the barrier below), then each lane
  // takes its own row
  float* stw = reinterpret_cast<float*>(sv) + wrp * (32 * 33);
  static_assert(sizeof(sv) >= 8 * 32 * 33 * sizeof(float), "z staging aliases the V transpose buffer");
  const int qw0 = blockIdx.x * QKV_TQ;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int t = t0 + tl[r];
    ok[r] = t < T && q < Q;
    const int n_q = t < T ? min(32, Q - qw0) : 0;             // valid bins of this warp's row
    const float4* src = reinterpret_cast<const float4*>(z + ((static_cast<int64_t>(b) * T + (t < T ? t : 0)) * Q + qw0) * TC);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = j * 32 + lane, pos = idx >> 3, c4 = idx & 7;
      const float4 v = pos < n_q ? __ldg(src + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      stw[pos * 33 + 4 * c4] = v.x; stw[pos * 33 + 4 * c4 + 1] = v.y; stw[pos * 33 + 4 * c4 + 2] = v.z; stw[pos * 33 + 4 * c4 + 3] = v.w;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 32; ++k) zv[r][k] = stw[lane * 33 + k];
  }
  __syncthreads();                                            // staging reads done before the V transpose writes
  // six groups of 8 outputs: Q (heads x E), K, then the four V heads
#pragma unroll 1
  for (int grp = 0; grp < 6; ++grp) {
    float acc[2][8];
#pragma unroll
    for (int o = 0; o < 8; ++o) { acc[0][o] = acc[1][o] = sb[grp * 8 + o]; }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float4* wrow = reinterpret_cast<const float4*>(sw + (grp * 8 + o) * 32);
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 w = wrow[k4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          acc[r][o] = fmaf(w.x, zv[r][4 * k4], acc[r][o]); acc[r][o] = fmaf(w.y, zv[r][4 * k4 + 1], acc[r][o]);
          acc[r][o] = fmaf(w.z, zv[r][4 * k4 + 2], acc[r][o]); acc[r][o] = fmaf(w.w, zv[r][4 * k4 + 3], acc[r][o]);
        }
      }
    }
    if (grp < 2) {
      // Q / K: channel o = head * 2 + e; PReLU(head) then normalise over the head's two channels
      __half* dst = grp == 0 ? Qh : Kh;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int hd = 0; hd < 4; ++hd) {
          const float slope = s_slope[grp * 4 + hd];
          float v0 = acc[r][2 * hd], v1 = acc[r][2 * hd + 1];
          v0 = v0 >= 0.f ? v0 : slope * v0; v1 = v1 >= 0.f ? v1 : slope * v1;
          const float mu = 0.5f * (v0 + v1), d0 = v0 - mu, d1 = v1 - mu;
          const float rs = rsqrtf(d0 * d0 + eps);                        // two elements: var = ((v0 - v1) / 2)^2 = d0^2
          if (ok[r]) {
            __half* row = dst + ((static_cast<int64_t>(b) * 4 + hd) * T + t0 + tl[r]) * ldf + q;
            row[0] = __float2half_rn(d0 * rs * s_ga[grp * 8 + 2 * hd] + s_be[grp * 8 + 2 * hd]);
            row[Q] = __float2half_rn(d1 * rs * s_ga[grp * 8 + 2 * hd + 1] + s_be[grp * 8 + 2 * hd + 1]);
          }
        }
      }
    } else {
      const int hd = grp - 2;
      const float slope = s_slope[8 + hd];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v[8], mu = 0.f, var = 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o) { v[o] = acc[r][o] >= 0.f ? acc[r][o] : slope * acc[r][o]; mu += v[o]; }
        mu *= 0.125f;
#pragma unroll
        for (int o = 0; o < 8; ++o) { v[o] -= mu; var = fmaf(v[o], v[o], var); }
        const float rs = rsqrtf(var * 0.125f + eps);
#pragma unroll
        for (int o = 0; o < 8; ++o)
          sv[((hd * 8 + o) * QKV_TQ + lane) * QKV_PITCH + tl[r]] = __float2half_rn(v[o] * rs * s_ga[16 + hd * 8 + o] + s_be[16 + hd * 8 + o]);
      }
    }
  }
  __syncthreads();
  // Vt rows (channel, bin): 16 frames = 32 contiguous bytes
  const int n_t = min(QKV_TT, T - t0);
  for (int row = threadIdx.x; row < 32 * QKV_TQ; row += 256) {
    const int c = row >> 5, qq = blockIdx.x * QKV_TQ + (row & 31);
    if (qq >= Q) continue;
    __half* dst = Vt + ((static_cast<int64_t>(b) * 4 + (c >> 3)) * (8 * Q) + static_cast<int64_t>(qq) * 8 + (c & 7)) * ldt + t0;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(sv + row * QKV_PITCH);
    if (n_t == QKV_TT) {
      reinterpret_cast<uint4*>(dst)[0] = make_uint4(src[0], src[1], src[2], src[3]);
      reinterpret_cast<uint4*>(dst)[1] = make_uint4(src[4], src[5], src[6], src[7]);
    } else {
      const __half* sh = sv + row * QKV_PITCH;
      for (int t = 0; t < n_t; ++t) dst[t] = sh[t];
    }
  }
}

// batched C[m, n] = scale * sum_k A[m, k] B[n, k]   (both operands K-contiguous fp16, fp32 accumulate, mma.sync m16n8k16)
// CTA tile 64 x 64, 4 warps of 32 x 32.  out_mode 0: fp32 C[batch][m][n];  1: fp16 C[batch][m][n];
// 2: attention output into the [B, T, Q, C] activation layout: batch = b * 4 + head, m = t, n = f * 8 + c8 (a thread's two
// adjacent columns are one 8-byte store, a quad's eight columns one 32-byte run).
struct GemmArgs {
  const __half* A; const __half* Bm; void* C;
  int M, N, K; int64_t sA, sB, sC; int lda, ldb, ldc; float scale; int out_mode; int Q;
};
__global__ void __launch_bounds__(128) gemm_tn_kernel(const GemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 64, PITCH = BK + 8, NLD = BM * (BK / 8) / 128;
  __shared__ __align__(16) __half sA[BM][PITCH];
  __shared__ __align__(16) __half sB[BN][PITCH];
  const int bz = blockIdx.z, m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const __half* A = g.A + bz * g.sA; const __half* Bm = g.Bm + bz * g.sB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }
  // 64 rows x 64 halves per operand as 16-byte pieces (K, lda, ldb are multiples of 8 halves; pad columns hold zeros); the next
  // K-slice is fetched into registers while the current one is multiplied
  uint4 ra[NLD], rb[NLD];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int e = threadIdx.x + u * 128, r = e >> 3, k = k0 + (e & 7) * 8;
      const int m = m0 + r, n = n0 + r;
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      ra[u] = (m < g.M && k < g.K) ? __ldg(reinterpret_cast<const uint4*>(A + static_cast<int64_t>(m) * g.lda + k)) : zero;
      rb[u] = (n < g.N && k < g.K) ? __ldg(reinterpret_cast<const uint4*>(Bm + static_cast<int64_t>(n) * g.ldb + k)) : zero;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int e = threadIdx.x + u * 128, r = e >> 3, kk = (e & 7) * 8;
      *reinterpret_cast<uint4*>(&sA[r][kk]) = ra[u];
      *reinterpret_cast<uint4*>(&sB[r][kk]) = rb[u];
    }
    __syncthreads();
    if (k0 + BK < g.K) fetch(k0 + BK);
#pragma unroll
    for (int ks = 0; ks < BK; ks += 16) {
      uint32_t af[2][4], bf[4][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = wm + i * 16 + gq;
        af[i][0] = *reinterpret_cast<const uint32_t*>(&sA[r][ks + 2 * tq]);
        af[i][1] = *reinterpret_cast<const uint32_t*>(&sA[r + 8][ks + 2 * tq]);
        af[i][2] = *reinterpret_cast<const uint32_t*>(&sA[r][ks + 2 * tq + 8]);
        af[i][3] = *reinterpret_cast<const uint32_t*>(&sA[r + 8][ks + 2 * tq + 8]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = wn + j * 8 + gq;
        bf[j][0] = *reinterpret_cast<const uint32_t*>(&sB[c][ks + 2 * tq]);
        bf[j][1] = *reinterpret_cast<const uint32_t*>(&sB[c][ks + 2 * tq + 8]);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816(acc[i][j], af[i], make_uint2(bf[j][0], bf[j][1]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        const int m = m0 + wm + i * 16 + gq + e2 * 8, n = n0 + wn + j * 8 + 2 * tq;      // columns n, n + 1 (N and ldc are even)
        if (m >= g.M || n >= g.N) continue;
        const float v0 = acc[i][j][2 * e2] * g.scale, v1 = acc[i][j][2 * e2 + 1] * g.scale;
        if (g.out_mode == 0) *reinterpret_cast<float2*>(reinterpret_cast<float*>(g.C) + bz * g.sC + static_cast<int64_t>(m) * g.ldc + n) = make_float2(v0, v1);
        else if (g.out_mode == 1) *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(g.C) + bz * g.sC + static_cast<int64_t>(m) * g.ldc + n) = __floats2half2_rn(v0, v1);
        else {
          const int b = bz >> 2, hd = bz & 3, f = n >> 3, c8 = n & 7;
          *reinterpret_cast<float2*>(reinterpret_cast<float*>(g.C) + ((static_cast<int64_t>(b) * g.M + m) * g.Q + f) * TC + hd * 8 + c8) = make_float2(v0, v1);
        }
      }
}

// soft-max over the last axis of S [rows][T] fp32 -> P fp16 (tfgridnet.py:405)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, __half* __restrict__ P, int64_t rows, int T, int ldt) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = blockIdx.x * 8ll + (threadIdx.x >> 5); r < rows; r += 8ll * gridDim.x) {
    const float* s = S + r * ldt;
    float mx = -3.0e38f;
    for (int i = lane; i < T; i += 32) mx = fmaxf(mx, s[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < T; i += 32) sum += __expf(s[i] - mx);
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int i = lane; i < ldt; i += 32) P[r * ldt + i] = __float2half_rn(i < T ? __expf(s[i] - mx) * inv : 0.f);
  }
}

// attention back: 1x1 conv C -> C, PReLU (one slope), LayerNorm over channels, + residual (tfgridnet.py:300-306, 424-427)
__global__ void __launch_bounds__(256)
tfg_attn_proj_kernel(const float* __restrict__ o, const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ slope,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ resid, int64_t n_pos, float eps,
                     float* __restrict__ out) {
  // thread = position: the 32 x 32 product, PReLU and LayerNorm are register-local (weights broadcast from shared memory).
  // Positions enter and leave through a shared-memory tile [256][33] so that every global access is a fully used 16-byte
  // piece of a contiguous 32 KB run (per-thread 128-byte rows straight from global memory ran at 2.2 TB/s).
  __shared__ __align__(16) float sw[32 * 32];
  __shared__ float sp[3 * 32];
  __shared__ float st[256][TC + 1];
  for (int i = threadIdx.x; i < 1024; i += 256) sw[i] = w[i];
  if (threadIdx.x < 32) { sp[threadIdx.x] = bias[threadIdx.x]; sp[32 + threadIdx.x] = gamma[threadIdx.x]; sp[64 + threadIdx.x] = beta[threadIdx.x]; }
  const float a = __ldg(slope);
  for (int64_t p0 = blockIdx.x * 256ll; p0 < n_pos; p0 += 256ll * gridDim.x) {
    __syncthreads();                                   // weights visible (first pass); previous tile's store phase done
    const int n_here = n_pos - p0 < 256 ? static_cast<int>(n_pos - p0) : 256;
    const float4* src = reinterpret_cast<const float4*>(o + p0 * TC);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = j * 256 + threadIdx.x, pos = idx >> 3, c4 = idx & 7;
      if (pos < n_here) { const float4 v = __ldg(src + idx); st[pos][4 * c4] = v.x; st[pos][4 * c4 + 1] = v.y; st[pos][4 * c4 + 2] = v.z; st[pos][4 * c4 + 3] = v.w; }
    }
    __syncthreads();
    if (threadIdx.x < n_here) {
      float x[32], acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) x[k] = st[threadIdx.x][k];
      float mu = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float s = sp[c];
        const float4* wrow = reinterpret_cast<const float4*>(sw + c * 32);
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 wv = wrow[k4];
          s = fmaf(wv.x, x[4 * k4], s); s = fmaf(wv.y, x[4 * k4 + 1], s); s = fmaf(wv.z, x[4 * k4 + 2], s); s = fmaf(wv.w, x[4 * k4 + 3], s);
        }
        s = s >= 0.f ? s : a * s;
        acc[c] = s; mu += s;
      }
      mu *= (1.0f / 32.0f);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) { acc[c] -= mu; var = fmaf(acc[c], acc[c], var); }
      const float rs = rsqrtf(var * (1.0f / 32.0f) + eps);
#pragma unroll
      for (int c = 0; c < 32; ++c) st[threadIdx.x][c] = acc[c] * rs * sp[32 + c] + sp[64 + c];      // own row: no hazard with other threads
    }
    __syncthreads();
    const float4* rsrc = reinterpret_cast<const float4*>(resid + p0 * TC);
    float4* dst = reinterpret_cast<float4*>(out + p0 * TC);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = j * 256 + threadIdx.x, pos = idx >> 3, c4 = idx & 7;
      if (pos < n_here) {
        const float4 r = __ldg(rsrc + idx);
        dst[idx] = make_float4(st[pos][4 * c4] + r.x, st[pos][4 * c4 + 1] + r.y, st[pos][4 * c4 + 2] + r.z, st[pos][4 * c4 + 3] + r.w);
      }
    }
  }
}

// output ConvTranspose2d(C -> 2, 3x3, padding 1) (tfgridnet.py:175, 221-226): out[o, t, f] = b[o] + sum h[c, t+1-dt, f+1-dq] W[c, o, dt, dq],
// written as complex [B,1,F,T].  A block owns 32 frames x 8 bins: the h patch (34 x 10 positions x 32 channels) is staged in
// shared memory (frame rows padded by 4 words: the per-lane 16-byte reads of a warp, whose lanes run along frames, then hit
// distinct banks), a thread computes one output position, and a warp stores 32 consecutive frames = 256 contiguous bytes.
constexpr int DC_TT = 32, DC_TQ = 8, DC_ROW = (DC_TQ + 2) * TC + 4;
__global__ void __launch_bounds__(256)
tfg_deconv_out_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias, int B, int T, int Q,
                      float2* __restrict__ out) {
  __shared__ __align__(16) float sh[(DC_TT + 2) * DC_ROW];
  __shared__ __align__(16) float sw[9 * 2 * TC];                  // [tap][o][c]
  for (int i = threadIdx.x; i < TC * 18; i += 256) { const int c = i / 18, o = (i / 9) % 2, tap = i % 9; sw[(tap * 2 + o) * TC + c] = w[i]; }
  const int b = blockIdx.z, t0 = blockIdx.y * DC_TT, q0 = blockIdx.x * DC_TQ;
  for (int i = threadIdx.x; i < (DC_TT + 2) * (DC_TQ + 2) * 8; i += 256) {
    const int c4 = i & 7, ql = (i >> 3) % (DC_TQ + 2), tl = (i >> 3) / (DC_TQ + 2);
    const int t = t0 + tl - 1, q = q0 + ql - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < T && q >= 0 && q < Q) v = __ldg(reinterpret_cast<const float4*>(h + ((static_cast<int64_t>(b) * T + t) * Q + q) * TC) + c4);
    *reinterpret_cast<float4*>(sh + tl * DC_ROW + ql * TC + c4 * 4) = v;
  }
  __syncthreads();
  const int tl = threadIdx.x & 31, ql = threadIdx.x >> 5;         // lanes along frames: the output's contiguous axis
  float re = __ldg(bias), im = __ldg(bias + 1);
#pragma unroll
  for (int dt = 0; dt < 3; ++dt)
#pragma unroll
    for (int dq = 0; dq < 3; ++dq) {
      // source position (t + 1 - dt, q + 1 - dq) = patch index (tl + 2 - dt, ql + 2 - dq)
      const float4* hp = reinterpret_cast<const float4*>(sh + (tl + 2 - dt) * DC_ROW + (ql + 2 - dq) * TC);
      const float4* wr = reinterpret_cast<const float4*>(sw + (dt * 3 + dq) * 2 * TC);
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 hv = hp[c4], w0 = wr[c4], w1 = wr[8 + c4];
        re = fmaf(hv.x, w0.x, re); re = fmaf(hv.y, w0.y, re); re = fmaf(hv.z, w0.z, re); re = fmaf(hv.w, w0.w, re);
        im = fmaf(hv.x, w1.x, im); im = fmaf(hv.y, w1.y, im); im = fmaf(hv.z, w1.z, im); im = fmaf(hv.w, w1.w, im);
      }
    }
  const int t = t0 + tl, q = q0 + ql;
  if (t < T && q < Q) out[(static_cast<int64_t>(b) * Q + q) * T + t] = make_float2(re, im);
}

int grid8(int64_t n_pos) { return static_cast<int>(std::min<int64_t>(ceil_div64(n_pos, 8), static_cast<int64_t>(num_sms()) * 16)); }

}  // namespace
}  // namespace fdbm

using namespace fdbm;

// ---- C ABI ------------------------------------------------------------------------------------------------------------
extern "C" int fdbm_tfg_pad_add_norm(const float* h, const float* emb, const float* gamma, const float* beta, int batch, int T, int Q, float eps,
                                     float* xp, void* xn, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(h && gamma && beta && xp && xn && batch > 0 && T > 0 && Q > 0, "fdbm_tfg_pad_add_norm: bad arguments");
  const int64_t n_pos = static_cast<int64_t>(batch) * (T + 2 * OLP) * (Q + 2 * OLP);
  pad_add_norm_kernel<<<grid8(n_pos), 256, 0, as_stream(stream)>>>(h, emb, gamma, beta, batch, T, Q, eps, xp, reinterpret_cast<__half*>(xn));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_tfg_sweep_post(const void* y_fw, const void* y_bw, const float* lin_bias, const float* resid, int batch, int T, int Q, int mode,
                                   const float* gamma, const float* beta, float eps, float* out_full, void* xn, float* out_crop, int xn_transposed,
                                   void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(y_fw && y_bw && lin_bias && resid && (mode == 0 ? (out_full && xn && gamma && beta) : out_crop != nullptr), "fdbm_tfg_sweep_post: bad arguments");
  const int64_t n_pos = mode == 0 ? static_cast<int64_t>(batch) * (T + 2 * OLP) * (Q + 2 * OLP) : static_cast<int64_t>(batch) * T * Q;
  sweep_post_kernel<<<grid8(n_pos), 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(y_fw), reinterpret_cast<const __half*>(y_bw), lin_bias,
                                                                 resid, batch, T, Q, mode, gamma, beta, eps, out_full, reinterpret_cast<__half*>(xn), out_crop,
                                                                 xn_transposed);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_tfg_input(const float* x, const float* y, const float* w, const float* bias, const float* gn_w, const float* gn_b, int batch,
                              int T, int Q, int Cin, float eps, double* sums, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x && w && bias && gn_w && gn_b && sums && out && (Cin == 2 || (Cin == 4 && y)), "fdbm_tfg_input: bad arguments");
  cudaStream_t s = as_stream(stream);
  FDBM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * batch, s));
  tfg_input_conv_kernel<<<dim3(ceil_div(Q, IC_TQ), ceil_div(T, IC_TT), batch), 256, 0, s>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(y), w, bias, batch, T, Q,
                                                        Cin, out, sums);
  FDBM_LAUNCH_CHECK();
  const int64_t per = static_cast<int64_t>(T) * Q * TC;
  tfg_groupnorm1_kernel<<<dim3(std::max(1, std::min<int>(static_cast<int>(ceil_div64(per, 256 * 8)), num_sms() * 8 / batch + 1)), batch), 256, 0, s>>>(
      out, sums, gn_w, gn_b, per, eps);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_tfg_time_embedding(const float* t, int t_stride, const float* fourier_w, const float* w1, const float* b1, const float* w2,
                                       const float* b2, const float* w_blocks, const float* b_blocks, int n_layers, int batch, float* emb,
                                       void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(t && fourier_w && w1 && b1 && w2 && b2 && w_blocks && b_blocks && emb && batch > 0 && n_layers > 0, "fdbm_tfg_time_embedding: bad arguments");
  tfg_temb_kernel<<<batch, 128, 0, as_stream(stream)>>>(t, t_stride, fourier_w, w1, b1, w2, b2, w_blocks, b_blocks, n_layers, emb, batch);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

// attention of one block: z [B,T,Q,32] -> out [B,T,Q,32] = LN(PReLU(conv(attn(z)))) + z.  workspace: see fdbm_tfg_attention_workspace_bytes
extern "C" int64_t fdbm_tfg_attention_workspace_bytes(int batch, int T, int Q) {
  const int64_t bh = static_cast<int64_t>(batch) * 4, ldf = (2 * Q + 7) / 8 * 8, ldt = (T + 7) / 8 * 8;
  return bh * T * ldf * 2 * 2 + bh * (8 * Q) * ldt * 2 + bh * T * ldt * 4 + bh * T * ldt * 2 + static_cast<int64_t>(batch) * T * Q * 32 * 4 + 4096;
}

extern "C" int fdbm_tfg_attention(const float* z, const float* const* params /* 22 device pointers, see tfgridnet.py */, int batch, int T, int Q,
                                  float eps, void* workspace, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(z && params && workspace && out && batch > 0 && T > 0 && Q > 0, "fdbm_tfg_attention: bad arguments");
  cudaStream_t s = as_stream(stream);
  const int64_t bh = static_cast<int64_t>(batch) * 4;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  FDBM_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "fdbm_tfg_attention: workspace must be 256-byte aligned");
  const int ldf = (2 * Q + 7) / 8 * 8, ldt = (T + 7) / 8 * 8;      // fp16 row pitches in multiples of 16 bytes; pad columns are zero
  __half* Qh = reinterpret_cast<__half*>(ws); ws += bh * T * ldf * 2;
  __half* Kh = reinterpret_cast<__half*>(ws); ws += bh * T * ldf * 2;
  __half* Vt = reinterpret_cast<__half*>(ws); ws += bh * (8 * Q) * ldt * 2;
  FDBM_CUDA(cudaMemsetAsync(Qh, 0, static_cast<size_t>(ws - reinterpret_cast<uint8_t*>(Qh)), s));
  float* S = reinterpret_cast<float*>(ws); ws += bh * T * ldt * 4;
  __half* P = reinterpret_cast<__half*>(ws); ws += bh * T * ldt * 2;
  ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~static_cast<uintptr_t>(255));
  float* O = reinterpret_cast<float*>(ws);
  const float* const* p = params;
  // order: wq bq wk bk wv bv | aq ak av | gq betq gk betk gv betv | wproj bproj slope gproj betproj
  const int64_t n_pos = static_cast<int64_t>(batch) * T * Q;
  tfg_qkv_kernel<<<dim3(ceil_div(Q, QKV_TQ), ceil_div(T, QKV_TT), batch), 256, 0, s>>>(z, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14], batch, T,
                                              Q, eps, ldf, ldt, Qh, Kh, Vt);
  FDBM_LAUNCH_CHECK();
  GemmArgs g;
  g.A = Qh; g.Bm = Kh; g.C = S; g.M = T; g.N = T; g.K = ldf; g.sA = static_cast<int64_t>(T) * ldf; g.sB = g.sA; g.sC = static_cast<int64_t>(T) * ldt;
  g.lda = ldf; g.ldb = ldf; g.ldc = ldt; g.scale = 1.0f / sqrtf(static_cast<float>(2 * Q)); g.out_mode = 0; g.Q = Q;
  gemm_tn_kernel<<<dim3(ceil_div(T, 64), ceil_div(T, 64), static_cast<unsigned>(bh)), 128, 0, s>>>(g);
  FDBM_LAUNCH_CHECK();
  softmax_rows_kernel<<<grid8(bh * T), 256, 0, s>>>(S, P, bh * T, T, ldt);
  FDBM_LAUNCH_CHECK();
  g.A = P; g.Bm = Vt; g.C = O; g.M = T; g.N = 8 * Q; g.K = ldt; g.sA = static_cast<int64_t>(T) * ldt; g.sB = static_cast<int64_t>(8) * Q * ldt; g.sC = 0;
  g.lda = ldt; g.ldb = ldt; g.ldc = 0; g.scale = 1.0f; g.out_mode = 2;
  gemm_tn_kernel<<<dim3(ceil_div(8 * Q, 64), ceil_div(T, 64), static_cast<unsigned>(bh)), 128, 0, s>>>(g);
  FDBM_LAUNCH_CHECK();
  tfg_attn_proj_kernel<<<static_cast<int>(std::min<int64_t>(ceil_div64(n_pos, 256), static_cast<int64_t>(num_sms()) * 8)), 256, 0, s>>>(O, p[15], p[16], p[17], p[18], p[19], z, n_pos, eps, out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_tfg_output(const float* h, const float* w, const float* bias, int batch, int T, int Q, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(h && w && bias && out && batch > 0, "fdbm_tfg_output: bad arguments");
  tfg_deconv_out_kernel<<<dim3(ceil_div(Q, DC_TQ), ceil_div(T, DC_TT), batch), 256, 0, as_stream(stream)>>>(h, w, bias, batch, T, Q, reinterpret_cast<float2*>(out));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}
