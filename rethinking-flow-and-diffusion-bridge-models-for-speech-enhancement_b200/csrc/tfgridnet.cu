// TF-GridNet backbones (fdbm/backbones/tfgridnet.py:126-229 TFGridNet, :236-427 GridNetV3Block, :430-484 the two
// normalisation layers; tfgridnet_predictive.py) -- the backbones config.yaml / config_predictive.yaml select.
//
// 97 % of the arithmetic is the ten bidirectional LSTM sweeps of a forward (intra: one sequence of 260 unfolded positions
// along frequency per (utterance, frame); inter: one sequence along time per (utterance, bin)); a 4 s utterance is
// ~250 GFLOP.  The recurrence is sequential in the position, so the sweep is a PERSISTENT kernel: one CTA owns 128 sequences
// of one direction for all steps, keeps the whole gate matrix [K = 128 inputs + 112 hidden] x [N = 416 gate columns]
// (fp16, pre-swizzled into mma B-fragments, 195 KB) and the ConvTranspose1d matrix (28 KB) in shared memory, the hidden
// state as mma A-fragments and the cell state in registers, and issues per step 780 + 112 tensor-core MMAs per warp:
//   * gate columns are ordered [unit group of 8][gate i,f,g,o][unit]: a thread's accumulator fragments of the four gate
//     tiles of a group hold all four gates of the same two units, so the point-wise cell update needs no shuffles, and the
//     new hidden values ARE next step's A-fragment (C-fragment columns 2tq,2tq+1 of group G = A-fragment columns of k-tile
//     G/2): the recurrent operand never leaves registers;
//   * the unfolded LSTM input (emb_ks = 4 neighbouring positions x 32 channels, tfgridnet.py:337-341) is a sliding window
//     over the LayerNorm-ed sequence: with the input columns ordered [tap][channel] a step shifts the window by two k-tiles,
//     so only the 32 channels of ONE new position are loaded per step (prefetched a step ahead);
//   * ConvTranspose1d (tfgridnet.py:346 / :371) is applied to the fresh hidden state in the same step (h_t W_lin ->
//     [4 taps x 32 channels], fp16 to HBM); a light pass adds the four shifted taps of both directions, the bias and the
//     residual, and produces the next LayerNorm.
// Everything else (3x3 input conv + GroupNorm, time embedding, the attention's 1x1 convs / PReLU-LayerNorms, two batched
// tensor-core GEMMs for Q K^T and P V, the output ConvTranspose2d) is small and HBM / latency bound.
#include <mma.h>
#include "common.cuh"

namespace fdbm {
namespace {

constexpr int TC = 32;            // emb_dim
constexpr int TKS = 4;            // emb_ks (emb_hs = 1)
constexpr int OLP = 3;            // emb_ks - emb_hs
constexpr int HPAD = 112;         // hidden units padded to 7 k-tiles
constexpr int NGRP = 13;          // unit groups of 8 (104 >= 100 units)
constexpr int KT_IN = 8, KT_H = 7, KT = KT_IN + KT_H;
constexpr int NT = NGRP * 4;      // 52 gate n-tiles
constexpr int NT_LIN = 16;        // 128 = 4 taps x 32 channels
constexpr int SEQ_PER_CTA = 128;
constexpr int LSTM_THREADS = 256;
constexpr size_t WG_BYTES = static_cast<size_t>(KT) * NT * 32 * 8;            // 199680
constexpr size_t WL_BYTES = static_cast<size_t>(KT_H) * NT_LIN * 32 * 8;      // 28672
constexpr size_t BIAS_BYTES = NT * 8 * 4;                                     // 1664
constexpr size_t LSTM_SMEM = WG_BYTES + WL_BYTES + BIAS_BYTES;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint2 b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ float tanh_fast(float x) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// ---- weight packing: reference layout -> mma B-fragments ------------------------------------------------------------
// gates: wfrag[kt][nt][lane] = (half2 B[k0..k0+1][n], half2 B[k0+8..k0+9][n]),  k0 = 16 kt + 2 (lane & 3), n = lane >> 2
//   k < 128: input feature f' = tap * 32 + c  <- weight_ih[row][c * 4 + tap];  k >= 128: hidden unit k - 128 <- weight_hh
//   column (nt, n): group G = nt / 4, gate = nt % 4, unit = 8 G + n  <- row = gate * H + unit  (PyTorch gate order i, f, g, o)
__global__ void pack_lstm_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                 const float* __restrict__ b_hh, const float* __restrict__ w_lin, int H, int dir,
                                 uint2* __restrict__ wg, uint2* __restrict__ wl, float* __restrict__ bias) {
  const int total = KT * NT * 32;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int lane = e & 31, nt = (e >> 5) % NT, kt = (e >> 5) / NT;
    const int n = lane >> 2, tq = lane & 3;
    const int G = nt >> 2, gate = nt & 3, unit = 8 * G + n;
    auto val = [&](int k) -> float {
      if (unit >= H) return 0.f;
      const int row = gate * H + unit;
      if (k < 128) { const int tap = k >> 5, c = k & 31; return w_ih[row * (TC * TKS) + c * TKS + tap]; }
      const int j = k - 128;
      return j < H ? w_hh[row * H + j] : 0.f;
    };
    const int k0 = 16 * kt + 2 * tq;
    wg[e] = make_uint2(pack_h2(val(k0), val(k0 + 1)), pack_h2(val(k0 + 8), val(k0 + 9)));
  }
  // ConvTranspose1d weight [2H, C, ks]: this direction's rows dir * H + unit; column = tap * 32 + c
  const int total_l = KT_H * NT_LIN * 32;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total_l; e += gridDim.x * blockDim.x) {
    const int lane = e & 31, nt = (e >> 5) % NT_LIN, kt = (e >> 5) / NT_LIN;
    const int n = lane >> 2, tq = lane & 3;
    const int col = nt * 8 + n, tap = col >> 5, c = col & 31;
    auto val = [&](int j) -> float { return j < H ? w_lin[((dir * H + j) * TC + c) * TKS + tap] : 0.f; };
    const int k0 = 16 * kt + 2 * tq;
    wl[e] = make_uint2(pack_h2(val(k0), val(k0 + 1)), pack_h2(val(k0 + 8), val(k0 + 9)));
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NT * 8; e += gridDim.x * blockDim.x) {
    const int nt = e >> 3, n = e & 7, G = nt >> 2, gate = nt & 3, unit = 8 * G + n;
    bias[e] = unit < H ? b_ih[gate * H + unit] + b_hh[gate * H + unit] : 0.f;
  }
}

// ---- the persistent BiLSTM sweep -------------------------------------------------------------------------------------
struct LstmArgs {
  const __half* xn;          // LayerNorm-ed input, fp16; element (seq, pos, c) at seq_base(seq) + pos * pos_stride + c
  int n_seq, n_inner;        // sequences; seq = outer * n_inner + inner
  int64_t outer_stride, inner_stride, pos_stride;
  int L;                     // steps (positions - 3)
  const uint2* wg[2]; const uint2* wl[2]; const float* bias[2];
  __half* y[2];              // per direction [n_seq][L][128] fp16: h_t W_lin
};

__global__ void __launch_bounds__(LSTM_THREADS, 1) lstm_sweep_kernel(const LstmArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* s_wg = reinterpret_cast<uint2*>(smem);
  uint2* s_wl = reinterpret_cast<uint2*>(smem + WG_BYTES);
  float* s_bias = reinterpret_cast<float*>(smem + WG_BYTES + WL_BYTES);
  const int dir = blockIdx.y;
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.wg[dir]);
    uint4* dst = reinterpret_cast<uint4*>(s_wg);
    for (int i = threadIdx.x; i < static_cast<int>(WG_BYTES / 16); i += LSTM_THREADS) dst[i] = __ldg(src + i);
    const uint4* src2 = reinterpret_cast<const uint4*>(a.wl[dir]);
    uint4* dst2 = reinterpret_cast<uint4*>(s_wl);
    for (int i = threadIdx.x; i < static_cast<int>(WL_BYTES / 16); i += LSTM_THREADS) dst2[i] = __ldg(src2 + i);
    for (int i = threadIdx.x; i < NT * 8; i += LSTM_THREADS) s_bias[i] = __ldg(a.bias[dir] + i);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int seq0 = blockIdx.x * SEQ_PER_CTA + warp * 16;
  int row_seq[2] = {seq0 + g, seq0 + g + 8};
  bool ok[2];
  const __half* base[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    ok[r] = row_seq[r] < a.n_seq;
    const int sq = ok[r] ? row_seq[r] : 0;
    base[r] = a.xn + static_cast<int64_t>(sq / a.n_inner) * a.outer_stride + static_cast<int64_t>(sq % a.n_inner) * a.inner_stride;
  }
  // A-fragments: u[kt] for the 8 input k-tiles (tap = kt / 2, channels 16 (kt & 1) ..), h[kt] for the 7 hidden k-tiles
  uint32_t u[KT_IN][4], h[KT_H][4];
  float c[NGRP][4];                       // cell state: rows g / g+8 x units (8G + 2tq, +1)
#pragma unroll
  for (int kt = 0; kt < KT_H; ++kt) { h[kt][0] = h[kt][1] = h[kt][2] = h[kt][3] = 0u; }
#pragma unroll
  for (int G = 0; G < NGRP; ++G) { c[G][0] = c[G][1] = c[G][2] = c[G][3] = 0.f; }
  // loads of one position's 32 channels as two k-tiles of A-fragments
  auto load_pos = [&](int pos, uint32_t (&t0)[4], uint32_t (&t1)[4]) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(base[r] + static_cast<int64_t>(pos) * a.pos_stride);
      // channels 2tq,2tq+1 | +8 of each 16-channel half
      const uint32_t v0 = ok[r] ? __ldg(p + tq) : 0u, v1 = ok[r] ? __ldg(p + tq + 4) : 0u;
      const uint32_t v2 = ok[r] ? __ldg(p + 8 + tq) : 0u, v3 = ok[r] ? __ldg(p + 12 + tq) : 0u;
      t0[r] = v0; t0[2 + r] = v1; t1[r] = v2; t1[2 + r] = v3;
    }
  };
  const int L = a.L;
  const bool rev = dir == 1;
  // first step's window: positions s .. s + 3
  {
    const int s = rev ? L - 1 : 0;
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) load_pos(s + tap, u[2 * tap], u[2 * tap + 1]);
  }
  uint32_t nx0[4], nx1[4];                // the one new position of the NEXT step
  for (int step = 0; step < L; ++step) {
    const int s = rev ? L - 1 - step : step;
    const bool more = step + 1 < L;
    if (more) load_pos(rev ? s - 1 : s + 4, nx0, nx1);
    uint32_t hn[KT_H][4];
#pragma unroll
    for (int kt = 0; kt < KT_H; ++kt) { hn[kt][0] = hn[kt][1] = hn[kt][2] = hn[kt][3] = 0u; }
#pragma unroll
    for (int G = 0; G < NGRP; ++G) {
      float acc[4][4];
#pragma unroll
      for (int gate = 0; gate < 4; ++gate) {
        const float2 b = *reinterpret_cast<const float2*>(s_bias + (G * 4 + gate) * 8 + 2 * tq);
        acc[gate][0] = b.x; acc[gate][1] = b.y; acc[gate][2] = b.x; acc[gate][3] = b.y;
      }
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const uint2* wrow = s_wg + (static_cast<size_t>(kt) * NT + G * 4) * 32 + lane;
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
          if (kt < KT_IN) mma16816(acc[gate], u[kt], wrow[gate * 32]);
          else mma16816(acc[gate], h[kt - KT_IN], wrow[gate * 32]);
        }
      }
      // point-wise cell update of units (8G + 2tq, +1), rows g and g+8  (tfgridnet uses nn.LSTM: i, f, g, o)
      float hv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float ig = sigmoid_fast(acc[0][e]), fg = sigmoid_fast(acc[1][e]), gg = tanh_fast(acc[2][e]), og = sigmoid_fast(acc[3][e]);
        const float cn = fmaf(fg, c[G][e], ig * gg);
        c[G][e] = cn;
        hv[e] = og * tanh_fast(cn);
      }
      // C-fragment (rows g | g+8, cols 2tq, 2tq+1 of group G) -> A-fragment of k-tile G / 2, left (G even) or right half
      hn[G >> 1][(G & 1) * 2 + 0] = pack_h2(hv[0], hv[1]);
      hn[G >> 1][(G & 1) * 2 + 1] = pack_h2(hv[2], hv[3]);
    }
#pragma unroll
    for (int kt = 0; kt < KT_H; ++kt) { h[kt][0] = hn[kt][0]; h[kt][1] = hn[kt][1]; h[kt][2] = hn[kt][2]; h[kt][3] = hn[kt][3]; }
    // y_s = h_s W_lin  -> fp16 [seq][s][tap * 32 + c]
#pragma unroll
    for (int nb = 0; nb < NT_LIN; nb += 4) {
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
#pragma unroll
      for (int kt = 0; kt < KT_H; ++kt) {
        const uint2* wrow = s_wl + (static_cast<size_t>(kt) * NT_LIN + nb) * 32 + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816(acc[j], h[kt], wrow[j * 32]);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (!ok[r]) continue;
        uint32_t* dst = reinterpret_cast<uint32_t*>(a.y[dir] + (static_cast<int64_t>(row_seq[r]) * L + s) * 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[(nb + j) * 4 + tq] = pack_h2(acc[j][2 * r], acc[j][2 * r + 1]);
      }
    }
    // slide the input window by one position
    if (more) {
      if (!rev) {
#pragma unroll
        for (int kt = 0; kt < KT_IN - 2; ++kt) { u[kt][0] = u[kt + 2][0]; u[kt][1] = u[kt + 2][1]; u[kt][2] = u[kt + 2][2]; u[kt][3] = u[kt + 2][3]; }
#pragma unroll
        for (int e = 0; e < 4; ++e) { u[KT_IN - 2][e] = nx0[e]; u[KT_IN - 1][e] = nx1[e]; }
      } else {
#pragma unroll
        for (int kt = KT_IN - 1; kt >= 2; --kt) { u[kt][0] = u[kt - 2][0]; u[kt][1] = u[kt - 2][1]; u[kt][2] = u[kt - 2][2]; u[kt][3] = u[kt - 2][3]; }
#pragma unroll
        for (int e = 0; e < 4; ++e) { u[0][e] = nx0[e]; u[1][e] = nx1[e]; }
      }
    }
  }
}

// ---- per-position passes ---------------------------------------------------------------------------------------------
// LayerNorm over the 32 channels of one position: one warp per position, lane = channel
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ln32(float v, float gamma, float beta, float eps) {
  const float mu = warp_sum(v) * (1.0f / 32.0f);
  const float d = v - mu;
  const float var = warp_sum(d * d) * (1.0f / 32.0f);
  return d * rsqrtf(var + eps) * gamma + beta;
}

// xp[b, t', q', c] = (inside ? h[b, t, q, c] + emb[b, c] : 0);  xn = fp16 LayerNorm(xp)   (tfgridnet.py:206, :329-334)
__global__ void __launch_bounds__(256)
pad_add_norm_kernel(const float* __restrict__ hcur, const float* __restrict__ emb, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int B, int T, int Q, float eps, float* __restrict__ xp, __half* __restrict__ xn) {
  const int Tp = T + 2 * OLP, Qp = Q + 2 * OLP;
  const int64_t n_pos = static_cast<int64_t>(B) * Tp * Qp;
  const int lane = threadIdx.x & 31;
  const float ga = __ldg(gamma + lane), be = __ldg(beta + lane);
  for (int64_t p = blockIdx.x * 8ll + (threadIdx.x >> 5); p < n_pos; p += 8ll * gridDim.x) {
    const int qp = static_cast<int>(p % Qp), tp = static_cast<int>((p / Qp) % Tp), b = static_cast<int>(p / (static_cast<int64_t>(Qp) * Tp));
    const int t = tp - OLP, q = qp - OLP;
    float v = 0.f;
    if (t >= 0 && t < T && q >= 0 && q < Q)
      v = hcur[((static_cast<int64_t>(b) * T + t) * Q + q) * TC + lane] + (emb ? __ldg(emb + b * TC + lane) : 0.f);
    xp[p * TC + lane] = v;
    xn[p * TC + lane] = __float2half_rn(ln32(v, ga, be, eps));
  }
}

// After a sweep: out[seq, pos, c] = bias[c] + sum_tap (yf + yb)[seq, pos - tap, tap * 32 + c] + resid[seq, pos, c]
// (ConvTranspose1d with stride 1 + residual, tfgridnet.py:346-350 / :371-375).  intra (mode 0): writes the full padded
// tensor and its LayerNorm for the inter sweep.  inter (mode 1): writes only the un-padded crop [B,T,Q,C] (:381).
__global__ void __launch_bounds__(256)
sweep_post_kernel(const __half* __restrict__ yf, const __half* __restrict__ yb, const float* __restrict__ lin_bias, const float* __restrict__ resid,
                  int B, int T, int Q, int mode, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                  float* __restrict__ out_full, __half* __restrict__ xn, float* __restrict__ out_crop) {
  const int Tp = T + 2 * OLP, Qp = Q + 2 * OLP;
  const int lane = threadIdx.x & 31;
  const float lb = __ldg(lin_bias + lane);
  const float ga = gamma ? __ldg(gamma + lane) : 1.f, be = beta ? __ldg(beta + lane) : 0.f;
  const int L = (mode == 0 ? Qp : Tp) - OLP;
  const int64_t n_pos = mode == 0 ? static_cast<int64_t>(B) * Tp * Qp : static_cast<int64_t>(B) * T * Q;
  for (int64_t p = blockIdx.x * 8ll + (threadIdx.x >> 5); p < n_pos; p += 8ll * gridDim.x) {
    int b, tp, qp;
    if (mode == 0) { qp = static_cast<int>(p % Qp); tp = static_cast<int>((p / Qp) % Tp); b = static_cast<int>(p / (static_cast<int64_t>(Qp) * Tp)); }
    else { qp = static_cast<int>(p % Q) + OLP; tp = static_cast<int>((p / Q) % T) + OLP; b = static_cast<int>(p / (static_cast<int64_t>(Q) * T)); }
    const int64_t seq = mode == 0 ? static_cast<int64_t>(b) * Tp + tp : static_cast<int64_t>(b) * Qp + qp;
    const int pos = mode == 0 ? qp : tp;
    float acc = lb;
#pragma unroll
    for (int tap = 0; tap < TKS; ++tap) {
      const int s = pos - tap;
      if (s >= 0 && s < L) {
        const int64_t o = (seq * L + s) * 128 + tap * 32 + lane;
        acc += __half2float(yf[o]) + __half2float(yb[o]);
      }
    }
    const int64_t pfull = (static_cast<int64_t>(b) * Tp + tp) * Qp + qp;
    acc += resid[pfull * TC + lane];
    if (mode == 0) {
      out_full[pfull * TC + lane] = acc;
      xn[pfull * TC + lane] = __float2half_rn(ln32(acc, ga, be, eps));
    } else {
      out_crop[p * TC + lane] = acc;
    }
  }
}

// input conv 3x3 (Cin = 4: x.re, x.im, y.re, y.im; or 2) over the [T, F] plane of complex [B,1,F,T] inputs (tfgridnet.py:152,201-214)
// + the sums for GroupNorm(1, C) (a LayerNorm over the whole (C, T, F) of an utterance)
__global__ void __launch_bounds__(256)
tfg_input_conv_kernel(const float2* __restrict__ x, const float2* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bias,
                      int B, int T, int Q, int Cin, float* __restrict__ out, double* __restrict__ sums) {
  __shared__ float sw[TC * 4 * 9];
  __shared__ double red[2][8];
  for (int i = threadIdx.x; i < TC * Cin * 9; i += 256) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int64_t n_pos = static_cast<int64_t>(T) * Q;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t p = blockIdx.x * 8ll + wrp; p < n_pos; p += 8ll * gridDim.x) {
    const int q = static_cast<int>(p % Q), t = static_cast<int>(p / Q);
    float acc = __ldg(bias + lane);
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int tt = t + dt - 1;
      if (tt < 0 || tt >= T) continue;
#pragma unroll
      for (int dq = 0; dq < 3; ++dq) {
        const int qq = q + dq - 1;
        if (qq < 0 || qq >= Q) continue;
        const int64_t si = (static_cast<int64_t>(b) * Q + qq) * T + tt;           // [B,1,F,T]
        const float2 xv = __ldg(x + si);
        // conv weight [C, Cin, kh = time, kw = freq] on input [B, Cin, T, F]
        const float* wp = sw + lane * Cin * 9 + dt * 3 + dq;
        acc = fmaf(xv.x, wp[0], acc); acc = fmaf(xv.y, wp[9], acc);
        if (Cin == 4) { const float2 yv = __ldg(y + si); acc = fmaf(yv.x, wp[18], acc); acc = fmaf(yv.y, wp[27], acc); }
      }
    }
    out[((static_cast<int64_t>(b) * T + t) * Q + q) * TC + lane] = acc;
    s1 += acc; s2 += static_cast<double>(acc) * acc;
  }
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
// (tfgridnet.py:383-385, 458-484).  One warp per position; outputs fp16 operands of the two batched GEMMs:
//   Qh, Kh [B*4][T][E*F] with feature e * F + f  (tfgridnet.py:392-396);  Vt [B*4][8*F][T] (feature c8 * F + f, frames contiguous)
__global__ void __launch_bounds__(256)
tfg_qkv_kernel(const float* __restrict__ z, const float* __restrict__ wq, const float* __restrict__ bq, const float* __restrict__ wk,
               const float* __restrict__ bk, const float* __restrict__ wv, const float* __restrict__ bv, const float* __restrict__ aq,
               const float* __restrict__ ak, const float* __restrict__ av, const float* __restrict__ gq, const float* __restrict__ betq,
               const float* __restrict__ gk, const float* __restrict__ betk, const float* __restrict__ gv, const float* __restrict__ betv,
               int B, int T, int Q, float eps, int ldf, int ldt, __half* __restrict__ Qh, __half* __restrict__ Kh, __half* __restrict__ Vt) {
  __shared__ float sw[48 * 32];
  __shared__ float sb[48];
  for (int i = threadIdx.x; i < 48 * 32; i += 256) sw[i] = i < 256 ? wq[i] : (i < 512 ? wk[i - 256] : wv[i - 512]);
  if (threadIdx.x < 48) sb[threadIdx.x] = threadIdx.x < 8 ? bq[threadIdx.x] : (threadIdx.x < 16 ? bk[threadIdx.x - 8] : bv[threadIdx.x - 16]);
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t n_pos = static_cast<int64_t>(B) * T * Q;
  const int EF = ldf, VF = 8 * Q;                  // row pitches: ldf >= 2 Q, ldt >= T, multiples of 8 (pad columns pre-zeroed)
  for (int64_t p = blockIdx.x * 8ll + wrp; p < n_pos; p += 8ll * gridDim.x) {
    const int q = static_cast<int>(p % Q), t = static_cast<int>((p / Q) % T), b = static_cast<int>(p / (static_cast<int64_t>(Q) * T));
    const float zv = z[p * TC + lane];
    // output o of this lane: lanes 0..15 -> Q/K channel `lane`, all lanes -> V channel `lane`
    float accv = sb[16 + lane], accqk = lane < 16 ? sb[lane] : 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float zk = __shfl_sync(0xffffffffu, zv, k);
      accv = fmaf(sw[(16 + lane) * 32 + k], zk, accv);
      if (lane < 16) accqk = fmaf(sw[lane * 32 + k], zk, accqk);
    }
    // Q / K: head = (lane & 7) / 2, E = 2: normalise over the pair (lane, lane ^ 1)
    {
      const int ch = lane & 7, hd = ch >> 1;
      const float slope = lane < 8 ? __ldg(aq + hd) : __ldg(ak + hd);
      float v = accqk >= 0.f ? accqk : slope * accqk;
      const float o = __shfl_xor_sync(0xffffffffu, v, 1);
      const float mu = 0.5f * (v + o), d = v - mu, var = d * d;        // two elements: var = ((v - o) / 2)^2
      const float ga = lane < 8 ? __ldg(gq + ch) : __ldg(gk + ch), be = lane < 8 ? __ldg(betq + ch) : __ldg(betk + ch);
      const float r = d * rsqrtf(var + eps) * ga + be;
      if (lane < 16) {
        __half* dst = lane < 8 ? Qh : Kh;
        const int e = ch & 1;
        dst[((static_cast<int64_t>(b) * 4 + hd) * T + t) * EF + e * Q + q] = __float2half_rn(r);
      }
    }
    // V: head = lane / 8, 8 channels per head
    {
      const int hd = lane >> 3, c8 = lane & 7;
      const float slope = __ldg(av + hd);
      const float v = accv >= 0.f ? accv : slope * accv;
      float s = v;
      s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
      const float mu = s * 0.125f, d = v - mu;
      float vs = d * d;
      vs += __shfl_xor_sync(0xffffffffu, vs, 1); vs += __shfl_xor_sync(0xffffffffu, vs, 2); vs += __shfl_xor_sync(0xffffffffu, vs, 4);
      const float r = d * rsqrtf(vs * 0.125f + eps) * __ldg(gv + lane) + __ldg(betv + lane);
      Vt[((static_cast<int64_t>(b) * 4 + hd) * VF + c8 * Q + q) * ldt + t] = __float2half_rn(r);
    }
  }
}

// batched C[m, n] = scale * sum_k A[m, k] B[n, k]   (both operands K-contiguous fp16, fp32 accumulate, mma.sync m16n8k16)
// CTA tile 64 x 64, 4 warps of 32 x 32.  out_mode 0: fp32 C[batch][m][n];  1: fp16 C[batch][m][n];
// 2: attention output scattered into the [B, T, Q, C] activation layout: batch = b * 4 + head, m = t, n = c8 * Q + f.
struct GemmArgs {
  const __half* A; const __half* Bm; void* C;
  int M, N, K; int64_t sA, sB, sC; int lda, ldb, ldc; float scale; int out_mode; int Q;
};
__global__ void __launch_bounds__(128) gemm_tn_kernel(const GemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 32, PITCH = BK + 8;
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
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    // 64 rows x 32 halves per operand as 16-byte pieces (K, lda, ldb are multiples of 8 halves; pad columns hold zeros)
    for (int e = threadIdx.x; e < BM * (BK / 8); e += 128) {
      const int r = e >> 2, kk = (e & 3) * 8;
      const int m = m0 + r, n = n0 + r, k = k0 + kk;
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(&sA[r][kk]) = (m < g.M && k < g.K) ? __ldg(reinterpret_cast<const uint4*>(A + static_cast<int64_t>(m) * g.lda + k)) : zero;
      *reinterpret_cast<uint4*>(&sB[r][kk]) = (n < g.N && k < g.K) ? __ldg(reinterpret_cast<const uint4*>(Bm + static_cast<int64_t>(n) * g.ldb + k)) : zero;
    }
    __syncthreads();
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
      for (int e = 0; e < 4; ++e) {
        const int m = m0 + wm + i * 16 + gq + (e >> 1) * 8, n = n0 + wn + j * 8 + 2 * tq + (e & 1);
        if (m >= g.M || n >= g.N) continue;
        const float v = acc[i][j][e] * g.scale;
        if (g.out_mode == 0) reinterpret_cast<float*>(g.C)[bz * g.sC + static_cast<int64_t>(m) * g.ldc + n] = v;
        else if (g.out_mode == 1) reinterpret_cast<__half*>(g.C)[bz * g.sC + static_cast<int64_t>(m) * g.ldc + n] = __float2half_rn(v);
        else {
          const int b = bz >> 2, hd = bz & 3, c8 = n / g.Q, f = n - c8 * g.Q;
          reinterpret_cast<float*>(g.C)[((static_cast<int64_t>(b) * g.M + m) * g.Q + f) * TC + hd * 8 + c8] = v;
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
  __shared__ float sw[32 * 33];
  for (int i = threadIdx.x; i < 1024; i += 256) sw[(i >> 5) * 33 + (i & 31)] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float a = __ldg(slope), ga = __ldg(gamma + lane), be = __ldg(beta + lane), bi = __ldg(bias + lane);
  for (int64_t p = blockIdx.x * 8ll + (threadIdx.x >> 5); p < n_pos; p += 8ll * gridDim.x) {
    const float ov = o[p * TC + lane];
    float acc = bi;
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = fmaf(sw[lane * 33 + k], __shfl_sync(0xffffffffu, ov, k), acc);
    acc = acc >= 0.f ? acc : a * acc;
    out[p * TC + lane] = ln32(acc, ga, be, eps) + resid[p * TC + lane];
  }
}

// output ConvTranspose2d(C -> 2, 3x3, padding 1) (tfgridnet.py:175, 221-226): out[o, t, f] = b[o] + sum h[c, t+1-dt, f+1-dq] W[c, o, dt, dq],
// written as complex [B,1,F,T]
__global__ void __launch_bounds__(256)
tfg_deconv_out_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias, int B, int T, int Q,
                      float2* __restrict__ out) {
  __shared__ float sw[TC * 2 * 9];
  for (int i = threadIdx.x; i < TC * 18; i += 256) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t n_pos = static_cast<int64_t>(B) * T * Q;
  for (int64_t p = blockIdx.x * 8ll + wrp; p < n_pos; p += 8ll * gridDim.x) {
    const int q = static_cast<int>(p % Q), t = static_cast<int>((p / Q) % T), b = static_cast<int>(p / (static_cast<int64_t>(Q) * T));
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int tt = t + 1 - dt;
      if (tt < 0 || tt >= T) continue;
#pragma unroll
      for (int dq = 0; dq < 3; ++dq) {
        const int qq = q + 1 - dq;
        if (qq < 0 || qq >= Q) continue;
        const float hv = h[((static_cast<int64_t>(b) * T + tt) * Q + qq) * TC + lane];
        re = fmaf(hv, sw[(lane * 2 + 0) * 9 + dt * 3 + dq], re);
        im = fmaf(hv, sw[(lane * 2 + 1) * 9 + dt * 3 + dq], im);
      }
    }
    re = warp_sum(re); im = warp_sum(im);
    if (lane == 0) out[(static_cast<int64_t>(b) * Q + q) * T + t] = make_float2(re + __ldg(bias), im + __ldg(bias + 1));
  }
}

int grid8(int64_t n_pos) { return static_cast<int>(std::min<int64_t>(ceil_div64(n_pos, 8), static_cast<int64_t>(num_sms()) * 16)); }

}  // namespace
}  // namespace fdbm

using namespace fdbm;

// ---- C ABI ------------------------------------------------------------------------------------------------------------
extern "C" int64_t fdbm_tfg_lstm_pack_bytes(void) { return static_cast<int64_t>(LSTM_SMEM); }

// pack one direction (dir 0 forward / 1 reverse) of one BiLSTM + its ConvTranspose1d into `packed` (fdbm_tfg_lstm_pack_bytes())
extern "C" int fdbm_tfg_lstm_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* w_lin,
                                  int hidden, int dir, void* packed, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(w_ih && w_hh && b_ih && b_hh && w_lin && packed, "fdbm_tfg_lstm_pack: null pointer");
  FDBM_REQUIRE(hidden >= 8 && hidden <= 104, "fdbm_tfg_lstm_pack: hidden units must be in 8..104 (got %d)", hidden);
  uint8_t* p = reinterpret_cast<uint8_t*>(packed);
  pack_lstm_kernel<<<64, 256, 0, as_stream(stream)>>>(w_ih, w_hh, b_ih, b_hh, w_lin, hidden, dir, reinterpret_cast<uint2*>(p),
                                                      reinterpret_cast<uint2*>(p + WG_BYTES), reinterpret_cast<float*>(p + WG_BYTES + WL_BYTES));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

// one bidirectional sweep.  xn fp16: element (seq = outer * n_inner + inner, pos, c) at outer * outer_stride + inner * inner_stride +
// pos * pos_stride + c;  y_fw / y_bw fp16 [n_seq][L][128]
extern "C" int fdbm_tfg_lstm_sweep(const void* xn, int n_seq, int n_inner, int64_t outer_stride, int64_t inner_stride, int64_t pos_stride,
                                   int L, const void* packed_fw, const void* packed_bw, void* y_fw, void* y_bw, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(xn && packed_fw && packed_bw && y_fw && y_bw && n_seq > 0 && n_inner > 0 && L > 0, "fdbm_tfg_lstm_sweep: bad arguments");
  static PerDeviceOnce attr_once;
  if (attr_once.first(current_device()))
    FDBM_CUDA(cudaFuncSetAttribute(lstm_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(LSTM_SMEM)));
  LstmArgs a;
  a.xn = reinterpret_cast<const __half*>(xn); a.n_seq = n_seq; a.n_inner = n_inner;
  a.outer_stride = outer_stride; a.inner_stride = inner_stride; a.pos_stride = pos_stride; a.L = L;
  const uint8_t* pk[2] = {reinterpret_cast<const uint8_t*>(packed_fw), reinterpret_cast<const uint8_t*>(packed_bw)};
  for (int d = 0; d < 2; ++d) {
    a.wg[d] = reinterpret_cast<const uint2*>(pk[d]); a.wl[d] = reinterpret_cast<const uint2*>(pk[d] + WG_BYTES);
    a.bias[d] = reinterpret_cast<const float*>(pk[d] + WG_BYTES + WL_BYTES);
  }
  a.y[0] = reinterpret_cast<__half*>(y_fw); a.y[1] = reinterpret_cast<__half*>(y_bw);
  dim3 grid(ceil_div(n_seq, SEQ_PER_CTA), 2);
  lstm_sweep_kernel<<<grid, LSTM_THREADS, LSTM_SMEM, as_stream(stream)>>>(a);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

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
                                   const float* gamma, const float* beta, float eps, float* out_full, void* xn, float* out_crop, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(y_fw && y_bw && lin_bias && resid && (mode == 0 ? (out_full && xn && gamma && beta) : out_crop != nullptr), "fdbm_tfg_sweep_post: bad arguments");
  const int64_t n_pos = mode == 0 ? static_cast<int64_t>(batch) * (T + 2 * OLP) * (Q + 2 * OLP) : static_cast<int64_t>(batch) * T * Q;
  sweep_post_kernel<<<grid8(n_pos), 256, 0, as_stream(stream)>>>(reinterpret_cast<const __half*>(y_fw), reinterpret_cast<const __half*>(y_bw), lin_bias,
                                                                 resid, batch, T, Q, mode, gamma, beta, eps, out_full, reinterpret_cast<__half*>(xn), out_crop);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_tfg_input(const float* x, const float* y, const float* w, const float* bias, const float* gn_w, const float* gn_b, int batch,
                              int T, int Q, int Cin, float eps, double* sums, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x && w && bias && gn_w && gn_b && sums && out && (Cin == 2 || (Cin == 4 && y)), "fdbm_tfg_input: bad arguments");
  cudaStream_t s = as_stream(stream);
  FDBM_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * batch, s));
  const int bx = std::max(1, std::min<int>(static_cast<int>(ceil_div64(static_cast<int64_t>(T) * Q, 8 * 8)), num_sms() * 8 / batch + 1));
  tfg_input_conv_kernel<<<dim3(bx, batch), 256, 0, s>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(y), w, bias, batch, T, Q,
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
  tfg_qkv_kernel<<<grid8(n_pos), 256, 0, s>>>(z, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14], batch, T,
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
  tfg_attn_proj_kernel<<<grid8(n_pos), 256, 0, s>>>(O, p[15], p[16], p[17], p[18], p[19], z, n_pos, eps, out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_tfg_output(const float* h, const float* w, const float* bias, int batch, int T, int Q, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(h && w && bias && out && batch > 0, "fdbm_tfg_output: bad arguments");
  tfg_deconv_out_kernel<<<grid8(static_cast<int64_t>(batch) * T * Q), 256, 0, as_stream(stream)>>>(h, w, bias, batch, T, Q, reinterpret_cast<float2*>(out));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}
