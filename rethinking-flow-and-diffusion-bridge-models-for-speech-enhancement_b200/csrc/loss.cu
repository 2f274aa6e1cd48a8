// Loss heads of the training step: BridgeModel._loss, loss_type "data_prediction_hybrid" (fdbm/model.py:187-218, the one
// config.yaml trains with) and "data_prediction" (model.py:163-185, the argparse default), both with pesq_weight = 0:
// forward AND the gradient w.r.t. the backbone output, without autograd.  Hybrid:
//
//   u = spec_back(z) = (|z| / f)^(1/e) e^{j angle z}                                  (data_module.py:188-199)
//   L = 70 mean((|u|^0.3 - |u^|^0.3)^2) + 30 sum |u/|u|^0.7 - u^/|u^|^0.7|^2 / N - mean_b log10 SI-SNR(istft u, istft u^)
//
// Passes: (1) spectral terms per element (loss partials + their gradient w.r.t. u^), (2) the two waveforms through
// the fused de-compress + iSTFT kernel, (3) per-utterance SI-SNR sums and the waveform gradient, (4) adjoint of the
// iSTFT = (c_k / n_fft) * STFT of the zero-padded waveform gradient (c_k = 1 for DC / Nyquist, 2 otherwise; the
// sqrt-Hann overlap-add envelope is 1), through the fused STFT kernel, (5) chain through spec_back, times loss scale.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <vector>

extern "C" int fdbm_stft_compress(const float* wave, int batch, int64_t n_samples, int64_t wave_stride, const float* window, int n_fft,
                                  int hop, int transform_type, float spec_factor, float abs_exponent, int pad_mode, int n_frames_out,
                                  float* spec, void* stream);
extern "C" int fdbm_decompress_istft(const float* spec, int batch, int n_frames, const float* window, int n_fft, int hop,
                                     int transform_type, float spec_factor, float abs_exponent, int64_t length, int64_t wave_stride,
                                     float* wave, void* stream);

namespace fdbm {
namespace {

struct LossScalars {          // device-side accumulators
  double mag, ri;             // sums of the two spectral terms
  double sisnr;               // sum_b log10(ratio_b)
};

__device__ __forceinline__ float2 spec_back_exp(float2 z, float inv_f_pow, float p) {
  // u = f^-p |z|^(p-1) z
  const float a = sqrtf(z.x * z.x + z.y * z.y);
  if (a == 0.f) return make_float2(0.f, 0.f);
  const float s = inv_f_pow * __powf(a, p - 1.0f);
  return make_float2(s * z.x, s * z.y);
}

// (1) per element: loss partials and d(70 L_mag + 30 L_ri)/d u^  (complex: dL/dRe + i dL/dIm)
__global__ void __launch_bounds__(256)
loss_spec_kernel(const float2* __restrict__ zh, const float2* __restrict__ zx, int64_t n, float inv_f_pow, float p, float inv_n,
                 float2* __restrict__ g_u, LossScalars* __restrict__ acc) {
  __shared__ double red[2][256];
  double lm = 0, lr = 0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float2 u = spec_back_exp(zx[i], inv_f_pow, p), uh = spec_back_exp(zh[i], inv_f_pow, p);
    const float2 v = make_float2(u.x + 1e-12f, u.y), vh = make_float2(uh.x + 1e-12f, uh.y);
    const float A = sqrtf(v.x * v.x + v.y * v.y), Ah = sqrtf(vh.x * vh.x + vh.y * vh.y);
    const float A3 = __powf(A, 0.3f), Ah3 = __powf(Ah, 0.3f);
    const float A7i = 1.0f / __powf(A, 0.7f), Ah7i = 1.0f / __powf(Ah, 0.7f);
    const float dm = Ah3 - A3;
    lm += static_cast<double>(dm) * dm;
    const float2 d = make_float2(uh.x * Ah7i - u.x * A7i, uh.y * Ah7i - u.y * A7i);
    lr += static_cast<double>(d.x) * d.x + static_cast<double>(d.y) * d.y;
    // gradients (see DESIGN notes in the header comment): mag: 2 dm 0.3 Ah^-0.7 vh / Ah; ri: 2 d Ah^-0.7 - 1.4 Re(conj(d) uh) Ah^-2.7 vh
    const float cm = 70.0f * inv_n * 0.6f * dm * Ah7i / Ah;
    const float re_du = d.x * uh.x + d.y * uh.y;
    const float cr = 30.0f * inv_n * 1.4f * re_du * Ah7i / (Ah * Ah);
    const float c2 = 30.0f * inv_n * 2.0f * Ah7i;
    g_u[i] = make_float2(cm * vh.x + c2 * d.x - cr * vh.x, cm * vh.y + c2 * d.y - cr * vh.y);
  }
  red[0][threadIdx.x] = lm; red[1][threadIdx.x] = lr;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) { red[0][threadIdx.x] += red[0][threadIdx.x + k]; red[1][threadIdx.x] += red[1][threadIdx.x + k]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { atomicAdd(&acc->mag, red[0][0]); atomicAdd(&acc->ri, red[1][0]); }
}

// (3) one block per utterance: SI-SNR sums (fixed order), log10 ratio, and the gradient of -mean_b log10(ratio_b) w.r.t. x^_td,
// written into the zero-padded buffer gpad[b][n_fft/2 + n]
__global__ void __launch_bounds__(1024)
sisnr_kernel(const float* __restrict__ x, const float* __restrict__ xh, int L, int B, int pad, float* __restrict__ gpad, LossScalars* __restrict__ acc) {
  __shared__ double red[3][1024];
  __shared__ double sh[6];
  const int b = blockIdx.x;
  const float* xb = x + static_cast<int64_t>(b) * L;
  const float* hb = xh + static_cast<int64_t>(b) * L;
  double a = 0, e = 0;
  for (int i = threadIdx.x; i < L; i += 1024) { a += static_cast<double>(xb[i]) * hb[i]; e += static_cast<double>(xb[i]) * xb[i]; }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = e;
  __syncthreads();
  for (int k = 512; k > 0; k >>= 1) {
    if (threadIdx.x < k) { red[0][threadIdx.x] += red[0][threadIdx.x + k]; red[1][threadIdx.x] += red[1][threadIdx.x + k]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sh[0] = red[0][0]; sh[1] = red[1][0]; }
  __syncthreads();
  const double dot = sh[0], xx = sh[1], E = xx + 1e-12;
  const double alpha = dot / E;
  double nn = 0, xr = 0;
  for (int i = threadIdx.x; i < L; i += 1024) { const double r = hb[i] - alpha * xb[i]; nn += r * r; xr += xb[i] * r; }
  __syncthreads();
  red[0][threadIdx.x] = nn; red[1][threadIdx.x] = xr;
  __syncthreads();
  for (int k = 512; k > 0; k >>= 1) {
    if (threadIdx.x < k) { red[0][threadIdx.x] += red[0][threadIdx.x + k]; red[1][threadIdx.x] += red[1][threadIdx.x + k]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sh[2] = red[0][0]; sh[3] = red[1][0]; }
  __syncthreads();
  const double S2 = alpha * alpha * xx, Nn = sh[2] + 1e-12, x_r = sh[3];
  const double ratio = S2 / Nn;
  const bool clamped = !(ratio > 1e-12);
  if (threadIdx.x == 0) atomicAdd(&acc->sisnr, log10(clamped ? 1e-12 : ratio));
  // d(-log10 ratio / B)/dx^ = -(1/(B ln 10)) (dS2/S2 - dNn/Nn);  dS2 = 2 alpha xx / E x;  dNn = 2 (x^ - s) - 2 x (x . (x^ - s)) / E
  const double k0 = clamped ? 0.0 : -1.0 / (B * 2.302585092994046);
  const double cs = S2 > 0 ? 2.0 * alpha * xx / (E * S2) : 0.0;
  float* gb = gpad + static_cast<int64_t>(b) * (L + 2 * pad) + pad;
  for (int i = threadIdx.x; i < L; i += 1024) {
    const double r = hb[i] - alpha * xb[i];
    const double dN = 2.0 * r - 2.0 * xb[i] * x_r / E;
    gb[i] = static_cast<float>(k0 * (cs * xb[i] - dN / Nn));
  }
}

// "data_prediction" (model.py:163-185): L = mean_b 0.5 sum_{f,t} |z^ - z|^2 / (F T)  +  l1_weight * mean_b 0.5 sum_n |x^_td - x_td| / L
// (1') the time-frequency term on the COMPRESSED spectrograms: sum |z^ - z|^2 (its gradient (z^ - z) / (F T B) is added in (5))
__global__ void __launch_bounds__(256)
loss_tf_kernel(const float2* __restrict__ zh, const float2* __restrict__ zx, int64_t n, LossScalars* __restrict__ acc) {
  __shared__ double red[256];
  double l = 0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float dx = zh[i].x - zx[i].x, dy = zh[i].y - zx[i].y;
    l += static_cast<double>(dx) * dx + static_cast<double>(dy) * dy;
  }
  red[threadIdx.x] = l;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(&acc->mag, red[0]);
}
// (3') one block per utterance: sum_n |x^_td - x_td| (fixed order) and its gradient coef * sign(x^_td - x_td) into gpad[b][pad + n]
__global__ void __launch_bounds__(1024)
l1_td_kernel(const float* __restrict__ x, const float* __restrict__ xh, int L, int pad, float coef, float* __restrict__ gpad,
             LossScalars* __restrict__ acc) {
  __shared__ double red[1024];
  const int b = blockIdx.x;
  const float* xb = x + static_cast<int64_t>(b) * L;
  const float* hb = xh + static_cast<int64_t>(b) * L;
  float* gb = gpad + static_cast<int64_t>(b) * (L + 2 * pad) + pad;
  double a = 0;
  for (int i = threadIdx.x; i < L; i += 1024) {
    const float d = hb[i] - xb[i];
    a += fabsf(d);
    gb[i] = d > 0.f ? coef : (d < 0.f ? -coef : 0.f);                   // torch.abs has gradient sign(d), 0 at d = 0
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int k = 512; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(&acc->ri, red[0]);
}
__global__ void loss_final_dp_kernel(const LossScalars* __restrict__ acc, double tf_scale, double l1_scale, float* __restrict__ loss) {
  *loss = static_cast<float>(acc->mag * tf_scale + acc->ri * l1_scale);
}

// (5) g_z = loss_scale * (chain through spec_back of (g_u + (c_k / n_fft) * Gt[b, k, t + 1]) + tf_coef (z^ - z)); g_u / zx may be null
__global__ void __launch_bounds__(256)
loss_chain_kernel(const float2* __restrict__ zh, const float2* __restrict__ g_u, const float2* __restrict__ Gt, int Fb, int T, int Tg,
                  int64_t n, float inv_f_pow, float p, float inv_nfft, float loss_scale, float2* __restrict__ g_out,
                  const float2* __restrict__ zx = nullptr, float tf_coef = 0.f) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const int t = static_cast<int>(i % T);
    const int k = static_cast<int>((i / T) % Fb);
    const int64_t b = i / (static_cast<int64_t>(T) * Fb);
    const float ck = (k == 0 || k == Fb - 1) ? 1.0f : 2.0f;
    const float2 gt = Gt[(b * Fb + k) * Tg + t + 1];
    float2 g = g_u ? g_u[i] : make_float2(0.f, 0.f);
    g.x += ck * inv_nfft * gt.x;
    g.y += (k == 0 || k == Fb - 1) ? 0.f : ck * inv_nfft * gt.y;        // irfft ignores the imaginary part of DC / Nyquist
    // u = f^-p |z|^(p-1) z:  g_z = f^-p |z|^(p-1) g + f^-p (p-1) |z|^(p-3) Re(conj(g) z) z
    const float2 z = zh[i];
    const float a = sqrtf(z.x * z.x + z.y * z.y);
    float2 o = make_float2(0.f, 0.f);
    if (a > 0.f) {
      const float s1 = inv_f_pow * __powf(a, p - 1.0f);
      const float s2 = inv_f_pow * (p - 1.0f) * __powf(a, p - 3.0f) * (g.x * z.x + g.y * z.y);
      o = make_float2(s1 * g.x + s2 * z.x, s1 * g.y + s2 * z.y);
    }
    if (zx) { o.x = fmaf(tf_coef, z.x - zx[i].x, o.x); o.y = fmaf(tf_coef, z.y - zx[i].y, o.y); }
    g_out[i] = make_float2(loss_scale * o.x, loss_scale * o.y);
  }
}

__global__ void loss_final_kernel(const LossScalars* __restrict__ acc, double inv_n, int B, float* __restrict__ loss) {
  *loss = static_cast<float>(70.0 * acc->mag * inv_n + 30.0 * acc->ri * inv_n - acc->sisnr / B);
}

int grid_for(int64_t n) { return static_cast<int>(std::min<int64_t>(ceil_div64(n, 256), static_cast<int64_t>(num_sms()) * 8)); }

}  // namespace
}  // namespace fdbm

using namespace fdbm;

extern "C" int64_t fdbm_hybrid_loss_workspace_bytes(int batch, int n_frames, int n_fft, int hop) {
  const int64_t Fb = n_fft / 2 + 1, L = static_cast<int64_t>(hop) * (n_frames - 1);
  const int64_t n = static_cast<int64_t>(batch) * Fb * n_frames;
  return 256 + n * 8 + 2 * batch * L * 4 + batch * (L + n_fft) * 4 + static_cast<int64_t>(batch) * Fb * (n_frames + 2) * 8 + 1024;
}

// x_hat (backbone output D) and x (clean), cplx [B,1,n_fft/2+1,T] compressed spectrograms -> *loss (device fp32) and
// g_out = loss_scale * dL/dx_hat, cplx [B,1,n_fft/2+1,T].  transform_type must be FDBM_TRANSFORM_EXPONENT.
extern "C" int fdbm_hybrid_loss(const float* x_hat, const float* x, int batch, int n_frames, const float* window, int n_fft, int hop,
                                int transform_type, float spec_factor, float abs_exponent, float loss_scale, void* workspace,
                                float* loss, float* g_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x_hat && x && window && workspace && loss && g_out && batch > 0 && n_frames > 1, "fdbm_hybrid_loss: bad arguments");
  FDBM_REQUIRE(transform_type == FDBM_TRANSFORM_EXPONENT, "fdbm_hybrid_loss: only the exponent transform is supported");
  FDBM_REQUIRE(n_fft == 2 * hop, "fdbm_hybrid_loss: the iSTFT adjoint assumes 50 %% overlap with a sqrt-Hann window (envelope 1)");
  cudaStream_t s = as_stream(stream);
  const int Fb = n_fft / 2 + 1;
  const int64_t L = static_cast<int64_t>(hop) * (n_frames - 1);
  const int64_t n = static_cast<int64_t>(batch) * Fb * n_frames;
  const int Tg = n_frames + 2;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  LossScalars* acc = reinterpret_cast<LossScalars*>(w); w += 256;
  float2* g_u = reinterpret_cast<float2*>(w); w += n * 8;
  float* x_td = reinterpret_cast<float*>(w); w += batch * L * 4;
  float* xh_td = reinterpret_cast<float*>(w); w += batch * L * 4;
  float* gpad = reinterpret_cast<float*>(w); w += batch * (L + n_fft) * 4;
  float2* Gt = reinterpret_cast<float2*>(w);
  const float p = 1.0f / abs_exponent;
  const float inv_f_pow = powf(spec_factor, -p);
  FDBM_CUDA(cudaMemsetAsync(acc, 0, sizeof(LossScalars), s));
  loss_spec_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const float2*>(x_hat), reinterpret_cast<const float2*>(x), n, inv_f_pow, p,
                                               1.0f / static_cast<float>(n), g_u, acc);
  FDBM_LAUNCH_CHECK();
  if (int rc = fdbm_decompress_istft(x, batch, n_frames, window, n_fft, hop, transform_type, spec_factor, abs_exponent, L, L, x_td, stream)) return rc;
  if (int rc = fdbm_decompress_istft(x_hat, batch, n_frames, window, n_fft, hop, transform_type, spec_factor, abs_exponent, L, L, xh_td, stream)) return rc;
  FDBM_CUDA(cudaMemsetAsync(gpad, 0, static_cast<size_t>(batch) * (L + n_fft) * 4, s));
  sisnr_kernel<<<batch, 1024, 0, s>>>(x_td, xh_td, static_cast<int>(L), batch, n_fft / 2, gpad, acc);
  FDBM_LAUNCH_CHECK();
  // adjoint of the iSTFT: frames 1..T of the centred STFT of the zero-padded gradient (no frame touches the reflection)
  if (int rc = fdbm_stft_compress(gpad, batch, L + n_fft, L + n_fft, window, n_fft, hop, FDBM_TRANSFORM_NONE, 1.0f, 1.0f, FDBM_PAD_ZERO, Tg,
                                  reinterpret_cast<float*>(Gt), stream)) return rc;
  loss_chain_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const float2*>(x_hat), g_u, Gt, Fb, n_frames, Tg, n, inv_f_pow, p,
                                                1.0f / n_fft, loss_scale, reinterpret_cast<float2*>(g_out));
  FDBM_LAUNCH_CHECK();
  loss_final_kernel<<<1, 1, 0, s>>>(acc, 1.0 / static_cast<double>(n), batch, loss);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

// loss_type "data_prediction" (fdbm/model.py:163-185, pesq_weight = 0): same arguments as fdbm_hybrid_loss plus l1_weight; the same
// workspace (fdbm_hybrid_loss_workspace_bytes).  target_len = (n_frames - 1) * hop.
extern "C" int fdbm_data_prediction_loss(const float* x_hat, const float* x, int batch, int n_frames, const float* window, int n_fft, int hop,
                                         int transform_type, float spec_factor, float abs_exponent, float l1_weight, float loss_scale,
                                         void* workspace, float* loss, float* g_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x_hat && x && window && workspace && loss && g_out && batch > 0 && n_frames > 1, "fdbm_data_prediction_loss: bad arguments");
  FDBM_REQUIRE(transform_type == FDBM_TRANSFORM_EXPONENT, "fdbm_data_prediction_loss: only the exponent transform is supported");
  FDBM_REQUIRE(n_fft == 2 * hop, "fdbm_data_prediction_loss: the iSTFT adjoint assumes 50 %% overlap with a sqrt-Hann window (envelope 1)");
  cudaStream_t s = as_stream(stream);
  const int Fb = n_fft / 2 + 1;
  const int64_t L = static_cast<int64_t>(hop) * (n_frames - 1);
  const int64_t n = static_cast<int64_t>(batch) * Fb * n_frames;
  const int Tg = n_frames + 2;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  LossScalars* acc = reinterpret_cast<LossScalars*>(w); w += 256;
  w += n * 8;                                                    // (the hybrid loss's g_u)
  float* x_td = reinterpret_cast<float*>(w); w += batch * L * 4;
  float* xh_td = reinterpret_cast<float*>(w); w += batch * L * 4;
  float* gpad = reinterpret_cast<float*>(w); w += batch * (L + n_fft) * 4;
  float2* Gt = reinterpret_cast<float2*>(w);
  const float p = 1.0f / abs_exponent;
  const float inv_f_pow = powf(spec_factor, -p);
  const double tf_scale = 0.5 / (static_cast<double>(Fb) * n_frames * batch), l1_scale = 0.5 * l1_weight / (static_cast<double>(L) * batch);
  FDBM_CUDA(cudaMemsetAsync(acc, 0, sizeof(LossScalars), s));
  loss_tf_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const float2*>(x_hat), reinterpret_cast<const float2*>(x), n, acc);
  FDBM_LAUNCH_CHECK();
  if (int rc = fdbm_decompress_istft(x, batch, n_frames, window, n_fft, hop, transform_type, spec_factor, abs_exponent, L, L, x_td, stream)) return rc;
  if (int rc = fdbm_decompress_istft(x_hat, batch, n_frames, window, n_fft, hop, transform_type, spec_factor, abs_exponent, L, L, xh_td, stream)) return rc;
  FDBM_CUDA(cudaMemsetAsync(gpad, 0, static_cast<size_t>(batch) * (L + n_fft) * 4, s));
  l1_td_kernel<<<batch, 1024, 0, s>>>(x_td, xh_td, static_cast<int>(L), n_fft / 2, static_cast<float>(l1_scale), gpad, acc);
  FDBM_LAUNCH_CHECK();
  if (int rc = fdbm_stft_compress(gpad, batch, L + n_fft, L + n_fft, window, n_fft, hop, FDBM_TRANSFORM_NONE, 1.0f, 1.0f, FDBM_PAD_ZERO, Tg,
                                  reinterpret_cast<float*>(Gt), stream)) return rc;
  loss_chain_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const float2*>(x_hat), nullptr, Gt, Fb, n_frames, Tg, n, inv_f_pow, p,
                                                1.0f / n_fft, loss_scale, reinterpret_cast<float2*>(g_out), reinterpret_cast<const float2*>(x),
                                                static_cast<float>(2.0 * tf_scale));
  FDBM_LAUNCH_CHECK();
  loss_final_dp_kernel<<<1, 1, 0, s>>>(acc, tf_scale, l1_scale, loss);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

// =====================================================================================================================
// "data_prediction_mel" / "data_prediction_melphase" (fdbm/model.py:220-251, fdbm/loss.py:9-33,213-289)
//
//   L = 0.5 mean |z^ - z|^2 + 0.1 sum_r mean | log10 clamp(M_r |STFT_r x^|, 1e-5)^2 - log10 clamp(M_r |STFT_r x|, 1e-5)^2 |  (+ 0.01 phase)
//
// over the seven resolutions n_fft = 32 .. 2048 (Hann window of n_fft, hop n_fft / 4, centred with reflection, n_mels =
// 5, 10, 20, 40, 80, 160, 210; model.py:77-92) of the two waveforms x^ = to_audio(z^), x = to_audio(z).  Per resolution three
// launches: (a) framing + window + radix-2 FFT in shared memory + magnitude + the (sparse, triangular) mel projection for
// both signals, the spectrum of x^ kept for (b); (b) the L1-of-log terms, the gradient back through log / clamp / mel / |.|
// and the adjoint FFT, giving d L / d frame; (c) overlap-add of the frame gradients with the reflection folded back, in a
// fixed order (no atomics), accumulated over the resolutions into the zero-padded waveform-gradient buffer that the shared
// iSTFT adjoint + spec_back chain of the other two heads consumes.  ~1 M complex points per utterance and resolution: the
// head is launch / latency bound (21 + 8 launches), a few hundred microseconds of a 60 ms step.
// Tables (windows, twiddles, librosa-convention Slaney mel filterbanks and their non-zero ranges) are built once by
// fdbm_mel_tables_init into a caller-owned device buffer: no library-global state.
// =====================================================================================================================
namespace fdbm {
namespace {

constexpr int kMelRes = 7;
constexpr int kMelNfft[kMelRes] = {32, 64, 128, 256, 512, 1024, 2048};
constexpr int kMelBands[kMelRes] = {5, 10, 20, 40, 80, 160, 210};
constexpr int kMelMaxN = 2048;
constexpr float kMelEps = 1e-5f;

struct MelResTables {           // device pointers into the tables buffer
  const float* window;          // [N] periodic Hann
  const float* basis;           // [n_mels][bins]
  const int* lo; const int* hi; // [n_mels] non-zero bin range of a filter
  const int* mlo; const int* mhi;  // [bins] filters that touch a bin (inclusive range; empty: mlo > mhi)
};
struct MelLayout { int64_t twiddle, window[kMelRes], basis[kMelRes], lo[kMelRes], hi[kMelRes], mlo[kMelRes], mhi[kMelRes], total; };

MelLayout mel_layout() {
  MelLayout l{};
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t at = o; o += (bytes + 255) / 256 * 256; return at; };
  l.twiddle = take(kMelMaxN / 2 * 8);
  for (int r = 0; r < kMelRes; ++r) {
    const int N = kMelNfft[r], bins = N / 2 + 1, M = kMelBands[r];
    l.window[r] = take(N * 4); l.basis[r] = take(static_cast<int64_t>(M) * bins * 4);
    l.lo[r] = take(M * 4); l.hi[r] = take(M * 4); l.mlo[r] = take(bins * 4); l.mhi[r] = take(bins * 4);
  }
  l.total = o;
  return l;
}
MelResTables mel_tables_of(const void* tables, const MelLayout& l, int r) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(tables);
  return MelResTables{reinterpret_cast<const float*>(b + l.window[r]), reinterpret_cast<const float*>(b + l.basis[r]),
                      reinterpret_cast<const int*>(b + l.lo[r]),       reinterpret_cast<const int*>(b + l.hi[r]),
                      reinterpret_cast<const int*>(b + l.mlo[r]),      reinterpret_cast<const int*>(b + l.mhi[r])};
}

// In-place radix-2 decimation-in-time FFTs of `frames` frames of N points (input in bit-reversed order) held in shared memory;
// sign = -1: forward e^{-j}, +1: adjoint e^{+j}.  tw = (cos, sin)(2 pi k / 2048), k < 1024.
__device__ __forceinline__ void smem_fft(float2* buf, int N, int logN, int frames, float sign, const float2* __restrict__ tw) {
  const int per = N >> 1, total = frames * per;
  for (int s = 0; s < logN; ++s) {
    const int half = 1 << s, tstep = kMelMaxN >> (s + 1);
    for (int j = threadIdx.x; j < total; j += blockDim.x) {
      const int f = j / per, q = j - f * per;
      const int pos = q & (half - 1), i0 = f * N + ((q >> s) << (s + 1)) + pos, i1 = i0 + half;
      const float2 w = tw[pos * tstep];
      const float wr = w.x, wi = sign * w.y;
      const float2 a = buf[i0], b = buf[i1];
      const float2 t = make_float2(b.x * wr - b.y * wi, b.x * wi + b.y * wr);
      buf[i0] = make_float2(a.x + t.x, a.y + t.y);
      buf[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
}
__device__ __forceinline__ int mel_frames_per_block(int N) { return N >= 512 ? 1 : 512 / N; }

// (a) grid (frame groups, B, 2 signals): mel[sig][b][t][m]; spectrum of x^ (sig 1) -> spec[b][t][k]
__global__ void __launch_bounds__(256)
mel_spec_kernel(const float* __restrict__ x_td, const float* __restrict__ xh_td, int L, int N, int logN, int hop, int n_frames, int n_mels,
                MelResTables tb, const float2* __restrict__ tw, float* __restrict__ mel, float2* __restrict__ spec) {
  __shared__ float2 buf[kMelMaxN];
  __shared__ float mag[1040];
  const int fpb = mel_frames_per_block(N), bins = N / 2 + 1;
  const int t0 = blockIdx.x * fpb, b = blockIdx.y, sig = blockIdx.z, B = gridDim.y;
  const int frames = min(fpb, n_frames - t0);
  const float* w = (sig ? xh_td : x_td) + static_cast<int64_t>(b) * L;
  for (int i = threadIdx.x; i < frames * N; i += blockDim.x) {
    const int f = i / N, n = i - f * N;
    int p = (t0 + f) * hop + n - N / 2;
    p = p < 0 ? -p : (p >= L ? 2 * (L - 1) - p : p);
    buf[f * N + (__brev(n) >> (32 - logN))] = make_float2(w[p] * tb.window[n], 0.f);
  }
  __syncthreads();
  smem_fft(buf, N, logN, frames, -1.f, tw);
  for (int i = threadIdx.x; i < frames * bins; i += blockDim.x) {
    const int f = i / bins, k = i - f * bins;
    const float2 v = buf[f * N + k];
    mag[i] = sqrtf(v.x * v.x + v.y * v.y);
    if (sig) spec[(static_cast<int64_t>(b) * n_frames + t0 + f) * bins + k] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < frames * n_mels; i += blockDim.x) {
    const int f = i / n_mels, m = i - f * n_mels;
    const float* row = tb.basis + static_cast<int64_t>(m) * bins;
    float a = 0.f;
    for (int k = tb.lo[m]; k < tb.hi[m]; ++k) a = fmaf(row[k], mag[f * bins + k], a);
    mel[((static_cast<int64_t>(sig) * B + b) * n_frames + t0 + f) * n_mels + m] = a;
  }
}

// (b) grid (frame groups, B): loss partial sum |D^ - D| * inv_count into acc->ri; frame gradients dfr[b][t][n]
__global__ void __launch_bounds__(256)
mel_grad_kernel(const float* __restrict__ mel, const float2* __restrict__ spec, int N, int logN, int n_frames, int n_mels, MelResTables tb,
                const float2* __restrict__ tw, double inv_count, float gcoef, float* __restrict__ dfr, LossScalars* __restrict__ acc) {
  __shared__ float2 buf[kMelMaxN];
  __shared__ float gm[1040];
  __shared__ double red[256];
  const int fpb = mel_frames_per_block(N), bins = N / 2 + 1;
  const int t0 = blockIdx.x * fpb, b = blockIdx.y, B = gridDim.y;
  const int frames = min(fpb, n_frames - t0);
  double part = 0;
  for (int i = threadIdx.x; i < frames * n_mels; i += blockDim.x) {
    const int f = i / n_mels, m = i - f * n_mels;
    const float mx = mel[((static_cast<int64_t>(b)) * n_frames + t0 + f) * n_mels + m];
    const float mh = mel[((static_cast<int64_t>(B) + b) * n_frames + t0 + f) * n_mels + m];
    const float cx = fmaxf(mx, kMelEps), ch = fmaxf(mh, kMelEps);
    const float d = log10f(ch * ch) - log10f(cx * cx);
    part += fabsf(d);
    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    gm[i] = mh >= kMelEps ? gcoef * sg * 0.8685889638065035f / mh : 0.f;        // d log10(m^2) / dm = 2 / (m ln 10)
  }
  for (int i = threadIdx.x; i < frames * N; i += blockDim.x) buf[i] = make_float2(0.f, 0.f);
  __syncthreads();
  for (int i = threadIdx.x; i < frames * bins; i += blockDim.x) {
    const int f = i / bins, k = i - f * bins;
    float dm = 0.f;
    for (int m = tb.mlo[k]; m <= tb.mhi[k]; ++m) dm = fmaf(tb.basis[static_cast<int64_t>(m) * bins + k], gm[f * n_mels + m], dm);
    const float2 v = spec[(static_cast<int64_t>(b) * n_frames + t0 + f) * bins + k];
    const float a = sqrtf(v.x * v.x + v.y * v.y);
    const float s = a > 0.f ? dm / a : 0.f;                                       // d|X| / d(Re, Im) = (Re, Im) / |X|, 0 at 0
    buf[f * N + (__brev(k) >> (32 - logN))] = make_float2(s * v.x, s * v.y);
  }
  __syncthreads();
  // d Re X_k / d frame[n] = w[n] cos(2 pi k n / N), d Im X_k / d frame[n] = -w[n] sin(..):  d frame[n] = w[n] Re sum_k G_k e^{+j 2 pi k n / N}
  smem_fft(buf, N, logN, frames, 1.f, tw);
  for (int i = threadIdx.x; i < frames * N; i += blockDim.x) {
    const int f = i / N, n = i - f * N;
    dfr[(static_cast<int64_t>(b) * n_frames + t0 + f) * N + n] = tb.window[n] * buf[i].x;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(&acc->ri, red[0] * inv_count);
}

// (c) gpad[b][pad + i] += P[i + N/2] + reflected images, P[p] = sum_t dfr[b][t][p - t hop] (fixed order)
__global__ void __launch_bounds__(256)
mel_ola_kernel(const float* __restrict__ dfr, int L, int N, int hop, int n_frames, int pad, float* __restrict__ gpad) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= L) return;
  const float* fr = dfr + static_cast<int64_t>(b) * n_frames * N;
  auto P = [&](int p) {
    const int t_hi = min(n_frames - 1, p / hop), t_lo = max(0, (p - N + hop) / hop);
    float a = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) a += fr[static_cast<int64_t>(t) * N + (p - t * hop)];
    return a;
  };
  const int h = N / 2;
  float g = P(i + h);
  if (i >= 1 && i <= h) g += P(h - i);                                  // left reflection: padded p = h - i holds sample i
  if (i <= L - 2 && i >= L - 1 - h) g += P(h + 2 * (L - 1) - i);        // right reflection
  gpad[static_cast<int64_t>(b) * (L + 2 * pad) + pad + i] += g;
}

// PhaseLoss (loss.py:9-33) on the compressed spectrograms: value (sum of the three |anti-wrapped differences|, into acc->sisnr) and
// g_out += loss_scale * coef * dL/dphase * d angle / d(Re, Im).  d[j] = p[j-1] - p[j], d[0] = -p[0] along frequency / frames.
__device__ __forceinline__ float anti_wrap(float v) { return v - 6.283185307179586f * rintf(v / 6.283185307179586f); }
__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
__global__ void __launch_bounds__(256)
phase_loss_kernel(const float2* __restrict__ zh, const float2* __restrict__ zx, int Fb, int T, int64_t n, float coef, float loss_scale,
                  float2* __restrict__ g_out, LossScalars* __restrict__ acc) {
  __shared__ double red[256];
  double part = 0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const int t = static_cast<int>(i % T), k = static_cast<int>((i / T) % Fb);
    auto ph = [&](const float2* z, int64_t j) { const float2 v = z[j]; return atan2f(v.y, v.x); };
    const float2 z = zh[i];
    const float pg = atan2f(z.y, z.x), pr = ph(zx, i);
    // differences owned by this element (index k along frequency, t along frames) and the next ones, which also contain pg
    const float pg_km = k > 0 ? ph(zh, i - T) : 0.f, pr_km = k > 0 ? ph(zx, i - T) : 0.f;
    const float pg_tm = t > 0 ? ph(zh, i - 1) : 0.f, pr_tm = t > 0 ? ph(zx, i - 1) : 0.f;
    const float u_ip = anti_wrap(pr - pg);
    const float u_gd = anti_wrap((pr_km - pr) - (pg_km - pg));
    const float u_td = anti_wrap((pr_tm - pr) - (pg_tm - pg));
    part += fabsf(u_ip) + fabsf(u_gd) + fabsf(u_td);
    float g = -sgnf(u_ip) + sgnf(u_gd) + sgnf(u_td);
    if (k + 1 < Fb) g -= sgnf(anti_wrap((pr - ph(zx, i + T)) - (pg - ph(zh, i + T))));
    if (t + 1 < T) g -= sgnf(anti_wrap((pr - ph(zx, i + 1)) - (pg - ph(zh, i + 1))));
    const float a2 = z.x * z.x + z.y * z.y;
    if (a2 > 0.f) {
      const float s = loss_scale * coef * g / a2;
      float2 o = g_out[i];
      o.x -= s * z.y; o.y += s * z.x;
      g_out[i] = o;
    }
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(&acc->sisnr, red[0]);
}
__global__ void loss_final_mel_kernel(const LossScalars* __restrict__ acc, double tf_scale, double phase_scale, float* __restrict__ loss) {
  *loss = static_cast<float>(acc->mag * tf_scale + 0.1 * acc->ri + phase_scale * acc->sisnr);
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin = 0, fmax = sr / 2) with its defaults (Slaney scale, norm = 'slaney'), float32 result
void mel_filterbank_host(int sr, int n_fft, int n_mels, std::vector<float>& basis) {
  const int bins = n_fft / 2 + 1;
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  auto hz_to_mel = [&](double f) { return f < min_log_hz ? f / f_sp : min_log_mel + log(f / min_log_hz) / logstep; };
  auto mel_to_hz = [&](double m) { return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m; };
  const double m_lo = hz_to_mel(0.0), m_hi = hz_to_mel(sr / 2.0);
  std::vector<double> f(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
  basis.assign(static_cast<size_t>(n_mels) * bins, 0.f);
  for (int m = 0; m < n_mels; ++m) {
    const double enorm = 2.0 / (f[m + 2] - f[m]);
    for (int k = 0; k < bins; ++k) {
      const double fk = (sr / 2.0) * k / (bins - 1);
      const double lower = (fk - f[m]) / (f[m + 1] - f[m]), upper = (f[m + 2] - fk) / (f[m + 2] - f[m + 1]);
      const double w = std::max(0.0, std::min(lower, upper));
      basis[static_cast<size_t>(m) * bins + k] = static_cast<float>(w * enorm);
    }
  }
}

int64_t mel_ws_extra(int batch, int64_t L, int64_t* mel_b, int64_t* spec_b, int64_t* dfr_b) {
  int64_t mx_mel = 0, mx_spec = 0, mx_dfr = 0;
  for (int r = 0; r < kMelRes; ++r) {
    const int64_t N = kMelNfft[r], nfr = 1 + L / (N / 4);
    mx_mel = std::max(mx_mel, 2 * batch * nfr * kMelBands[r] * 4);
    mx_spec = std::max(mx_spec, batch * nfr * (N / 2 + 1) * 8);
    mx_dfr = std::max(mx_dfr, batch * nfr * N * 4);
  }
  auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
  *mel_b = up(mx_mel); *spec_b = up(mx_spec); *dfr_b = up(mx_dfr);
  return *mel_b + *spec_b + *dfr_b;
}

}  // namespace
}  // namespace fdbm

extern "C" int64_t fdbm_mel_tables_bytes(void) { return mel_layout().total; }

extern "C" int fdbm_mel_tables_init(void* tables, int sample_rate, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(tables && sample_rate > 0, "fdbm_mel_tables_init: bad arguments");
  const MelLayout l = mel_layout();
  std::vector<uint8_t> host(static_cast<size_t>(l.total), 0);
  const double two_pi = 6.283185307179586476925;
  float* tw = reinterpret_cast<float*>(host.data() + l.twiddle);
  for (int k = 0; k < kMelMaxN / 2; ++k) { tw[2 * k] = static_cast<float>(cos(two_pi * k / kMelMaxN)); tw[2 * k + 1] = static_cast<float>(sin(two_pi * k / kMelMaxN)); }
  std::vector<float> basis;
  for (int r = 0; r < kMelRes; ++r) {
    const int N = kMelNfft[r], bins = N / 2 + 1, M = kMelBands[r];
    float* win = reinterpret_cast<float*>(host.data() + l.window[r]);
    for (int n = 0; n < N; ++n) win[n] = static_cast<float>(0.5 - 0.5 * cos(two_pi * n / N));       // torch.hann_window (periodic)
    mel_filterbank_host(sample_rate, N, M, basis);
    memcpy(host.data() + l.basis[r], basis.data(), basis.size() * 4);
    int* lo = reinterpret_cast<int*>(host.data() + l.lo[r]); int* hi = reinterpret_cast<int*>(host.data() + l.hi[r]);
    int* mlo = reinterpret_cast<int*>(host.data() + l.mlo[r]); int* mhi = reinterpret_cast<int*>(host.data() + l.mhi[r]);
    for (int k = 0; k < bins; ++k) { mlo[k] = 1; mhi[k] = 0; }
    for (int m = 0; m < M; ++m) {
      lo[m] = hi[m] = 0;
      bool any = false;
      for (int k = 0; k < bins; ++k) {
        if (basis[static_cast<size_t>(m) * bins + k] == 0.f) continue;
        if (!any) { lo[m] = k; any = true; }
        hi[m] = k + 1;
        if (mlo[k] > mhi[k]) mlo[k] = m;
        mhi[k] = m;
      }
    }
  }
  cudaStream_t s = as_stream(stream);
  FDBM_CUDA(cudaMemcpyAsync(tables, host.data(), host.size(), cudaMemcpyHostToDevice, s));
  FDBM_CUDA(cudaStreamSynchronize(s));
  return FDBM_OK;
}

extern "C" int64_t fdbm_mel_loss_workspace_bytes(int batch, int n_frames, int n_fft, int hop) {
  int64_t a, b, c;
  return (fdbm_hybrid_loss_workspace_bytes(batch, n_frames, n_fft, hop) + 255) / 256 * 256 + mel_ws_extra(batch, static_cast<int64_t>(hop) * (n_frames - 1), &a, &b, &c);
}

extern "C" int fdbm_mel_loss(const float* x_hat, const float* x, int batch, int n_frames, const float* window, int n_fft, int hop,
                             int transform_type, float spec_factor, float abs_exponent, int with_phase, float loss_scale,
                             const void* tables, void* workspace, float* loss, float* g_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x_hat && x && window && tables && workspace && loss && g_out && batch > 0 && n_frames > 1, "fdbm_mel_loss: bad arguments");
  FDBM_REQUIRE(transform_type == FDBM_TRANSFORM_EXPONENT, "fdbm_mel_loss: only the exponent transform is supported");
  FDBM_REQUIRE(n_fft == 2 * hop, "fdbm_mel_loss: the iSTFT adjoint assumes 50 %% overlap with a sqrt-Hann window (envelope 1)");
  cudaStream_t s = as_stream(stream);
  const int Fb = n_fft / 2 + 1;
  const int64_t L = static_cast<int64_t>(hop) * (n_frames - 1);
  FDBM_REQUIRE(L > kMelMaxN / 2 && L < (1ll << 30) && batch <= 65535,
               "fdbm_mel_loss: target_len must exceed 1024 samples (reflection of the 2048-point resolution), batch <= 65535");
  const int64_t n = static_cast<int64_t>(batch) * Fb * n_frames;
  const int Tg = n_frames + 2;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  LossScalars* acc = reinterpret_cast<LossScalars*>(w); w += 256;
  w += n * 8;                                                    // (the hybrid loss's g_u)
  float* x_td = reinterpret_cast<float*>(w); w += batch * L * 4;
  float* xh_td = reinterpret_cast<float*>(w); w += batch * L * 4;
  float* gpad = reinterpret_cast<float*>(w); w += batch * (L + n_fft) * 4;
  float2* Gt = reinterpret_cast<float2*>(w);
  w = reinterpret_cast<uint8_t*>(workspace) + (fdbm_hybrid_loss_workspace_bytes(batch, n_frames, n_fft, hop) + 255) / 256 * 256;
  int64_t mel_b, spec_b, dfr_b;
  mel_ws_extra(batch, L, &mel_b, &spec_b, &dfr_b);
  float* mel = reinterpret_cast<float*>(w); w += mel_b;
  float2* spec = reinterpret_cast<float2*>(w); w += spec_b;
  float* dfr = reinterpret_cast<float*>(w);
  const float p = 1.0f / abs_exponent;
  const float inv_f_pow = powf(spec_factor, -p);
  const double tf_scale = 0.5 / static_cast<double>(n);
  const MelLayout lay = mel_layout();
  const float2* tw = reinterpret_cast<const float2*>(reinterpret_cast<const uint8_t*>(tables) + lay.twiddle);

  FDBM_CUDA(cudaMemsetAsync(acc, 0, sizeof(LossScalars), s));
  loss_tf_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const float2*>(x_hat), reinterpret_cast<const float2*>(x), n, acc);
  FDBM_LAUNCH_CHECK();
  if (int rc = fdbm_decompress_istft(x, batch, n_frames, window, n_fft, hop, transform_type, spec_factor, abs_exponent, L, L, x_td, stream)) return rc;
  if (int rc = fdbm_decompress_istft(x_hat, batch, n_frames, window, n_fft, hop, transform_type, spec_factor, abs_exponent, L, L, xh_td, stream)) return rc;
  FDBM_CUDA(cudaMemsetAsync(gpad, 0, static_cast<size_t>(batch) * (L + n_fft) * 4, s));
  for (int r = 0; r < kMelRes; ++r) {
    const int N = kMelNfft[r], hp = N / 4, M = kMelBands[r], nfr = 1 + static_cast<int>(L / hp);
    int logN = 0; while ((1 << logN) < N) ++logN;
    const int fpb = N >= 512 ? 1 : 512 / N;
    const MelResTables tb = mel_tables_of(tables, lay, r);
    const dim3 grid_a(ceil_div(nfr, fpb), batch, 2), grid_b(ceil_div(nfr, fpb), batch, 1);
    mel_spec_kernel<<<grid_a, 256, 0, s>>>(x_td, xh_td, static_cast<int>(L), N, logN, hp, nfr, M, tb, tw, mel, spec);
    FDBM_LAUNCH_CHECK();
    const double inv_count = 1.0 / (static_cast<double>(batch) * M * nfr);
    mel_grad_kernel<<<grid_b, 256, 0, s>>>(mel, spec, N, logN, nfr, M, tb, tw, inv_count, static_cast<float>(0.1 * inv_count), dfr, acc);
    FDBM_LAUNCH_CHECK();
    mel_ola_kernel<<<dim3(ceil_div(static_cast<int>(L), 256), batch), 256, 0, s>>>(dfr, static_cast<int>(L), N, hp, nfr, n_fft / 2, gpad);
    FDBM_LAUNCH_CHECK();
  }
  if (int rc = fdbm_stft_compress(gpad, batch, L + n_fft, L + n_fft, window, n_fft, hop, FDBM_TRANSFORM_NONE, 1.0f, 1.0f, FDBM_PAD_ZERO, Tg,
                                  reinterpret_cast<float*>(Gt), stream)) return rc;
  loss_chain_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const float2*>(x_hat), nullptr, Gt, Fb, n_frames, Tg, n, inv_f_pow, p,
                                                1.0f / n_fft, loss_scale, reinterpret_cast<float2*>(g_out), reinterpret_cast<const float2*>(x),
                                                static_cast<float>(2.0 * tf_scale));
  FDBM_LAUNCH_CHECK();
  if (with_phase) {
    phase_loss_kernel<<<grid_for(n), 256, 0, s>>>(reinterpret_cast<const float2*>(x_hat), reinterpret_cast<const float2*>(x), Fb, n_frames, n,
                                                  static_cast<float>(0.01 / static_cast<double>(n)), loss_scale, reinterpret_cast<float2*>(g_out), acc);
    FDBM_LAUNCH_CHECK();
  }
  loss_final_mel_kernel<<<1, 1, 0, s>>>(acc, tf_scale, with_phase ? 0.01 / static_cast<double>(n) : 0.0, loss);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}
