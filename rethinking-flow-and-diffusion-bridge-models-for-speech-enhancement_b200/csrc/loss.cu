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
