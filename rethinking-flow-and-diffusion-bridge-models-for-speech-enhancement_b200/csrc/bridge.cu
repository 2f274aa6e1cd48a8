// Bridge sampler arithmetic: prior sample and the per-step update, one HBM-bound elementwise
// kernel each.  Replaces fdbm/bridge.py:45-49 (prior_sampling) and the loop bodies :73-85 (ODE)
// and :96-111 (SDE).  Algorithmic traffic: 32 B per complex element (ODE: read x, d, y, write x),
// 24 B (SDE with in-kernel Philox noise).
#include "common.cuh"

namespace fdbm {
namespace {

// Philox4x32-10 (Salmon et al. 2011), counter = (element index, stream offset), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// complex standard normal: real and imaginary parts ~ N(0, 1/2) (torch.randn_like on complex64)
__device__ __forceinline__ float2 complex_normal(uint64_t seed, uint64_t offset, uint64_t idx) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32),
                                           static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32)),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const float u1 = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  const float u2 = (static_cast<float>(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-logf(u1));                                              // sqrt(-2 ln u1) * sqrt(1/2)
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(rad * c, rad * s);
}

__global__ void __launch_bounds__(256)
prior_sample_kernel(const float2* __restrict__ y, const float2* __restrict__ z, float b, float sigma, uint64_t seed,
                    uint64_t offset, int64_t n, float2* __restrict__ x) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float2 yy = y[i];
    float2 zz = make_float2(0.f, 0.f);
    if (z) zz = z[i];
    else if (sigma != 0.f) zz = complex_normal(seed, offset, static_cast<uint64_t>(i));
    // y*b + z*sigma, each product rounded separately (no FMA): bridge.py:48
    x[i] = make_float2(__fadd_rn(__fmul_rn(yy.x, b), __fmul_rn(zz.x, sigma)),
                       __fadd_rn(__fmul_rn(yy.y, b), __fmul_rn(zz.y, sigma)));
  }
}

// x <- (wx*x + ws*d) + w3*{y|z}.  float4 = two complex elements per thread-iteration.
template <int KIND>
__global__ void __launch_bounds__(256)
bridge_step_kernel(float4* __restrict__ x, const float4* __restrict__ d, const float4* __restrict__ yz,
                   const float* __restrict__ coef, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ rng,
                   int64_t n_pairs) {
  const float wx = coef[0], ws = coef[1], w3 = coef[2];
  if (rng) { seed = rng[0]; offset += rng[1]; }           // graph replays read the seed from device memory
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_pairs; i += 256ll * gridDim.x) {
    const float4 xv = x[i], dv = d[i];
    float4 t;
    if (KIND == FDBM_STEP_ODE || yz != nullptr) {
      t = yz[i];
    } else {
      const float2 z0 = complex_normal(seed, offset, static_cast<uint64_t>(2 * i));
      const float2 z1 = complex_normal(seed, offset, static_cast<uint64_t>(2 * i + 1));
      t = make_float4(z0.x, z0.y, z1.x, z1.y);
    }
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fmul_rn(wx, xv.x), __fmul_rn(ws, dv.x)), __fmul_rn(w3, t.x));
    o.y = __fadd_rn(__fadd_rn(__fmul_rn(wx, xv.y), __fmul_rn(ws, dv.y)), __fmul_rn(w3, t.y));
    o.z = __fadd_rn(__fadd_rn(__fmul_rn(wx, xv.z), __fmul_rn(ws, dv.z)), __fmul_rn(w3, t.z));
    o.w = __fadd_rn(__fadd_rn(__fmul_rn(wx, xv.w), __fmul_rn(ws, dv.w)), __fmul_rn(w3, t.w));
    x[i] = o;
  }
}

// Predictor-corrector sampler arithmetic (bridge.py:142-166, util/predictors.py:39-51, util/correctors.py:36-81): every
// update of that sampler is  x_mean = c0*x + c1*d + c2*y ;  x = x_mean + c3*z  with per-step scalars (DEVICE coef[4]).
__global__ void __launch_bounds__(256)
update4_kernel(float4* __restrict__ x, const float4* __restrict__ d, const float4* __restrict__ y, const float4* __restrict__ z,
               const float* __restrict__ coef, uint64_t seed, uint64_t offset, int64_t n_pairs, float4* __restrict__ x_mean_out) {
  const float c0 = coef[0], c1 = coef[1], c2 = coef[2], c3 = coef[3];
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_pairs; i += 256ll * gridDim.x) {
    const float4 xv = x[i], dv = d[i], yv = y[i];
    float4 zv;
    if (z) {
      zv = z[i];
    } else if (c3 != 0.f) {
      const float2 z0 = complex_normal(seed, offset, static_cast<uint64_t>(2 * i));
      const float2 z1 = complex_normal(seed, offset, static_cast<uint64_t>(2 * i + 1));
      zv = make_float4(z0.x, z0.y, z1.x, z1.y);
    } else {
      zv = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 m;
    m.x = fmaf(c2, yv.x, fmaf(c1, dv.x, c0 * xv.x)); m.y = fmaf(c2, yv.y, fmaf(c1, dv.y, c0 * xv.y));
    m.z = fmaf(c2, yv.z, fmaf(c1, dv.z, c0 * xv.z)); m.w = fmaf(c2, yv.w, fmaf(c1, dv.w, c0 * xv.w));
    if (x_mean_out) x_mean_out[i] = m;
    x[i] = make_float4(fmaf(c3, zv.x, m.x), fmaf(c3, zv.y, m.y), fmaf(c3, zv.z, m.z), fmaf(c3, zv.w, m.w));
  }
}

// Langevin corrector (correctors.py:44-52): per-utterance sums of |x - a d - b y|^2 and |z|^2 (z regenerated from the same
// Philox counters the update kernel will use when no noise tensor is given) ...
__global__ void __launch_bounds__(256)
langevin_norms_kernel(const float2* __restrict__ x, const float2* __restrict__ d, const float2* __restrict__ y,
                      const float2* __restrict__ z, float a, float b, uint64_t seed, uint64_t offset, int64_t n_per_utt,
                      double* __restrict__ sums) {
  const int u = blockIdx.y;
  const int64_t base = static_cast<int64_t>(u) * n_per_utt;
  double sg = 0.0, sz = 0.0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n_per_utt; i += 256ll * gridDim.x) {
    const float2 xv = x[base + i], dv = d[base + i], yv = y[base + i];
    const float2 zv = z ? z[base + i] : complex_normal(seed, offset, static_cast<uint64_t>(base + i));
    const float gr = xv.x - (a * dv.x + b * yv.x), gi = xv.y - (a * dv.y + b * yv.y);
    sg += static_cast<double>(gr) * gr + static_cast<double>(gi) * gi;
    sz += static_cast<double>(zv.x) * zv.x + static_cast<double>(zv.y) * zv.y;
  }
  __shared__ double red[2][256];
  red[0][threadIdx.x] = sg; red[1][threadIdx.x] = sz;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) { red[0][threadIdx.x] += red[0][threadIdx.x + k]; red[1][threadIdx.x] += red[1][threadIdx.x + k]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { atomicAdd(&sums[2 * u], red[0][0]); atomicAdd(&sums[2 * u + 1], red[1][0]); }
}
// ... and the step size step = 2 (snr * mean_b |z_b| / (mean_b |grad_b| + 1e-8))^2 turned into the four update weights
__global__ void langevin_coef_kernel(const double* __restrict__ sums, int batch, float a, float b, float sigma, float snr,
                                     float* __restrict__ coef) {
  if (threadIdx.x != 0) return;
  const float k = 1.0f / (sigma * sigma + 1e-8f);
  float gn = 0.f, zn = 0.f;
  for (int u = 0; u < batch; ++u) { gn += k * sqrtf(static_cast<float>(sums[2 * u])); zn += sqrtf(static_cast<float>(sums[2 * u + 1])); }
  gn /= batch; zn /= batch;
  const float r = snr * zn / (gn + 1e-8f);
  const float step = r * r * 2.0f;
  coef[0] = 1.0f - step * k; coef[1] = step * k * a; coef[2] = step * k * b; coef[3] = sqrtf(step * 2.0f);
}

// Adaptive ODE sampler (bridge.py:115-140, scipy RK45): linear combinations of up to 8 state-sized tensors and the
// scaled RMS error norm of scipy.integrate's step-size controller.
struct LinComb { const float4* src[8]; float c[8]; int n; };
__global__ void __launch_bounds__(256) lincomb_kernel(LinComb a, float4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += 256ll * gridDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < a.n) {
        const float4 v = a.src[k][i]; const float c = a.c[k];
        acc.x = fmaf(c, v.x, acc.x); acc.y = fmaf(c, v.y, acc.y); acc.z = fmaf(c, v.z, acc.z); acc.w = fmaf(c, v.w, acc.w);
      }
    }
    out[i] = acc;
  }
}
struct ErrNorm { const float2* src[8]; float c[8]; int n; };
// sum over complex elements of |sum_k c_k src_k|^2 / (atol + rtol max(|y|, |y_new|))^2   (RK45._estimate_error_norm)
__global__ void __launch_bounds__(256)
rk_error_norm_kernel(ErrNorm a, const float2* __restrict__ y, const float2* __restrict__ y_new, float rtol, float atol, int64_t n,
                     double* __restrict__ out) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    float er = 0.f, ei = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < a.n) { const float2 v = a.src[k][i]; er = fmaf(a.c[k], v.x, er); ei = fmaf(a.c[k], v.y, ei); }
    const float2 p = y[i], q = y_new[i];
    const float sc = atol + rtol * fmaxf(hypotf(p.x, p.y), hypotf(q.x, q.y));
    s += (static_cast<double>(er) * er + static_cast<double>(ei) * ei) / (static_cast<double>(sc) * sc);
  }
  __shared__ double red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) atomicAdd(out, red[0]);
}

}  // namespace

static int grid_for(int64_t n_threads_needed) {
  const int64_t blocks = ceil_div64(n_threads_needed, 256);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;       // 16 resident 256-thread CTAs cover the SM
  return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace fdbm

using namespace fdbm;

extern "C" int fdbm_prior_sample(const float* y, const float* z, float b, float sigma, uint64_t seed, uint64_t offset,
                                 int64_t n_complex, float* x, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(y && x && n_complex > 0, "fdbm_prior_sample: null pointer or empty");
  prior_sample_kernel<<<grid_for(n_complex), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float2*>(y), reinterpret_cast<const float2*>(z), b, sigma, seed, offset, n_complex,
      reinterpret_cast<float2*>(x));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

static int bridge_step_impl(float* x, const float* d, const float* y_or_z, const float* coef, int kind, uint64_t seed,
                            uint64_t offset, const uint64_t* rng, int64_t n_complex, cudaStream_t stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x && d && coef && n_complex > 0, "fdbm_bridge_step: null pointer or empty");
  FDBM_REQUIRE(n_complex % 2 == 0, "fdbm_bridge_step: n_complex must be even (16-byte vectors)");
  FDBM_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(y_or_z)) % 16 == 0,
               "fdbm_bridge_step: pointers must be 16-byte aligned");
  FDBM_REQUIRE(kind == FDBM_STEP_ODE || kind == FDBM_STEP_SDE, "fdbm_bridge_step: bad kind %d", kind);
  FDBM_REQUIRE(kind == FDBM_STEP_SDE || y_or_z, "fdbm_bridge_step: ODE step needs y");
  const int64_t n_pairs = n_complex / 2;
  const int grid = grid_for(n_pairs);
  if (kind == FDBM_STEP_ODE)
    bridge_step_kernel<FDBM_STEP_ODE><<<grid, 256, 0, stream>>>(
        reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(d), reinterpret_cast<const float4*>(y_or_z), coef,
        seed, offset, rng, n_pairs);
  else
    bridge_step_kernel<FDBM_STEP_SDE><<<grid, 256, 0, stream>>>(
        reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(d), reinterpret_cast<const float4*>(y_or_z), coef,
        seed, offset, rng, n_pairs);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

namespace fdbm {
int launch_bridge_step_rng(float* x, const float* d, const float* third, const float* coef, int kind,
                           const uint64_t* rng, uint64_t offset, int64_t n_complex, cudaStream_t s) {
  return bridge_step_impl(x, d, third, coef, kind, 0, offset, rng, n_complex, s);
}
}  // namespace fdbm

extern "C" int fdbm_bridge_step(float* x, const float* d, const float* y_or_z, const float* coef, int kind,
                                uint64_t seed, uint64_t offset, int64_t n_complex, void* stream) {
  return bridge_step_impl(x, d, y_or_z, coef, kind, seed, offset, nullptr, n_complex, as_stream(stream));
}

extern "C" int fdbm_bridge_update4(float* x, const float* d, const float* y, const float* z, const float* coef, uint64_t seed,
                                   uint64_t offset, int64_t n_complex, float* x_mean_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x && d && y && coef && n_complex > 0 && n_complex % 2 == 0, "fdbm_bridge_update4: null pointer, empty or odd size");
  FDBM_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z) |
                reinterpret_cast<uintptr_t>(x_mean_out)) % 16 == 0, "fdbm_bridge_update4: pointers must be 16-byte aligned");
  const int64_t n_pairs = n_complex / 2;
  update4_kernel<<<grid_for(n_pairs), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(d), reinterpret_cast<const float4*>(y),
      reinterpret_cast<const float4*>(z), coef, seed, offset, n_pairs, reinterpret_cast<float4*>(x_mean_out));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_langevin_coef(const float* x, const float* d, const float* y, const float* z, float a, float b, float sigma,
                                  float snr, uint64_t seed, uint64_t offset, int batch, int64_t n_per_utt, double* scratch,
                                  float* coef_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(x && d && y && scratch && coef_out && batch > 0 && n_per_utt > 0, "fdbm_langevin_coef: bad arguments");
  cudaStream_t s = as_stream(stream);
  FDBM_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * batch, s));
  const int bx = std::max(1, std::min<int>(static_cast<int>(ceil_div64(n_per_utt, 256 * 8)), num_sms() * 8 / batch + 1));
  langevin_norms_kernel<<<dim3(bx, batch), 256, 0, s>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(d),
                                                        reinterpret_cast<const float2*>(y), reinterpret_cast<const float2*>(z), a, b,
                                                        seed, offset, n_per_utt, scratch);
  FDBM_LAUNCH_CHECK();
  langevin_coef_kernel<<<1, 32, 0, s>>>(scratch, batch, a, b, sigma, snr, coef_out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_lincomb(float* out, const float* const* srcs, const float* coefs, int n_terms, int64_t n_floats, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(out && srcs && coefs && n_terms >= 1 && n_terms <= 8 && n_floats > 0 && n_floats % 4 == 0, "fdbm_lincomb: bad arguments");
  LinComb a{};
  a.n = n_terms;
  for (int k = 0; k < n_terms; ++k) {
    FDBM_REQUIRE(srcs[k] && reinterpret_cast<uintptr_t>(srcs[k]) % 16 == 0, "fdbm_lincomb: source %d is null or not 16-byte aligned", k);
    a.src[k] = reinterpret_cast<const float4*>(srcs[k]); a.c[k] = coefs[k];
  }
  FDBM_REQUIRE(reinterpret_cast<uintptr_t>(out) % 16 == 0, "fdbm_lincomb: out must be 16-byte aligned");
  lincomb_kernel<<<grid_for(n_floats / 4), 256, 0, as_stream(stream)>>>(a, reinterpret_cast<float4*>(out), n_floats / 4);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_rk_error_norm(const float* const* srcs, const float* coefs, int n_terms, const float* y, const float* y_new,
                                  float rtol, float atol, int64_t n_complex, double* out_sumsq, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(srcs && coefs && n_terms >= 1 && n_terms <= 8 && y && y_new && out_sumsq && n_complex > 0, "fdbm_rk_error_norm: bad arguments");
  ErrNorm a{};
  a.n = n_terms;
  for (int k = 0; k < n_terms; ++k) { FDBM_REQUIRE(srcs[k], "fdbm_rk_error_norm: source %d is null", k); a.src[k] = reinterpret_cast<const float2*>(srcs[k]); a.c[k] = coefs[k]; }
  cudaStream_t s = as_stream(stream);
  FDBM_CUDA(cudaMemsetAsync(out_sumsq, 0, sizeof(double), s));
  rk_error_norm_kernel<<<grid_for(n_complex), 256, 0, s>>>(a, reinterpret_cast<const float2*>(y), reinterpret_cast<const float2*>(y_new), rtol,
                                                           atol, n_complex, out_sumsq);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}
