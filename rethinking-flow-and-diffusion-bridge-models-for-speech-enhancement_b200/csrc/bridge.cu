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
