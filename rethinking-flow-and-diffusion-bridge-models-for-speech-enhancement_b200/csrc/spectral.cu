// Fused STFT + amplitude compression (+ pad_spec) and its inverse (de-compression + iSTFT
// overlap-add).  Replaces fdbm/data_module.py:173-229 and fdbm/util/other.py:76-90.
//
// One block transforms 16 consecutive frames of one utterance.  Two real frames are packed into
// one 512-point complex FFT (real -> re, second frame -> im) and separated by Hermitian symmetry,
// so a warp produces two spectra per shared-memory radix-2 pass.  Output is staged in shared memory
// and written as 128-byte runs along the (contiguous) frame axis of the reference layout
// [B,1,257,T].  HBM-bound: 4 B/sample read, 8 B/bin written.
#include "common.cuh"

namespace fdbm {
namespace {

constexpr int NFFT = 512;
constexpr int NBIN = NFFT / 2 + 1;
constexpr int FR = 16;            // frames (STFT) / hop segments (iSTFT) per block
constexpr int WARPS = 8;
constexpr int MAXR = 4;           // n_fft / hop <= 4
constexpr int NF_MAX = FR + MAXR; // frames an iSTFT block touches (even)

__device__ __forceinline__ int brev9(int k) { return __brev(static_cast<unsigned>(k)) >> 23; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// forward transform W = exp(-2 pi i k / 512): natural-order input, bit-reversed output (DIF).
__device__ __forceinline__ void fft512_dif(float2* z, const float2* tw, int lane) {
#pragma unroll 1
  for (int half = 256; half >= 1; half >>= 1) {
    const int tstride = 256 / half;
    for (int j = lane; j < 256; j += 32) {
      const int k = j & (half - 1);
      const int i0 = ((j - k) << 1) + k;
      const int i1 = i0 + half;
      const float2 a = z[i0], b = z[i1];
      z[i0] = make_float2(a.x + b.x, a.y + b.y);
      z[i1] = cmul(make_float2(a.x - b.x, a.y - b.y), tw[k * tstride]);
    }
    __syncwarp();
  }
}

// inverse transform (unnormalised, conj twiddles): bit-reversed input, natural-order output (DIT).
__device__ __forceinline__ void ifft512_dit(float2* z, const float2* tw, int lane) {
#pragma unroll 1
  for (int half = 1; half <= 256; half <<= 1) {
    const int tstride = 256 / half;
    for (int j = lane; j < 256; j += 32) {
      const int k = j & (half - 1);
      const int i0 = ((j - k) << 1) + k;
      const int i1 = i0 + half;
      float2 w = tw[k * tstride];
      w.y = -w.y;
      const float2 a = z[i0], t = cmul(z[i1], w);
      z[i0] = make_float2(a.x + t.x, a.y + t.y);
      z[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ float2 compress(float2 z, int transform, float factor, float expo) {
  if (transform == FDBM_TRANSFORM_NONE) return z;
  const float mag = sqrtf(z.x * z.x + z.y * z.y);
  float s;
  if (transform == FDBM_TRANSFORM_EXPONENT) {
    if (expo == 1.0f) s = factor;
    else if (mag > 0.f) s = (expo == 0.5f ? rsqrtf(mag) : powf(mag, expo) / mag) * factor;
    else s = 0.f;
  } else {  // log
    s = mag > 0.f ? log1pf(mag) / mag * factor : 0.f;
  }
  return make_float2(z.x * s, z.y * s);
}

__device__ __forceinline__ float2 decompress(float2 z, int transform, float factor, float expo) {
  if (transform == FDBM_TRANSFORM_NONE) return z;
  z.x = z.x / factor;
  z.y = z.y / factor;
  const float mag = sqrtf(z.x * z.x + z.y * z.y);
  float s;
  if (transform == FDBM_TRANSFORM_EXPONENT) {
    if (expo == 1.0f) s = 1.f;
    else if (mag > 0.f) s = (expo == 0.5f ? mag : powf(mag, 1.0f / expo) / mag);
    else s = 0.f;
  } else {
    s = mag > 0.f ? expm1f(mag) / mag : 0.f;
  }
  return make_float2(z.x * s, z.y * s);
}

// frame index a padded output frame copies from (pad_spec, other.py:76-90); -1 = zeros
__device__ __forceinline__ int pad_source(int t, int M, int pad_mode) {
  if (t < M) return t;
  if (pad_mode == FDBM_PAD_REFLECTION) return 2 * (M - 1) - t;
  if (pad_mode == FDBM_PAD_REPLICATION) return M - 1;
  return -1;
}

struct StftSmem {
  float2 tw[256];
  float2 work[WARPS][NFFT];
  float2 tile[NBIN][FR];
};

__global__ void __launch_bounds__(WARPS * 32)
stft_compress_kernel(const float* __restrict__ wave, int64_t n_samples, int64_t wave_stride,
                     const float* __restrict__ window, int hop, int transform, float factor, float expo,
                     int pad_mode, int M, int n_frames_out, float2* __restrict__ spec) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const float* x = wave + static_cast<int64_t>(b) * wave_stride;

  if (threadIdx.x < 256) {
    float s, c;
    sincospif(static_cast<float>(threadIdx.x) / 256.0f, &s, &c);
    sm.tw[threadIdx.x] = make_float2(c, -s);
  }
  __syncthreads();

  float2* z = sm.work[warp];
  const int ta = t0 + 2 * warp, tb = ta + 1;
  const int sa = ta < n_frames_out ? pad_source(ta, M, pad_mode) : -1;
  const int sb = tb < n_frames_out ? pad_source(tb, M, pad_mode) : -1;
  if (sa >= 0 || sb >= 0) {
    for (int n = lane; n < NFFT; n += 32) {
      const float w = window[n];
      float va = 0.f, vb = 0.f;
      if (sa >= 0) {
        int64_t p = static_cast<int64_t>(sa) * hop + n - NFFT / 2;       // centred framing, reflect pad
        if (p < 0) p = -p;
        if (p >= n_samples) p = 2 * (n_samples - 1) - p;
        va = x[p] * w;
      }
      if (sb >= 0) {
        int64_t p = static_cast<int64_t>(sb) * hop + n - NFFT / 2;
        if (p < 0) p = -p;
        if (p >= n_samples) p = 2 * (n_samples - 1) - p;
        vb = x[p] * w;
      }
      z[n] = make_float2(va, vb);
    }
    __syncwarp();
    fft512_dif(z, sm.tw, lane);
  }
  for (int k = lane; k < NBIN; k += 32) {
    float2 A = make_float2(0.f, 0.f), Bv = A;
    if (sa >= 0 || sb >= 0) {
      const float2 p = z[brev9(k)];
      const float2 q = z[brev9((NFFT - k) & (NFFT - 1))];
      A = make_float2(0.5f * (p.x + q.x), 0.5f * (p.y - q.y));
      Bv = make_float2(0.5f * (p.y + q.y), -0.5f * (p.x - q.x));
      A = sa >= 0 ? compress(A, transform, factor, expo) : make_float2(0.f, 0.f);     // zero-padded frames stay exact zeros
      Bv = sb >= 0 ? compress(Bv, transform, factor, expo) : make_float2(0.f, 0.f);
    }
    sm.tile[k][2 * warp] = A;
    sm.tile[k][2 * warp + 1] = Bv;
  }
  __syncthreads();
  float2* out = spec + static_cast<int64_t>(b) * NBIN * n_frames_out;
  for (int i = threadIdx.x; i < NBIN * FR; i += WARPS * 32) {
    const int k = i / FR, fl = i % FR;
    if (t0 + fl < n_frames_out) out[static_cast<int64_t>(k) * n_frames_out + t0 + fl] = sm.tile[k][fl];
  }
}

struct IstftSmem {
  float2 tw[256];
  float2 work[WARPS][NFFT];
  float2 spec[NBIN][NF_MAX];
  float frames[NF_MAX][NFFT];
};

__global__ void __launch_bounds__(WARPS * 32)
decompress_istft_kernel(const float2* __restrict__ spec, int M, const float* __restrict__ window, int hop,
                        int transform, float factor, float expo, int64_t length, int64_t wave_stride,
                        float* __restrict__ wave) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  IstftSmem& sm = *reinterpret_cast<IstftSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int R = NFFT / hop;
  const int s0 = blockIdx.x * FR;              // first hop segment (padded timeline) of this block
  const int m0 = s0 - R + 1;                   // first frame touched
  const int NF = (FR + R - 1 + 1) & ~1;        // frames touched, rounded up to even

  if (threadIdx.x < 256) {
    float s, c;
    sincospif(static_cast<float>(threadIdx.x) / 256.0f, &s, &c);
    sm.tw[threadIdx.x] = make_float2(c, -s);
  }
  const float2* in = spec + static_cast<int64_t>(b) * NBIN * M;
  for (int i = threadIdx.x; i < NBIN * NF; i += WARPS * 32) {
    const int k = i / NF, fi = i % NF;
    const int m = m0 + fi;
    float2 v = make_float2(0.f, 0.f);
    if (m >= 0 && m < M) v = decompress(in[static_cast<int64_t>(k) * M + m], transform, factor, expo);
    sm.spec[k][fi] = v;
  }
  __syncthreads();

  float2* z = sm.work[warp];
  for (int pair = warp; pair < NF / 2; pair += WARPS) {
    const int fa = 2 * pair, fb = fa + 1;
    // Hermitian extension of both spectra, Z = A + iB, scattered to bit-reversed positions.
    // irfft semantics: the imaginary parts of the DC and Nyquist bins are ignored.
    for (int k = lane; k < NFFT; k += 32) {
      float2 A, Bv;
      if (k <= NFFT / 2) {
        A = sm.spec[k][fa];
        Bv = sm.spec[k][fb];
        if (k == 0 || k == NFFT / 2) { A.y = 0.f; Bv.y = 0.f; }
      } else {
        A = sm.spec[NFFT - k][fa];
        Bv = sm.spec[NFFT - k][fb];
        A.y = -A.y;
        Bv.y = -Bv.y;
      }
      z[brev9(k)] = make_float2(A.x - Bv.y, A.y + Bv.x);
    }
    __syncwarp();
    ifft512_dit(z, sm.tw, lane);
    for (int n = lane; n < NFFT; n += 32) {
      const float w = window[n] * (1.0f / NFFT);
      sm.frames[fa][n] = z[n].x * w;
      sm.frames[fb][n] = z[n].y * w;
    }
    __syncwarp();
  }
  __syncthreads();

  float* out = wave + static_cast<int64_t>(b) * wave_stride;
  for (int i = threadIdx.x; i < FR * hop; i += WARPS * 32) {
    const int seg = i / hop, j = i % hop;
    const int64_t n = static_cast<int64_t>(s0 + seg) * hop + j - NFFT / 2;
    if (n < 0 || n >= length) continue;
    float acc = 0.f, env = 0.f;
    for (int r = R - 1; r >= 0; --r) {           // oldest frame first (fold order)
      const int m = s0 + seg - r;
      if (m < 0 || m >= M) continue;
      const float w = window[j + r * hop];
      acc += sm.frames[m - m0][j + r * hop];
      env += w * w;
    }
    out[n] = env > 1e-11f ? acc / env : 0.f;
  }
}

// stand-alone spec_fwd / spec_back (data_module.py:173-199) and pad_spec (other.py:76-90) for callers
// that use the reference's unfused API
__global__ void __launch_bounds__(256)
spec_transform_kernel(const float2* __restrict__ in, float2* __restrict__ out, int64_t n, int transform, float factor,
                      float expo, int inverse) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x)
    out[i] = inverse ? decompress(in[i], transform, factor, expo) : compress(in[i], transform, factor, expo);
}

__global__ void __launch_bounds__(256)
pad_spec_kernel(const float2* __restrict__ in, float2* __restrict__ out, int64_t rows, int T, int T_out, int pad_mode) {
  const int64_t total = rows * T_out;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int t = static_cast<int>(i % T_out);
    const int64_t r = i / T_out;
    const int src = pad_source(t, T, pad_mode);
    out[i] = src >= 0 ? in[r * T + src] : make_float2(0.f, 0.f);
  }
}

}  // namespace

}  // namespace fdbm

using namespace fdbm;

extern "C" int fdbm_spec_transform(const float* in, float* out, int64_t n_complex, int transform_type, float spec_factor,
                                   float abs_exponent, int inverse, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(in && out && n_complex > 0, "fdbm_spec_transform: bad arguments");
  FDBM_REQUIRE(transform_type >= 0 && transform_type <= 2, "fdbm_spec_transform: bad transform");
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(n_complex, 256), static_cast<int64_t>(num_sms()) * 16));
  spec_transform_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out),
                                                              n_complex, transform_type, spec_factor, abs_exponent, inverse);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_pad_spec(const float* in, int64_t rows, int n_frames, int pad_mode, int n_frames_out, float* out,
                             void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(in && out && rows > 0 && n_frames > 0 && n_frames_out >= n_frames, "fdbm_pad_spec: bad arguments");
  FDBM_REQUIRE(pad_mode >= 0 && pad_mode <= 2, "fdbm_pad_spec: bad pad mode");
  FDBM_REQUIRE(pad_mode != FDBM_PAD_REFLECTION || n_frames_out - n_frames < n_frames, "fdbm_pad_spec: reflection pad wider than the input");
  const int64_t total = rows * n_frames_out;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(total, 256), static_cast<int64_t>(num_sms()) * 16));
  pad_spec_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out),
                                                        rows, n_frames, n_frames_out, pad_mode);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_stft_compress(const float* wave, int batch, int64_t n_samples, int64_t wave_stride,
                                  const float* window, int n_fft, int hop, int transform_type, float spec_factor,
                                  float abs_exponent, int pad_mode, int n_frames_out, float* spec, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(n_fft == NFFT, "fdbm_stft_compress: n_fft must be 512 (got %d)", n_fft);
  FDBM_REQUIRE(hop > 0 && NFFT % hop == 0 && NFFT / hop <= MAXR && NFFT / hop >= 1, "fdbm_stft_compress: hop %d unsupported", hop);
  FDBM_REQUIRE(batch > 0 && n_samples > NFFT / 2, "fdbm_stft_compress: need batch > 0 and n_samples > n_fft/2");
  FDBM_REQUIRE(transform_type >= 0 && transform_type <= 2 && pad_mode >= 0 && pad_mode <= 2, "fdbm_stft_compress: bad enum");
  const int M = 1 + static_cast<int>(n_samples / hop);
  FDBM_REQUIRE(n_frames_out >= M, "fdbm_stft_compress: n_frames_out %d < frame count %d", n_frames_out, M);
  FDBM_REQUIRE(pad_mode != FDBM_PAD_REFLECTION || n_frames_out - M < M, "fdbm_stft_compress: reflection pad wider than the input");
  FDBM_REQUIRE(wave && window && spec && wave_stride >= n_samples, "fdbm_stft_compress: null pointer or bad stride");
  static bool attr_set = false;
  if (!attr_set) {
    FDBM_CUDA(cudaFuncSetAttribute(stft_compress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftSmem)));
    attr_set = true;
  }
  dim3 grid(ceil_div(n_frames_out, FR), batch);
  stft_compress_kernel<<<grid, WARPS * 32, sizeof(StftSmem), as_stream(stream)>>>(
      wave, n_samples, wave_stride, window, hop, transform_type, spec_factor, abs_exponent, pad_mode, M, n_frames_out,
      reinterpret_cast<float2*>(spec));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

extern "C" int fdbm_decompress_istft(const float* spec, int batch, int n_frames, const float* window, int n_fft, int hop,
                                     int transform_type, float spec_factor, float abs_exponent, int64_t length,
                                     int64_t wave_stride, float* wave, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(n_fft == NFFT, "fdbm_decompress_istft: n_fft must be 512 (got %d)", n_fft);
  FDBM_REQUIRE(hop > 0 && NFFT % hop == 0 && NFFT / hop <= MAXR, "fdbm_decompress_istft: hop %d unsupported", hop);
  FDBM_REQUIRE(batch > 0 && n_frames > 0 && length > 0 && wave_stride >= length, "fdbm_decompress_istft: bad sizes");
  FDBM_REQUIRE(transform_type >= 0 && transform_type <= 2, "fdbm_decompress_istft: bad transform");
  FDBM_REQUIRE(spec && window && wave, "fdbm_decompress_istft: null pointer");
  static bool attr_set = false;
  if (!attr_set) {
    FDBM_CUDA(cudaFuncSetAttribute(decompress_istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IstftSmem)));
    attr_set = true;
  }
  const int64_t n_seg = ceil_div64(length + NFFT / 2, hop);      // hop segments covering [0, length + n_fft/2)
  dim3 grid(static_cast<unsigned>(ceil_div64(n_seg, FR)), batch);
  decompress_istft_kernel<<<grid, WARPS * 32, sizeof(IstftSmem), as_stream(stream)>>>(
      reinterpret_cast<const float2*>(spec), n_frames, window, hop, transform_type, spec_factor, abs_exponent, length,
      wave_stride, wave);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}
