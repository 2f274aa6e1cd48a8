// Shared host/device helpers for libfdbm_b200: error reporting, launch checks, small math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <algorithm>
#include "../../include/fdbm_b200.h"

namespace fdbm {

// thread-local message behind fdbm_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define FDBM_CUDA(expr)                                                          \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return ::fdbm::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define FDBM_REQUIRE(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      ::fdbm::set_error(__VA_ARGS__);          \
      return FDBM_EINVAL;                      \
    }                                          \
  } while (0)

// Every launch is followed by this: catches bad launch configurations immediately.
#define FDBM_LAUNCH_CHECK() FDBM_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int require_sm100();   // FDBM_OK or FDBM_EARCH (cached per device)
int num_sms();

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// --------------------------------------------------------------------------------------------
// Internal launchers (device pointers, no argument validation beyond what kernels need).
// The C ABI in api.cu validates and forwards; the backbone plan calls these directly.
// --------------------------------------------------------------------------------------------
int launch_fir_resample(const float* in, int B, int T, int F, int C, int mode, float* out, cudaStream_t s);
int launch_channel_stats(const float* in, int B, int T, int F, int C, double* sums, cudaStream_t s);
int launch_groupnorm_act(const float* src1, const double* sums1, int C1, const float* src2, const double* sums2,
                         int C2, const float* gamma, const float* beta, int B, int T, int F, int silu, int mode,
                         __nv_bfloat16* act_out, __nv_bfloat16* raw_out, cudaStream_t s);
int launch_pack_input(const float* x, const float* y, int B, int T, int F_in, int F, int Cin, float* out,
                      cudaStream_t s);
int launch_conv_in(const float* in, int Cin, const float* w, const float* bias, int B, int T, int F, int Cout,
                   float* out, cudaStream_t s);
int launch_combine(float* h, const float* pyr, int Cp, const float* w, const float* bias, int B, int T, int F,
                   int C, cudaStream_t s);
int launch_pyramid_conv(const __nv_bfloat16* act, int C, const float* w, const float* bias, const float* prev,
                        int Cp, int B, int T, int F, float* out, cudaStream_t s);
int launch_output_layer(const float* pyr, int Cp, const float* w, const float* bias, int B, int T, int F,
                        int F_out, float* out, cudaStream_t s);
int launch_temb(const float* t, const float* fourier_w, int nf, const float* w1, const float* b1, const float* w2,
                const float* b2, int B, int t_stride, float* temb_act, cudaStream_t s);
int launch_dense_all(const float* temb_act, const float* w, const float* bias, int B, int K, int rows, float* out,
                     cudaStream_t s);
int launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, int ld, int B, int L,
                     int C, __nv_bfloat16* o, int ldo, cudaStream_t s);

struct ConvArgs {
  const __nv_bfloat16* in1; int C1; int ksize;
  const __nv_bfloat16* in2; int C2;
  const __nv_bfloat16* wpack;
  const float* bias; const float* bias_b; int bias_b_stride; const float* residual;
  float scale; int B, T, F, Cout;
  float* out_f32; __nv_bfloat16* out_bf16; int out_ld;      // out_ld: row stride (elements) of outputs, >= Cout
  double* sums;
};
int launch_conv_igemm(const ConvArgs& a, cudaStream_t s);
int64_t conv_wpack_bytes(int C1, int ksize, int C2, int Cout);
int launch_pack_conv_weights(const float* w1, int C1, int ksize, const float* w2, int C2, int Cout, int n_rows_total,
                             int row_offset, __nv_bfloat16* wpack, cudaStream_t s);

}  // namespace fdbm
