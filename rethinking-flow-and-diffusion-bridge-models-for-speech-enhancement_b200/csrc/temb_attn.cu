// Time embedding (layerspp.py:32-41, ncsnpp_v2.py:108-113,252-270), the 49 per-block Dense_0
// projections (layerspp.py:263) as one batched GEMV, and the attention core of AttnBlockpp
// (layerspp.py:82-86).  All tiny next to the convolutions (< 0.1 % of the FLOPs at 4 s).
#include "common.cuh"

namespace fdbm {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// One block per batch item.  emb = [sin, cos](log(t) * W * 2 * pi)  -> Linear -> SiLU -> Linear -> SiLU.
// The result is SiLU(temb): every consumer applies the activation first (layerspp.py:263).
__global__ void __launch_bounds__(512)
temb_kernel(const float* __restrict__ t, const float* __restrict__ fw, int nf, const float* __restrict__ w1,
            const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
            int t_stride, float* __restrict__ out) {
  extern __shared__ float sm[];          // emb[2nf] | h1[4nf]
  float* emb = sm;
  float* h1 = sm + 2 * nf;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // fp32 op order of the reference: ((log(t) * W) * 2) * pi, pi rounded to fp32
  const float lt = static_cast<float>(log(static_cast<double>(t[b * t_stride])));
  for (int j = threadIdx.x; j < nf; j += blockDim.x) {
    const float proj = __fmul_rn(__fmul_rn(__fmul_rn(lt, fw[j]), 2.0f), 3.14159274101257324f);
    emb[j] = sinf(proj);
    emb[nf + j] = cosf(proj);
  }
  __syncthreads();
  const int D = 4 * nf;
  for (int r = warp; r < D; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < 2 * nf; k += 32) acc = fmaf(w1[r * 2 * nf + k], emb[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) h1[r] = silu_f(acc + b1[r]);
  }
  __syncthreads();
  for (int r = warp; r < D; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(w2[r * D + k], h1[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[static_cast<int64_t>(b) * D + r] = silu_f(acc + b2[r]);
  }
}

// out[b][r] = sum_k act[b][k] * w[r][k] + bias[r];  warp per row, up to 8 batch items per pass.
__global__ void __launch_bounds__(256)
dense_all_kernel(const float* __restrict__ act, const float* __restrict__ w, const float* __restrict__ bias, int B,
                 int K, int rows, float* __restrict__ out) {
  extern __shared__ float sa[];          // [8][K]
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = min(8, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < nb * K; i += 256) sa[i] = act[static_cast<int64_t>(b0) * K + i];
    __syncthreads();
    for (int r = blockIdx.x * 8 + warp; r < rows; r += 8 * gridDim.x) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float wv = w[static_cast<int64_t>(r) * K + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nb) acc[j] = fmaf(wv, sa[j * K + k], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nb) {
          const float v = warp_sum(acc[j]);
          if (lane == 0) out[static_cast<int64_t>(b0 + j) * rows + r] = v + bias[r];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Attention: one warp per query.  Pass 1 scores all keys (lane = key), pass 2 soft-max, pass 3
// o = P V (lane = 8 channels).  K/V stay L1/L2 resident (L*C*2 bytes = 128 KB at L=256, C=256).
// ------------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 8;

__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_kernel(const op_t* __restrict__ q, const op_t* __restrict__ k,
                 const op_t* __restrict__ v, int ld, int L, int C, float scale,
                 op_t* __restrict__ o, int ldo) {
  extern __shared__ float sm[];                       // per warp: q[C] | p[L]
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int b = blockIdx.y;
  const int qi = blockIdx.x * ATT_WARPS + warp;
  if (qi >= L) return;
  float* sq = sm + warp * (C + L);
  float* sp = sq + C;
  const int64_t base = static_cast<int64_t>(b) * L;
  for (int c = lane; c < C; c += 32) sq[c] = op2f(q[(base + qi) * ld + c]) * scale;
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) {
    const uint4* kr = reinterpret_cast<const uint4*>(k + (base + j) * ld);
    float acc = 0.f;
    for (int c8 = 0; c8 < C / 8; ++c8) {
      const uint4 raw = kr[c8];
      const op2_t* h2 = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f2 = op22f2(h2[u]);
        acc = fmaf(f2.x, sq[c8 * 8 + 2 * u], acc);
        acc = fmaf(f2.y, sq[c8 * 8 + 2 * u + 1], acc);
      }
    }
    sp[j] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float e = __expf(sp[j] - mx);
    sp[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __syncwarp();
  // lane owns channels [lane*8, lane*8+8) (+256 per repeat)
  for (int c0 = lane * 8; c0 < C; c0 += 256) {
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    for (int j = 0; j < L; ++j) {
      const float pj = sp[j];
      const uint4 raw = *reinterpret_cast<const uint4*>(v + (base + j) * ld + c0);
      const op2_t* h2 = reinterpret_cast<const op2_t*>(&raw);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f2 = op22f2(h2[u]);
        acc[2 * u] = fmaf(pj, f2.x, acc[2 * u]);
        acc[2 * u + 1] = fmaf(pj, f2.y, acc[2 * u + 1]);
      }
    }
    op2_t ov[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ov[u] = f2op2(acc[2 * u] * inv, acc[2 * u + 1] * inv);
    *reinterpret_cast<uint4*>(o + (base + qi) * ldo + c0) = *reinterpret_cast<uint4*>(ov);
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core attention (C = 256): flash-style, one block = 64 queries (4 warps x 16), keys/values in tiles
// of 64 staged in shared memory, S = Q K^T and O += P V on mma.sync m16n8k16 (16-bit operands, fp32
// accumulators), online soft-max in fp32 (exp2 with the 1/sqrt(C) scale folded in).  ldmatrix feeds the
// fragments; rows are padded by 16 bytes so that the 8 rows of an 8x8 matrix fall in different banks.
// tcgen05 is not used here on purpose: 0.04 % of the FLOPs, tiles of 16 queries, no TMEM round trip.
// ------------------------------------------------------------------------------------------------
constexpr int AT_BQ = 64, AT_BK = 64, AT_C = 256, AT_PITCH = AT_C + 8;     // elements

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef FDBM_OPERAND_BF16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#endif
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows [r0, r0 + 64) of a [*, ld] 16-bit matrix -> shared tile [64][AT_PITCH]; rows >= L are zero
__device__ __forceinline__ void load_tile64(const op_t* __restrict__ src, int64_t base_row, int r0, int L, int ld, op_t* dst) {
  for (int i = threadIdx.x; i < 64 * (AT_C / 8); i += 128) {
    const int r = i / (AT_C / 8), c8 = i % (AT_C / 8);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < L) val = __ldg(reinterpret_cast<const uint4*>(src + (base_row + r0 + r) * ld) + c8);
    *reinterpret_cast<uint4*>(dst + r * AT_PITCH + c8 * 8) = val;
  }
}

__global__ void __launch_bounds__(128)
attention_mma_kernel(const op_t* __restrict__ q, const op_t* __restrict__ k, const op_t* __restrict__ v, int ld, int L,
                     float scale_log2e, op_t* __restrict__ o, int ldo) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  op_t* sQ = reinterpret_cast<op_t*>(att_smem);
  op_t* sK = sQ + AT_BQ * AT_PITCH;
  op_t* sV = sK + AT_BK * AT_PITCH;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, q0 = blockIdx.x * AT_BQ;
  const int64_t base = static_cast<int64_t>(b) * L;

  load_tile64(q, base, q0, L, ld, sQ);

  float oacc[AT_C / 8][4];
#pragma unroll
  for (int i = 0; i < AT_C / 8; ++i) { oacc[i][0] = 0.f; oacc[i][1] = 0.f; oacc[i][2] = 0.f; oacc[i][3] = 0.f; }
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};      // rows g and g + 8 of this warp's 16 queries

  // ldmatrix lane addresses (bytes, shared space)
  const uint32_t sQ_a = static_cast<uint32_t>(__cvta_generic_to_shared(sQ)) +
                        ((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * AT_PITCH + 8 * (lane >> 4)) * 2;
  const uint32_t sK_a = static_cast<uint32_t>(__cvta_generic_to_shared(sK)) +
                        (((lane >> 4) * 8 + (lane & 7)) * AT_PITCH + 8 * ((lane >> 3) & 1)) * 2;
  const uint32_t sV_a = static_cast<uint32_t>(__cvta_generic_to_shared(sV)) +
                        (((lane & 7) + 8 * ((lane >> 3) & 1)) * AT_PITCH + 8 * (lane >> 4)) * 2;

  const int n_kt = (L + AT_BK - 1) / AT_BK;
  for (int kt = 0; kt < n_kt; ++kt) {
    __syncthreads();                                   // previous tile fully consumed (and Q stored, first pass)
    load_tile64(k, base, kt * AT_BK, L, ld, sK);
    load_tile64(v, base, kt * AT_BK, L, ld, sV);
    __syncthreads();

    // S = Q K^T : 16 queries x 64 keys per warp
    float sacc[AT_BK / 8][4];
#pragma unroll
    for (int i = 0; i < AT_BK / 8; ++i) { sacc[i][0] = 0.f; sacc[i][1] = 0.f; sacc[i][2] = 0.f; sacc[i][3] = 0.f; }
#pragma unroll 4
    for (int kk = 0; kk < AT_C / 16; ++kk) {
      uint32_t a[4];
      ldsm_x4(sQ_a + kk * 32, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int p = 0; p < AT_BK / 16; ++p) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4(sK_a + (p * 16 * AT_PITCH) * 2 + kk * 32, r0, r1, r2, r3);
        mma16816(sacc[2 * p], a, r0, r1);
        mma16816(sacc[2 * p + 1], a, r2, r3);
      }
    }
    // mask keys beyond L, online soft-max
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < AT_BK / 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kt * AT_BK + i * 8 + 2 * t + (e & 1);
        if (key >= L) sacc[i][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sacc[i][e]);
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(mrow[r], mx[r]);
      alpha[r] = exp2f((mrow[r] - mnew) * scale_log2e);
      mrow[r] = mnew;
      lrow[r] *= alpha[r];
    }
    uint32_t pa[AT_BK / 16][4];
#pragma unroll
    for (int i = 0; i < AT_BK / 8; ++i) {
      const float p0 = exp2f((sacc[i][0] - mrow[0]) * scale_log2e), p1 = exp2f((sacc[i][1] - mrow[0]) * scale_log2e);
      const float p2 = exp2f((sacc[i][2] - mrow[1]) * scale_log2e), p3 = exp2f((sacc[i][3] - mrow[1]) * scale_log2e);
      lrow[0] += p0 + p1;
      lrow[1] += p2 + p3;
      pa[i >> 1][(i & 1) * 2] = pack_op2(p0, p1);
      pa[i >> 1][(i & 1) * 2 + 1] = pack_op2(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < AT_C / 8; ++i) {
      oacc[i][0] *= alpha[0]; oacc[i][1] *= alpha[0]; oacc[i][2] *= alpha[1]; oacc[i][3] *= alpha[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < AT_BK / 16; ++kk) {
#pragma unroll
      for (int cp = 0; cp < AT_C / 16; ++cp) {
        uint32_t r0, r1, r2, r3;
        ldsm_x4_t(sV_a + (kk * 16 * AT_PITCH + cp * 16) * 2, r0, r1, r2, r3);
        mma16816(oacc[2 * cp], pa[kk], r0, r1);
        mma16816(oacc[2 * cp + 1], pa[kk], r2, r3);
      }
    }
  }
  // normalise, stage the 64 x 256 output tile in sQ (this warp only touches its own 16 rows), coalesced stores
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
    lrow[r] = 1.0f / lrow[r];
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < AT_C / 8; ++i) {
    op_t* d0 = sQ + (warp * 16 + g) * AT_PITCH + i * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(d0) = pack_op2(oacc[i][0] * lrow[0], oacc[i][1] * lrow[0]);
    *reinterpret_cast<uint32_t*>(d0 + 8 * AT_PITCH) = pack_op2(oacc[i][2] * lrow[1], oacc[i][3] * lrow[1]);
  }
  __syncwarp();
  for (int i = lane; i < 16 * (AT_C / 8); i += 32) {
    const int r = i / (AT_C / 8), c8 = i % (AT_C / 8);
    const int qi = q0 + warp * 16 + r;
    if (qi < L)
      *(reinterpret_cast<uint4*>(o + (base + qi) * ldo) + c8) = *reinterpret_cast<const uint4*>(sQ + (warp * 16 + r) * AT_PITCH + c8 * 8);
  }
}

}  // namespace

int launch_temb(const float* t, const float* fourier_w, int nf, const float* w1, const float* b1, const float* w2,
                const float* b2, int B, int t_stride, float* temb_act, cudaStream_t s) {
  temb_kernel<<<B, 512, sizeof(float) * 6 * nf, s>>>(t, fourier_w, nf, w1, b1, w2, b2, t_stride, temb_act);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_dense_all(const float* temb_act, const float* w, const float* bias, int B, int K, int rows, float* out,
                     cudaStream_t s) {
  const int grid = std::min(ceil_div(rows, 8), num_sms() * 8);
  dense_all_kernel<<<grid, 256, sizeof(float) * 8 * K, s>>>(temb_act, w, bias, B, K, rows, out);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_attention(const op_t* q, const op_t* k, const op_t* v, int ld, int B, int L,
                     int C, op_t* o, int ldo, cudaStream_t s, int fp32_probs) {
  FDBM_REQUIRE(C % 8 == 0 && ld % 8 == 0 && ldo % 8 == 0, "attention: channels / strides must be multiples of 8");
  if (C == AT_C && !fp32_probs) {                      // the backbone's case: tensor-core kernel
    constexpr int kSmem = (AT_BQ + 2 * AT_BK) * AT_PITCH * 2;
    static PerDeviceOnce attr_once;
    if (attr_once.first(current_device()))
      FDBM_CUDA(cudaFuncSetAttribute(attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    dim3 grid(ceil_div(L, AT_BQ), B);
    attention_mma_kernel<<<grid, 128, kSmem, s>>>(q, k, v, ld, L, 1.4426950408889634f / sqrtf(static_cast<float>(C)), o, ldo);
    FDBM_LAUNCH_CHECK();
    return FDBM_OK;
  }
  const size_t smem = sizeof(float) * ATT_WARPS * (C + L);
  FDBM_REQUIRE(smem <= 200 * 1024, "attention: sequence length %d too long for the shared-memory score buffer", L);
  static PerDeviceOnce attr_big;
  if (smem > 48 * 1024 && attr_big.first(current_device()))
    FDBM_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid(ceil_div(L, ATT_WARPS), B);
  attention_kernel<<<grid, ATT_WARPS * 32, smem, s>>>(q, k, v, ld, L, C, 1.0f / sqrtf(static_cast<float>(C)), o, ldo);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

}  // namespace fdbm

using namespace fdbm;

// time embedding (layerspp.py:32-41, ncsnpp_v2.py:252-270): out[b] = SiLU(W2 SiLU(W1 [sin, cos](2 pi W log t_b) + b1) + b2), [B, 4 nf]
// (every consumer is `Dense_0(act(temb))`, so the activated vector is what is kept)
extern "C" int fdbm_time_embedding(const float* t, const float* fourier_w, int nf, const float* w1, const float* b1, const float* w2,
                                   const float* b2, int batch, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(t && fourier_w && w1 && b1 && w2 && b2 && out && batch > 0 && nf > 0 && nf <= 1024, "fdbm_time_embedding: bad arguments");
  return launch_temb(t, fourier_w, nf, w1, b1, w2, b2, batch, 1, out, as_stream(stream));
}
// the FiLM rows of all residual blocks at once (layerspp.py:263 `Dense_0(act(temb))`, 49 layers): out[b][r] = w[r] . act[b] + bias[r]
extern "C" int fdbm_film_rows(const float* temb_act, const float* weight, const float* bias, int batch, int k, int rows, float* out, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(temb_act && weight && bias && out && batch > 0 && k > 0 && k <= 1536 && rows > 0, "fdbm_film_rows: bad arguments");
  return launch_dense_all(temb_act, weight, bias, batch, k, rows, out, as_stream(stream));
}

extern "C" int fdbm_attention(const void* q, const void* k, const void* v, int batch, int L, int C, void* o,
                              void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(q && k && v && o && batch > 0 && L > 0, "fdbm_attention: bad arguments");
  return launch_attention(reinterpret_cast<const op_t*>(q), reinterpret_cast<const op_t*>(k),
                          reinterpret_cast<const op_t*>(v), C, batch, L, C,
                          reinterpret_cast<op_t*>(o), C, as_stream(stream), 0);
}
