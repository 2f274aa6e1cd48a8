// Implicit-GEMM convolution for the NCSN++ backbone on Blackwell tensor cores.
//
//   out[b,t,f,n] = scale * ( sum_{dt,df,c} W[n,c,df,dt] * in1[b,t+dt-1,f+df-1,c]        (3x3 or 1x1)
//                            (+ sum_c W2[n,c] * in2[b,t,f,c])                             (fused 1x1 shortcut)
//                            + bias[n] (+ bias_b[b,n]) (+ residual[b,t,f,n]) )
//
// GEMM view: M = pixels, N = Cout, K = taps*C1 (+ C2).  Replaces nn.Conv2d / NIN of
// fdbm/backbones/ncsnpp_utils/layers.py:100-124,546-555 and the adds of layerspp.py:263,270-274.
//
// Design (B200, sm_100a):
//   * activations are bf16 [B,T,F,C]; one M-tile = 16 frames x 8 bins = 128 pixels = one UMMA M.
//   * per 64-channel K-block ONE TMA box of 18 x 10 pixels (tile + halo, zero-filled outside the
//     image = the convolution's zero padding) lands in shared memory with the 128-byte swizzle; all
//     nine taps are fed from it by shifting the UMMA descriptor start address by (dt*10 + df) pixels
//     (row pitch 10 pixels -> stride-byte-offset 1280).  Validated on hardware by tools/umma_probe.cu.
//     A traffic per K-block is 180 pixel rows instead of 9 x 128.
//   * a CTA owns 2 M-tiles x 128 output channels: every weight tile (128 x 64 bf16, TMA, 6-deep ring)
//     is used by two MMAs; accumulators live in TMEM (2 stages x 2 tiles x 128 columns = 512).
//   * warp-specialised, persistent: warp0 = A producer, warp1 = B producer, warp2 = MMA issuer,
//     warp3 = TMEM owner, warps4-7 = epilogue (tcgen05.ld -> bias / FiLM / residual / scale -> global),
//     overlapping the next tile's main loop through the second accumulator stage.
#include <cuda.h>
#include <mutex>
#include "common.cuh"
#include "tc05.cuh"

namespace fdbm {
namespace {

using namespace tc05;

constexpr int MT = 2;                       // M-tiles per CTA tile
constexpr int TILE_T = 16, TILE_F = 8;
constexpr int HALO_T = TILE_T + 2, HALO_F = TILE_F + 2;
constexpr int A_TILE_BYTES = HALO_T * HALO_F * 128;          // 23040 bytes landed per TMA box
constexpr int A_TILE_STRIDE = 23552;                         // rounded up to 1024
constexpr int A_STAGES = 2;
constexpr int B_STAGES = 6;
constexpr int BN = 128;
constexpr int B_TILE_BYTES = BN * 128;
constexpr int A_SBO = HALO_F * 128;                          // 1280: distance between 8-pixel row groups
constexpr int NUM_THREADS = 256;
constexpr int SMEM_BYTES = 1024 + A_STAGES * MT * A_TILE_STRIDE + B_STAGES * B_TILE_BYTES + 256;

struct ConvParams {
  int B, T, F, Cout;
  int kb1, taps1, kb2;
  int tiles_t, tiles_f, n_mtiles, n_nblocks, n_items;
  const float* bias;
  const float* bias_b;
  int bias_b_stride;
  const float* residual;
  float scale;
  float* out_f32;
  op_t* out_h16;
};

struct TileCoord { int b, t0, f0; bool valid; };

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int mi) {
  TileCoord c;
  c.valid = mi < p.n_mtiles;
  const int tt = mi % p.tiles_t;
  const int rest = mi / p.tiles_t;
  c.t0 = tt * TILE_T;
  c.f0 = (rest % p.tiles_f) * TILE_F;
  c.b = rest / p.tiles_f;
  return c;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                  const __grid_constant__ CUtensorMap map_b, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_STAGES * MT * A_TILE_STRIDE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + B_STAGES * B_TILE_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* acc_full = b_empty + B_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kb = p.kb1 + p.kb2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    if (p.kb2) tma_prefetch_desc(&map_a2);
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int ct = item / p.n_nblocks;
        TileCoord tc[MT];
        int n_valid = 0;
        for (int j = 0; j < MT; ++j) { tc[j] = decode_tile(p, ct * MT + j); n_valid += tc[j].valid; }
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(a_empty + stage, phase ^ 1);
          mbar_expect_tx(a_full + stage, n_valid * A_TILE_BYTES);
          const CUtensorMap* map = kb < p.kb1 ? &map_a1 : &map_a2;
          const int c0 = (kb < p.kb1 ? kb : kb - p.kb1) * 64;
          for (int j = 0; j < MT; ++j) {
            if (!tc[j].valid) continue;
            tma_load_4d(sA + (stage * MT + j) * A_TILE_STRIDE, map, a_full + stage, c0, tc[j].f0 - 1, tc[j].t0 - 1,
                        tc[j].b);
          }
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ B (weight) producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int n0 = (item % p.n_nblocks) * BN;
        int kt = 0;
        for (int kb = 0; kb < n_kb; ++kb) {
          const int ntaps = kb < p.kb1 ? p.taps1 : 1;
          for (int tap = 0; tap < ntaps; ++tap, ++kt) {
            mbar_wait(b_empty + stage, phase ^ 1);
            mbar_expect_tx(b_full + stage, B_TILE_BYTES);
            tma_load_2d(sB + stage * B_TILE_BYTES, &map_b, b_full + stage, 0, kt * p.Cout + n0);
            if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, BN, kOperandIsBf16);
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pacc = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int ct = item / p.n_nblocks;
        bool valid[MT];
        for (int j = 0; j < MT; ++j) valid[j] = (ct * MT + j) < p.n_mtiles;
        mbar_wait(acc_empty + as, pacc ^ 1);
        fence_after_sync();
        bool first = true;
        for (int kb = 0; kb < n_kb; ++kb) {
          const int ntaps = kb < p.kb1 ? p.taps1 : 1;
          mbar_wait(a_full + sa, pa);
          fence_after_sync();
          for (int tap = 0; tap < ntaps; ++tap) {
            mbar_wait(b_full + sb, pb);
            fence_after_sync();
            // tap -> (df, dt); a 1x1 conv reads the centre of the halo box
            const int df = ntaps == 9 ? tap / 3 : 1, dt = ntaps == 9 ? tap % 3 : 1;
            const uint32_t a_off = (dt * HALO_F + df) * 128;
            const uint32_t b_addr = smem_u32(sB + sb * B_TILE_BYTES);
#pragma unroll
            for (int j = 0; j < MT; ++j) {
              if (!valid[j]) continue;
              const uint32_t a_addr = smem_u32(sA + (sa * MT + j) * A_TILE_STRIDE) + a_off;
              const uint32_t d_tmem = tmem_base + (as * MT + j) * BN;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                mma_f16(d_tmem, make_desc_sw128(a_addr + k * 32, A_SBO), make_desc_sw128(b_addr + k * 32, 1024), idesc,
                         (first && k == 0) ? 0u : 1u);
              }
            }
            first = false;
            mma_commit(b_empty + sb);
            if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
          }
          mma_commit(a_empty + sa);
          if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
        }
        mma_commit(acc_full + as);
        if (++as == 2) { as = 0; pacc ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;                              // TMEM lane quadrant of this warp
    const int m = q * 32 + lane;                         // pixel row inside the M-tile
    const int r = m >> 3, c = m & 7;
    uint32_t as = 0, pacc = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int ct = item / p.n_nblocks;
      const int n0 = (item % p.n_nblocks) * BN;
      mbar_wait(acc_full + as, pacc);
      fence_after_sync();
      for (int j = 0; j < MT; ++j) {
        const TileCoord tc = decode_tile(p, ct * MT + j);
        if (!tc.valid) continue;                          // warp-uniform
        const int t = tc.t0 + r, f = tc.f0 + c;
        const bool ok = t < p.T && f < p.F;
        const int64_t pix = (static_cast<int64_t>(tc.b) * p.T + t) * p.F + f;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + (as * MT + j) * BN + ch * 32, v);
          tmem_ld_wait();
          if (ok) {
            const int nb = n0 + ch * 32;
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias + nb);
            const float4* bb4 = p.bias_b ? reinterpret_cast<const float4*>(p.bias_b + static_cast<int64_t>(tc.b) * p.bias_b_stride + nb) : nullptr;
            const float4* res4 = p.residual ? reinterpret_cast<const float4*>(p.residual + pix * p.Cout + nb) : nullptr;
            float4* of = p.out_f32 ? reinterpret_cast<float4*>(p.out_f32 + pix * p.Cout + nb) : nullptr;
            uint2* ob = p.out_h16 ? reinterpret_cast<uint2*>(p.out_h16 + pix * p.Cout + nb) : nullptr;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 o = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                     __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
              const float4 bv = __ldg(bias4 + g);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              if (bb4) { const float4 e = __ldg(bb4 + g); o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
              if (res4) { const float4 e = res4[g]; o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w; }
              o.x *= p.scale; o.y *= p.scale; o.z *= p.scale; o.w *= p.scale;
              if (of) of[g] = o;
              if (ob) {
                ob[g] = make_uint2(pack_op2(o.x, o.y), pack_op2(o.z, o.w));
              }
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + as);
      if (++as == 2) { as = 0; pacc ^= 1; }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, 512);
}

// weight packing: fp32 (OIHW, or [in][out] for NIN) -> bf16 [kt][rows_total][64]
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w1, int C1, int ksize, int io_layout, const float* __restrict__ w2,
                    int C2, int Cout, int rows_total, int row_offset, op_t* __restrict__ out) {
  const int taps = ksize * ksize;
  const int n_kt = (C1 / 64) * taps + C2 / 64;
  const int64_t total = static_cast<int64_t>(n_kt) * Cout * 64;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const int j = static_cast<int>(i % 64);
    const int co = static_cast<int>((i / 64) % Cout);
    const int kt = static_cast<int>(i / (64ll * Cout));
    float v;
    if (kt < (C1 / 64) * taps) {
      const int kb = kt / taps, tap = kt % taps;
      const int ci = kb * 64 + j;
      if (io_layout) v = w1[static_cast<int64_t>(ci) * Cout + co];
      else {
        const int kf = ksize == 3 ? tap / 3 : 0, ktm = ksize == 3 ? tap % 3 : 0;
        v = w1[((static_cast<int64_t>(co) * C1 + ci) * ksize + kf) * ksize + ktm];
      }
    } else {
      const int ci = (kt - (C1 / 64) * taps) * 64 + j;
      v = w2[static_cast<int64_t>(co) * C2 + ci];
    }
    out[(static_cast<int64_t>(kt) * rows_total + row_offset + co) * 64 + j] = f2op(v);
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

int make_act_map(CUtensorMap* map, const op_t* ptr, int B, int T, int F, int C) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return FDBM_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * F, (cuuint64_t)C * 2 * F * T};
  cuuint32_t box[4] = {64, HALO_F, HALO_T, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(map, (kOperandIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 4, const_cast<op_t*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation [%d,%d,%d,%d]) failed: %d", B, T, F, C, (int)r); return FDBM_ECUDA; }
  return FDBM_OK;
}

int make_weight_map(CUtensorMap* map, const op_t* ptr, int64_t rows) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return FDBM_ECUDA; }
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, BN};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(map, (kOperandIsBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16), 2, const_cast<op_t*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights, %lld rows) failed: %d", (long long)rows, (int)r); return FDBM_ECUDA; }
  return FDBM_OK;
}

}  // namespace

int64_t conv_wpack_bytes(int C1, int ksize, int C2, int Cout) {
  return (static_cast<int64_t>(C1 / 64) * ksize * ksize + C2 / 64) * Cout * 64 * 2;
}

int launch_pack_conv_weights(const float* w1, int C1, int ksize, const float* w2, int C2, int Cout, int n_rows_total,
                             int row_offset, op_t* wpack, cudaStream_t s) {
  FDBM_REQUIRE(C1 % 64 == 0 && C2 % 64 == 0 && (ksize == 1 || ksize == 3 || ksize == -1),
               "pack_conv_weights: channels must be multiples of 64, ksize 1 or 3");
  const int io = ksize == -1;                     // ksize -1: NIN weight, [in][out] layout, 1x1
  const int k = io ? 1 : ksize;
  const int64_t total = (static_cast<int64_t>(C1 / 64) * k * k + C2 / 64) * Cout * 64;
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div64(total, 256), 4096));
  pack_weights_kernel<<<grid, 256, 0, s>>>(w1, C1, k, io, w2, C2, Cout, n_rows_total, row_offset, wpack);
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_conv_igemm(const ConvArgs& a, cudaStream_t s) {
  FDBM_REQUIRE(a.C1 > 0 && a.C1 % 64 == 0 && a.C2 % 64 == 0, "conv_igemm: channels must be multiples of 64 (%d, %d)", a.C1, a.C2);
  FDBM_REQUIRE(a.ksize == 1 || a.ksize == 3, "conv_igemm: ksize must be 1 or 3");
  FDBM_REQUIRE(a.Cout % BN == 0, "conv_igemm: Cout must be a multiple of %d (got %d)", BN, a.Cout);
  FDBM_REQUIRE((a.C2 == 0) == (a.in2 == nullptr), "conv_igemm: in2 / C2 mismatch");
  FDBM_REQUIRE(a.out_f32 || a.out_h16, "conv_igemm: no output");
  static bool attr_set = false;
  if (!attr_set) {
    FDBM_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap map_a1, map_a2, map_b;
  if (int rc = make_act_map(&map_a1, a.in1, a.B, a.T, a.F, a.C1)) return rc;
  if (a.in2) { if (int rc = make_act_map(&map_a2, a.in2, a.B, a.T, a.F, a.C2)) return rc; }
  else map_a2 = map_a1;
  ConvParams p;
  p.B = a.B; p.T = a.T; p.F = a.F; p.Cout = a.Cout;
  p.kb1 = a.C1 / 64; p.taps1 = a.ksize * a.ksize; p.kb2 = a.C2 / 64;
  const int n_kt = p.kb1 * p.taps1 + p.kb2;
  if (int rc = make_weight_map(&map_b, a.wpack, static_cast<int64_t>(n_kt) * a.Cout)) return rc;
  p.tiles_t = ceil_div(a.T, TILE_T); p.tiles_f = ceil_div(a.F, TILE_F);
  p.n_mtiles = a.B * p.tiles_t * p.tiles_f;
  p.n_nblocks = a.Cout / BN;
  p.n_items = ceil_div(p.n_mtiles, MT) * p.n_nblocks;
  p.bias = a.bias; p.bias_b = a.bias_b; p.bias_b_stride = a.bias_b_stride; p.residual = a.residual; p.scale = a.scale;
  p.out_f32 = a.out_f32; p.out_h16 = a.out_h16;
  const int grid = std::min(p.n_items, num_sms());
  conv_igemm_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(map_a1, map_a2, map_b, p);
  FDBM_LAUNCH_CHECK();
  if (a.sums) {
    FDBM_REQUIRE(a.out_f32, "conv_igemm: channel sums need the fp32 output");
    return launch_channel_stats(a.out_f32, a.B, a.T, a.F, a.Cout, a.sums, s);
  }
  return FDBM_OK;
}

}  // namespace fdbm

using namespace fdbm;

extern "C" int fdbm_pack_conv_weights(const float* w1, int C1, int ksize, const float* w2, int C2, int Cout,
                                      void* wpack, int64_t* bytes, void* stream) {
  const int k = ksize == -1 ? 1 : ksize;
  if (bytes) *bytes = conv_wpack_bytes(C1, k, C2, Cout);
  if (!wpack) return FDBM_OK;
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(w1 && ((C2 == 0) == (w2 == nullptr)), "fdbm_pack_conv_weights: null pointer");
  return launch_pack_conv_weights(w1, C1, ksize, w2, C2, Cout, Cout, 0, reinterpret_cast<op_t*>(wpack),
                                  as_stream(stream));
}

extern "C" int fdbm_conv_igemm(const void* in1, int C1, int ksize, const void* in2, int C2, const void* wpack,
                               const float* bias, const float* bias_b, const float* residual, float scale, int batch,
                               int T, int F, int Cout, float* out_f32, void* out_h16, double* sums, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(in1 && wpack && bias && batch > 0 && T > 0 && F > 0, "fdbm_conv_igemm: bad arguments");
  ConvArgs a;
  a.in1 = reinterpret_cast<const op_t*>(in1); a.C1 = C1; a.ksize = ksize;
  a.in2 = reinterpret_cast<const op_t*>(in2); a.C2 = C2;
  a.wpack = reinterpret_cast<const op_t*>(wpack);
  a.bias = bias; a.bias_b = bias_b; a.bias_b_stride = Cout; a.residual = residual; a.scale = scale;
  a.B = batch; a.T = T; a.F = F; a.Cout = Cout;
  a.out_f32 = out_f32; a.out_h16 = reinterpret_cast<op_t*>(out_h16); a.out_ld = Cout; a.sums = sums;
  return launch_conv_igemm(a, as_stream(stream));
}
