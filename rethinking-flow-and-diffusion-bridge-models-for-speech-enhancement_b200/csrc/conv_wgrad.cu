// Weight gradient of the backbone's convolutions on Blackwell tensor cores (training step, SURVEY.md section 8 A10;
// the autograd of nn.Conv2d / NIN in fdbm/backbones/ncsnpp_utils/layers.py:100-124,546-555):
//
//   dW[n, c, df, dt] += sum_{b,t,f} dY[b,t,f,n] * X[b, t+dt-1, f+df-1, c]          (3x3; 1x1 uses the centre tap only)
//
// GEMM view per tap: D[M = Cout block of 128, N = Cin block of 128 or 64] over K = pixels.  Both operands live in
// HBM as [B,T,F,C] (channels innermost), i.e. the GEMM's K (pixel) index is the ROW index of the shared-memory
// tile and M / N run along the 128-byte rows: MN-major UMMA operands.  One TMA box per 64 channels lands with
// the 128-byte swizzle; in the canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) (units of 16 B)
//   LBO = distance between the 64-channel boxes, SBO = distance between groups of 8 pixel rows,
// so the 18x10-pixel halo box of X serves all taps by shifting the descriptor start address by (dt*10+df) rows
// (SBO = 10 rows = 1280 B), exactly as in the forward kernel, and dY is a dense 16x8-pixel box (SBO = 1024 B).
//
// Work split: grid = (pixel splits, tap groups {df = 0,1,2}, Cout blocks x Cin blocks).  A CTA keeps its three
// taps' accumulators (3 x 128 TMEM columns) for its whole pixel range (split-K), then writes them as fp32 partials
// [split][tap][Cout][Cin] to a workspace; wgrad_reduce_kernel sums the splits in a fixed order (deterministic)
// and accumulates into dW (OIHW fp32).
#include <cuda.h>
#include "common.cuh"
#include "tc05.cuh"

namespace fdbm {
namespace {

using namespace tc05;

constexpr int TILE_T = 16, TILE_F = 8;
constexpr int HALO_T = TILE_T + 2, HALO_F = TILE_F + 2;
constexpr int DY_BOX_BYTES = TILE_T * TILE_F * 128;          // 16384: 128 pixel rows x 64 channels
constexpr int X_BOX_BYTES = HALO_T * HALO_F * 128;           // 23040
constexpr int X_BOX_STRIDE = 23552;                          // rounded up to 1024
constexpr int STAGES = 2;
constexpr int STAGE_BYTES = 2 * DY_BOX_BYTES + 2 * X_BOX_STRIDE;     // dY: 128 Cout; X: up to 128 Cin
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256;
constexpr int NUM_THREADS = 192;                             // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue

struct WgradParams {
  int B, T, F, Cout, Cin;
  int taps;                    // 9 or 1
  int n_cin;                   // channels of the Cin block handled per CTA: 128 or 64
  int tiles_t, tiles_f, n_tiles, n_splits, tiles_per_split;
  int cin_blocks;
  int dy_coff, x_coff;         // channel offsets inside the (wider) dY / X tensors
  float* partial;              // [split][tap][Cout][Cin]
};

// idesc for kind::f16 with BOTH operands MN-major (bits 15 / 16), fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N, int is_bf16) {
  return make_idesc_f16(M, N, is_bf16) | (1u << 15) | (1u << 16);
}
// MN-major SW128 descriptor: LBO = stride between 64-element groups along M/N, SBO = stride between 8-row K groups
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

  const int split = blockIdx.x;
  const int group = blockIdx.y;                           // df for 3x3; 0 for 1x1
  const int co_blk = blockIdx.z / p.cin_blocks, ci_blk = blockIdx.z % p.cin_blocks;
  const int n_taps_cta = p.taps == 9 ? 3 : 1;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(p.n_tiles, tile_begin + p.tiles_per_split);
  const int n_x_boxes = p.n_cin / 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  // programmatic dependent launch: the prologue above overlapped the tail of the previous kernel; its results (and buffers it still
  // read) are only touched from here on.  The split-K reduction behind this kernel may become resident (it waits the same way).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int tt = tile % p.tiles_t;
        const int rest = tile / p.tiles_t;
        const int t0 = tt * TILE_T, f0 = (rest % p.tiles_f) * TILE_F, b = rest / p.tiles_f;
        mbar_wait(empty + stage, phase ^ 1);
        mbar_expect_tx(full + stage, 2 * DY_BOX_BYTES + n_x_boxes * X_BOX_BYTES);
        uint8_t* s = smem + stage * STAGE_BYTES;
        tma_load_4d(s, &map_dy, full + stage, p.dy_coff + co_blk * 128, f0, t0, b);
        tma_load_4d(s + DY_BOX_BYTES, &map_dy, full + stage, p.dy_coff + co_blk * 128 + 64, f0, t0, b);
        for (int h = 0; h < n_x_boxes; ++h)
          tma_load_4d(s + 2 * DY_BOX_BYTES + h * X_BOX_STRIDE, &map_x, full + stage, p.x_coff + ci_blk * p.n_cin + h * 64, f0 - 1, t0 - 1, b);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp converged, one elected lane issues (operands must be warp-uniform)
    const uint32_t idesc = make_idesc_mn(128, p.n_cin, kOperandIsBf16);
    uint32_t stage = 0, phase = 0;
    bool first = true;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(full + stage, phase);
      fence_after_sync();
      const uint32_t s_dy = smem_u32(smem + stage * STAGE_BYTES);
      const uint32_t s_x = s_dy + 2 * DY_BOX_BYTES;
      if (elect_one()) {
        for (int tl = 0; tl < n_taps_cta; ++tl) {
          const int df = p.taps == 9 ? group : 1, dt = p.taps == 9 ? tl : 1;
#pragma unroll
          for (int k = 0; k < TILE_T / 2; ++k) {            // 16 pixels = frames 2k, 2k+1 of the tile
            const uint64_t da = make_desc_mn(s_dy + k * 2048, DY_BOX_BYTES, 1024);
            const uint64_t db = make_desc_mn(s_x + ((2 * k + dt) * HALO_F + df) * 128, X_BOX_STRIDE, HALO_F * 128);
            mma_f16(tmem_base + tl * 128, da, db, idesc, (first && k == 0) ? 0u : 1u);
          }
        }
        mma_commit(empty + stage);
      }
      __syncwarp();
      first = false;
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) mma_commit(done);
    __syncwarp();
  } else {
    // epilogue: TMEM lane = output channel (row of dW), columns = input channels
    const int q = warp & 3;
    mbar_wait(done, 0);
    fence_after_sync();
    const bool any = tile_end > tile_begin;
    const int co = co_blk * 128 + q * 32 + lane;
    for (int tl = 0; tl < n_taps_cta; ++tl) {
      const int tap = p.taps == 9 ? group * 3 + tl : 0;
      float* dst = p.partial + ((static_cast<int64_t>(split) * p.taps + tap) * p.Cout + co) * p.Cin + ci_blk * p.n_cin;
      for (int c0 = 0; c0 < p.n_cin; c0 += 32) {
        uint32_t v[32];
        if (any) {
          tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + tl * 128 + c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(dst + c0)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                               __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dW += scale * sum over splits of the partials [split][tap][Cout][Cin], written in the layout the parameter has:
//   0: OIHW [Cout][Cin_total][k][k], this call covering input channels ci_off..  (tap = df*3 + dt)
//   1: NIN matrix [Cin_total][Cout]
//   2: first conv: D[co][k], k = tap*aux + ci (im2col K-block)  -> OIHW [Cout][aux][3][3]
//   3: pyramid conv: D[c][k], k = tap'*aux + o, tap = 8 - tap'   -> OIHW [aux][Cout][3][3]   (see train_small.cu)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int n_splits, int taps, int Cout, int Cin, float scale, int layout,
                    int Cin_total, int ci_off, int aux, float* __restrict__ dw) {
  const int64_t n = static_cast<int64_t>(taps) * Cout * Cin;
  // programmatic dependent launch: this grid may be resident before the wgrad kernel in front of it has finished; its partials are
  // complete and visible from here on (a no-op when launched without the attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    float acc = 0.f;
    for (int s = 0; s < n_splits; ++s) acc += partial[s * n + i];
    const int ci = static_cast<int>(i % Cin);
    const int co = static_cast<int>((i / Cin) % Cout);
    const int tap = static_cast<int>(i / (static_cast<int64_t>(Cin) * Cout));
    int64_t o;
    if (layout == 0) o = (static_cast<int64_t>(co) * Cin_total + ci_off + ci) * taps + tap;
    else if (layout == 1) o = static_cast<int64_t>(ci_off + ci) * Cout + co;
    else if (layout == 2) { if (ci >= 9 * aux) continue; o = (static_cast<int64_t>(co) * aux + ci % aux) * 9 + ci / aux; }
    else { if (ci >= 9 * aux) continue; o = (static_cast<int64_t>(ci % aux) * Cout + co) * 9 + (8 - ci / aux); }
    dw[o] += scale * acc;
  }
}

int pick_splits(int n_tiles, int groups, int blocks) {
  const int per = std::max(1, num_sms() / (groups * blocks));
  return std::max(1, std::min(per, n_tiles));
}

}  // namespace

int64_t conv_wgrad_workspace_bytes(int Cout, int Cin, int ksize, int B, int T, int F) {
  const int taps = ksize * ksize;
  const int n_tiles = B * ceil_div(T, TILE_T) * ceil_div(F, TILE_F);
  const int n_cin = Cin % 128 == 0 ? 128 : 64;
  const int splits = pick_splits(n_tiles, taps == 9 ? 3 : 1, (Cout / 128) * (Cin / n_cin));
  return static_cast<int64_t>(splits) * taps * Cout * Cin * 4;
}

// dw += scale * wgrad(dy, x).  dy h16 [B,T,F,dy_ld] (channels dy_coff.. +Cout), x h16 [B,T,F,x_ld] (channels x_coff.. +Cin),
// dw in `layout` (see wgrad_reduce_kernel).  workspace: conv_wgrad_workspace_bytes().
int launch_conv_wgrad_ex(const WgradCall& c, cudaStream_t s) {
  FDBM_REQUIRE(c.Cout % 128 == 0 && c.Cin % 64 == 0, "conv_wgrad: Cout %% 128 and Cin %% 64 required (got %d, %d)", c.Cout, c.Cin);
  FDBM_REQUIRE(c.ksize == 1 || c.ksize == 3, "conv_wgrad: ksize must be 1 or 3");
  FDBM_REQUIRE(c.layout == 0 || c.ksize == 1, "conv_wgrad: layouts 1..3 are for 1x1 only");
  FDBM_REQUIRE(c.dy_coff % 64 == 0 && c.x_coff % 64 == 0, "conv_wgrad: channel offsets must be multiples of 64");
  static PerDeviceOnce attr_once;
  if (attr_once.first(current_device()))
    FDBM_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  WgradParams p;
  p.B = c.B; p.T = c.T; p.F = c.F; p.Cout = c.Cout; p.Cin = c.Cin; p.taps = c.ksize * c.ksize;
  p.n_cin = c.Cin % 128 == 0 ? 128 : 64;
  p.cin_blocks = c.Cin / p.n_cin;
  p.tiles_t = ceil_div(c.T, TILE_T); p.tiles_f = ceil_div(c.F, TILE_F);
  p.n_tiles = c.B * p.tiles_t * p.tiles_f;
  const int groups = p.taps == 9 ? 3 : 1;
  const int blocks = (c.Cout / 128) * p.cin_blocks;
  p.n_splits = pick_splits(p.n_tiles, groups, blocks);
  p.tiles_per_split = ceil_div(p.n_tiles, p.n_splits);
  p.dy_coff = c.dy_coff; p.x_coff = c.x_coff;
  p.partial = c.workspace;
  CUtensorMap map_dy, map_x;
  if (int rc = make_act_tile_map(&map_dy, c.dy, c.B, c.T, c.F, c.dy_ld, TILE_F, TILE_T)) return rc;
  if (int rc = make_act_tile_map(&map_x, c.x, c.B, c.T, c.F, c.x_ld, HALO_F, HALO_T)) return rc;
  dim3 grid(p.n_splits, groups, blocks);
  FDBM_CUDA(launch_maybe_pdl(conv_wgrad_kernel, grid, dim3(NUM_THREADS), SMEM_BYTES, s, pdl_enabled() && c.B <= pdl_batch_limit(), map_dy, map_x, p));
  FDBM_LAUNCH_CHECK();
  const int64_t n = static_cast<int64_t>(p.taps) * c.Cout * c.Cin;
  // the reduction is launched with programmatic stream serialisation: its launch latency and ramp overlap the tail of the wgrad
  // kernel (228 such pairs per training step); FDBM_PDL=0 launches it plainly
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(std::min<int64_t>(ceil_div64(n, 256), 2048))); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  FDBM_CUDA(cudaLaunchKernelEx(&cfg, wgrad_reduce_kernel, static_cast<const float*>(c.workspace), p.n_splits, p.taps, c.Cout, c.Cin, c.scale, c.layout,
                               c.Cin_total ? c.Cin_total : c.Cin, c.ci_off, c.aux, c.dw));
  FDBM_LAUNCH_CHECK();
  return FDBM_OK;
}

int launch_conv_wgrad(const op_t* dy, int Cout, const op_t* x, int Cin, int ksize, int B, int T, int F, float scale,
                      int io_layout, float* dw, float* workspace, cudaStream_t s) {
  WgradCall c;
  c.dy = dy; c.dy_ld = Cout; c.Cout = Cout; c.x = x; c.x_ld = Cin; c.Cin = Cin; c.ksize = ksize; c.B = B; c.T = T; c.F = F;
  c.scale = scale; c.layout = io_layout ? 1 : 0; c.dw = dw; c.workspace = workspace;
  return launch_conv_wgrad_ex(c, s);
}

}  // namespace fdbm

using namespace fdbm;

extern "C" int64_t fdbm_conv_wgrad_workspace_bytes(int Cout, int Cin, int ksize, int batch, int T, int F) {
  if (Cout <= 0 || Cin <= 0 || batch <= 0 || T <= 0 || F <= 0 || (ksize != 1 && ksize != 3)) return 0;
  return conv_wgrad_workspace_bytes(Cout, Cin, ksize, batch, T, F);
}

extern "C" int fdbm_conv_wgrad(const void* dy, int Cout, const void* x, int Cin, int ksize, int batch, int T, int F,
                               float scale, float* dw, float* workspace, void* stream) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(dy && x && dw && workspace && batch > 0 && T > 0 && F > 0, "fdbm_conv_wgrad: bad arguments");
  return launch_conv_wgrad(reinterpret_cast<const op_t*>(dy), Cout, reinterpret_cast<const op_t*>(x), Cin, ksize, batch, T,
                           F, scale, 0, dw, workspace, as_stream(stream));
}
