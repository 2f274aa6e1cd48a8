// The NCSN++ forward as a static launch plan (fdbm/backbones/ncsnpp_v2.py:241-401 and
// ncsnpp_v2_predictive.py:222-362), plus the N-step sampler loop of fdbm/bridge.py:66-113.
//
// fdbm_plan_create walks the reference constructor's module list (ncsnpp_v2.py:95-239) once for a
// fixed (batch, n_frames) and records, in execution order, every kernel launch with its buffers
// resolved inside one device arena (first-fit with explicit frees, so the skip stack and the
// temporaries of a residual block reuse memory).  Forward / sampler replay that list; the sampler
// captures all N steps (backbone + bridge update) into one CUDA graph per argument set.
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "common.cuh"

using namespace fdbm;

namespace {

struct Mod {
  enum Kind { FOURIER, LINEAR, CONV3, RES, ATTN, COMBINE, GN } kind;
  int idx, cin, cout;
  bool up, down;
};

struct Slot { int64_t off; int64_t numel; bool loaded; };     // fp32 parameter inside `params`

struct Act {                 // one residual-stream tensor [B,T,F,C]: fp32 master, 16-bit GEMM-operand copy, channel sums
  float* data = nullptr;
  op_t* h16 = nullptr;
  double* sums = nullptr;
  int C = 0, T = 0, F = 0;
};

struct GraphKey {
  const void* p[6]; int n_steps, kind; uint64_t seed;
  bool operator==(const GraphKey& o) const {
    for (int i = 0; i < 6; ++i) if (p[i] != o.p[i]) return false;
    return n_steps == o.n_steps && kind == o.kind && seed == o.seed;
  }
};

// first-fit arena over one cudaMalloc
class Arena {
 public:
  void reset(int64_t cap) { free_.clear(); free_[0] = cap; cap_ = cap; peak_ = 0; }
  int64_t alloc(int64_t bytes) {
    bytes = (bytes + 1023) / 1024 * 1024;
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= bytes) {
        const int64_t off = it->first, rest = it->second - bytes;
        free_.erase(it);
        if (rest) free_[off + bytes] = rest;
        live_[off] = bytes;
        peak_ = std::max(peak_, off + bytes);
        return off;
      }
    }
    return -1;
  }
  void release(int64_t off) {
    auto it = live_.find(off);
    if (it == live_.end()) return;
    int64_t o = off, n = it->second;
    live_.erase(it);
    auto nx = free_.lower_bound(o);
    if (nx != free_.end() && o + n == nx->first) { n += nx->second; nx = free_.erase(nx); }
    if (nx != free_.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == o) { o = pv->first; n += pv->second; free_.erase(pv); }
    }
    free_[o] = n;
  }
  int64_t peak() const { return peak_; }
 private:
  std::map<int64_t, int64_t> free_, live_;
  int64_t cap_ = 0, peak_ = 0;
};

__global__ void add_vec_kernel(const float* a, const float* b, float* o, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) o[i] = a[i] + (b ? b[i] : 0.f);
}

}  // namespace

struct fdbm_plan {
  fdbm_arch arch;
  int B = 0, T = 0, F = 0, F_io = 257, Cin = 4;
  std::vector<Mod> mods;
  // parameters
  std::unordered_map<std::string, Slot> slots;
  int64_t params_numel = 0;
  float* params = nullptr;            // fp32 originals + derived fp32 (combined biases)
  op_t* wpacked = nullptr;   // packed conv weights
  int64_t wpacked_bytes = 0;
  int dense_rows = 0;
  bool weights_ready = false;
  // activations
  uint8_t* arena = nullptr;
  int64_t arena_bytes = 0;
  // per-call arguments the recorded ops read
  const float* cur_x = nullptr; const float* cur_y = nullptr; const float* cur_t = nullptr; int cur_t_stride = 1;
  float* cur_out = nullptr;
  float* d_buf = nullptr;             // backbone output inside the sampler loop
  std::vector<std::function<int(cudaStream_t)>> ops;        // one forward: exactly one kernel launch per entry
  std::vector<int> op_kind;                                  // FDBM_OP_* of every entry
  std::vector<double> op_flops;                              // algorithmic FLOPs (2*MAC) of every entry (convs)
  std::vector<std::function<int(cudaStream_t)>> pack_ops;   // weight packing after load_weights
  int n_launches = 0;
  // graph cache
  bool have_graph = false; GraphKey graph_key{}; cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t capture_stream = nullptr;   // private stream: capture works even when the caller is on the legacy stream
};

void fdbm_plan_release_sampler_state(fdbm_plan* plan);

namespace {

std::vector<Mod> build_modules(const fdbm_arch& a) {
  std::vector<Mod> m;
  auto add = [&](Mod::Kind k, int cin, int cout, bool up = false, bool down = false) {
    m.push_back(Mod{k, static_cast<int>(m.size()), cin, cout, up, down});
  };
  const int nf = a.nf, C = a.predictive ? 2 : 4, L = a.n_levels;
  if (!a.predictive) { add(Mod::FOURIER, 0, nf); add(Mod::LINEAR, 2 * nf, 4 * nf); add(Mod::LINEAR, 4 * nf, 4 * nf); }
  add(Mod::CONV3, C, nf);
  std::vector<int> hs_c{nf};
  int in_ch = nf;
  for (int lvl = 0; lvl < L; ++lvl) {
    const int res = a.image_size >> lvl;
    for (int b = 0; b < a.num_res_blocks; ++b) {
      const int out_ch = nf * a.ch_mult[lvl];
      add(Mod::RES, in_ch, out_ch);
      in_ch = out_ch;
      if (res == a.attn_resolution) add(Mod::ATTN, in_ch, in_ch);
      hs_c.push_back(in_ch);
    }
    if (lvl != L - 1) {
      add(Mod::RES, in_ch, in_ch, false, true);
      add(Mod::COMBINE, C, in_ch);
      hs_c.push_back(in_ch);
    }
  }
  in_ch = hs_c.back();
  add(Mod::RES, in_ch, in_ch); add(Mod::ATTN, in_ch, in_ch); add(Mod::RES, in_ch, in_ch);
  for (int lvl = L - 1; lvl >= 0; --lvl) {
    const int res = a.image_size >> lvl;
    for (int b = 0; b < a.num_res_blocks + 1; ++b) {
      const int out_ch = nf * a.ch_mult[lvl];
      add(Mod::RES, in_ch + hs_c.back(), out_ch);
      hs_c.pop_back();
      in_ch = out_ch;
    }
    if (res == a.attn_resolution) add(Mod::ATTN, in_ch, in_ch);
    add(Mod::GN, in_ch, in_ch);
    add(Mod::CONV3, in_ch, C);
    if (lvl != 0) add(Mod::RES, in_ch, in_ch, true, false);
  }
  return m;
}

struct Builder {
  fdbm_plan* P;
  Arena arena;
  int64_t wp_off = 0;                  // running offset (bytes) into wpacked
  int dense_off = 0;                   // running row offset into the Dense_0 table
  bool dry = true;                     // first pass: sizes only

  // ---------------- parameters
  int64_t param(const std::string& name, int64_t numel) {
    auto it = P->slots.find(name);
    if (it != P->slots.end()) return it->second.off;
    const int64_t off = P->params_numel;
    P->slots[name] = Slot{off, numel, false};
    P->params_numel += (numel + 3) / 4 * 4;              // keep every tensor 16-byte aligned
    return off;
  }
  int64_t derived(int64_t numel) {                       // fp32 scratch parameter (not loaded by name)
    const int64_t off = P->params_numel;
    P->params_numel += (numel + 3) / 4 * 4;
    return off;
  }
  float* pp(int64_t off) const { return P->params + off; }     // call OUTSIDE recorded lambdas only
  std::string pre(const Mod& m) const { return "all_modules." + std::to_string(m.idx) + "."; }

  // ---------------- activations
  template <typename Tp> Tp* alloc(int64_t n_elems) {
    const int64_t off = arena.alloc(n_elems * static_cast<int64_t>(sizeof(Tp)));
    if (off < 0) return nullptr;
    return reinterpret_cast<Tp*>(P->arena + off);
  }
  void release(const void* p) { if (p) arena.release(reinterpret_cast<const uint8_t*>(p) - P->arena); }
  Act new_act(int C, int T, int F) {
    Act a; a.C = C; a.T = T; a.F = F;
    a.data = alloc<float>(static_cast<int64_t>(P->B) * T * F * C);
    a.h16 = alloc<op_t>(static_cast<int64_t>(P->B) * T * F * C);
    a.sums = alloc<double>(static_cast<int64_t>(P->B) * C * 2);
    return a;
  }
  void free_act(Act& a) { release(a.data); release(a.h16); release(a.sums); a.data = nullptr; a.h16 = nullptr; a.sums = nullptr; }
  // (scale, shift) table of a GroupNorm over the channel concatenation x1 (+ x2): a tiny launch; the consuming
  // convolution applies it while the operand tile sits in shared memory
  float2* norm_table(const double* q1, int C1, const double* q2, int C2, const float* gamma, const float* beta, int T, int F) {
    const int B = P->B;
    float2* tab = alloc<float2>(static_cast<int64_t>(B) * (C1 + C2));
    const int64_t px = static_cast<int64_t>(T) * F;
    op([=](cudaStream_t s) { return launch_gn_finalize(q1, C1, q2, C2, gamma, beta, B, px, tab, s); }, FDBM_OP_STATS);
    return tab;
  }
  static ConvSeg seg(const op_t* in, int C, int taps, const float2* tab = nullptr, int tab_stride = 0, int act = 0) {
    ConvSeg sg; sg.in = in; sg.C = C; sg.taps = taps; sg.norm_tab = tab; sg.tab_stride = tab_stride; sg.act = act;
    return sg;
  }

  void op(std::function<int(cudaStream_t)> f, int kind, double flops = 0.0) {
    if (dry) return;
    P->ops.push_back(std::move(f));
    P->op_kind.push_back(kind);
    P->op_flops.push_back(flops);
    P->n_launches += 1;
  }
  // convolution; the channel statistics of its output come out of the same kernel's epilogue
  void conv_op(ConvArgs c, double flops = 0.0) {
    if (flops == 0.0) {
      double k = 0.0;
      for (int i = 0; i < c.n_seg; ++i) k += static_cast<double>(c.seg[i].taps) * c.seg[i].C;
      flops = 2.0 * c.B * c.T * c.F * c.Cout * k;
    }
    op([=](cudaStream_t s) { return launch_conv_igemm(c, s); }, FDBM_OP_CONV, flops);
  }
  void pack_op(std::function<int(cudaStream_t)> f) { if (!dry) P->pack_ops.push_back(std::move(f)); }

  // packed weights for conv (w1: name, C1, ksize (3, 1 or -1 = NIN [in][out])), optional fused 1x1 w2
  op_t* pack(const std::string& w1, int C1, int ksize, const std::string& w2, int C2, int Cout,
                      int rows_total = 0, int row_off = 0, op_t* into = nullptr) {
    const int k = ksize == -1 ? 1 : ksize;
    if (rows_total == 0) rows_total = Cout;
    op_t* dst = into;
    if (!dst) {
      dst = reinterpret_cast<op_t*>(reinterpret_cast<uint8_t*>(P->wpacked) + wp_off);
      wp_off += (conv_wpack_bytes(C1, k, C2, rows_total) + 1023) / 1024 * 1024;
    }
    const float* p1 = pp(param(w1, static_cast<int64_t>(Cout) * C1 * k * k));
    const float* p2 = C2 ? pp(param(w2, static_cast<int64_t>(Cout) * C2)) : nullptr;
    pack_op([=](cudaStream_t s) {
      return launch_pack_conv_weights(p1, C1, ksize, p2, C2, Cout, rows_total, row_off, dst, s);
    });
    return dst;
  }

  // ---------------- layers
  // ResnetBlockBigGANpp (layerspp.py:242-274) on the concatenation of x1 (and x2)
  struct CombineArgs { const float* pyr; const float* w; const float* b; int Cp; };
  Act resblock(const Mod& m, const Act& x1, const Act* x2, const float* dense, int dense_stride,
               const CombineArgs* comb = nullptr) {
    const int B = P->B, Cin = m.cin, Cout = m.cout;
    const int mode = m.down ? 1 : (m.up ? 2 : 0);
    const int T = x1.T, F = x1.F;
    const int To = mode == 1 ? T / 2 : (mode == 2 ? T * 2 : T), Fo = mode == 1 ? F / 2 : (mode == 2 ? F * 2 : F);
    const bool shortcut = (Cin != Cout) || m.up || m.down;
    const std::string p = pre(m);
    const float* g0w = pp(param(p + "GroupNorm_0.weight", Cin)); const float* g0b = pp(param(p + "GroupNorm_0.bias", Cin));
    const float* c0b = pp(param(p + "Conv_0.bias", Cout));
    const float* g1w = pp(param(p + "GroupNorm_1.weight", Cout)); const float* g1b = pp(param(p + "GroupNorm_1.bias", Cout));
    const float* c1b = pp(param(p + "Conv_1.bias", Cout));
    op_t* w0 = pack(p + "Conv_0.weight", Cin, 3, "", 0, Cout);
    op_t* w1 = shortcut ? pack(p + "Conv_1.weight", Cout, 3, p + "Conv_2.weight", Cin, Cout)
                                 : pack(p + "Conv_1.weight", Cout, 3, "", 0, Cout);
    const float* bias1 = c1b;
    if (shortcut) {                                   // Conv_1.bias + Conv_2.bias, summed once at load time
      const float* c2b = pp(param(p + "Conv_2.bias", Cout));
      float* bsum = pp(derived(Cout));
      bias1 = bsum;
      pack_op([=](cudaStream_t s) {
        add_vec_kernel<<<ceil_div(Cout, 256), 256, 0, s>>>(c1b, c2b, bsum, Cout);
        FDBM_LAUNCH_CHECK();
        return FDBM_OK;
      });
    }
    int dense_row = -1;
    if (!P->arch.predictive) {
      param(p + "Dense_0.weight", static_cast<int64_t>(Cout) * 4 * P->arch.nf);   // slots are laid out by build_dense()
      dense_row = dense_off;
      dense_off += Cout;
    }

    const int64_t npx = static_cast<int64_t>(B) * To * Fo;
    const int C1 = x1.C, C2 = x2 ? x2->C : 0;
    // Conv_0 output only feeds GroupNorm_1: keep it in the 16-bit operand format (its statistics are taken
    // from the fp32 accumulators in the conv epilogue, before rounding)
    op_t* h1 = alloc<op_t>(npx * Cout);
    double* h1_sums = alloc<double>(static_cast<int64_t>(B) * Cout * 2);
    op_t* a0 = nullptr; op_t* xr = nullptr; float2* tab0 = nullptr;
    ConvArgs c0;
    c0.wpack = w0; c0.bias = c0b; c0.bias_b = dense_row >= 0 ? dense + dense_row : nullptr; c0.bias_b_stride = dense_stride;
    c0.B = B; c0.T = To; c0.F = Fo; c0.Cout = Cout; c0.out_h16 = h1; c0.sums = h1_sums;
    if (mode == 0) {
      // GroupNorm_0 + SiLU applied by Conv_0 on load, straight from the 16-bit copies of the residual stream
      tab0 = norm_table(x1.sums, C1, x2 ? x2->sums : nullptr, C2, g0w, g0b, T, F);
      c0.seg[0] = seg(x1.h16, C1, 9, tab0, Cin, 1); c0.n_seg = 1;
      if (x2) { c0.seg[1] = seg(x2->h16, C2, 9, tab0 + C1, Cin, 1); c0.n_seg = 2; }
    } else {
      // resampling blocks: one pass does GroupNorm_0 + SiLU + FIR up/down of h and the FIR of the raw shortcut operand
      a0 = alloc<op_t>(npx * Cin);
      xr = alloc<op_t>(npx * Cin);
      const float* s1 = x1.data; const double* q1 = x1.sums;
      const float* s2 = x2 ? x2->data : nullptr; const double* q2 = x2 ? x2->sums : nullptr;
      op([=](cudaStream_t s) {
        return launch_groupnorm_act(s1, 0, q1, C1, s2, q2, C2, g0w, g0b, B, T, F, 1, mode, a0, xr, s);
      }, FDBM_OP_NORM);
      c0.seg[0] = seg(a0, Cin, 9); c0.n_seg = 1;
    }
    conv_op(c0);
    release(a0); release(tab0);
    float2* tab1 = norm_table(h1_sums, Cout, nullptr, 0, g1w, g1b, To, Fo);
    Act out = new_act(Cout, To, Fo);
    {
      ConvArgs c;
      c.seg[0] = seg(h1, Cout, 9, tab1, Cout, 1); c.n_seg = 1;
      if (shortcut) {
        if (mode == 0) {
          c.seg[c.n_seg++] = seg(x1.h16, C1, 1);
          if (x2) c.seg[c.n_seg++] = seg(x2->h16, C2, 1);
        } else {
          c.seg[c.n_seg++] = seg(xr, Cin, 1);
        }
      }
      c.wpack = w1; c.bias = bias1; c.residual = shortcut ? nullptr : x1.data;
      c.scale = 0.70710678118654752f; c.B = B; c.T = To; c.F = Fo; c.Cout = Cout;
      c.out_f32 = out.data; c.out_h16 = out.h16; c.sums = out.sums;
      if (comb) { c.comb_pyr = comb->pyr; c.comb_w = comb->w; c.comb_b = comb->b; c.comb_C = comb->Cp; }
      conv_op(c);
    }
    release(h1); release(h1_sums); release(tab1);
    release(xr);
    return out;
  }

  // AttnBlockpp (layerspp.py:75-91)
  Act attn(const Mod& m, const Act& x) {
    const int B = P->B, C = m.cin, T = x.T, F = x.F;
    const std::string p = pre(m);
    const float* gw = pp(param(p + "GroupNorm_0.weight", C)); const float* gb = pp(param(p + "GroupNorm_0.bias", C));
    // q, k, v projections as one GEMM with 3C outputs; biases are three consecutive slots
    const float* bq = pp(param(p + "NIN_0.b", C));
    param(p + "NIN_1.b", C); param(p + "NIN_2.b", C);
    const float* b3 = pp(param(p + "NIN_3.b", C));
    op_t* wqkv = pack(p + "NIN_0.W", C, -1, "", 0, C, 3 * C, 0);
    pack(p + "NIN_1.W", C, -1, "", 0, C, 3 * C, C, wqkv);
    pack(p + "NIN_2.W", C, -1, "", 0, C, 3 * C, 2 * C, wqkv);
    op_t* w3 = pack(p + "NIN_3.W", C, -1, "", 0, C);
    const int64_t npx = static_cast<int64_t>(B) * T * F;
    float2* tab = norm_table(x.sums, C, nullptr, 0, gw, gb, T, F);
    op_t* qkv = alloc<op_t>(npx * 3 * C);
    {
      ConvArgs c;
      c.seg[0] = seg(x.h16, C, 1, tab, C, 0); c.n_seg = 1;      // GroupNorm_0 (no activation) on load
      c.wpack = wqkv; c.bias = bq; c.B = B; c.T = T; c.F = F; c.Cout = 3 * C;
      c.out_h16 = qkv;
      conv_op(c);
    }
    release(tab);
    op_t* o = alloc<op_t>(npx * C);
    op([=](cudaStream_t s) { return launch_attention(qkv, qkv + C, qkv + 2 * C, 3 * C, B, T * F, C, o, C, s); }, FDBM_OP_ATTN);
    release(qkv);
    Act out = new_act(C, T, F);
    {
      ConvArgs c;
      c.seg[0] = seg(o, C, 1); c.n_seg = 1;
      c.wpack = w3; c.bias = b3; c.residual = x.data; c.scale = 0.70710678118654752f;
      c.B = B; c.T = T; c.F = F; c.Cout = C; c.out_f32 = out.data; c.out_h16 = out.h16; c.sums = out.sums;
      conv_op(c);
    }
    release(o);
    return out;
  }

  int build() {
    fdbm_plan& pl = *P;
    const fdbm_arch& A = pl.arch;
    const int B = pl.B, nf = A.nf, Cp = pl.Cin, L = A.n_levels;
    size_t mi = 0;
    auto next = [&]() -> const Mod& { return pl.mods[mi++]; };
    wp_off = 0; dense_off = 0;

    // ---- time embedding + all Dense_0 projections (one table [B, dense_rows])
    float* temb_act = nullptr; float* dense = nullptr;
    if (!A.predictive) {
      const Mod& mf = next(); const Mod& l1 = next(); const Mod& l2 = next();
      const float* fw = pp(param(pre(mf) + "W", nf));
      const float* w1 = pp(param(pre(l1) + "weight", static_cast<int64_t>(4 * nf) * 2 * nf)); const float* b1 = pp(param(pre(l1) + "bias", 4 * nf));
      const float* w2 = pp(param(pre(l2) + "weight", static_cast<int64_t>(4 * nf) * 4 * nf)); const float* b2 = pp(param(pre(l2) + "bias", 4 * nf));
      // Dense_0 weights / biases of all residual blocks, contiguous and in execution order
      int rows = 0; int64_t dw0 = -1, db0 = -1;
      for (const Mod& m : pl.mods) if (m.kind == Mod::RES) {
        const int64_t o = param(pre(m) + "Dense_0.weight", static_cast<int64_t>(m.cout) * 4 * nf);
        if (dw0 < 0) dw0 = o;
        rows += m.cout;
      }
      for (const Mod& m : pl.mods) if (m.kind == Mod::RES) {
        const int64_t o = param(pre(m) + "Dense_0.bias", m.cout);
        if (db0 < 0) db0 = o;
      }
      pl.dense_rows = rows;
      temb_act = alloc<float>(static_cast<int64_t>(B) * 4 * nf);
      dense = alloc<float>(static_cast<int64_t>(B) * rows);
      const float* dwp = pp(dw0); const float* dbp = pp(db0);
      fdbm_plan* plp = P;
      op([=](cudaStream_t s) {
        return launch_temb(plp->cur_t, fw, nf, w1, b1, w2, b2, B, plp->cur_t_stride, temb_act, s);
      }, FDBM_OP_SMALL);
      op([=](cudaStream_t s) { return launch_dense_all(temb_act, dwp, dbp, B, 4 * nf, rows, dense, s); }, FDBM_OP_SMALL);
    }
    const int dstride = pl.dense_rows;

    // ---- input packing and first convolution
    int T = pl.T, F = pl.F;
    float* pyr_in = alloc<float>(static_cast<int64_t>(B) * T * F * Cp);
    {
      fdbm_plan* plp = P; const int Tc = T, Fc = F; float* dst = pyr_in; const bool pred = A.predictive;
      op([=](cudaStream_t s) { return launch_pack_input(plp->cur_x, pred ? nullptr : plp->cur_y, B, Tc, plp->F_io, Fc, Cp, dst, s); }, FDBM_OP_SKINNY);
    }
    std::vector<Act> hs;
    {
      const Mod& m = next();
      const float* w = pp(param(pre(m) + "weight", static_cast<int64_t>(nf) * Cp * 9)); const float* b = pp(param(pre(m) + "bias", nf));
      // first conv (4 -> nf, 3x3) on the tensor cores: im2col to one 64-wide K-block, then a K=64 GEMM
      op_t* wp = reinterpret_cast<op_t*>(reinterpret_cast<uint8_t*>(P->wpacked) + wp_off);
      wp_off += (static_cast<int64_t>(nf) * 64 * 2 + 1023) / 1024 * 1024;
      pack_op([=](cudaStream_t s) { return launch_pack_conv_weights(w, Cp, -2, nullptr, 0, nf, nf, 0, wp, s); });
      Act h0 = new_act(nf, T, F);
      const int Tc = T, Fc = F; float* src = pyr_in;
      op_t* cols = alloc<op_t>(static_cast<int64_t>(B) * T * F * 64);
      op([=](cudaStream_t s) { return launch_im2col_input(src, Cp, B, Tc, Fc, cols, s); }, FDBM_OP_SKINNY);
      ConvArgs c;
      c.seg[0] = seg(cols, 64, 1); c.n_seg = 1;
      c.wpack = wp; c.bias = b; c.B = B; c.T = T; c.F = F; c.Cout = nf;
      c.out_f32 = h0.data; c.out_h16 = h0.h16; c.sums = h0.sums;
      conv_op(c, 2.0 * B * T * F * nf * 9.0 * Cp);
      release(cols);
      hs.push_back(h0);
    }
    // ---- down path
    for (int lvl = 0; lvl < L; ++lvl) {
      for (int blk = 0; blk < A.num_res_blocks; ++blk) {
        Act h = resblock(next(), hs.back(), nullptr, dense, dstride);
        if (h.F == A.attn_resolution) { Act h2 = attn(next(), h); free_act(h); h = h2; }
        hs.push_back(h);
      }
      if (lvl != L - 1) {
        // input pyramid one level down, then the down-sampling block whose Conv_1 epilogue also applies the
        // Combine (layerspp.py:52-59): h = resblock(...) + conv1x1(4 -> C)(pyramid)
        const Mod& mr = next();
        float* pyr_next = alloc<float>(static_cast<int64_t>(B) * (T / 2) * (F / 2) * Cp);
        {
          const int Tc = T, Fc = F; float* src = pyr_in;
          op([=](cudaStream_t s) { return launch_fir_resample(src, B, Tc, Fc, Cp, 1, pyr_next, s); }, FDBM_OP_SKINNY);
        }
        release(pyr_in);
        pyr_in = pyr_next;
        const Mod& m = pl.mods[mi];               // the Combine module follows the block
        CombineArgs cb;
        cb.pyr = pyr_in; cb.Cp = Cp;
        cb.w = pp(param(pre(m) + "Conv_0.weight", static_cast<int64_t>(m.cout) * Cp)); cb.b = pp(param(pre(m) + "Conv_0.bias", m.cout));
        Act h = resblock(mr, hs.back(), nullptr, dense, dstride, &cb);
        next();
        T /= 2; F /= 2;
        hs.push_back(h);
      }
    }
    release(pyr_in);
    // ---- bottleneck
    Act h = hs.back();                      // still owned by the skip stack
    {
      Act a = resblock(next(), h, nullptr, dense, dstride);
      Act b2 = attn(next(), a); free_act(a);
      Act c = resblock(next(), b2, nullptr, dense, dstride); free_act(b2);
      h = c;
    }
    // ---- up path
    float* pyramid = nullptr;
    for (int lvl = L - 1; lvl >= 0; --lvl) {
      for (int blk = 0; blk < A.num_res_blocks + 1; ++blk) {
        Act skip = hs.back(); hs.pop_back();
        Act o = resblock(next(), h, &skip, dense, dstride);
        free_act(h); free_act(skip);
        h = o;
      }
      if (h.F == A.attn_resolution) { Act h2 = attn(next(), h); free_act(h); h = h2; }
      {
        const Mod& mg = next(); const Mod& mc = next();
        const int C = mg.cin, Tc = h.T, Fc = h.F;
        const float* gw = pp(param(pre(mg) + "weight", C)); const float* gb = pp(param(pre(mg) + "bias", C));
        const float* w = pp(param(pre(mc) + "weight", static_cast<int64_t>(Cp) * C * 9)); const float* b = pp(param(pre(mc) + "bias", Cp));
        // C -> 4 conv on the tensor cores: weight rows padded with zeros to one 128-wide N block
        op_t* wp = reinterpret_cast<op_t*>(reinterpret_cast<uint8_t*>(P->wpacked) + wp_off);
        const int64_t wbytes = conv_wpack_bytes(C, 3, 0, 128);
        wp_off += (wbytes + 1023) / 1024 * 1024;
        pack_op([=](cudaStream_t s) {
          FDBM_CUDA(cudaMemsetAsync(wp, 0, wbytes, s));
          return launch_pack_conv_weights(w, C, 3, nullptr, 0, Cp, 128, 0, wp, s);
        });
        float* pyr_new = alloc<float>(static_cast<int64_t>(B) * Tc * Fc * Cp);
        float* prev = pyramid;
        float2* tab = norm_table(h.sums, C, nullptr, 0, gw, gb, Tc, Fc);
        ConvArgs c;
        c.seg[0] = seg(h.h16, C, 9, tab, C, 1); c.n_seg = 1;     // GroupNorm + SiLU on load
        c.wpack = wp; c.bias = b; c.B = B; c.T = Tc; c.F = Fc; c.Cout = 128;
        c.pyr_out = pyr_new; c.pyr_prev = prev; c.pyr_C = Cp;
        conv_op(c, 2.0 * B * Tc * Fc * Cp * 9.0 * C);
        release(tab);
        release(pyramid);
        pyramid = pyr_new;
      }
      if (lvl != 0) { Act o = resblock(next(), h, nullptr, dense, dstride); free_act(h); h = o; }
    }
    free_act(h);
    if (mi != pl.mods.size() || !hs.empty()) { set_error("plan: module walk out of sync (%zu of %zu)", mi, pl.mods.size()); return FDBM_EINVAL; }
    {
      const float* w = pp(param("output_layer.weight", 2 * Cp)); const float* b = pp(param("output_layer.bias", 2));
      const int Tc = pl.T, Fc = pl.F; fdbm_plan* plp = P; float* pyr = pyramid;
      op([=](cudaStream_t s) { return launch_output_layer(pyr, Cp, w, b, B, Tc, Fc, plp->F_io, plp->cur_out, s); }, FDBM_OP_SKINNY);
    }
    release(pyramid);
    return FDBM_OK;
  }
};

int run_ops(fdbm_plan* plan, cudaStream_t s) {
  for (auto& f : plan->ops) if (int rc = f(s)) return rc;
  return FDBM_OK;
}

}  // namespace

extern "C" int fdbm_plan_create(const fdbm_arch* arch, int batch, int n_frames, fdbm_plan** out) {
  if (int rc = require_sm100()) return rc;
  FDBM_REQUIRE(arch && out, "fdbm_plan_create: null pointer");
  FDBM_REQUIRE(batch > 0 && n_frames > 0, "fdbm_plan_create: batch and n_frames must be positive");
  FDBM_REQUIRE(arch->n_levels >= 1 && arch->n_levels <= 8, "fdbm_plan_create: n_levels out of range");
  FDBM_REQUIRE(arch->nf % 64 == 0 && arch->nf >= 128, "fdbm_plan_create: nf must be a multiple of 64 and >= 128 (got %d)", arch->nf);
  FDBM_REQUIRE(arch->image_size == 256, "fdbm_plan_create: image_size must be 256 (n_fft = 512 with the Nyquist bin dropped)");
  const int down = 1 << (arch->n_levels - 1);
  FDBM_REQUIRE(n_frames % down == 0, "fdbm_plan_create: n_frames %d must be a multiple of %d (pad_spec pads to 64)", n_frames, down);
  for (int i = 0; i < arch->n_levels; ++i)
    FDBM_REQUIRE((arch->nf * arch->ch_mult[i]) % 128 == 0, "fdbm_plan_create: level %d channel count must be a multiple of 128", i);

  fdbm_plan* P = new fdbm_plan();
  P->arch = *arch; P->B = batch; P->T = n_frames; P->F = arch->image_size; P->F_io = arch->image_size + 1;
  P->Cin = arch->predictive ? 2 : 4;
  P->mods = build_modules(*arch);
  auto fail = [&](int rc) { fdbm_plan_destroy(P); return rc; };

  // pass 1 (dry): sizes.  A 1 TiB virtual arena never fails; its peak is the real requirement.
  Builder b1{P};
  b1.dry = true;
  P->arena = reinterpret_cast<uint8_t*>(uintptr_t(1) << 30);      // fake non-null base for the sizing pass
  b1.arena.reset(int64_t(1) << 40);
  if (int rc = b1.build()) return fail(rc);
  P->arena_bytes = b1.arena.peak();
  P->arena = nullptr;
  P->wpacked_bytes = b1.wp_off;
  const int64_t params_numel = P->params_numel;
  // d_buf for the sampler
  const int64_t spec_elems = static_cast<int64_t>(batch) * P->F_io * n_frames * 2;
  cudaError_t e;
  if ((e = cudaMalloc(&P->arena, P->arena_bytes)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc(arena)", __FILE__, __LINE__));
  if ((e = cudaMalloc(&P->params, params_numel * sizeof(float))) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc(params)", __FILE__, __LINE__));
  if ((e = cudaMalloc(&P->wpacked, P->wpacked_bytes)) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc(wpacked)", __FILE__, __LINE__));
  if ((e = cudaMalloc(&P->d_buf, spec_elems * sizeof(float))) != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc(d_buf)", __FILE__, __LINE__));
  if ((e = cudaMemset(P->params, 0, params_numel * sizeof(float))) != cudaSuccess) return fail(cuda_fail(e, "cudaMemset", __FILE__, __LINE__));

  // pass 2: identical walk, now recording launches against real addresses
  P->params_numel = 0;
  P->slots.clear();
  Builder b2{P};
  b2.dry = false;
  b2.arena.reset(P->arena_bytes);
  if (int rc = b2.build()) return fail(rc);
  if (P->params_numel != params_numel || b2.wp_off != P->wpacked_bytes) {
    set_error("plan: second pass diverged from the sizing pass");
    return fail(FDBM_EINVAL);
  }
  *out = P;
  return FDBM_OK;
}

extern "C" int fdbm_plan_destroy(fdbm_plan* plan) {
  if (!plan) return FDBM_OK;
  if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
  if (plan->capture_stream) cudaStreamDestroy(plan->capture_stream);
  fdbm_plan_release_sampler_state(plan);
  cudaFree(plan->arena); cudaFree(plan->params); cudaFree(plan->wpacked); cudaFree(plan->d_buf);
  delete plan;
  return FDBM_OK;
}

extern "C" int fdbm_plan_load_weights(fdbm_plan* plan, const fdbm_tensor_ref* tensors, int n_tensors, void* stream) {
  FDBM_REQUIRE(plan && tensors && n_tensors > 0, "fdbm_plan_load_weights: bad arguments");
  cudaStream_t s = as_stream(stream);
  for (int i = 0; i < n_tensors; ++i) {
    FDBM_REQUIRE(tensors[i].name && tensors[i].data, "fdbm_plan_load_weights: tensor %d has a null field", i);
    auto it = plan->slots.find(tensors[i].name);
    FDBM_REQUIRE(it != plan->slots.end(), "fdbm_plan_load_weights: unexpected tensor '%s'", tensors[i].name);
    FDBM_REQUIRE(it->second.numel == tensors[i].numel, "fdbm_plan_load_weights: '%s' has %lld elements, expected %lld",
                 tensors[i].name, (long long)tensors[i].numel, (long long)it->second.numel);
    FDBM_CUDA(cudaMemcpyAsync(plan->params + it->second.off, tensors[i].data, sizeof(float) * tensors[i].numel,
                              cudaMemcpyDeviceToDevice, s));
    it->second.loaded = true;
  }
  for (auto& kv : plan->slots)
    FDBM_REQUIRE(kv.second.loaded, "fdbm_plan_load_weights: tensor '%s' was never provided", kv.first.c_str());
  for (auto& f : plan->pack_ops) if (int rc = f(s)) return rc;
  plan->weights_ready = true;
  return FDBM_OK;
}

extern "C" int64_t fdbm_plan_device_bytes(const fdbm_plan* plan) {
  if (!plan) return 0;
  return plan->arena_bytes + plan->params_numel * 4 + plan->wpacked_bytes +
         static_cast<int64_t>(plan->B) * plan->F_io * plan->T * 8;
}

extern "C" int fdbm_plan_num_launches(const fdbm_plan* plan) { return plan ? plan->n_launches : 0; }

extern "C" int fdbm_ncsnpp_forward(fdbm_plan* plan, const float* x, const float* y, const float* t, float* out,
                                   void* stream) {
  FDBM_REQUIRE(plan && x && out, "fdbm_ncsnpp_forward: null pointer");
  if (!plan->weights_ready) { set_error("fdbm_ncsnpp_forward: weights not loaded"); return FDBM_ESTATE; }
  FDBM_REQUIRE(plan->arch.predictive || (y && t), "fdbm_ncsnpp_forward: y and t are required");
  plan->cur_x = x; plan->cur_y = y; plan->cur_t = t; plan->cur_t_stride = 1; plan->cur_out = out;
  return run_ops(plan, as_stream(stream));
}

extern "C" int fdbm_plan_profile_forward(fdbm_plan* plan, const float* x, const float* y, const float* t, float* out,
                                         float* ms, int* kinds, double* flops, int max_ops, void* stream) {
  FDBM_REQUIRE(plan && x && out && ms && kinds && flops, "fdbm_plan_profile_forward: null pointer");
  if (!plan->weights_ready) { set_error("fdbm_plan_profile_forward: weights not loaded"); return FDBM_ESTATE; }
  FDBM_REQUIRE(plan->arch.predictive || (y && t), "fdbm_plan_profile_forward: y and t are required");
  const int n = static_cast<int>(plan->ops.size());
  FDBM_REQUIRE(max_ops >= n, "fdbm_plan_profile_forward: need room for %d entries", n);
  cudaStream_t s = as_stream(stream);
  plan->cur_x = x; plan->cur_y = y; plan->cur_t = t; plan->cur_t_stride = 1; plan->cur_out = out;
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) FDBM_CUDA(cudaEventCreate(&e));
  int rc = FDBM_OK;
  FDBM_CUDA(cudaEventRecord(ev[0], s));
  for (int i = 0; i < n && rc == FDBM_OK; ++i) {
    rc = plan->ops[i](s);
    if (rc == FDBM_OK && cudaEventRecord(ev[i + 1], s) != cudaSuccess) rc = FDBM_ECUDA;
  }
  if (rc == FDBM_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "cudaStreamSynchronize", __FILE__, __LINE__);
  for (int i = 0; i < n && rc == FDBM_OK; ++i) {
    cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    kinds[i] = plan->op_kind[i];
    flops[i] = plan->op_flops[i];
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc == FDBM_OK ? n : rc;
}

namespace {

struct SamplerState {               // plan-owned staging so that one captured graph serves every call
  float* y = nullptr; float* x = nullptr; float* noise = nullptr; int64_t noise_elems = 0;
  float* times = nullptr; float* coef = nullptr; uint64_t* rng = nullptr; int cap_steps = 0;
  std::vector<float> h_times, h_coef;
};
std::map<fdbm_plan*, SamplerState>& sampler_states() { static std::map<fdbm_plan*, SamplerState> m; return m; }

}  // namespace

namespace fdbm { int launch_bridge_step_rng(float* x, const float* d, const float* third, const float* coef, int kind,
                                            const uint64_t* rng, uint64_t offset, int64_t n_complex, cudaStream_t s); }

extern "C" int fdbm_sampler_run(fdbm_plan* plan, const float* y, float* x, const float* times, const float* coef,
                                int n_steps, int kind, const float* noise, uint64_t seed, void* stream) {
  FDBM_REQUIRE(plan && y && x && times && coef && n_steps > 0, "fdbm_sampler_run: bad arguments");
  FDBM_REQUIRE(!plan->arch.predictive, "fdbm_sampler_run: predictive plans have no sampling loop");
  FDBM_REQUIRE(kind == FDBM_STEP_ODE || kind == FDBM_STEP_SDE, "fdbm_sampler_run: bad kind");
  if (!plan->weights_ready) { set_error("fdbm_sampler_run: weights not loaded"); return FDBM_ESTATE; }
  cudaStream_t s = as_stream(stream);
  const int64_t n_complex = static_cast<int64_t>(plan->B) * plan->F_io * plan->T;
  const size_t spec_bytes = static_cast<size_t>(n_complex) * 8;
  SamplerState& st = sampler_states()[plan];
  if (!st.y) {
    FDBM_CUDA(cudaMalloc(&st.y, spec_bytes));
    FDBM_CUDA(cudaMalloc(&st.x, spec_bytes));
    FDBM_CUDA(cudaMalloc(&st.rng, 2 * sizeof(uint64_t)));
  }
  if (st.cap_steps < n_steps) {
    cudaFree(st.times); cudaFree(st.coef);
    FDBM_CUDA(cudaMalloc(&st.times, sizeof(float) * n_steps));
    FDBM_CUDA(cudaMalloc(&st.coef, sizeof(float) * 3 * n_steps));
    st.cap_steps = n_steps;
    st.h_times.clear();
  }
  const bool want_noise = kind == FDBM_STEP_SDE && noise != nullptr;
  if (want_noise && st.noise_elems < 2 * n_complex * n_steps) {
    cudaFree(st.noise);
    FDBM_CUDA(cudaMalloc(&st.noise, spec_bytes * n_steps));
    st.noise_elems = 2 * n_complex * n_steps;
    plan->have_graph = false;
  }
  // stage the call's arguments into the plan-owned buffers (all stream-ordered)
  FDBM_CUDA(cudaMemcpyAsync(st.y, y, spec_bytes, cudaMemcpyDeviceToDevice, s));
  FDBM_CUDA(cudaMemcpyAsync(st.x, x, spec_bytes, cudaMemcpyDeviceToDevice, s));
  if (want_noise) FDBM_CUDA(cudaMemcpyAsync(st.noise, noise, spec_bytes * n_steps, cudaMemcpyDeviceToDevice, s));
  const std::vector<float> ht(times, times + n_steps), hc(coef, coef + 3 * n_steps);
  if (ht != st.h_times || hc != st.h_coef) {
    FDBM_CUDA(cudaMemcpyAsync(st.times, times, sizeof(float) * n_steps, cudaMemcpyHostToDevice, s));
    FDBM_CUDA(cudaMemcpyAsync(st.coef, coef, sizeof(float) * 3 * n_steps, cudaMemcpyHostToDevice, s));
    FDBM_CUDA(cudaStreamSynchronize(s));                 // the host arrays may be temporaries of the caller
    st.h_times = ht; st.h_coef = hc;
  }
  const uint64_t rng_host[2] = {seed, 0};
  FDBM_CUDA(cudaMemcpyAsync(st.rng, rng_host, sizeof(rng_host), cudaMemcpyHostToDevice, s));

  auto body = [&](cudaStream_t cs) -> int {
    for (int i = 0; i < n_steps; ++i) {
      plan->cur_x = st.x; plan->cur_y = st.y; plan->cur_t = st.times + i; plan->cur_t_stride = 0; plan->cur_out = plan->d_buf;
      if (int rc = run_ops(plan, cs)) return rc;
      const float* third = kind == FDBM_STEP_ODE ? st.y : (want_noise ? st.noise + 2 * n_complex * i : nullptr);
      if (int rc = launch_bridge_step_rng(st.x, plan->d_buf, third, st.coef + 3 * i, kind, st.rng,
                                          static_cast<uint64_t>(i) + 1, n_complex, cs))
        return rc;
    }
    return FDBM_OK;
  };
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  FDBM_CUDA(cudaStreamIsCapturing(s, &cap));
  if (cap != cudaStreamCaptureStatusNone) {                       // caller is capturing: record into its graph
    if (int rc = body(s)) return rc;
  } else {
    GraphKey key{{st.y, st.x, st.times, st.coef, want_noise ? st.noise : nullptr, nullptr}, n_steps, kind, 0};
    if (!(plan->have_graph && plan->graph_key == key)) {
      if (plan->graph_exec) { cudaGraphExecDestroy(plan->graph_exec); plan->graph_exec = nullptr; plan->have_graph = false; }
      cudaGraph_t graph = nullptr;
      if (!plan->capture_stream) FDBM_CUDA(cudaStreamCreateWithFlags(&plan->capture_stream, cudaStreamNonBlocking));
      cudaStream_t cs = plan->capture_stream;
      FDBM_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      const int rc = body(cs);
      const cudaError_t e = cudaStreamEndCapture(cs, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
      const cudaError_t e2 = cudaGraphInstantiate(&plan->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e2 != cudaSuccess) return cuda_fail(e2, "cudaGraphInstantiate", __FILE__, __LINE__);
      plan->graph_key = key; plan->have_graph = true;
    }
    FDBM_CUDA(cudaGraphLaunch(plan->graph_exec, s));
  }
  FDBM_CUDA(cudaMemcpyAsync(x, st.x, spec_bytes, cudaMemcpyDeviceToDevice, s));
  return FDBM_OK;
}

void fdbm_plan_release_sampler_state(fdbm_plan* plan) {
  auto& m = sampler_states();
  auto it = m.find(plan);
  if (it == m.end()) return;
  SamplerState& st = it->second;
  cudaFree(st.y); cudaFree(st.x); cudaFree(st.noise); cudaFree(st.times); cudaFree(st.coef); cudaFree(st.rng);
  m.erase(it);
}
