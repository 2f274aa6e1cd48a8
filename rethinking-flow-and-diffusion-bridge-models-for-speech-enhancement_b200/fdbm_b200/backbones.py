"""NCSN++ backbones behind the reference's BackboneRegistry API.

`BackboneRegistry.get_by_name("ncsnpp_v2")(**kwargs)` returns an nn.Module whose parameters have
the reference's names and shapes (`all_modules.<i>.<Layer>.weight`, `output_layer.*`; 647 tensors for
the default model) so reference checkpoints load unchanged, and whose `forward(x, y, t)` /
`forward(y)` (predictive) runs the whole U-Net in libfdbm_b200 (fdbm/backbones/ncsnpp_v2.py:36-401,
ncsnpp_v2_predictive.py:36-362).  The module itself holds no arithmetic: it owns the parameters and
a cache of C-side plans keyed by (device, batch, n_frames); packed weights are refreshed whenever a
parameter's (storage, version) pair changes (optimizer step, load_state_dict) and on every
train()/eval() switch (the reference swaps EMA weights there through `param.data`, which bumps no
version counter -- fdbm/model.py:146-160).

Training drop-in: in train() mode with autograd enabled, `forward(x, y, t)` of the bridge backbone
returns a tensor with a grad_fn; `loss.backward()` runs fdbm_ncsnpp_backward on a training plan and
delivers the parameter gradients to autograd (so `param.grad` is filled exactly as for the
reference module and BridgeModel.training_step / Lightning / torch.optim work unchanged,
fdbm/model.py:258-282).  In eval() mode, or under no_grad, the output carries no graph.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import Arch, TensorRef, check, current_stream, on_device, ptr
from .registry import BackboneRegistry


def _module_table(nf, ch_mult, num_res_blocks, attn_resolutions, image_size, predictive):
    """(kind, cin, cout, up, down) per entry of the reference's `all_modules` (ncsnpp_v2.py:95-239)."""
    mods = []
    C_io = 2 if predictive else 4
    if not predictive:
        mods += [("fourier", 0, nf, False, False), ("linear", 2 * nf, 4 * nf, False, False),
                 ("linear", 4 * nf, 4 * nf, False, False)]
    mods.append(("conv3", C_io, nf, False, False))
    hs_c, in_ch, L = [nf], nf, len(ch_mult)
    for lvl in range(L):
        for _ in range(num_res_blocks):
            out_ch = nf * ch_mult[lvl]
            mods.append(("res", in_ch, out_ch, False, False))
            in_ch = out_ch
            if image_size // (2 ** lvl) in attn_resolutions:
                mods.append(("attn", in_ch, in_ch, False, False))
            hs_c.append(in_ch)
        if lvl != L - 1:
            mods.append(("res", in_ch, in_ch, False, True))
            mods.append(("combine", C_io, in_ch, False, False))
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    mods += [("res", in_ch, in_ch, False, False), ("attn", in_ch, in_ch, False, False), ("res", in_ch, in_ch, False, False)]
    for lvl in reversed(range(L)):
        for _ in range(num_res_blocks + 1):
            out_ch = nf * ch_mult[lvl]
            mods.append(("res", in_ch + hs_c.pop(), out_ch, False, False))
            in_ch = out_ch
        if image_size // (2 ** lvl) in attn_resolutions:
            mods.append(("attn", in_ch, in_ch, False, False))
        mods.append(("gn", in_ch, in_ch, False, False))
        mods.append(("conv3", in_ch, C_io, False, False))
        if lvl != 0:
            mods.append(("res", in_ch, in_ch, True, False))
    assert not hs_c
    return mods


def _fan_avg_uniform(shape, scale=1.0):
    """default_init (layers.py:54-91): variance_scaling(scale, 'fan_avg', 'uniform'), scale 0 -> 1e-10."""
    scale = 1e-10 if scale == 0 else scale
    rf = np.prod(shape) / shape[0] / shape[1]
    var = scale / ((shape[1] * rf + shape[0] * rf) / 2)
    return (torch.rand(*shape) * 2.0 - 1.0) * math.sqrt(3 * var)


@torch.no_grad()
def sensitise_(net: nn.Module, seed: int = 0) -> nn.Module:
    """Re-randomise a backbone in place so that every kernel influences the output: the reference's
    default initialisation zeroes 56 weight tensors (init_scale=0 -> 1e-10), which makes the network
    output a constant ('two things a fresh reader must know').  Weights get the same
    fan_avg-uniform law with scale 1, biases N(0, 0.02), GroupNorm affine 1 + N(0, 0.1) / N(0, 0.1).
    Used for benchmarks and smoke tests with random-init weights; never needed for trained checkpoints."""
    g = torch.Generator().manual_seed(seed)
    for name, p in net.named_parameters():
        if name.endswith(".W") and p.dim() == 1:
            continue                                                     # Fourier features keep their draw
        if p.dim() >= 2:
            shape = tuple(p.shape)
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            var = 1.0 / ((shape[1] * rf + shape[0] * rf) / 2)
            new = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3 * var)
        elif "GroupNorm" in name or (name.count(".") == 2 and name.endswith(("weight", "bias")) and p.dim() == 1
                                     and _is_gn_entry(net, name)):
            new = torch.randn(p.shape, generator=g) * 0.1 + (1.0 if name.endswith("weight") else 0.0)
        else:
            new = torch.randn(p.shape, generator=g) * 0.02
        p.copy_(new.to(p.device, p.dtype))
    return net


def _pad_channel_blocks(name, w, real=96, block=128):
    """Zero-padded copy of a parameter of an nf = 96 network in the channel layout of the nf = 128 plan that runs it
    (fdbm_arch.channel_block_real): along every dimension whose size is a multiple of 96, real channel r moves to
    block * (r // 96) + r % 96 and the 32 channels that end each block are zero.  Concatenated inputs (skip connections) need no
    special case: every tensor of the network is a whole number of 96-channel blocks.  Zero weights, biases, FiLM rows and
    GroupNorm gains keep the padding channels exactly zero through every layer, so the result is that of the nf = 96 network;
    the two places where a channel count enters the arithmetic are the GroupNorm group structure (handled in gn_finalize) and the
    attention logit scale C^-1/2 (layerspp.py:83), folded into the query projection NIN_0 here."""
    if name.endswith("NIN_0.W") or name.endswith("NIN_0.b"):
        w = w * (block / real) ** 0.5
    for dim, size in enumerate(w.shape):
        if size % real or size == 0:
            continue
        nb = size // real
        shape = list(w.shape)
        shape[dim:dim + 1] = [nb, real]
        v = w.reshape(shape)
        shape[dim + 1] = block
        out = w.new_zeros(shape)
        out.narrow(dim + 1, 0, real).copy_(v)
        shape[dim:dim + 2] = [nb * block]
        w = out.reshape(shape)
    return w.contiguous()


def _is_gn_entry(net, name):
    parts = name.split(".")
    if parts[0] != "all_modules":
        return False
    return net._table[int(parts[1])][0] == "gn"


class _Holder(nn.Module):
    """Parameter container; sub-holders are created on demand so dotted names match the reference."""

    def add(self, dotted, tensor, requires_grad=True):
        head, _, rest = dotted.partition(".")
        if rest:
            if not hasattr(self, head):
                self.add_module(head, _Holder())
            getattr(self, head).add(rest, tensor, requires_grad)
        else:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=requires_grad))


class _NCSNppBase(nn.Module):
    predictive = False

    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--nf", type=int, default=128)
        parser.add_argument("--ch_mult", type=int, nargs="+", default=[1, 1, 2, 2, 2, 2, 2])
        parser.add_argument("--num_res_blocks", type=int, default=2)
        parser.add_argument("--attn_resolutions", type=int, nargs="+", default=[16])
        return parser

    def __init__(self, nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,),
                 init_scale=0.0, fourier_scale=16, image_size=256, **unused_kwargs):
        super().__init__()
        for key, want in (("nonlinearity", "swish"), ("resblock_type", "biggan"), ("progressive", "output_skip"),
                          ("progressive_input", "input_skip"), ("progressive_combine", "sum"),
                          ("embedding_type", "fourier"), ("fir", True), ("skip_rescale", True),
                          ("resamp_with_conv", True), ("fir_kernel", [1, 3, 3, 1]), ("dropout", 0.0)):
            got = unused_kwargs.get(key, want)
            if isinstance(want, list):
                got = list(got)
            if got != want:
                raise NotImplementedError(f"fdbm_b200 implements the reference's default NCSN++ variant only ({key}={want!r})")
        attn = [r for r in attn_resolutions if r > 0]
        if len(attn) > 1:
            raise NotImplementedError("at most one attention resolution is supported")
        self.nf, self.ch_mult, self.num_res_blocks = nf, tuple(ch_mult), num_res_blocks
        self.attn_resolutions, self.image_size = tuple(attn), image_size
        self.num_resolutions = len(ch_mult)
        self._table = _module_table(nf, self.ch_mult, num_res_blocks, self.attn_resolutions, image_size, self.predictive)
        C_io = 2 if self.predictive else 4

        # parameters, with the reference's initialisation (ncsnpp_v2.py:93-239, layerspp.py, layers.py:88-124)
        self.output_layer = nn.Conv2d(C_io, 2, 1)
        holders = []
        for kind, cin, cout, up, down in self._table:
            h = _Holder()
            if kind == "fourier":
                h.add("W", torch.randn(cout) * fourier_scale, requires_grad=False)
            elif kind == "linear":
                h.add("weight", _fan_avg_uniform((cout, cin))); h.add("bias", torch.zeros(cout))
            elif kind == "conv3":
                scale = init_scale if cout == C_io else 1.0        # pyramid output convs use init_scale
                h.add("weight", _fan_avg_uniform((cout, cin, 3, 3), scale)); h.add("bias", torch.zeros(cout))
            elif kind == "gn":
                h.add("weight", torch.ones(cin)); h.add("bias", torch.zeros(cin))
            elif kind == "combine":
                h.add("Conv_0.weight", _fan_avg_uniform((cout, cin, 1, 1))); h.add("Conv_0.bias", torch.zeros(cout))
            elif kind == "attn":
                h.add("GroupNorm_0.weight", torch.ones(cin)); h.add("GroupNorm_0.bias", torch.zeros(cin))
                for i in range(4):
                    h.add(f"NIN_{i}.W", _fan_avg_uniform((cin, cin), init_scale if i == 3 else 0.1))
                    h.add(f"NIN_{i}.b", torch.zeros(cin))
            elif kind == "res":
                h.add("GroupNorm_0.weight", torch.ones(cin)); h.add("GroupNorm_0.bias", torch.zeros(cin))
                h.add("Conv_0.weight", _fan_avg_uniform((cout, cin, 3, 3))); h.add("Conv_0.bias", torch.zeros(cout))
                if not self.predictive:
                    h.add("Dense_0.weight", _fan_avg_uniform((cout, 4 * nf))); h.add("Dense_0.bias", torch.zeros(cout))
                h.add("GroupNorm_1.weight", torch.ones(cout)); h.add("GroupNorm_1.bias", torch.zeros(cout))
                h.add("Conv_1.weight", _fan_avg_uniform((cout, cout, 3, 3), init_scale)); h.add("Conv_1.bias", torch.zeros(cout))
                if cin != cout or up or down:
                    h.add("Conv_2.weight", _fan_avg_uniform((cout, cin, 1, 1))); h.add("Conv_2.bias", torch.zeros(cout))
            holders.append(h)
        self.all_modules = nn.ModuleList(holders)
        self._plans = {}          # (device index, batch, n_frames) -> [plan handle, weight version]; insertion order = LRU order
        self._train_plans = {}    # same key -> [training plan handle, weight version] (autograd drop-in, train() mode)
        self._param_cache = None  # [(name, parameter)], rebuilt after .to()/.cuda()
        self.grad_loss_scale = 1024.0   # activation gradients are 16-bit GEMM operands: dL/dD is scaled by this in backward
        # device memory all cached plans may hold together (each plan owns an arena + a packed copy of the weights); folders
        # with many distinct padded lengths would otherwise grow without bound -- least recently used plans are destroyed
        self.max_plan_bytes = 96 << 30

    # ---- plan / weight management --------------------------------------------------------------
    def _arch(self) -> Arch:
        a = Arch()
        a.nf, a.n_levels, a.num_res_blocks = self.nf, len(self.ch_mult), self.num_res_blocks
        if self.nf == 96:                       # ncsnpp_v2_5M / _37M: run as nf = 128 with 96 real channels per 128-channel block
            a.nf, a.channel_block_real = 128, 96
        for i, m in enumerate(self.ch_mult):
            a.ch_mult[i] = m
        a.attn_resolution = self.attn_resolutions[0] if self.attn_resolutions else 0
        a.predictive, a.image_size = int(self.predictive), self.image_size
        return a

    def _named(self):
        if self._param_cache is None:
            self._param_cache = list(self.named_parameters())
        return self._param_cache

    def _apply(self, fn, *a, **k):
        self._param_cache = None                                    # .to() / .cuda() may replace the Parameter objects
        return super()._apply(fn, *a, **k)

    def _weight_version(self):
        """Identity of the current weight state: (storage address, in-place version) of every parameter.  Catches
        optimizer steps, load_state_dict and re-allocation; writes through `param.data` bump no counter and need
        `invalidate_weights()` (train()/eval() call it)."""
        return tuple((p.data_ptr(), p._version) for _, p in self._named())

    def _load_into(self, handle, device):
        named = self._named()
        refs = (TensorRef * len(named))()
        keep = []
        for i, (n, p) in enumerate(named):
            if p.device != device or p.dtype != torch.float32:
                raise RuntimeError(f"parameter {n} must be fp32 on {device}")
            d = p.detach().contiguous()
            if self.nf == 96:
                d = _pad_channel_blocks(n, d)
            keep.append(d)
            refs[i].name, refs[i].data, refs[i].numel = n.encode(), d.data_ptr(), d.numel()
        check(_lib.load().fdbm_plan_load_weights(handle, refs, len(named), current_stream()), "fdbm_plan_load_weights")

    def _plan(self, device, batch, n_frames, train=False):
        """Plan for (device, batch, n_frames), created on first use, with the module's current weights loaded.  Must be
        called with `device` current (the public entry points are wrapped in `on_device`)."""
        lib = _lib.load()
        key = (device.index, batch, n_frames)
        cache = self._train_plans if train else self._plans
        entry = cache.pop(key, None)
        if entry is None:
            handle = C.c_void_p()
            arch = self._arch()
            create = lib.fdbm_plan_create_train if train else lib.fdbm_plan_create
            check(create(C.byref(arch), batch, n_frames, C.byref(handle)), "fdbm_plan_create")
            entry = [handle, None]
            everything = list(self._plans.items()) + list(self._train_plans.items())
            used = sum(int(lib.fdbm_plan_device_bytes(h)) for _, (h, _) in everything) + int(lib.fdbm_plan_device_bytes(handle))
            while used > self.max_plan_bytes and (self._plans or self._train_plans):
                victims = self._plans if self._plans else self._train_plans
                old_key = next(iter(victims))
                old_handle, _ = victims.pop(old_key)
                used -= int(lib.fdbm_plan_device_bytes(old_handle))
                torch.cuda.synchronize(device)                      # the evicted plan's arena may still be in flight
                lib.fdbm_plan_destroy(old_handle)
        cache[key] = entry                                          # (re-)insert as most recently used
        version = self._weight_version()
        if entry[1] != version:
            self._load_into(entry[0], device)
            entry[1] = version
        return entry[0]

    def invalidate_weights(self):
        """Force a re-pack on the next call.  Needed after writes through `param.data` (e.g. torch_ema's
        in-place swap, fdbm/model.py:146-160), which do not bump the parameter's version counter."""
        for entry in list(self._plans.values()) + list(self._train_plans.values()):
            entry[1] = None

    def train(self, mode: bool = True):
        self.invalidate_weights()          # the reference swaps EMA weights in train()/eval()
        return super().train(mode)

    def release_plans(self):
        lib = _lib.load()
        for handle, _ in list(self._plans.values()) + list(self._train_plans.values()):
            lib.fdbm_plan_destroy(handle)                           # switches to the plan's own device internally
        self._plans.clear()
        self._train_plans.clear()

    def __del__(self):
        try:
            self.release_plans()
        except Exception:
            pass

    def plan_info(self, batch, n_frames, device=None):
        device = device or next(self.parameters()).device
        with torch.cuda.device(device):
            h = self._plan(device, batch, n_frames)
        lib = _lib.load()
        return {"device_bytes": lib.fdbm_plan_device_bytes(h), "launches": lib.fdbm_plan_num_launches(h)}

    @on_device
    def profile_forward(self, x, y=None, t=None, max_ops=4096):
        """One forward with a CUDA event pair around every kernel launch (bench.py's roofline leg).
        Returns [(milliseconds, kind, algorithmic_flops), ...] in launch order; kinds are FDBM_OP_*."""
        x = self._check_spec(x, "x")
        plan = self._plan(x.device, x.shape[0], x.shape[3])
        out = torch.empty_like(x)
        ms = (C.c_float * max_ops)(); kinds = (C.c_int * max_ops)(); flops = (C.c_double * max_ops)()
        y = None if y is None else self._check_spec(y, "y")
        t = None if t is None else t.to(device=x.device, dtype=torch.float32).contiguous()
        n = _lib.load().fdbm_plan_profile_forward(plan, ptr(x), ptr(y), ptr(t), ptr(out), ms, kinds, flops, max_ops,
                                                  current_stream())
        if n < 0:
            check(n, "fdbm_plan_profile_forward")
        return [(ms[i], kinds[i], flops[i]) for i in range(n)]

    def _check_spec(self, s, name):
        if not (s.is_cuda and s.dtype == torch.complex64 and s.dim() == 4 and s.shape[1] == 1
                and s.shape[2] == self.image_size + 1):
            raise RuntimeError(f"{name} must be a complex64 CUDA tensor [B,1,{self.image_size + 1},T], got "
                               f"{tuple(s.shape)} {s.dtype} on {s.device}")
        return s.contiguous()

    # ---- sampler hook used by fdbm_b200.bridge.Bridge -----------------------------------------
    @on_device
    def run_sampler(self, y, x, times, table, kind, noise, seed):
        """In-place N-step sampler on x (fdbm/bridge.py:66-113) as one CUDA graph."""
        if self.predictive:
            raise RuntimeError("predictive backbones have no sampling loop")
        y = self._check_spec(y, "y")
        if not x.is_contiguous():
            raise RuntimeError("x must be contiguous")
        if times.is_cuda or table.is_cuda or times.dtype != torch.float32 or table.dtype != torch.float32:
            raise RuntimeError("times / table must be fp32 host tensors")
        times, table = times.contiguous(), table.contiguous()
        plan = self._plan(y.device, y.shape[0], y.shape[3])
        check(_lib.load().fdbm_sampler_run(plan, ptr(y), ptr(x), ptr(times), ptr(table), times.numel(), kind, ptr(noise),
                                           seed, current_stream()), "fdbm_sampler_run")
        return x


class _DevBuf:
    """A plan-owned flat fp32 device buffer as a zero-copy torch tensor (`torch.as_tensor(_DevBuf(ptr, n), device=...)`)."""

    def __init__(self, pointer: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (pointer, False), "version": 2}


class _BackboneFunction(torch.autograd.Function):
    """D = dnn(x_t, y, t) with parameter gradients from the library's own backward pass (tcgen05 dgrad / wgrad, GroupNorm /
    FIR / attention backward): the autograd seam that makes `BridgeModel.training_step` + `loss.backward()`
    (fdbm/model.py:258-282) work on the CUDA backbone.  Gradients w.r.t. x_t, y and t are not produced (the reference's
    `_step` never needs them: x_t is sampled, not learned)."""

    @staticmethod
    def forward(ctx, net, names, x, y, t, *params):
        lib = _lib.load()
        plan = net._plan(x.device, x.shape[0], x.shape[3], train=True)
        out = torch.empty_like(x)
        check(lib.fdbm_ncsnpp_forward(plan, ptr(x), ptr(y), ptr(t), ptr(out), current_stream()), "fdbm_ncsnpp_forward")
        ctx.net, ctx.names, ctx.plan, ctx.device = net, names, plan, x.device
        ctx.keep = (x, y, t)                                         # the plan reads them again in backward
        ctx.shapes = [p.shape for p in params]
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, g_out):
        if g_out is None:
            return (None,) * (5 + len(ctx.shapes))
        lib = _lib.load()
        net, scale = ctx.net, float(ctx.net.grad_loss_scale)
        with torch.cuda.device(ctx.device):
            g = (g_out.to(torch.complex64) * scale).contiguous()
            check(lib.fdbm_ncsnpp_backward(ctx.plan, ptr(torch.view_as_real(g)), scale, 0, current_stream()), "fdbm_ncsnpp_backward")
            pp_, gp_, ep_, n_ = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
            check(lib.fdbm_plan_buffers(ctx.plan, C.byref(pp_), C.byref(gp_), C.byref(ep_), C.byref(n_)), "fdbm_plan_buffers")
            flat = torch.as_tensor(_DevBuf(gp_.value, n_.value), device=ctx.device)
            grads = []
            off, num = C.c_int64(), C.c_int64()
            for name, shape in zip(ctx.names, ctx.shapes):
                check(lib.fdbm_plan_param_info(ctx.plan, name.encode(), C.byref(off), C.byref(num)), "fdbm_plan_param_info")
                grads.append(flat[off.value:off.value + num.value].view(shape).clone())
        return (None, None, None, None, None, *grads)


@BackboneRegistry.register("ncsnpp_v2")
class NCSNpp_v2(_NCSNppBase):
    """fdbm/backbones/ncsnpp_v2.py:36-401."""

    @on_device
    def forward(self, x, y, t):
        x, y = self._check_spec(x, "x"), self._check_spec(y, "y")
        t = t.to(device=x.device, dtype=torch.float32).contiguous()
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for _, p in self._named()):
            names = [n for n, p in self._named() if p.requires_grad]
            return _BackboneFunction.apply(self, tuple(names), x, y, t, *[p for _, p in self._named() if p.requires_grad])
        out = torch.empty_like(x)
        plan = self._plan(x.device, x.shape[0], x.shape[3])
        check(_lib.load().fdbm_ncsnpp_forward(plan, ptr(x), ptr(y), ptr(t), ptr(out), current_stream()), "fdbm_ncsnpp_forward")
        return out


@BackboneRegistry.register("ncsnpp_v2_16M")
class NCSNpp_v2_16M(NCSNpp_v2):
    """fdbm/backbones/ncsnpp_v2.py:418-433: nf = 64 (64 / 128 channels), no attention besides the bottleneck's."""

    def __init__(self, **kwargs):
        for k in ("nf", "ch_mult", "num_res_blocks", "attn_resolutions"):
            kwargs.pop(k, None)
        super().__init__(nf=64, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=[0], **kwargs)

    @staticmethod
    def add_argparse_args(parser):
        return parser


class _Nf96Variant(NCSNpp_v2):
    """ncsnpp_v2_5M / ncsnpp_v2_37M (fdbm/backbones/ncsnpp_v2.py:404-415, 436-448) have nf = 96: 96- and 192-channel tensors, which the
    tensor-core convolution (64-channel K-blocks, 64 / 128-channel N tiles) does not tile.  They run on an nf = 128 plan with
    zero-padded weights (`_pad_channel_blocks`): (128 / 96)^2 = 1.78 x the algorithmic FLOPs, same results.  Inference only: the
    training plan (flat gradient buffers in the parameters' own layout) is not built for the padded layout."""
    _variant = {}

    def __init__(self, **kwargs):
        for k in ("nf", "ch_mult", "num_res_blocks", "attn_resolutions"):
            kwargs.pop(k, None)
        super().__init__(**self._variant, **kwargs)

    def _plan(self, device, batch, n_frames, train=False):
        if train:
            raise NotImplementedError(f"{type(self).__name__}: training plans are built for nf = 64 / 128 (the nf = 96 variants run "
                                      "inference on a zero-padded nf = 128 plan)")
        return super()._plan(device, batch, n_frames)

    @staticmethod
    def add_argparse_args(parser):
        return parser


@BackboneRegistry.register("ncsnpp_v2_5M")
class NCSNpp_v2_5M(_Nf96Variant):
    _variant = dict(nf=96, ch_mult=(1, 1, 1, 1), num_res_blocks=1, attn_resolutions=[0])


@BackboneRegistry.register("ncsnpp_v2_37M")
class NCSNpp_v2_37M(_Nf96Variant):
    _variant = dict(nf=96, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2, attn_resolutions=[16])


@BackboneRegistry.register("ncsnpp_v2_predictive")
class NCSNpp_v2_predictive(_NCSNppBase):
    """fdbm/backbones/ncsnpp_v2_predictive.py:36-362."""
    predictive = True

    @on_device
    def forward(self, x):
        x = self._check_spec(x, "x")
        out = torch.empty_like(x)
        plan = self._plan(x.device, x.shape[0], x.shape[3])
        check(_lib.load().fdbm_ncsnpp_forward(plan, ptr(x), None, None, ptr(out), current_stream()), "fdbm_ncsnpp_forward")
        return out
