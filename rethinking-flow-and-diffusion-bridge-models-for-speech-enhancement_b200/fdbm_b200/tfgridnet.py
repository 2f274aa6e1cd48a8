"""TF-GridNet backbones behind the reference's BackboneRegistry API: `tfgridnet_5l32c100` (config.yaml's default),
`tfgridnet_4l32c80` and `tfgridnet_5l32c100_predictive` (config_predictive.yaml's default)
(fdbm/backbones/tfgridnet.py:126-229, 236-427, 430-484, 487-510; tfgridnet_predictive.py).

The module tree below consists of stock torch.nn layers used ONLY as parameter containers, arranged exactly like the
reference's, so `state_dict()` keys / shapes and the default initialisation are the reference's and its checkpoints load
unchanged; none of those layers is ever called.  `forward` runs the network as libfdbm_b200 kernels (csrc/tfgridnet.cu):
per block one fused pad + time-embedding-add + LayerNorm pass, the tcgen05 BiLSTM sweep along frequency (csrc/tfg_lstm_tc.cu,
ConvTranspose1d folded in), a pass that adds both directions + residual and takes the next LayerNorm, the sweep along time, the crop, and the
full-band self-attention (1x1 convs + PReLU-LayerNorm front, two batched tensor-core GEMMs, projection + PReLU + LayerNorm).
Inference only (outputs carry no autograd graph); there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, current_stream, on_device, ptr
from .registry import BackboneRegistry


class _Affine(nn.Module):
    """gamma / beta of the reference's LayerNormalization (tfgridnet.py:430-455)."""

    def __init__(self, shape):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(*shape))
        self.beta = nn.Parameter(torch.zeros(*shape))


class _HeadNorm(_Affine):
    """AllHeadPReLULayerNormalization4DC (tfgridnet.py:458-484): [1,H,E,1,1] affine + a per-head PReLU."""

    def __init__(self, H, E):
        super().__init__((1, H, E, 1, 1))
        self.act = nn.PReLU(num_parameters=H, init=0.25)


class _Fourier(nn.Module):
    def __init__(self, size, scale):
        super().__init__()
        self.W = nn.Parameter(torch.randn(size) * scale, requires_grad=False)      # layerspp.py:32-41


class _Block(nn.Module):
    """Parameters of GridNetV3Block (tfgridnet.py:241-317)."""

    def __init__(self, C_, ks, hs, H, n_head, E):
        super().__init__()
        if ks == hs:
            raise NotImplementedError("fdbm_b200 implements emb_ks != emb_hs (ConvTranspose1d de-embedding), the shipped configuration")
        for name in ("intra", "inter"):
            setattr(self, f"{name}_norm", nn.LayerNorm(C_))
            setattr(self, f"{name}_rnn", nn.LSTM(C_ * ks, H, 1, batch_first=True, bidirectional=True))
            setattr(self, f"{name}_linear", nn.ConvTranspose1d(H * 2, C_, ks, stride=hs))
        self.attn_conv_Q = nn.Conv2d(C_, n_head * E, 1)
        self.attn_norm_Q = _HeadNorm(n_head, E)
        self.attn_conv_K = nn.Conv2d(C_, n_head * E, 1)
        self.attn_norm_K = _HeadNorm(n_head, E)
        self.attn_conv_V = nn.Conv2d(C_, C_, 1)
        self.attn_norm_V = _HeadNorm(n_head, C_ // n_head)
        self.attn_concat_proj = nn.Sequential(nn.Conv2d(C_, C_, 1), nn.PReLU(), _Affine((1, C_, 1, 1)))


class _TFGridNetBase(nn.Module):
    predictive = False

    @staticmethod
    def add_argparse_args(parser):
        return parser

    def __init__(self, n_layers=6, emb_dim=48, lstm_hidden_units=200, n_srcs=1, n_imics=None, attn_n_head=4, attn_qk_output_channel=2,
                 emb_ks=4, emb_hs=1, activation="prelu", eps=1.0e-5, time_embedding_type="fourier", fourier_scale=16, **kwargs):
        super().__init__()
        n_imics = (1 if self.predictive else 2) if n_imics is None else n_imics
        if (emb_dim, emb_ks, emb_hs, attn_n_head, attn_qk_output_channel, n_srcs) != (32, 4, 1, 4, 2, 1) or activation != "prelu":
            raise NotImplementedError("fdbm_b200 implements the shipped TF-GridNet geometry: emb_dim 32, emb_ks 4, emb_hs 1, 4 heads, E = 2")
        if not 8 <= lstm_hidden_units <= 112:
            raise NotImplementedError("lstm_hidden_units must be in 8..112 (the gate matrix lives in the shared memory of a two-SM cluster)")
        if n_imics != (1 if self.predictive else 2) or (not self.predictive and time_embedding_type != "fourier"):
            raise NotImplementedError("unsupported n_imics / time_embedding_type")
        self.n_layers, self.emb_dim, self.hidden, self.eps = n_layers, emb_dim, lstm_hidden_units, eps
        self.n_srcs, self.n_imics = n_srcs, n_imics
        self.conv = nn.Sequential(nn.Conv2d(2 * n_imics, emb_dim, (3, 3), padding=(1, 1)), nn.GroupNorm(1, emb_dim, eps=eps))
        self.blocks = nn.ModuleList([_Block(emb_dim, emb_ks, emb_hs, lstm_hidden_units, attn_n_head, attn_qk_output_channel)
                                     for _ in range(n_layers)])
        self.deconv = nn.ConvTranspose2d(emb_dim, n_srcs * 2, (3, 3), padding=(1, 1))
        if not self.predictive:
            self.get_time_emb = _Fourier(emb_dim, fourier_scale)
            self.time_emb_fc = nn.Sequential(nn.Linear(2 * emb_dim, emb_dim * 4), nn.SiLU(), nn.Linear(emb_dim * 4, emb_dim * 4), nn.SiLU())
            self.time_emb_blocks = nn.ModuleList([nn.Linear(emb_dim * 4, emb_dim) for _ in range(n_layers)])
        self._packed = None          # (weight version, dict of packed / flattened device tensors)
        self._ws = {}                # (device index, B, T, F) -> workspace tensors

    # ---- weights -----------------------------------------------------------------------------------------------------
    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_weights(self):
        self._packed = None

    def train(self, mode: bool = True):
        self.invalidate_weights()
        return super().train(mode)

    def _pack(self, dev):
        ver = self._version()
        if self._packed is not None and self._packed[0] == ver:
            return self._packed[1]
        lib = _lib.load()
        sd = {k: v.detach() for k, v in self.named_parameters()}
        for k, v in sd.items():
            if v.device != dev or v.dtype != torch.float32:
                raise RuntimeError(f"parameter {k} must be fp32 on {dev}")
        nbytes = int(lib.fdbm_tfg_lstm_pack_bytes())
        P = {"lstm": {}, "keep": []}
        for b in range(self.n_layers):
            for name in ("intra", "inter"):
                pre = f"blocks.{b}.{name}_rnn."
                wl = sd[f"blocks.{b}.{name}_linear.weight"].contiguous()
                w = [sd[pre + f"{n}_l0{sfx}"].contiguous() for sfx in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
                buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                arr = (C.c_void_p * 8)(*[t.data_ptr() for t in w])
                check(lib.fdbm_tfg_lstm_pack(arr, ptr(wl), self.hidden, ptr(buf), current_stream()), "fdbm_tfg_lstm_pack")
                P["lstm"][(b, name)] = buf
                P["keep"] += w + [wl]
        if not self.predictive:
            P["wb"] = torch.stack([sd[f"time_emb_blocks.{b}.weight"] for b in range(self.n_layers)]).contiguous()
            P["bb"] = torch.stack([sd[f"time_emb_blocks.{b}.bias"] for b in range(self.n_layers)]).contiguous()
        P["attn"] = []
        for b in range(self.n_layers):
            p = f"blocks.{b}."
            names = [p + "attn_conv_Q.weight", p + "attn_conv_Q.bias", p + "attn_conv_K.weight", p + "attn_conv_K.bias",
                     p + "attn_conv_V.weight", p + "attn_conv_V.bias", p + "attn_norm_Q.act.weight", p + "attn_norm_K.act.weight",
                     p + "attn_norm_V.act.weight", p + "attn_norm_Q.gamma", p + "attn_norm_Q.beta", p + "attn_norm_K.gamma",
                     p + "attn_norm_K.beta", p + "attn_norm_V.gamma", p + "attn_norm_V.beta", p + "attn_concat_proj.0.weight",
                     p + "attn_concat_proj.0.bias", p + "attn_concat_proj.1.weight", p + "attn_concat_proj.2.gamma",
                     p + "attn_concat_proj.2.beta"]
            tensors = [sd[n].contiguous() for n in names]
            arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
            P["attn"].append((arr, tensors))
        P["sd"] = {k: v.contiguous() for k, v in sd.items()}
        self._packed = (ver, P)
        return P

    def _workspace(self, dev, B, T, F):
        key = (dev.index, B, T, F)
        ws = self._ws.get(key)
        if ws is None:
            lib = _lib.load()
            Tp, Qp = T + 6, F + 6
            f32 = dict(dtype=torch.float32, device=dev)
            f16 = dict(dtype=torch.float16, device=dev)
            n_y = B * Tp * Qp * 32
            ws = dict(h=torch.empty(B, T, F, 32, **f32), h2=torch.empty(B, T, F, 32, **f32), xp=torch.empty(B, Tp, Qp, 32, **f32),
                      xp2=torch.empty(B, Tp, Qp, 32, **f32), xn=torch.empty(B, Tp, Qp, 32, **f16), yf=torch.empty(n_y, **f16),
                      yb=torch.empty(n_y, **f16), sums=torch.empty(2 * B, dtype=torch.float64, device=dev),
                      emb=torch.empty(self.n_layers, B, 32, **f32),
                      attn=torch.empty(int(lib.fdbm_tfg_attention_workspace_bytes(B, T, F)) + 256, dtype=torch.uint8, device=dev))
            self._ws.clear()                                                    # one geometry at a time (the buffers are GBs at B = 32)
            self._ws[key] = ws
        return ws

    @staticmethod
    def _sweep(lib, P, b, name, ws, n_seq, L, st):
        """One bidirectional sweep over `n_seq` contiguous sequences [n_seq][L + 3][32] of ws["xn"] -> ws["yf"], ws["yb"] (same shape)."""
        check(lib.fdbm_tfg_lstm_sweep(ptr(ws["xn"]), n_seq, L, ptr(P["lstm"][(b, name)]), ptr(ws["yf"]), ptr(ws["yb"]), st), "fdbm_tfg_lstm_sweep")

    # ---- forward -----------------------------------------------------------------------------------------------------
    def _run(self, x, y, t):
        if not (x.is_cuda and x.dtype == torch.complex64 and x.dim() == 4 and x.shape[1] == 1):
            raise RuntimeError(f"expected a complex64 CUDA tensor [B,1,F,T], got {tuple(x.shape)} {x.dtype} on {x.device}")
        lib = _lib.load()
        dev = x.device
        x = x.contiguous()
        B, _, F, T = x.shape
        P = self._pack(dev)
        sd = P["sd"]
        ws = self._workspace(dev, B, T, F)
        st = current_stream()
        Tp, Qp, eps = T + 6, F + 6, float(self.eps)
        if not self.predictive:
            y = y.contiguous()
            t = t.to(device=dev, dtype=torch.float32).contiguous()
            check(lib.fdbm_tfg_time_embedding(ptr(t), 1, ptr(sd["get_time_emb.W"]), ptr(sd["time_emb_fc.0.weight"]), ptr(sd["time_emb_fc.0.bias"]),
                                              ptr(sd["time_emb_fc.2.weight"]), ptr(sd["time_emb_fc.2.bias"]), ptr(P["wb"]), ptr(P["bb"]),
                                              self.n_layers, B, ptr(ws["emb"]), st), "fdbm_tfg_time_embedding")
        check(lib.fdbm_tfg_input(ptr(x), None if self.predictive else ptr(y), ptr(sd["conv.0.weight"]), ptr(sd["conv.0.bias"]),
                                 ptr(sd["conv.1.weight"]), ptr(sd["conv.1.bias"]), B, T, F, 2 if self.predictive else 4, eps, ptr(ws["sums"]),
                                 ptr(ws["h"]), st), "fdbm_tfg_input")
        h, h2 = ws["h"], ws["h2"]
        for b in range(self.n_layers):
            p = f"blocks.{b}."
            emb = None if self.predictive else ws["emb"][b]
            check(lib.fdbm_tfg_pad_add_norm(ptr(h), ptr(emb), ptr(sd[p + "intra_norm.weight"]), ptr(sd[p + "intra_norm.bias"]), B, T, F, eps,
                                            ptr(ws["xp"]), ptr(ws["xn"]), st), "fdbm_tfg_pad_add_norm")
            # intra: sequences along frequency, one per (utterance, padded frame): xn [B, T', Q', 32]
            self._sweep(lib, P, b, "intra", ws, B * Tp, Qp - 3, st)
            check(lib.fdbm_tfg_sweep_post(ptr(ws["yf"]), ptr(ws["yb"]), ptr(sd[p + "intra_linear.bias"]), ptr(ws["xp"]), B, T, F, 0,
                                          ptr(sd[p + "inter_norm.weight"]), ptr(sd[p + "inter_norm.bias"]), eps, ptr(ws["xp2"]), ptr(ws["xn"]),
                                          None, 1, st), "fdbm_tfg_sweep_post")
            # inter: sequences along time, one per (utterance, padded bin): xn transposed to [B, Q', T', 32]
            self._sweep(lib, P, b, "inter", ws, B * Qp, Tp - 3, st)
            check(lib.fdbm_tfg_sweep_post(ptr(ws["yf"]), ptr(ws["yb"]), ptr(sd[p + "inter_linear.bias"]), ptr(ws["xp2"]), B, T, F, 1, None, None,
                                          eps, None, None, ptr(h2), 0, st), "fdbm_tfg_sweep_post")
            base = ws["attn"].data_ptr()
            aligned = (base + 255) // 256 * 256
            check(lib.fdbm_tfg_attention(ptr(h2), P["attn"][b][0], B, T, F, eps, aligned, ptr(h), st), "fdbm_tfg_attention")
        out = torch.empty(B, 1, F, T, dtype=torch.complex64, device=dev)
        check(lib.fdbm_tfg_output(ptr(h), ptr(sd["deconv.weight"]), ptr(sd["deconv.bias"]), B, T, F, ptr(out), st), "fdbm_tfg_output")
        return out


class _TFGridNetBridge(_TFGridNetBase):
    @on_device
    def forward(self, x, y, t):
        with torch.no_grad():
            return self._run(x, y, t)


@BackboneRegistry.register("tfgridnet_5l32c100")
class TFGridNet_5l32c100(_TFGridNetBridge):
    """fdbm/backbones/tfgridnet.py:487-497."""

    def __init__(self, **kwargs):
        super().__init__(n_layers=5, emb_dim=32, lstm_hidden_units=100, **kwargs)


@BackboneRegistry.register("tfgridnet_4l32c80")
class TFGridNet_4l32c80(_TFGridNetBridge):
    """fdbm/backbones/tfgridnet.py:500-510."""

    def __init__(self, **kwargs):
        super().__init__(n_layers=4, emb_dim=32, lstm_hidden_units=80, **kwargs)


@BackboneRegistry.register("tfgridnet_5l32c100_predictive")
class TFGridNet_5l32c100_predictive(_TFGridNetBase):
    """fdbm/backbones/tfgridnet_predictive.py:449-459."""
    predictive = True

    def __init__(self, **kwargs):
        super().__init__(n_layers=5, emb_dim=32, lstm_hidden_units=100, **kwargs)

    @on_device
    def forward(self, y):
        with torch.no_grad():
            return self._run(y, None, None)
