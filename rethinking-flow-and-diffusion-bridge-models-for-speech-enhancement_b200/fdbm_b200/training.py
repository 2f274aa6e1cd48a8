"""Training step of the bridge model (fdbm/model.py:258-275 `_step` / `sample_prior`, :187-218 hybrid loss,
:101 Adam, :129-132 EMA, train.py:161 gradient_clip_val=3.0) on the CUDA backbone.

    x_t = a_t x + b_t y + sigma_t z            Bridge.probability_path (bridge.py:40-43)
    D   = dnn(x_t, y, t)                       fdbm_ncsnpp_forward on a TRAINING plan (all activations kept)
    L   = 70 MSE(|X|^0.3) + 30 MSE(X/|X|^0.7) - mean log10 SI-SNR(istft)     (data_prediction_hybrid)
    dL/dparams                                 fdbm_ncsnpp_backward: tcgen05 dgrad / wgrad + GroupNorm / FIR / attention backward
    all-reduce(grads) / world                  torch.distributed (NCCL over NVLink), one flat 262 MB buffer
    Adam + clip + EMA                          fdbm_plan_optimizer_step on the flat buffers, weights re-packed

The loss head and its gradient are one library call (fdbm_hybrid_loss: spectral terms, fused de-compress + iSTFT,
SI-SNR, iSTFT adjoint through the fused STFT kernel).  Nothing in the step goes through torch autograd.
Gradients of activations are 16-bit GEMM operands, so dL/dD is multiplied by `loss_scale` first (divided out of
the fp32 parameter gradients).  A non-finite gradient norm skips the update on the device: parameters, Adam
moments, the EMA and the update counter (Adam bias correction, torch_ema's num_updates) stay untouched and a skip
counter is incremented; `update_loss_scale()` reads that counter every `scale_check_interval` steps, halves the
loss scale after a skip and doubles it after `scale_growth_interval` clean steps (torch.cuda.amp.GradScaler's rule).

EMA (fdbm/model.py:56,129-160): the shadow parameters follow torch_ema.ExponentialMovingAverage.update with
use_num_updates=True, decay = min(ema_decay, (1 + n) / (10 + n)); `ema_state_dict()` / `load_ema_state_dict()` use
torch_ema's state_dict layout (the 'ema' entry of the reference's checkpoints), `swap_in_ema()` / `restore_params()`
are store + copy_to / restore of model.py:146-160.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check, current_stream, ptr


from .backbones import _DevBuf

def allreduce_gradients_(flat_grads: torch.Tensor, group=None) -> int:
    """DDP gradient exchange (fdbm/model.py trains under Lightning DDP): ONE all-reduce(sum) of the flat gradient buffer
    (NCCL over NVLink on the GPU box, gloo in the CPU test).  Returns the world size; the mean's 1/world is folded into
    the optimiser's `grad_div`, so the buffer keeps the SUM."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return world


def _on_plan_device(fn):
    """Run a TrainStep method with the plan's device current (the library works on the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        dev = self.device
        if dev.index == torch.cuda.current_device():
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapper


class TrainStep:
    """One data-parallel optimisation step; `dnn` is the fdbm_b200 NCSNpp_v2 whose parameters are trained."""

    def __init__(self, dnn, bridge, data_module, batch: int, n_frames: int = 256, lr: float = 1e-4, ema_decay: float = 0.999,
                 clip_norm: float = 3.0, t_eps: float = 0.03, loss_scale: float = 4096.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 ema_warmup: bool = True, dynamic_loss_scale: bool = True, scale_check_interval: int = 50,
                 scale_growth_interval: int = 2000, loss_type: str = "data_prediction_hybrid", l1_weight: float = 0.001,
                 pesq_weight: float = 0.0, sr: int = 16000):
        # BridgeModel.__init__(loss_type, l1_weight, pesq_weight, sr) (model.py:41): the four loss types of `_loss` (model.py:162-254)
        if loss_type not in ("data_prediction_hybrid", "data_prediction", "data_prediction_mel", "data_prediction_melphase"):
            raise ValueError("Invalid loss type: {}".format(loss_type))
        if pesq_weight != 0.0:
            raise NotImplementedError("pesq_weight > 0 (torch_pesq loss term) is not implemented")
        self.loss_type, self.l1_weight, self.sr = loss_type, float(l1_weight), int(sr)
        self._mel_tables = None
        self.dnn, self.bridge, self.dm = dnn, bridge, data_module
        self.batch, self.n_frames = batch, n_frames
        self.lr, self.ema_decay, self.clip_norm, self.t_eps, self.loss_scale = lr, ema_decay, clip_norm, t_eps, loss_scale
        self.betas, self.eps = betas, eps
        self.ema_warmup = ema_warmup                    # torch_ema's use_num_updates=True (what the reference constructs)
        self.dynamic_loss_scale = dynamic_loss_scale
        self.scale_check_interval, self.scale_growth_interval = scale_check_interval, scale_growth_interval
        self.steps_issued = 0                           # optimizer_step() calls, applied or skipped
        self._seen_skipped, self._clean_since = 0, 0
        self.lib = _lib.load()
        dev = next(dnn.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs the backbone on a CUDA device (there is no CPU path)")
        self.device = dev
        with torch.cuda.device(dev):
            handle = C.c_void_p()
            arch = dnn._arch()
            check(self.lib.fdbm_plan_create_train(C.byref(arch), batch, n_frames, C.byref(handle)), "fdbm_plan_create_train")
            self.plan = handle
            self._load_from_module()
            p, g, e, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
            check(self.lib.fdbm_plan_buffers(self.plan, C.byref(p), C.byref(g), C.byref(e), C.byref(n)), "fdbm_plan_buffers")
            self.numel = n.value
            self.flat_params = torch.as_tensor(_DevBuf(p.value, n.value), device=dev)
            self.flat_grads = torch.as_tensor(_DevBuf(g.value, n.value), device=dev)
            self.flat_ema = torch.as_tensor(_DevBuf(e.value, n.value), device=dev)
            self.broadcast_parameters()
            check(self.lib.fdbm_plan_reset_optimizer(self.plan, current_stream()), "fdbm_plan_reset_optimizer")   # EMA <- params

    def broadcast_parameters(self, src: int = 0):
        """DDP start-up semantics: every replica starts from rank `src`'s parameters."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.flat_params, src)
            check(self.lib.fdbm_plan_repack_weights(self.plan, current_stream()), "fdbm_plan_repack_weights")

    def _load_from_module(self):
        named = list(self.dnn.named_parameters())
        refs = (_lib.TensorRef * len(named))()
        keep = []
        for i, (n, p) in enumerate(named):
            d = p.detach().contiguous()
            keep.append(d)
            refs[i].name, refs[i].data, refs[i].numel = n.encode(), d.data_ptr(), d.numel()
        check(self.lib.fdbm_plan_load_weights(self.plan, refs, len(named), current_stream()), "fdbm_plan_load_weights")
        torch.cuda.current_stream().synchronize()

    def close(self):
        if self.plan is not None:
            self.lib.fdbm_plan_destroy(self.plan)
            self.plan = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pieces -------------------------------------------------------------------------------------------------
    def _slot(self, name: str):
        off, n = C.c_int64(), C.c_int64()
        check(self.lib.fdbm_plan_param_info(self.plan, name.encode(), C.byref(off), C.byref(n)), "fdbm_plan_param_info")
        return off.value, n.value

    def grads(self) -> Dict[str, torch.Tensor]:
        """Parameter gradients of the last backward, by reference-style name (views into the flat buffer)."""
        out = {}
        for name, p in self.dnn.named_parameters():
            off, n = self._slot(name)
            out[name] = self.flat_grads[off:off + n].view(p.shape)
        return out

    def params(self) -> Dict[str, torch.Tensor]:
        out = {}
        for name, p in self.dnn.named_parameters():
            off, n = self._slot(name)
            out[name] = self.flat_params[off:off + n].view(p.shape)
        return out

    def sync_module(self):
        """Copy the trained parameters back into the nn.Module (state_dict / checkpoint compatibility)."""
        with torch.no_grad():
            for name, p in self.dnn.named_parameters():
                off, n = self._slot(name)
                p.copy_(self.flat_params[off:off + n].view(p.shape))
        self.dnn.invalidate_weights()

    # ---- EMA (fdbm/model.py:56, 129-160; torch_ema.ExponentialMovingAverage) ------------------------------------------
    def ema_params(self) -> Dict[str, torch.Tensor]:
        """The EMA shadow of every parameter, by reference-style name (views into the plan's flat EMA buffer)."""
        out = {}
        for name, p in self.dnn.named_parameters():
            off, n = self._slot(name)
            out[name] = self.flat_ema[off:off + n].view(p.shape)
        return out

    def optimizer_state(self) -> Dict[str, float]:
        """{'applied', 'skipped', 'grad_norm'} read from the device (synchronises the stream)."""
        buf = (C.c_double * 4)()
        check(self.lib.fdbm_plan_optimizer_state(self.plan, buf, current_stream()), "fdbm_plan_optimizer_state")
        return {"applied": int(buf[0]), "skipped": int(buf[1]), "grad_norm": float(buf[2])}

    def ema_state_dict(self) -> dict:
        """torch_ema.ExponentialMovingAverage.state_dict() layout: what on_save_checkpoint stores under 'ema'
        (model.py:143-144).  shadow_params lists the requires_grad parameters in `parameters()` order."""
        st = self.optimizer_state()
        shadow = [self.ema_params()[n].clone() for n, p in self.dnn.named_parameters() if p.requires_grad]
        return {"decay": self.ema_decay, "num_updates": st["applied"] if self.ema_warmup else None,
                "shadow_params": shadow, "collected_params": None}

    def load_ema_state_dict(self, state: dict):
        """model.py:134-141 on_load_checkpoint."""
        self.ema_decay = float(state["decay"])
        names = [n for n, p in self.dnn.named_parameters() if p.requires_grad]
        if len(names) != len(state["shadow_params"]):
            raise RuntimeError(f"EMA state has {len(state['shadow_params'])} shadow tensors, the backbone {len(names)} trainable ones")
        dst = self.ema_params()
        with torch.no_grad():
            for n, s in zip(names, state["shadow_params"]):
                dst[n].copy_(s.to(self.device, torch.float32))
        if state.get("num_updates") is not None:
            st = self.optimizer_state()
            check(self.lib.fdbm_plan_set_optimizer_state(self.plan, float(state["num_updates"]), float(st["skipped"]), current_stream()),
                  "fdbm_plan_set_optimizer_state")

    def swap_in_ema(self):
        """eval(): ema.store(parameters) + ema.copy_to(parameters) (model.py:149-152); packed weights rebuilt."""
        check(self.lib.fdbm_plan_swap_ema(self.plan, 1, current_stream()), "fdbm_plan_swap_ema")

    def restore_params(self):
        """train(): ema.restore(parameters) (model.py:154-156)."""
        check(self.lib.fdbm_plan_swap_ema(self.plan, 0, current_stream()), "fdbm_plan_swap_ema")

    def update_loss_scale(self):
        """GradScaler's rule on the device-side skip counter: halve after a skipped step, double after
        `scale_growth_interval` consecutive applied steps.  Costs one 32-byte read-back (a stream sync)."""
        st = self.optimizer_state()
        if st["skipped"] > self._seen_skipped:
            self.loss_scale = max(self.loss_scale * 0.5 ** (st["skipped"] - self._seen_skipped), 1.0)
            self._seen_skipped, self._clean_since = st["skipped"], st["applied"]
        elif st["applied"] - self._clean_since >= self.scale_growth_interval:
            self.loss_scale = min(self.loss_scale * 2.0, 65536.0)
            self._clean_since = st["applied"]
        return self.loss_scale

    def sample_prior(self, x, y, t=None, z=None):
        """fdbm/model.py:267-275."""
        if z is None:
            z = torch.randn_like(x)
        if t is None:
            t = torch.rand(x.shape[0], device=x.device) * (self.bridge.T - self.t_eps) + self.t_eps
        mean, std = self.bridge.probability_path(x, y, t)
        return t, mean, z, mean + std[:, None, None, None] * z

    def forward(self, x_t, y, t):
        out = torch.empty_like(x_t)
        check(self.lib.fdbm_ncsnpp_forward(self.plan, ptr(x_t.contiguous()), ptr(y.contiguous()), ptr(t.float().contiguous()), ptr(out),
                                           current_stream()), "fdbm_ncsnpp_forward")
        return out

    def backward(self, g_out: torch.Tensor, accumulate: bool = False):
        g = (g_out * self.loss_scale).contiguous()
        check(self.lib.fdbm_ncsnpp_backward(self.plan, ptr(torch.view_as_real(g)), self.loss_scale, int(accumulate), current_stream()),
              "fdbm_ncsnpp_backward")
        self._keep = g

    def loss_and_backward(self, x, y, t=None, z=None) -> torch.Tensor:
        """`_step` (model.py:258-265) + backward.  x, y: complex64 [B,1,257,T] clean / noisy compressed spectrograms."""
        t, _, _, x_t = self.sample_prior(x, y, t, z)
        self._x_t, self._y, self._t = x_t.contiguous(), y.contiguous(), t.float().contiguous()
        D = self.forward(self._x_t, self._y, self._t)
        loss, g = self.loss_and_grad(D, x)
        check(self.lib.fdbm_ncsnpp_backward(self.plan, ptr(torch.view_as_real(g)), self.loss_scale, 0, current_stream()),
              "fdbm_ncsnpp_backward")
        self._keep = g
        return loss

    def loss_and_grad(self, D: torch.Tensor, x: torch.Tensor):
        """fdbm/model.py:163-218 (`_loss`, the configured head) and its gradient w.r.t. D, already multiplied by the loss scale: one library call."""
        from ._lib import FDBM_TRANSFORM
        B, _, Fb, T = D.shape
        dm = self.dm
        mel = self.loss_type in ("data_prediction_mel", "data_prediction_melphase")
        need = (self.lib.fdbm_mel_loss_workspace_bytes if mel else self.lib.fdbm_hybrid_loss_workspace_bytes)(B, T, dm.n_fft, dm.hop_length)
        if getattr(self, "_loss_ws", None) is None or self._loss_ws.numel() < need:
            self._loss_ws = torch.empty(need, dtype=torch.uint8, device=D.device)
        loss = torch.empty((), device=D.device)
        g = torch.empty_like(D)
        D, x = D.contiguous(), x.contiguous()
        if mel:
            if self._mel_tables is None:              # windows, twiddles, mel filterbanks of the seven resolutions: built once
                self._mel_tables = torch.empty(self.lib.fdbm_mel_tables_bytes(), dtype=torch.uint8, device=D.device)
                check(self.lib.fdbm_mel_tables_init(ptr(self._mel_tables), self.sr, current_stream()), "fdbm_mel_tables_init")
            check(self.lib.fdbm_mel_loss(ptr(torch.view_as_real(D)), ptr(torch.view_as_real(x)), B, T, ptr(dm._get_window(D)), dm.n_fft,
                                         dm.hop_length, FDBM_TRANSFORM[dm.transform_type], float(dm.spec_factor), float(dm.spec_abs_exponent),
                                         int(self.loss_type == "data_prediction_melphase"), self.loss_scale, ptr(self._mel_tables),
                                         ptr(self._loss_ws), ptr(loss), ptr(torch.view_as_real(g)), current_stream()), "fdbm_mel_loss")
            return loss, g
        if self.loss_type == "data_prediction":
            check(self.lib.fdbm_data_prediction_loss(ptr(torch.view_as_real(D)), ptr(torch.view_as_real(x)), B, T, ptr(dm._get_window(D)), dm.n_fft,
                                                     dm.hop_length, FDBM_TRANSFORM[dm.transform_type], float(dm.spec_factor),
                                                     float(dm.spec_abs_exponent), self.l1_weight, self.loss_scale, ptr(self._loss_ws), ptr(loss),
                                                     ptr(torch.view_as_real(g)), current_stream()), "fdbm_data_prediction_loss")
            return loss, g
        check(self.lib.fdbm_hybrid_loss(ptr(torch.view_as_real(D)), ptr(torch.view_as_real(x)), B, T, ptr(dm._get_window(D)), dm.n_fft,
                                        dm.hop_length, FDBM_TRANSFORM[dm.transform_type], float(dm.spec_factor), float(dm.spec_abs_exponent),
                                        self.loss_scale, ptr(self._loss_ws), ptr(loss), ptr(torch.view_as_real(g)), current_stream()),
              "fdbm_hybrid_loss")
        return loss, g

    def optimizer_step(self):
        """DDP gradient all-reduce (mean) over the flat buffer, then Adam + clip + EMA and the weight re-pack."""
        world = allreduce_gradients_(self.flat_grads)            # sum over ranks; the mean's 1/world goes into grad_div
        self.steps_issued += 1
        # step = 0: Adam's bias correction and the EMA warm-up use the device-side count of APPLIED steps
        check(self.lib.fdbm_plan_optimizer_step(self.plan, float(world), self.clip_norm, self.lr, self.betas[0], self.betas[1], self.eps,
                                                0, self.ema_decay, int(self.ema_warmup), current_stream()), "fdbm_plan_optimizer_step")
        if self.dynamic_loss_scale and self.steps_issued % self.scale_check_interval == 0:
            self.update_loss_scale()

    def training_step(self, x, y) -> torch.Tensor:
        loss = self.loss_and_backward(x, y)
        self.optimizer_step()
        return loss


for _name in ("broadcast_parameters", "close", "optimizer_state", "load_ema_state_dict", "swap_in_ema", "restore_params",
              "update_loss_scale", "forward", "backward", "loss_and_backward", "loss_and_grad", "optimizer_step", "training_step"):
    setattr(TrainStep, _name, _on_plan_device(getattr(TrainStep, _name)))
