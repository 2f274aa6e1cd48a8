"""Enhancement glue: the call sequence of BridgeModel.enhance (fdbm/model.py:391-406),
infer_single.py:80-99 and infer_folder.py:100-121, batched, plus the utterance sharding of
infer_folder.py:149-152 for one-process-per-GPU runs.

The reference enhances one file at a time (B=1).  Here `enhance_batch` runs a whole batch of
equal-length utterances through   peak-normalise -> fused STFT+compress+pad -> N-step sampler
(CUDA graph) -> fused decompress+iSTFT -> rescale   with no host synchronisation in between.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from .bridge import Bridge
from .data_module import SpecsDataModule, padded_frames
from .registry import BackboneRegistry


def si_sdr(s: np.ndarray, s_hat: np.ndarray) -> float:
    """fdbm/util/other.py:64-68."""
    alpha = np.dot(s_hat, s) / np.linalg.norm(s) ** 2
    return float(10 * np.log10(np.linalg.norm(alpha * s) ** 2 / np.linalg.norm(alpha * s - s_hat) ** 2))


def split_list(lst: Sequence, n: int) -> List[list]:
    """infer_folder.py:149-152: n contiguous chunks, the first len % n one element longer."""
    k, m = divmod(len(lst), n)
    return [list(lst[i * k + min(i, m):(i + 1) * k + min(i + 1, m)]) for i in range(n)]


def shard_for_rank(items: Sequence, rank: int, world_size: int) -> list:
    """The contiguous chunk of `items` rank `rank` enhances (weights replicated, no collective)."""
    return split_list(items, world_size)[rank]


def gather_waveforms(local: torch.Tensor, counts: Sequence[int], group=None) -> Optional[torch.Tensor]:
    """Final gather of the enhanced waveforms [n_local, Ts] of every rank onto every rank, in utterance
    order (NCCL over NVLink on the GPU box, gloo in the CPU tests).  `counts[r]` = utterances of rank r."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_max = max(counts)
    pad = torch.zeros(n_max, local.shape[1], dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


class EnhancementModel(nn.Module):
    """Inference-side equivalent of the reference's BridgeModel (fdbm/model.py:25-411): holds `dnn`,
    `bridge`, `data_module` under the same attribute names and offers `forward`, `enhance`, `to_audio`,
    `_stft`, `_istft`, `_forward_transform`, `_backward_transform`."""

    def __init__(self, backbone="ncsnpp_v2", bridge="sb", backbone_kwargs=None, bridge_kwargs=None,
                 data_module_kwargs=None, pad_mode=None):
        super().__init__()
        self.backbone = backbone
        self.dnn = BackboneRegistry.get_by_name(backbone)(**(backbone_kwargs or {}))
        self.bridge = Bridge(bridge, **(bridge_kwargs or {})) if bridge is not None else None
        dm = dict(n_fft=512, hop_length=256, num_frames=256, window="sqrthann")
        dm.update(data_module_kwargs or {})
        self.data_module = SpecsDataModule(**dm)
        # infer_single.py:64-69: reflection padding for exactly 'ncsnpp_v2', zeros otherwise
        self.pad_mode = pad_mode or ("reflection" if backbone == "ncsnpp_v2" else "zero_pad")

    # reference-named helpers (fdbm/model.py:356-389)
    def forward(self, x_t, y, t):
        return self.dnn(x_t, y, t)

    def to_audio(self, spec, length=None):
        return self.data_module.to_audio(spec, length)

    def _forward_transform(self, spec):
        return self.data_module.spec_fwd(spec)

    def _backward_transform(self, spec):
        return self.data_module.spec_back(spec)

    def _stft(self, sig):
        return self.data_module.stft(sig)

    def _istft(self, spec, length=None):
        return self.data_module.istft(spec, length)

    def _sample(self, Y):
        return self.bridge.sampler(self, Y)

    @torch.no_grad()
    def enhance_batch(self, y: torch.Tensor, clip_rescale: Optional[float] = None) -> torch.Tensor:
        """y fp32 CUDA [B, Ts] (un-normalised) -> enhanced fp32 [B, Ts].  Per utterance exactly the
        arithmetic of infer_single.py:80-99; `clip_rescale` (0.5 there, 0.95 in infer_folder.py:120)
        applies the optional peak rescale, None skips it as model.py:391-406 does."""
        if self.data_module.normalize == "noisy":
            norm = y.abs().amax(dim=1, keepdim=True)
        elif self.data_module.normalize == "std":
            norm = y.std(dim=1, keepdim=True)
        else:
            norm = torch.ones(y.shape[0], 1, device=y.device)
        T_orig = y.shape[1]
        Y = self.data_module.stft_compress(y / norm, pad_mode=self.pad_mode)
        sample = self._sample(Y)
        x_hat = self.data_module.to_audio(sample[:, 0], T_orig) * norm
        if clip_rescale is not None:
            peak = x_hat.abs().amax(dim=1, keepdim=True)
            x_hat = torch.where(peak > 1.0, x_hat / peak * clip_rescale, x_hat)
        return x_hat

    @torch.no_grad()
    def enhance(self, y: torch.Tensor, **sampler_kwargs) -> np.ndarray:
        """fdbm/model.py:391-406: y [1, Ts] (any device) -> numpy [Ts]."""
        dev = next(self.dnn.parameters()).device
        return self.enhance_batch(y.to(dev, torch.float32)).squeeze(0).cpu().numpy()

    @torch.no_grad()
    def enhance_many(self, waves: torch.Tensor, micro_batch: int = 8) -> torch.Tensor:
        """Equal-length utterances [N, Ts] in micro-batches (bounded activation memory); the last
        micro-batch is padded so that a single plan / CUDA graph serves every launch."""
        out = torch.empty_like(waves)
        N = waves.shape[0]
        for i in range(0, N, micro_batch):
            chunk = waves[i:i + micro_batch]
            n = chunk.shape[0]
            if n < micro_batch:
                chunk = torch.cat([chunk, chunk[-1:].expand(micro_batch - n, -1)], dim=0)
            out[i:i + n] = self.enhance_batch(chunk)[:n]
        return out


class PredictiveEnhancementModel(EnhancementModel):
    """PredictiveModel (fdbm/model.py:414-439): one backbone pass, no sampling loop."""

    def __init__(self, backbone="ncsnpp_v2_predictive", **kw):
        kw.setdefault("bridge", None)
        super().__init__(backbone=backbone, **kw)
        self.pad_mode = kw.get("pad_mode") or "zero_pad"          # model.py:431 calls pad_spec(Y) with its default

    def forward(self, y):
        return self.dnn(y)

    def _sample(self, Y):
        return self.dnn(Y)
