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


def length_buckets(n_samples: Sequence[int], hop: int, micro_batch: int) -> List[tuple]:
    """Batches for utterances of different lengths: [(padded_frames, [indices ...]), ...].  Utterances share a batch only
    if pad_spec (other.py:76-90) gives them the same padded frame count -- GroupNorm and attention see the whole padded
    image, so any other grouping would change results -- and every batch has at most `micro_batch` members, shortest
    first."""
    buckets = {}
    for i, n in enumerate(n_samples):
        buckets.setdefault(padded_frames(1 + int(n) // hop), []).append(i)
    out = []
    for T_pad, idx in sorted(buckets.items()):
        idx.sort(key=lambda i: (n_samples[i], i))
        for j0 in range(0, len(idx), micro_batch):
            out.append((T_pad, idx[j0:j0 + micro_batch]))
    return out


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
    def enhance_batch(self, y: torch.Tensor, clip_rescale: Optional[float] = None, pad_mode: Optional[str] = None) -> torch.Tensor:
        """y fp32 CUDA [B, Ts] (un-normalised) -> enhanced fp32 [B, Ts].  Per utterance exactly the
        arithmetic of infer_single.py:80-99; `clip_rescale` (0.5 there, 0.95 in infer_folder.py:120)
        applies the optional peak rescale, None skips it as model.py:391-406 does.  `pad_mode` defaults to the
        infer scripts' choice (`self.pad_mode`: reflection for exactly 'ncsnpp_v2', zeros otherwise)."""
        pad_mode = pad_mode or self.pad_mode
        dm = self.data_module
        y = y if y.stride(1) == 1 else y.contiguous()
        # norm_factor per utterance; the division / re-multiplication live inside the STFT / iSTFT kernels
        if dm.normalize == "noisy":
            norm = dm.wave_absmax(y)
        elif dm.normalize == "std":
            norm = y.std(dim=1).contiguous()
        else:
            norm = None
        Y = dm.stft_compress(y, pad_mode=pad_mode, norm=norm)
        sample = self._sample(Y)
        x_hat, peak = dm.to_audio_ex(sample[:, 0], y.shape[1], norm=norm, want_peak=clip_rescale is not None)
        if clip_rescale is not None:
            dm.clip_rescale_(x_hat, peak, clip_rescale)
        return x_hat

    @torch.no_grad()
    def enhance(self, y: torch.Tensor, pad_mode: str = "zero_pad", **sampler_kwargs) -> np.ndarray:
        """fdbm/model.py:391-406: y [1, Ts] (any device) -> numpy [Ts].  Like `BridgeModel.enhance` (the validation-time
        path) this pads with `pad_spec`'s default, zeros; the infer scripts' reflection padding for 'ncsnpp_v2'
        (infer_single.py:64-69) is `pad_mode="reflection"` here and the default of `enhance_batch` / `enhance_list`."""
        dev = next(self.dnn.parameters()).device
        return self.enhance_batch(y.to(dev, torch.float32), pad_mode=pad_mode).squeeze(0).cpu().numpy()

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


    # ------------------------------------------------------------------ callers' edge: lists of files / waveforms
    def _staging(self, slot: int, rows: int, cols: int):
        """Persistent pinned host staging (input, output) of one pipeline slot, grown on demand; allocating pinned memory per
        bucket (cudaHostAlloc is a synchronising driver call) cost more than the copies themselves."""
        ring = self.__dict__.setdefault("_pinned_ring", {})
        need = rows * cols
        bufs = ring.get(slot)
        if bufs is None or bufs[0].numel() < need:
            bufs = (torch.empty(need, dtype=torch.float32).pin_memory(), torch.empty(need, dtype=torch.float32).pin_memory())
            ring[slot] = bufs
        return bufs[0][:need].view(rows, cols), bufs[1][:need].view(rows, cols)

    @torch.no_grad()
    def enhance_list(self, waves: Sequence, micro_batch: int = 32, clip_rescale: Optional[float] = 0.95,
                     lengths: Optional[Sequence[int]] = None, on_done=None) -> List[np.ndarray]:
        """The per-file loop of infer_folder.py:91-121 for waveforms of DIFFERENT lengths, batched: utterances are
        bucketed by their padded frame count (pad_spec rounds to 64 frames; GroupNorm and attention see the whole padded
        image, so only utterances with the same padded length may share a batch without changing any result), each
        bucket runs in micro-batches through peak-normalise -> variable-length STFT -> sampler -> variable-length iSTFT
        -> rescale -> clip rule (0.95, infer_folder.py:119-120).  Returns the enhanced waveforms in input order.

        Two-slot software pipeline: while the GPU works on micro-batch k the host packs micro-batch k + 1 into the other
        pinned staging slot and unpacks the results of micro-batch k - 1; the only host waits are on per-slot events.

        Streaming form (used by `enhance_files`): with `lengths` given, an entry of `waves` may be a `concurrent.futures.Future`
        that resolves to the waveform (it is awaited when its micro-batch is packed), and `on_done(i, x)` is called as soon as
        utterance i has left the GPU, so decoding, the GPU and encoding overlap."""
        if self.data_module.normalize != "noisy":
            raise NotImplementedError("enhance_list implements the default normalize='noisy'")
        dev = next(self.dnn.parameters()).device
        hop = self.data_module.hop_length
        dm = self.data_module
        def as_array(w):
            if hasattr(w, "result"):
                w = w.result()
            if torch.is_tensor(w):
                return w.detach().to("cpu", torch.float32).reshape(-1).contiguous().numpy()
            return np.ascontiguousarray(np.asarray(w, dtype=np.float32).reshape(-1))

        if lengths is None:
            ws = [as_array(w) for w in waves]
            n_samples = [w.shape[0] for w in ws]
        else:
            ws = list(waves)                                       # resolved lazily, micro-batch by micro-batch
            n_samples = [int(n) for n in lengths]
            if len(n_samples) != len(ws):
                raise ValueError("enhance_list: `lengths` must have one entry per waveform")
        for i, n in enumerate(n_samples):
            if n <= dm.n_fft // 2:
                raise RuntimeError(f"utterance {i} is shorter than n_fft/2 samples")
        out: List[Optional[np.ndarray]] = [None] * len(ws)
        pending = []                                               # (event, host_out view, group indices, lengths)

        def drain(keep: int):
            while len(pending) > keep:
                ev, host_out, grp, lens = pending.pop(0)
                ev.synchronize()
                for r, i in enumerate(grp):
                    out[i] = host_out[r, :lens[r]].numpy().copy()
                    if on_done is not None:
                        on_done(i, out[i])

        with torch.cuda.device(dev):
            for k, (T_pad, grp) in enumerate(length_buckets(n_samples, hop, micro_batch)):
                n = len(grp)
                lens = [n_samples[i] for i in grp]
                # a partial batch is padded (last utterance repeated) to the next multiple of 8: at most four plans / graphs
                # per padded length instead of one per batch size, at most 7 wasted slots
                mb = micro_batch if n == micro_batch else min(micro_batch, -(-n // 8) * 8)
                lens_full = lens + [lens[-1]] * (mb - n)
                max_len, min_len = max(lens_full), min(lens_full)
                drain(1)                                           # slot k % 2 was used by micro-batch k - 2: its results are out
                host_in, host_out = self._staging(k % 2, mb, max_len)
                hin = host_in.numpy()
                for r in range(mb):
                    i = grp[min(r, n - 1)]
                    w = ws[i] = as_array(ws[i])
                    if w.shape[0] != n_samples[i]:
                        raise RuntimeError(f"utterance {i}: {w.shape[0]} samples, `lengths` said {n_samples[i]}")
                    hin[r, :w.shape[0]] = w
                    hin[r, w.shape[0]:] = 0.0
                for i in grp:
                    ws[i] = None                                    # the staged copy is the only one needed from here on
                y = host_in.to(dev, non_blocking=True)
                lengths = torch.tensor(lens_full, dtype=torch.int32).to(dev, non_blocking=True)
                norm = dm.wave_absmax(y, lengths)                    # rows are zero-padded beyond their length
                Y = dm.stft_compress_var(y, lengths, min_len, max_len, pad_mode=self.pad_mode, n_frames_out=T_pad, norm=norm)
                sample = self._sample(Y)
                x_hat, peak = dm.to_audio_ex(sample[:, 0], max_len, lengths=lengths, norm=norm, want_peak=clip_rescale is not None)
                if clip_rescale is not None:
                    dm.clip_rescale_(x_hat, peak, clip_rescale, lengths)
                host_out.copy_(x_hat, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending.append((ev, host_out, grp, lens))
            drain(0)
        return out  # type: ignore[return-value]

    def enhance_files(self, paths: Sequence[str], out_paths: Optional[Sequence[str]] = None, micro_batch: int = 32,
                      target_sr: int = 16000, io_threads: int = 8) -> List[np.ndarray]:
        """infer_folder.py:91-146 for a list of mono WAV files (16-bit / 32-bit PCM or float): decode, enhance in
        length buckets, optionally write 16-bit PCM WAVs.  Decoding and encoding run on a thread pool (file I/O and the numpy
        conversions release the GIL).  Resampling (librosa in the reference) is not part of the hot path: files at another
        sample rate are rejected."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        from scipy.io import wavfile

        def decode(p):
            sr, x = wavfile.read(p)
            if sr != target_sr:
                raise RuntimeError(f"{p}: sample rate {sr} != {target_sr} (resample before calling enhance_files)")
            if x.ndim != 1:
                raise RuntimeError(f"{p}: expected a mono file")
            if x.dtype == np.int16:
                return x.astype(np.float32) / 32768.0                      # torchaudio.load's normalisation
            if x.dtype == np.int32:
                return x.astype(np.float32) / 2147483648.0
            return x.astype(np.float32)

        def encode(args):
            p, x = args
            os.makedirs(os.path.dirname(os.path.abspath(p)), exist_ok=True)
            wavfile.write(p, target_sr, np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16))

        def n_samples_of(p):                                       # header only (memory-mapped): the buckets need the lengths first
            sr, x = wavfile.read(p, mmap=True)
            return int(x.shape[0])

        hop = self.data_module.hop_length
        with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool:
            lens = list(pool.map(n_samples_of, paths))
            # decode in the order the buckets consume the files; encode every utterance as soon as it has left the GPU
            order = [i for _, grp in length_buckets(lens, hop, micro_batch) for i in grp]
            fut = {i: pool.submit(decode, paths[i]) for i in order}
            writes = []
            on_done = None if out_paths is None else (lambda i, x: writes.append(pool.submit(encode, (out_paths[i], x))))
            enhanced = self.enhance_list([fut[i] for i in range(len(paths))], micro_batch=micro_batch, lengths=lens, on_done=on_done)
            for w in writes:
                w.result()
        return enhanced


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
