"""Spectral front / back end behind the reference's SpecsDataModule transform API.

Mirrors the transform half of fdbm/data_module.py (`stft`, `istft`, `spec_fwd`, `spec_back`,
attributes `n_fft, hop_length, num_frames, spec_factor, spec_abs_exponent, transform_type, normalize`,
:112-229) and `pad_spec` of fdbm/util/other.py:76-90.  All arithmetic runs in libfdbm_b200's CUDA
kernels; the dataset / DataLoader half of the reference class is out of scope.

Besides the reference's unfused call sequence this class offers the fused forms the hot path uses:
`stft_compress` (= pad_spec(spec_fwd(stft(y)))) and `to_audio` (= istft(spec_back(spec))).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import FDBM_PAD, FDBM_TRANSFORM, check, current_stream, on_device, ptr


def get_window(window_type: str, window_length: int) -> torch.Tensor:
    """fdbm/data_module.py:13-19."""
    if window_type == "sqrthann":
        return torch.sqrt(torch.hann_window(window_length, periodic=True))
    if window_type == "hann":
        return torch.hann_window(window_length, periodic=True)
    raise NotImplementedError(f"Window type {window_type} not implemented!")


def padded_frames(n_frames: int, multiple: int = 64) -> int:
    """Frame count after pad_spec (other.py:77-81)."""
    return n_frames + (multiple - n_frames % multiple) % multiple


def _as_cfloat(spec: torch.Tensor) -> torch.Tensor:
    if spec.dtype != torch.complex64 or not spec.is_cuda:
        raise RuntimeError("fdbm_b200 expects complex64 CUDA spectrograms")
    return spec.contiguous()


@on_device
def pad_spec(Y: torch.Tensor, mode: str = "zero_pad") -> torch.Tensor:
    """fdbm/util/other.py:76-90 on the GPU: right-pad the frame axis of [B,1,F,T] to a multiple of 64."""
    if mode not in FDBM_PAD:
        raise NotImplementedError("This function hasn't been implemented yet.")
    Y = _as_cfloat(Y)
    T = Y.size(3)
    T_out = padded_frames(T)
    if T_out == T:
        return Y
    out = torch.empty(*Y.shape[:3], T_out, dtype=Y.dtype, device=Y.device)
    rows = Y.numel() // T
    check(_lib.load().fdbm_pad_spec(ptr(Y), rows, T, FDBM_PAD[mode], T_out, ptr(out), current_stream()), "fdbm_pad_spec")
    return out


class SpecsDataModule:
    """Transform half of the reference's SpecsDataModule (fdbm/data_module.py:112-229)."""

    def __init__(self, base_dir=None, format="default", batch_size=8, n_fft=510, hop_length=128, num_frames=256,
                 window="hann", num_data_per_epoch=None, num_workers=4, dummy=False, spec_factor=0.15,
                 spec_abs_exponent=0.5, gpu=True, normalize="noisy", transform_type="exponent", **kwargs):
        self.base_dir = base_dir
        self.format = format
        self.batch_size = batch_size
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.num_frames = num_frames
        self.window = get_window(window, self.n_fft)
        self.windows = {}
        self.num_workers = num_workers
        self.dummy = dummy
        self.spec_factor = spec_factor
        self.spec_abs_exponent = spec_abs_exponent
        self.gpu = gpu
        self.normalize = normalize
        self.transform_type = transform_type
        self.num_data_per_epoch = num_data_per_epoch
        self.kwargs = kwargs

    # ---- reference API -------------------------------------------------------------------------
    @property
    def stft_kwargs(self):
        return {**self.istft_kwargs, "return_complex": True}

    @property
    def istft_kwargs(self):
        return dict(n_fft=self.n_fft, hop_length=self.hop_length, window=self.window, center=True)

    def _get_window(self, x):
        """fdbm/data_module.py:212-221: one cached window tensor per device."""
        w = self.windows.get(x.device, None)
        if w is None:
            w = self.window.to(x.device).contiguous()
            self.windows[x.device] = w
        return w

    def _transform_args(self):
        if self.transform_type not in FDBM_TRANSFORM:
            raise NotImplementedError(self.transform_type)
        return FDBM_TRANSFORM[self.transform_type], float(self.spec_factor), float(self.spec_abs_exponent)

    def stft(self, sig: torch.Tensor) -> torch.Tensor:
        """fdbm/data_module.py:223-225: [..., Ts] fp32 -> [..., F, M] complex64 (no compression)."""
        return self._stft_impl(sig, FDBM_TRANSFORM["none"], 1.0, 1.0, "zero_pad", None)

    def istft(self, spec: torch.Tensor, length=None) -> torch.Tensor:
        """fdbm/data_module.py:227-229."""
        return self._istft_impl(spec, FDBM_TRANSFORM["none"], 1.0, 1.0, length)

    def spec_fwd(self, spec: torch.Tensor) -> torch.Tensor:
        """fdbm/data_module.py:173-186."""
        return self._spec_transform(spec, inverse=0)

    def spec_back(self, spec: torch.Tensor) -> torch.Tensor:
        """fdbm/data_module.py:188-199."""
        return self._spec_transform(spec, inverse=1)

    # ---- fused forms ---------------------------------------------------------------------------
    def stft_compress(self, sig: torch.Tensor, pad_mode: str = "zero_pad", n_frames_out=None, norm=None) -> torch.Tensor:
        """pad_spec(spec_fwd(stft(sig / norm)))[:, None] in one kernel: [B, Ts] -> [B, 1, F, T_pad].  `norm`: optional fp32 [B]
        (the peak normalisation of infer_single.py:83-87, applied to the spectrum -- the STFT is linear)."""
        if sig.dim() != 2:
            raise RuntimeError("stft_compress expects a [B, n_samples] batch")
        M = 1 + sig.shape[-1] // self.hop_length
        T_out = padded_frames(M) if n_frames_out is None else n_frames_out
        return self._stft_ex(sig, None, sig.shape[-1], sig.shape[-1], pad_mode, T_out, norm)

    def stft_compress_var(self, sig: torch.Tensor, lengths: torch.Tensor, min_len: int, max_len: int,
                          pad_mode: str = "zero_pad", n_frames_out=None, norm=None) -> torch.Tensor:
        """Variable-length batch: row b of `sig` [B, >= max_len] holds lengths[b] samples (device int32 [B]); every row is
        framed / reflect-padded / frame-padded from its own length, i.e. equals stft_compress of that utterance alone."""
        if lengths.dtype != torch.int32 or not lengths.is_cuda or lengths.numel() != sig.shape[0]:
            raise RuntimeError("stft_compress_var expects device int32 lengths [B]")
        T_out = padded_frames(1 + max_len // self.hop_length) if n_frames_out is None else n_frames_out
        return self._stft_ex(sig, lengths, min_len, max_len, pad_mode, T_out, norm)

    @on_device
    def _stft_ex(self, sig, lengths, min_len, max_len, pad_mode, T_out, norm):
        if sig.dim() != 2 or not sig.is_cuda or sig.dtype != torch.float32:
            raise RuntimeError("fdbm_b200 expects an fp32 CUDA [B, n_samples] batch")
        tr, fac, e = self._transform_args()
        x = sig if sig.stride(1) == 1 else sig.contiguous()
        B = x.shape[0]
        if norm is not None and (norm.dtype != torch.float32 or norm.numel() != B or not norm.is_contiguous()):
            raise RuntimeError("norm must be a contiguous fp32 tensor with one entry per utterance")
        spec = torch.empty(B, self.n_fft // 2 + 1, T_out, dtype=torch.complex64, device=x.device)
        check(_lib.load().fdbm_stft_compress_ex(ptr(x), B, ptr(lengths), int(min_len), int(max_len), x.stride(0),
                                                ptr(self._get_window(x)), ptr(norm), self.n_fft, self.hop_length, tr, fac, e,
                                                FDBM_PAD[pad_mode], T_out, ptr(spec), current_stream()), "fdbm_stft_compress_ex")
        return spec[:, None]

    @on_device
    def wave_absmax(self, sig: torch.Tensor, lengths=None) -> torch.Tensor:
        """max_n |sig[b, n]| per utterance, fp32 [B] (norm_factor = y.abs().max(), infer_single.py:83-84)."""
        if sig.dim() != 2 or not sig.is_cuda or sig.dtype != torch.float32 or sig.stride(1) != 1:
            raise RuntimeError("wave_absmax expects an fp32 CUDA [B, n_samples] batch")
        out = torch.empty(sig.shape[0], dtype=torch.float32, device=sig.device)
        check(_lib.load().fdbm_wave_absmax(ptr(sig), sig.shape[0], sig.shape[1], ptr(lengths), sig.stride(0), ptr(out),
                                           current_stream()), "fdbm_wave_absmax")
        return out

    @on_device
    def clip_rescale_(self, wave: torch.Tensor, peak: torch.Tensor, rescale: float, lengths=None) -> torch.Tensor:
        """In place: rows whose peak exceeds 1 become wave / peak * rescale (infer_single.py:95-97, infer_folder.py:119-120)."""
        check(_lib.load().fdbm_clip_rescale(ptr(wave), wave.shape[0], wave.shape[1], ptr(lengths), wave.stride(0), ptr(peak),
                                            float(rescale), current_stream()), "fdbm_clip_rescale")
        return wave

    @on_device
    def to_audio_ex(self, spec: torch.Tensor, length: int, lengths=None, norm=None, want_peak=False):
        """istft(spec_back(spec)) * norm in one kernel: [B, F, T] -> [B, length]; rows of a variable-length batch are valid up to
        lengths[b] (zeros beyond).  Returns (wave, peak) with peak[b] = max |wave[b]| when `want_peak`."""
        tr, fac, e = self._transform_args()
        s = _as_cfloat(spec)
        B, F, M = s.shape
        if F != self.n_fft // 2 + 1:
            raise RuntimeError(f"expected {self.n_fft // 2 + 1} frequency bins, got {F}")
        alloc = torch.zeros if lengths is not None else torch.empty
        wave = alloc(B, int(length), dtype=torch.float32, device=s.device)
        peak = torch.empty(B, dtype=torch.float32, device=s.device) if want_peak else None
        check(_lib.load().fdbm_decompress_istft_ex(ptr(s), B, M, ptr(self._get_window(s)), self.n_fft, self.hop_length, tr, fac, e,
                                                   ptr(lengths), int(length), wave.stride(0), ptr(norm), ptr(peak), ptr(wave),
                                                   current_stream()), "fdbm_decompress_istft_ex")
        return wave, peak

    def to_audio_var(self, spec: torch.Tensor, lengths: torch.Tensor, max_len: int) -> torch.Tensor:
        """Variable-length to_audio: [B, F, T] -> [B, max_len], row b valid up to lengths[b] (zeros beyond)."""
        return self.to_audio_ex(spec, max_len, lengths=lengths)[0]

    def to_audio(self, spec: torch.Tensor, length=None) -> torch.Tensor:
        """istft(spec_back(spec), length) in one kernel (fdbm/model.py:376-377)."""
        tr, fac, e = self._transform_args()
        return self._istft_impl(spec, tr, fac, e, length)

    # ---- kernels -------------------------------------------------------------------------------
    @on_device
    def _stft_impl(self, sig, tr, fac, e, pad_mode, n_frames_out):
        if not sig.is_cuda or sig.dtype != torch.float32:
            raise RuntimeError("fdbm_b200 expects fp32 CUDA waveforms")
        lead = sig.shape[:-1]
        x = sig.reshape(-1, sig.shape[-1]).contiguous()
        B, Ts = x.shape
        M = 1 + Ts // self.hop_length
        T_out = M if n_frames_out is None else n_frames_out
        F = self.n_fft // 2 + 1
        spec = torch.empty(B, F, T_out, dtype=torch.complex64, device=x.device)
        check(_lib.load().fdbm_stft_compress(ptr(x), B, Ts, x.stride(0), ptr(self._get_window(x)), self.n_fft,
                                             self.hop_length, tr, fac, e, FDBM_PAD[pad_mode], T_out, ptr(spec),
                                             current_stream()), "fdbm_stft_compress")
        return spec.reshape(*lead, F, T_out)

    @on_device
    def _istft_impl(self, spec, tr, fac, e, length):
        spec = _as_cfloat(spec)
        F, M = spec.shape[-2], spec.shape[-1]
        if F != self.n_fft // 2 + 1:
            raise RuntimeError(f"expected {self.n_fft // 2 + 1} frequency bins, got {F}")
        lead = spec.shape[:-2]
        s = spec.reshape(-1, F, M)
        B = s.shape[0]
        if length is None:
            length = self.hop_length * (M - 1)
        wave = torch.empty(B, length, dtype=torch.float32, device=spec.device)
        check(_lib.load().fdbm_decompress_istft(ptr(s), B, M, ptr(self._get_window(spec)), self.n_fft, self.hop_length,
                                                tr, fac, e, length, wave.stride(0), ptr(wave), current_stream()),
              "fdbm_decompress_istft")
        return wave.reshape(*lead, length)

    @on_device
    def _spec_transform(self, spec, inverse):
        tr, fac, e = self._transform_args()
        spec = _as_cfloat(spec)
        out = torch.empty_like(spec)
        check(_lib.load().fdbm_spec_transform(ptr(spec), ptr(out), spec.numel(), tr, fac, e, inverse, current_stream()),
              "fdbm_spec_transform")
        return out
