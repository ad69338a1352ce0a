"""Mel / STFT loss tail of the generator step on the CUDA library.

Host-side mirror of the reference functions that consume the decoder's waveform:
  * ``mel_spectrogram_torch(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)``
    -- vits/mel_processing.py:115-142 (same value as ``spec_to_mel_torch(spectrogram_torch_audio(y, ...), ...)``,
    mel_processing.py:76-112, which is what vits/light/vcvits.py:64-76,96-108 calls);
  * ``mel_l1_loss(y_hat, y_mel, ..., c_mel)`` = ``F.l1_loss(mel_spectrogram_torch(y_hat, ...), y_mel) * c_mel``
    (vcvits.py:115) with its backward: one library call returns the loss AND d loss / d y_hat, the decoder's
    upstream gradient.

Like the reference's module-level ``mel_basis`` / ``hann_window`` caches, plans are cached per
(parameters, device).  The Slaney filterbank the reference takes from ``librosa.filters.mel`` (third-party, not in
the reference tree; pinned by requirements.txt as ``librosa``) is restated here from its published definition
(Slaney's Auditory Toolbox mel scale: linear below 1 kHz at 200/3 Hz per mel, logarithmic above with
log(6.4)/27 per mel; triangles normalised to unit area by 2 / bandwidth).

No CPU or PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib

__all__ = ["slaney_mel_filterbank", "mel_spectrogram_torch", "mel_l1_loss", "MelLossTail"]


def _hz_to_mel(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: Optional[float] = None) -> np.ndarray:
    """``librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)`` with its defaults (htk=False, norm='slaney'):
    fp32 ``[n_mels, n_fft // 2 + 1]``."""
    fmax = float(sr) / 2 if not fmax else float(fmax)
    fftfreqs = np.linspace(0.0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(np.float64(fmin)), _hz_to_mel(np.float64(fmax)), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


class _MelConfig(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop", C.c_int32), ("win", C.c_int32), ("n_mel", C.c_int32)]


class MelLossTail:
    """One plan (DFT basis, window, filterbank on the device) + a cached workspace."""

    def __init__(self, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int, win_size: int, fmin: float, fmax,
                 device: torch.device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("vcvits_b200.mel runs on CUDA devices only (no CPU fallback)")
        self.n_fft, self.num_mels, self.hop, self.win = int(n_fft), int(num_mels), int(hop_size), int(win_size)
        lib = _lib.load()
        fb = np.ascontiguousarray(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
        cfg = _MelConfig(self.n_fft, self.hop, self.win, self.num_mels)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.vcd_mel_plan_create(C.byref(cfg), fb.ctypes.data_as(C.c_void_p), C.byref(handle)),
                       "vcd_mel_plan_create")
        self._plan = handle
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if getattr(self, "_plan", None):
                _lib.load().vcd_mel_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def use_gemm_path(self, on: bool) -> None:
        """Debug / tests: force the dense-GEMM implementation (the only one when n_fft is not a power of two)."""
        _lib.check(_lib.load().vcd_mel_debug_path(self._plan, 1 if on else 0), "vcd_mel_debug_path")

    def frames(self, T: int) -> int:
        return int(_lib.load().vcd_mel_frames(self._plan, int(T)))

    def _workspace(self, B: int, T: int) -> torch.Tensor:
        need = int(_lib.load().vcd_mel_workspace_bytes(self._plan, B, T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _as_2d(y: torch.Tensor) -> torch.Tensor:
        if y.dim() == 3 and y.shape[1] == 1:
            y = y[:, 0]
        if y.dim() != 2:
            raise RuntimeError(f"expected audio of shape [B, T] or [B, 1, T], got {tuple(y.shape)}")
        return y.contiguous().float()

    def spectrogram(self, y: torch.Tensor) -> torch.Tensor:
        y2 = self._as_2d(y.detach())
        if y2.device != self.device:
            raise RuntimeError(f"audio on {y2.device}, plan on {self.device}")
        B, T = y2.shape
        with torch.cuda.device(self.device):
            ws = self._workspace(B, T)
            out = torch.empty(B, self.num_mels, max(self.frames(T), 0), dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.load().vcd_mel_spectrogram(self._plan, y2.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), B,
                                                       T, stream), "vcd_mel_spectrogram")
        return out

    def loss_and_grad(self, y_hat: torch.Tensor, y_mel: torch.Tensor, c_mel: float, want_grad: bool = True,
                      ids_slice: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """``ids_slice`` (int64 [B], device): ``y_mel`` is the whole-utterance mel and frames
        ``ids_slice[b] : ids_slice[b] + frames`` are the target (``commons.slice_segments`` folded into the read)."""
        y2 = self._as_2d(y_hat.detach())
        if y2.device != self.device:
            raise RuntimeError(f"audio on {y2.device}, plan on {self.device}")
        B, T = y2.shape
        F_ = self.frames(T)
        tgt = y_mel.detach().contiguous().float()
        if ids_slice is None:
            if tuple(tgt.shape) != (B, self.num_mels, F_):
                raise RuntimeError(f"mel target of shape {tuple(tgt.shape)}, expected {(B, self.num_mels, F_)}")
        else:
            if tgt.dim() != 3 or tuple(tgt.shape[:2]) != (B, self.num_mels) or tgt.shape[2] < F_:
                raise RuntimeError(f"full-length mel of shape {tuple(tgt.shape)}, expected {(B, self.num_mels)} + (>= {F_},)")
            ids_slice = ids_slice.detach().to(device=self.device, dtype=torch.int64).contiguous()
            if tuple(ids_slice.shape) != (B,):
                raise RuntimeError(f"ids_slice of shape {tuple(ids_slice.shape)}, expected {(B,)}")
        with torch.cuda.device(self.device):
            ws = self._workspace(B, T)
            loss = torch.empty((), dtype=torch.float32, device=self.device)
            dy = torch.empty(B, T, dtype=torch.float32, device=self.device) if want_grad else None
            stream = torch.cuda.current_stream().cuda_stream
            dyp = dy.data_ptr() if dy is not None else None
            if ids_slice is None:
                _lib.check(_lib.load().vcd_mel_loss(self._plan, y2.data_ptr(), tgt.data_ptr(), float(c_mel), loss.data_ptr(), dyp,
                                                    ws.data_ptr(), ws.numel(), B, T, stream), "vcd_mel_loss")
            else:
                _lib.check(_lib.load().vcd_mel_loss_sliced(self._plan, y2.data_ptr(), tgt.data_ptr(), int(tgt.shape[2]),
                                                           ids_slice.data_ptr(), float(c_mel), loss.data_ptr(), dyp, ws.data_ptr(),
                                                           ws.numel(), B, T, stream), "vcd_mel_loss_sliced")
        return loss, dy


_plans: Dict[tuple, MelLossTail] = {}


def _plan(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, device) -> MelLossTail:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (int(n_fft), int(num_mels), int(sampling_rate), int(hop_size), int(win_size), float(fmin),
           None if not fmax else float(fmax), str(device))
    p = _plans.get(key)
    if p is None:
        p = _plans[key] = MelLossTail(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, device)
    return p


def mel_spectrogram_torch(y: torch.Tensor, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int, win_size: int,
                          fmin, fmax, center: bool = False) -> torch.Tensor:
    """Same signature and value as vits/mel_processing.py:115 (``center`` must be False, as at every call site)."""
    if center:
        raise RuntimeError("center=True is not used by the reference and not implemented")
    return _plan(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, y.device).spectrogram(y)


class _MelL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, y_mel, plan: MelLossTail, c_mel: float, ids_slice=None):
        loss, dy = plan.loss_and_grad(y_hat, y_mel, c_mel, want_grad=y_hat.requires_grad, ids_slice=ids_slice)
        ctx.shape = y_hat.shape
        ctx.in_dtype = y_hat.dtype
        ctx.save_for_backward(dy if dy is not None else torch.empty(0, device=y_hat.device))
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dy,) = ctx.saved_tensors
        if dy.numel() == 0:
            return None, None, None, None, None
        return (dy * grad_out).reshape(ctx.shape).to(ctx.in_dtype), None, None, None, None


def mel_l1_loss(y_hat: torch.Tensor, y_mel: torch.Tensor, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int,
                win_size: int, fmin, fmax, c_mel: float = 1.0, ids_slice: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.l1_loss(spec_to_mel_torch(spectrogram_torch_audio(y_hat, ...), ...), y_mel) * c_mel`` (vcvits.py:96-115);
    differentiable w.r.t. ``y_hat`` (the gradient is computed in the same library call as the loss).  With ``ids_slice``
    (the second result of ``commons.rand_slice_segments``), ``y_mel`` is the whole-utterance mel and
    ``commons.slice_segments(y_mel, ids_slice, frames)`` (vcvits.py:110) happens inside the kernel's target read."""
    plan = _plan(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, y_hat.device)
    return _MelL1.apply(y_hat, y_mel, plan, float(c_mel), ids_slice)
