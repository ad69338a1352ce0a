"""CPU restatement of the reference's log-mel spectrogram (the parity metric named by the north star).

TEST INFRASTRUCTURE ONLY (used by tests/ and bench.py's checks; never by vcvits_b200/).

Follows ``vits/mel_processing.py:115-142`` (``mel_spectrogram_torch``): reflect-pad (n_fft-hop)/2 (line 131),
``torch.stft`` with a Hann window and ``center=False`` (134-135), magnitude ``sqrt(re^2 + im^2 + 1e-6)`` (137),
Slaney mel filterbank matmul (126, 139) and ``log(clamp(x, 1e-5))`` (``dynamic_range_compression_torch``, 22-28, 140).
The reference builds the filterbank with ``librosa.filters.mel`` (librosa is not installed here); the published
Slaney-scale / Slaney-norm definition is restated through ``torchaudio.functional.melscale_fbanks`` with
``norm="slaney", mel_scale="slaney"``.  Pinned against the reference function (imported with a librosa stub that
supplies exactly this filterbank) by ``oracle/make_golden.py`` -> ``tests/golden/mel_probe.npz``.
"""
import torch
import torchaudio


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax) -> torch.Tensor:
    fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax if fmax else sr / 2), n_mels, sr,
                                               norm="slaney", mel_scale="slaney")
    return fb.T.contiguous()  # [n_mels, n_fft/2+1] like librosa.filters.mel


def log_mel(y: torch.Tensor, n_fft=2048, num_mels=256, sampling_rate=48000, hop_size=512, win_size=2048, fmin=0.0,
            fmax=None) -> torch.Tensor:
    """y: [B, T] in [-1, 1] -> [B, num_mels, frames].  Defaults = configs/base.json:31-37."""
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    window = torch.hann_window(win_size, dtype=y.dtype, device=y.device)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window, center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-6)
    mel = torch.matmul(mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax).to(y.dtype).to(y.device), mag)
    return torch.log(torch.clamp(mel, min=1e-5))


def log_mel_l1_relative(y: torch.Tensor, y_ref: torch.Tensor, **kw) -> float:
    """mean |logmel(y) - logmel(y_ref)| / mean |logmel(y_ref)| -- the "log-mel L1 within 1 %" criterion."""
    a, b = log_mel(y.double(), **kw), log_mel(y_ref.double(), **kw)
    return float((a - b).abs().mean() / b.abs().mean())
