"""CPU: host logic of the mel loss tail (vcvits_b200/mel.py) -- the Slaney filterbank against the oracle's, and the
no-fallback rule."""
import numpy as np
import pytest
import torch

from oracle import mel_oracle as M
from vcvits_b200 import mel as V


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", [
    (48000, 2048, 256, 0.0, None),     # configs/base.json:31-37
    (48000, 2048, 128, 0.0, None),     # configs/48k_base.json
    (22050, 1024, 80, 0.0, 8000.0),
    (16000, 512, 40, 50.0, 7000.0),
])
def test_slaney_filterbank_matches_oracle(sr, n_fft, n_mels, fmin, fmax):
    a = V.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    b = M.mel_filterbank(sr, n_fft, n_mels, fmin, fmax).numpy()
    assert a.dtype == np.float32 and a.shape == b.shape == (n_mels, n_fft // 2 + 1)
    assert float(np.abs(a - b).max()) <= 1e-5 * float(np.abs(b).max())
    # every filter is a non-negative triangle of unit area in Hz (Slaney normalisation)
    assert (a >= 0).all()
    hz_per_bin = sr / n_fft
    areas = a.sum(axis=1) * hz_per_bin
    wide = (a > 0).sum(axis=1) >= 12          # the sampled area of a triangle narrower than a few bins is meaningless
    assert wide.any() and np.allclose(areas[wide], 1.0, atol=0.1)


def test_cpu_tensors_raise():
    y = torch.zeros(1, 4096)
    with pytest.raises(RuntimeError):
        V.mel_spectrogram_torch(y, 2048, 256, 48000, 512, 2048, 0.0, None)
    with pytest.raises(RuntimeError):
        V.mel_l1_loss(y, torch.zeros(1, 256, 8), 2048, 256, 48000, 512, 2048, 0.0, None)


def test_c_abi_fails_loudly_without_a_device_and_on_null_plans():
    import ctypes as C
    from vcvits_b200 import _lib
    lib = _lib.load()
    assert lib.vcd_set_deterministic(None, 1) != 0 and b"null plan" in lib.vcd_last_error()
    assert lib.vcd_mel_debug_path(None, 1) != 0 and b"null plan" in lib.vcd_last_error()
    assert lib.vcd_mel_frames(None, 4096) == 0 and lib.vcd_mel_workspace_bytes(None, 1, 4096) == 0
    if torch.cuda.is_available():
        pytest.skip("GPU present: plan creation succeeds")
    fb = np.ascontiguousarray(V.slaney_mel_filterbank(48000, 2048, 256, 0.0, None))
    cfg = V._MelConfig(2048, 512, 2048, 256)
    handle = C.c_void_p()
    assert lib.vcd_mel_plan_create(C.byref(cfg), fb.ctypes.data_as(C.c_void_p), C.byref(handle)) != 0
    assert b"no CUDA device" in lib.vcd_last_error()
