"""CPU: the log-mel restatement against the reference's mel_spectrogram_torch (vits/mel_processing.py:115-142)."""
import os

import numpy as np
import torch

from oracle import mel_oracle as M


def test_log_mel_matches_reference_fixture(golden_dir):
    d = np.load(os.path.join(golden_dir, "mel_probe.npz"))
    m = M.log_mel(torch.from_numpy(d["y"]))
    assert m.shape == (2, 256, 32)  # 16384 samples, hop 512 (configs/base.json:33)
    assert float((m - torch.from_numpy(d["logmel"])).abs().max()) <= 1e-5


def test_log_mel_l1_relative_is_zero_on_identical_and_scales():
    torch.manual_seed(0)
    y = 0.1 * torch.randn(1, 16384)
    assert M.log_mel_l1_relative(y, y) == 0.0
    assert 0.0 < M.log_mel_l1_relative(y * 1.05, y) < 0.05
