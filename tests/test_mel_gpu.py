"""GPU: bf16-mode parity bar of the north star -- max-abs <= 1e-2 AND log-mel L1 within 1 % (through the
restated vits/mel_processing.py::mel_spectrogram_torch), on configs/base.json shapes."""
import pytest
import torch

from oracle import hifigan_oracle as O
from oracle import mel_oracle as M
from tests.helpers import b200_run, oracle_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gain", [1.0, 1.3])
def test_base_config_bf16_waveform_and_log_mel(gain):
    # gain 1.0 = default init (|y| ~ 0.02); gain 1.3 drives the output into tanh's non-linear range
    sd = O.seeded_state_dict(O.BASE_CFG, 1234, gain=gain)
    torch.manual_seed(7)
    x, g = torch.randn(2, 256, 32), torch.randn(2, 256, 1)
    y, _, _ = b200_run(O.BASE_CFG, sd, x, g, mode="bf16")
    y_ref, _ = oracle_run(O.BASE_CFG, sd, x, g, dtype=torch.float32)
    assert float((y - y_ref).abs().max()) <= 1e-2
    rel = M.log_mel_l1_relative(y[:, 0], y_ref[:, 0], num_mels=256)
    assert rel <= 0.01, rel


def test_48k_config_bf16_log_mel():
    sd = O.seeded_state_dict(O.BASE48K_CFG, 1234, gain=1.3)
    torch.manual_seed(8)
    x, g = torch.randn(1, 128, 32), torch.randn(1, 256, 1)
    y, _, _ = b200_run(O.BASE48K_CFG, sd, x, g, mode="bf16")
    y_ref, _ = oracle_run(O.BASE48K_CFG, sd, x, g, dtype=torch.float32)
    assert float((y - y_ref).abs().max()) <= 1e-2
    assert M.log_mel_l1_relative(y[:, 0], y_ref[:, 0], num_mels=128) <= 0.01


def test_full_size_step_properties():
    """BASELINE.json configs[1] size (batch 16 x 32 frames): size-independent checks instead of the slow oracle --
    bf16 vs fp32 mode agreement of the waveform and of the global gradient direction."""
    from vcvits_b200 import Generator
    sd = O.seeded_state_dict(O.BASE_CFG, 1234, gain=1.2)
    torch.manual_seed(11)
    x = torch.randn(16, 256, 32, device="cuda")
    g = torch.randn(16, 256, 1, device="cuda")
    dy = torch.randn(16, 1, 16384, device="cuda")
    out = {}
    for mode in ("fp32", "bf16"):
        m = Generator(**O.BASE_CFG, mode=mode)
        m.load_state_dict(sd)
        m = m.cuda()
        y = m(x, g)
        y.backward(dy)
        out[mode] = (y.detach(), torch.cat([p.grad.flatten() for p in m.parameters()]))
        del m
    (y32, g32), (y16, g16) = out["fp32"], out["bf16"]
    assert float((y32 - y16).abs().max()) <= 1e-2
    cos = float(torch.dot(g32, g16) / (g32.norm() * g16.norm()))
    assert cos >= 0.995, cos
    assert M.log_mel_l1_relative(y16[:, 0].cpu(), y32[:, 0].cpu(), num_mels=256) <= 0.01
