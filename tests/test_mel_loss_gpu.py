"""GPU: the mel / STFT loss tail (include/vcd.h vcd_mel_*; vcvits_b200/mel.py) against the CPU oracle
(oracle/mel_oracle.py, pinned to the reference's mel_spectrogram_torch by tests/golden/mel_probe.npz).

Tolerances (fp32 path, fp64 oracle): log-mel max-abs <= 2e-4 where the mel energy is above the 1e-5 clamp,
loss rel <= 1e-5, gradient rel-L2 <= 1e-4.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle as M
from vcvits_b200 import mel as V

pytestmark = pytest.mark.gpu

BASE = dict(n_fft=2048, num_mels=256, sampling_rate=48000, hop_size=512, win_size=2048, fmin=0.0, fmax=None)
B48K = dict(BASE, num_mels=128)
SMALL = dict(n_fft=256, num_mels=40, sampling_rate=16000, hop_size=64, win_size=256, fmin=0.0, fmax=None)
SHORTWIN = dict(n_fft=512, num_mels=64, sampling_rate=22050, hop_size=128, win_size=400, fmin=30.0, fmax=8000.0)
NONPOW2 = dict(n_fft=400, num_mels=40, sampling_rate=16000, hop_size=100, win_size=400, fmin=0.0, fmax=None)  # GEMM path only
PATHS = ("fft", "gemm")   # per-frame shared-memory FFT kernel (default) | dense DFT GEMMs (any n_fft)


@pytest.fixture
def path(request):
    """Select the implementation for every plan the test touches; always restore the default afterwards."""
    which = request.param
    touched = []

    def select(kw):
        plan = V._plan(device="cuda", **kw)
        pow2 = kw["n_fft"] & (kw["n_fft"] - 1) == 0
        if which == "fft" and not pow2:
            pytest.skip("the FFT path needs a power-of-two n_fft")
        plan.use_gemm_path(which == "gemm")
        touched.append((plan, pow2))
    yield select
    for plan, pow2 in touched:
        if pow2:
            plan.use_gemm_path(False)


def audio(B, T, seed, amp=0.3):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float64) / 48000.0
    tone = torch.sin(2 * np.pi * 220.0 * t)[None] * torch.rand(B, 1, generator=g, dtype=torch.float64)
    return (amp * (0.5 * tone + 0.5 * torch.randn(B, T, generator=g, dtype=torch.float64))).clamp(-1, 1)


def oracle_loss(y, tgt, c_mel, kw):
    yy = y.clone().double().requires_grad_(True)
    loss = torch.nn.functional.l1_loss(M.log_mel(yy, **kw), tgt.double()) * c_mel
    loss.backward()
    return float(loss), yy.grad


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_reference_fixture_log_mel(golden_dir):
    d = np.load(os.path.join(golden_dir, "mel_probe.npz"))
    m = V.mel_spectrogram_torch(torch.from_numpy(d["y"]).cuda(), 2048, 256, 48000, 512, 2048, 0.0, None)
    assert m.shape == (2, 256, 32)
    assert float((m.cpu() - torch.from_numpy(d["logmel"])).abs().max()) <= 2e-4


@pytest.mark.parametrize("kw,B,T", [
    (BASE, 16, 16384),     # configs[1]: segment_size 16384 (base.json:20), B = 16
    (B48K, 32, 16384),     # configs[2]
    (BASE, 3, 48000),      # a whole second: the y_mel of the data side (vcvits.py:64-76)
    (BASE, 2, 8192 + 100), # hop does not divide T
    (BASE, 1, 769),        # shortest legal input: T = pad + 1 -> 1 frame... (pad = 768)
    (SMALL, 5, 1000),
    (SHORTWIN, 4, 3000),   # win_length < n_fft: window centred in the frame
    (NONPOW2, 3, 2500),
])
@pytest.mark.parametrize("path", PATHS, indirect=True)
def test_log_mel_matches_oracle(kw, B, T, path):
    path(kw)
    y = audio(B, T, seed=B * 1000 + T)
    ref = M.log_mel(y, **kw)
    got = V.mel_spectrogram_torch(y.float().cuda(), **kw).cpu()
    assert got.shape == ref.shape
    above = ref > np.log(1e-5) + 1e-3
    assert float((got.double() - ref).abs()[above].max()) <= 2e-4
    assert float((got.double() - ref).abs().max()) <= 5e-3


@pytest.mark.parametrize("kw,B,T", [(BASE, 16, 16384), (B48K, 32, 16384), (BASE, 2, 8192 + 100), (SMALL, 5, 1000),
                                    (SHORTWIN, 4, 3000), (NONPOW2, 3, 2500)])
@pytest.mark.parametrize("path", PATHS, indirect=True)
def test_loss_and_gradient_match_oracle(kw, B, T, path):
    path(kw)
    y_hat = audio(B, T, seed=7)
    y_real = audio(B, T, seed=8, amp=0.4)
    tgt = M.log_mel(y_real, **kw)
    c_mel = 45.0                                    # configs/base.json: c_mel
    loss_ref, dy_ref = oracle_loss(y_hat, tgt, c_mel, kw)
    yc = y_hat.float().cuda().unsqueeze(1).requires_grad_(True)      # [B, 1, T] like the decoder output
    loss = V.mel_l1_loss(yc, tgt.float().cuda(), c_mel=c_mel, **kw)
    loss.backward()
    assert abs(float(loss) - loss_ref) <= 1e-5 * abs(loss_ref)
    assert yc.grad.shape == yc.shape
    assert rel_l2(yc.grad.cpu()[:, 0], dy_ref) <= 1e-4
    # per-item bound as well (one wrong frame / pad position would hide in the global norm)
    for b in range(B):
        assert rel_l2(yc.grad.cpu()[b, 0], dy_ref[b]) <= 3e-4
    # the reflect-padded edges specifically
    pad = (kw["n_fft"] - kw["hop_size"]) // 2
    assert rel_l2(yc.grad.cpu()[:, 0, :pad + 2], dy_ref[:, :pad + 2]) <= 3e-4
    assert rel_l2(yc.grad.cpu()[:, 0, -pad - 2:], dy_ref[:, -pad - 2:]) <= 3e-4


def test_upstream_gradient_scaling_and_no_grad():
    kw, B, T = SMALL, 3, 2000
    y = audio(B, T, seed=3).float().cuda()
    tgt = V.mel_spectrogram_torch(audio(B, T, seed=4).float().cuda(), **kw)
    a = y.clone().requires_grad_(True)
    V.mel_l1_loss(a, tgt, c_mel=1.0, **kw).backward()
    b = y.clone().requires_grad_(True)
    (2.5 * V.mel_l1_loss(b, tgt, c_mel=2.0, **kw)).backward()
    assert torch.allclose(b.grad, 5.0 * a.grad, rtol=1e-5, atol=0)
    with torch.no_grad():
        l0 = V.mel_l1_loss(y, tgt, c_mel=1.0, **kw)
    assert float(l0) > 0 and not l0.requires_grad
    # identical input and target: zero loss, zero gradient (sign(0) = 0 like F.l1_loss)
    c = y.clone().requires_grad_(True)
    lz = V.mel_l1_loss(c, V.mel_spectrogram_torch(y, **kw), c_mel=1.0, **kw)
    lz.backward()
    assert float(lz) == 0.0 and float(c.grad.abs().max()) == 0.0


def test_the_two_implementations_agree():
    kw, B, T = BASE, 16, 16384
    plan = V._plan(device="cuda", **kw)
    y = audio(B, T, seed=31).float().cuda()
    tgt = V.mel_spectrogram_torch(audio(B, T, seed=32).float().cuda(), **kw)
    res = {}
    try:
        for name in PATHS:
            plan.use_gemm_path(name == "gemm")
            res[name] = (V.mel_spectrogram_torch(y, **kw),) + plan.loss_and_grad(y, tgt, 45.0)
    finally:
        plan.use_gemm_path(False)
    assert float((res["fft"][0] - res["gemm"][0]).abs().max()) <= 2e-3     # log of tiny energies amplifies fp32 rounding
    assert abs(float(res["fft"][1]) - float(res["gemm"][1])) <= 1e-5 * float(res["gemm"][1])
    assert rel_l2(res["fft"][2], res["gemm"][2]) <= 3e-4     # each is within 1e-4 of the fp64 oracle


@pytest.mark.parametrize("path", PATHS, indirect=True)
def test_bit_identical_from_run_to_run(path):
    kw, B, T = BASE, 16, 16384
    path(kw)
    y = audio(B, T, seed=11).float().cuda()
    tgt = V.mel_spectrogram_torch(audio(B, T, seed=12).float().cuda(), **kw)
    outs = []
    for _ in range(3):
        a = y.clone().requires_grad_(True)
        loss = V.mel_l1_loss(a, tgt, c_mel=45.0, **kw)
        loss.backward()
        outs.append((loss.detach().clone(), a.grad.clone()))
    for l, g in outs[1:]:
        assert torch.equal(l, outs[0][0]) and torch.equal(g, outs[0][1])


def test_error_paths():
    kw = BASE
    with pytest.raises(RuntimeError, match="exceed the reflect pad"):
        V.mel_spectrogram_torch(torch.zeros(1, 768).cuda(), **kw)
    with pytest.raises(RuntimeError, match="mel target of shape"):
        V.mel_l1_loss(torch.zeros(1, 4096).cuda(), torch.zeros(1, 256, 7).cuda(), **kw)
    with pytest.raises(RuntimeError, match="center"):
        V.mel_spectrogram_torch(torch.zeros(1, 4096).cuda(), center=True, **kw)
    with pytest.raises(RuntimeError, match="n_mel"):
        V.mel_spectrogram_torch(torch.zeros(1, 4096).cuda(), **dict(kw, num_mels=30))


def test_decoder_into_loss_tail_end_to_end():
    """Generator (bf16 tensor-core mode) -> mel loss tail -> decoder backward, against the fp64 oracle of both."""
    from oracle import hifigan_oracle as O
    from vcvits_b200 import Generator
    cfg = O.SMALL_CFG
    kw = dict(n_fft=64, num_mels=16, sampling_rate=16000, hop_size=16, win_size=64, fmin=0.0, fmax=None)
    sd = O.seeded_state_dict(cfg, 99, gain=1.3)
    torch.manual_seed(5)
    x, g = torch.randn(2, cfg["initial_channel"], 24), torch.randn(2, cfg["gin_channels"], 1)
    ref = O.build(cfg, {k: v.double() for k, v in sd.items()}, dtype=torch.float64)
    xr = x.double().requires_grad_(True)
    yr = ref(xr, g.double())
    tgt = M.log_mel(audio(2, yr.shape[-1], seed=21, amp=0.05), **kw)
    lr = torch.nn.functional.l1_loss(M.log_mel(yr[:, 0], **kw), tgt) * 45.0
    lr.backward()
    m = Generator(**cfg, mode="fp32")
    m.load_state_dict(sd)
    m = m.cuda()
    xc = x.cuda().requires_grad_(True)
    y = m(xc, g.cuda())
    loss = V.mel_l1_loss(y, tgt.float().cuda(), c_mel=45.0, **kw)
    loss.backward()
    assert abs(float(loss) - float(lr)) <= 1e-4 * abs(float(lr))
    assert rel_l2(xc.grad.cpu(), xr.grad) <= 2e-3


@pytest.mark.parametrize("path", PATHS, indirect=True)
def test_sliced_target_equals_slice_segments_then_loss(path):
    """``ids_slice`` folds ``commons.slice_segments(y_mel, ids_slice, frames)`` (vits/commons.py:48-55) into the target read."""
    kw, B, T, T_full = BASE, 6, 16384, 60000
    path(kw)
    y_hat = audio(B, T, seed=41).float().cuda().requires_grad_(True)
    mel_full = V.mel_spectrogram_torch(audio(B, T_full, seed=42).float().cuda(), **kw)
    frames = T // kw["hop_size"]
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, mel_full.shape[2] - frames + 1, (B,), generator=g)
    ids[0], ids[1] = 0, mel_full.shape[2] - frames            # both ends
    sliced = torch.stack([mel_full[b, :, int(ids[b]):int(ids[b]) + frames] for b in range(B)])
    a = V.mel_l1_loss(y_hat, sliced, c_mel=45.0, **kw)
    (ga,) = torch.autograd.grad(a, y_hat)
    b = V.mel_l1_loss(y_hat, mel_full, c_mel=45.0, ids_slice=ids.cuda(), **kw)
    (gb,) = torch.autograd.grad(b, y_hat)
    assert torch.equal(a, b) and torch.equal(ga, gb)
    with pytest.raises(RuntimeError, match="full-length mel"):
        V.mel_l1_loss(y_hat, mel_full[:, :, :frames - 1], c_mel=1.0, ids_slice=ids.cuda(), **kw)
