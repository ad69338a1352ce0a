"""GPU parity tests proper: the CUDA path through the C ABI vs. the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star + SURVEY.md §8c):
  fp32 mode : max-abs waveform error <= 1e-4 AND <= 1e-5 * absmax(y);  gradients rel-L2 <= 1e-3 vs fp64 oracle
  bf16 mode : max-abs <= 1e-2 (log-mel check lives in test_mel_gpu.py)
"""
import os

import pytest
import torch

from oracle import hifigan_oracle as O
from tests.helpers import b200_run, load_golden, oracle_run, rel_l2

pytestmark = pytest.mark.gpu

FP32_ABS = 1e-4
FP32_REL = 1e-5
GRAD_REL = 1e-3


def _check_fp32_forward(y, y_ref):
    err = float((y.double() - y_ref.double()).abs().max())
    amp = float(y_ref.abs().max())
    assert err <= FP32_ABS, f"max-abs {err}"
    assert err <= max(FP32_REL * amp, 2e-7), f"max-abs {err} vs amplitude {amp}"


@pytest.mark.parametrize("name,cfg", [("generator_tiny", O.TINY_CFG), ("generator_tiny2", O.TINY2_CFG)])
def test_golden_tiny_fp32_forward_backward(golden_dir, name, cfg):
    sd, ggrads, rest = load_golden(os.path.join(golden_dir, name + ".npz"))
    x, g, dy = (torch.from_numpy(rest[k]) for k in ("x", "g", "dy"))
    y, grads, _ = b200_run(cfg, sd, x, g, dy, mode="fp32")
    _check_fp32_forward(y, torch.from_numpy(rest["y64"]))
    assert rel_l2(grads["__x__"], torch.from_numpy(rest["grad_x"])) <= GRAD_REL
    assert rel_l2(grads["__g__"], torch.from_numpy(rest["grad_g"])) <= GRAD_REL
    for n, ref in ggrads.items():
        assert rel_l2(grads[n], ref) <= GRAD_REL, (n, rel_l2(grads[n], ref))


@pytest.mark.parametrize("B,T", [(1, 1), (2, 5), (3, 33), (1, 70)])
def test_tiny_ragged_shapes_fp32(B, T):
    """Edge shapes: single frame, odd lengths, lengths that do not fill a tile."""
    sd = O.seeded_state_dict(O.TINY_CFG, 11, gain=1.5)
    torch.manual_seed(B * 100 + T)
    x = torch.randn(B, 16, T)
    g = torch.randn(B, 8, 1)
    dy = torch.randn(B, 1, T * 8)
    y, grads, _ = b200_run(O.TINY_CFG, sd, x, g, dy, mode="fp32")
    y_ref, gref = oracle_run(O.TINY_CFG, sd, x, g, dy)
    _check_fp32_forward(y, y_ref)
    for n, ref in gref.items():
        assert rel_l2(grads[n], ref) <= GRAD_REL, (n, rel_l2(grads[n], ref))


def test_no_speaker_conditioning_and_strided_input():
    """dec(z) without g (synthesizer_svc.py:87) on a non-contiguous slice (synthesizer_tts.py:166)."""
    sd = O.seeded_state_dict(O.TINY_CFG, 5, gain=1.5)
    torch.manual_seed(0)
    z = torch.randn(2, 16, 40)
    x = z[:, :, :23]
    assert not x.is_contiguous()
    from vcvits_b200 import Generator
    m = Generator(**O.TINY_CFG, mode="fp32")
    m.load_state_dict(sd)
    m = m.cuda()
    with torch.no_grad():
        y = m(z.cuda()[:, :, :23]).cpu()
    y_ref, _ = oracle_run(O.TINY_CFG, sd, x, None)
    _check_fp32_forward(y, y_ref)


def test_base_config_fp32_forward_matches_oracle(golden_dir):
    """configs[0]: base.json decoder forward, batch 1, 32-frame segment (CPU-runnable reference case)."""
    import numpy as np
    d = np.load(os.path.join(golden_dir, "generator_base_probe.npz"))
    sd = O.seeded_state_dict(O.BASE_CFG, 1234)
    x, g = torch.from_numpy(d["x"]), torch.from_numpy(d["g"])
    y, _, _ = b200_run(O.BASE_CFG, sd, x, g, mode="fp32")
    y_ref, _ = oracle_run(O.BASE_CFG, sd, x, g)
    _check_fp32_forward(y, y_ref)
    # committed probe generated from the reference ResBlock classes
    assert float((y[0, 0, ::16] - torch.from_numpy(d["y_probe"])).abs().max()) <= 1e-6


def test_base_config_fp32_gradients():
    sd = O.seeded_state_dict(O.BASE_CFG, 1234, gain=1.25)
    torch.manual_seed(3)
    x, g = torch.randn(2, 256, 8), torch.randn(2, 256, 1)
    dy = torch.randn(2, 1, 8 * 512)
    y, grads, _ = b200_run(O.BASE_CFG, sd, x, g, dy, mode="fp32")
    y_ref, gref = oracle_run(O.BASE_CFG, sd, x, g, dy)
    _check_fp32_forward(y, y_ref)
    total_num = sum(float((grads[n].double() - gref[n]).pow(2).sum()) for n in gref)
    total_den = sum(float(gref[n].pow(2).sum()) for n in gref)
    assert (total_num / total_den) ** 0.5 <= GRAD_REL
    # per tensor: <= max(1e-3, 2x PyTorch-fp32's own error for that tensor) -- SURVEY.md F8 / §8c: PyTorch fp32
    # itself reaches ~2e-3 on a few weight_v tensors whose gradient is dominated by the cancelling radial part
    _, g32 = oracle_run(O.BASE_CFG, sd, x, g, dy, dtype=torch.float32)
    report = []
    for n in gref:
        mine, torch32 = rel_l2(grads[n], gref[n]), rel_l2(g32[n], gref[n])
        report.append((mine / max(1e-3, 2.0 * torch32), mine, torch32, n))
    worst = max(report)
    assert worst[0] <= 1.0, worst


@pytest.mark.parametrize("cfg", [O.TINY_CFG, O.TINY2_CFG])
def test_bf16_mode_close_to_oracle(cfg):
    sd = O.seeded_state_dict(cfg, 21, gain=1.5)
    torch.manual_seed(1)
    x, g = torch.randn(2, 16, 37), torch.randn(2, 8, 1)
    dy = torch.randn(2, 1, 37 * 8)
    y, grads, _ = b200_run(cfg, sd, x, g, dy, mode="bf16")
    y_ref, gref = oracle_run(cfg, sd, x, g, dy)
    assert float((y.double() - y_ref).abs().max()) <= 1e-2
    # bf16 operand rounding: gradients within a few percent (global), informational bound
    num = sum(float((grads[n].double() - gref[n]).pow(2).sum()) for n in gref)
    den = sum(float(gref[n].pow(2).sum()) for n in gref)
    assert (num / den) ** 0.5 <= 0.12  # 16/8-channel toy config, bf16 activations AND gradients (all-bf16 PyTorch: 5-19%)


def test_linearity_of_backward_in_dy():
    """Size-independent property: parameter gradients are linear in the upstream gradient."""
    from vcvits_b200 import Generator
    sd = O.seeded_state_dict(O.TINY_CFG, 2, gain=1.5)
    m = Generator(**O.TINY_CFG, mode="fp32")
    m.load_state_dict(sd)
    m = m.cuda()
    torch.manual_seed(5)
    x = torch.randn(2, 16, 64, device="cuda")
    g = torch.randn(2, 8, 1, device="cuda")
    dy1, dy2 = torch.randn(2, 1, 512, device="cuda"), torch.randn(2, 1, 512, device="cuda")

    def grads_for(dy):
        m.zero_grad(set_to_none=True)
        m(x, g).backward(dy)
        return torch.cat([p.grad.flatten() for p in m.parameters()])

    g1, g2, g12 = grads_for(dy1), grads_for(dy2), grads_for(2.0 * dy1 - 0.5 * dy2)
    assert rel_l2(g12.cpu(), (2.0 * g1 - 0.5 * g2).cpu()) <= 1e-5


def test_inference_and_training_workspaces_agree():
    """save_for_backward=0 (recycled buffers, infer.py path) must give the same waveform as the training path."""
    from vcvits_b200 import Generator
    sd = O.seeded_state_dict(O.TINY_CFG, 9, gain=1.5)
    m = Generator(**O.TINY_CFG, mode="fp32")
    m.load_state_dict(sd)
    m = m.cuda()
    x = torch.randn(2, 16, 50, device="cuda")
    g = torch.randn(2, 8, 1, device="cuda")
    with torch.no_grad():
        y_inf = m(x, g)
    y_tr = m(x, g)
    assert y_tr.requires_grad
    assert torch.equal(y_inf, y_tr.detach())
    y_host = m.synthesize_host(x.cpu(), g.cpu())
    assert torch.equal(y_host, y_inf.cpu())


def test_long_utterance_inference_bf16_and_fp32():
    """infer.py shape (configs[3], scaled): 938-frame (10 s) latent, forward only, recycled workspaces."""
    from oracle import mel_oracle as M
    sd = O.seeded_state_dict(O.BASE48K_CFG, 1234, gain=1.2)
    torch.manual_seed(42)
    x, g = torch.randn(1, 128, 938), torch.randn(1, 256, 1)
    y_ref, _ = oracle_run(O.BASE48K_CFG, sd, x, g, dtype=torch.float32)
    from vcvits_b200 import Generator
    for mode in ("fp32", "bf16"):
        m = Generator(**O.BASE48K_CFG, mode=mode)
        m.load_state_dict(sd)
        m = m.cuda()
        with torch.no_grad():
            y = m(x.cuda(), g.cuda()).cpu()
        assert y.shape == (1, 1, 938 * 512)
        err = float((y - y_ref).abs().max())
        if mode == "fp32":
            assert err <= 1e-4 and err <= max(1e-5 * float(y_ref.abs().max()), 5e-7), err
        else:
            assert err <= 1e-2, err
            assert M.log_mel_l1_relative(y[:, 0], y_ref[:, 0], num_mels=128) <= 0.01
        del m


@pytest.mark.parametrize("cfg_name", ["SMALL_CFG", "SMALL2_CFG"])
def test_tensor_core_path_against_ffma_path(cfg_name):
    """bf16 mode (tcgen05 kernels for every conv / wgrad of this config) vs fp32 FFMA mode on the GPU and vs the
    oracle: forward within bf16 noise, parameter-gradient direction preserved.  Covers ResBlock1 and ResBlock2."""
    cfg = getattr(O, cfg_name)
    sd = O.seeded_state_dict(cfg, 77, gain=1.3)
    torch.manual_seed(4)
    x, g = torch.randn(3, 64, 45), torch.randn(3, 16, 1)
    dy = torch.randn(3, 1, 45 * 16)
    y16, g16, m = b200_run(cfg, sd, x, g, dy, mode="bf16")
    assert any("tcgen05" in s for s in m.layer_paths())
    y32, g32, _ = b200_run(cfg, sd, x, g, dy, mode="fp32")
    y_ref, gref = oracle_run(cfg, sd, x, g, dy)
    assert float((y32.double() - y_ref).abs().max()) <= 1e-4
    assert float((y16.double() - y_ref).abs().max()) <= 1e-2
    a = torch.cat([g16[n].flatten() for n in sorted(gref)]).double()
    b = torch.cat([gref[n].flatten() for n in sorted(gref)]).double()
    assert float(torch.dot(a, b) / (a.norm() * b.norm())) >= 0.99


def test_error_paths_return_messages():
    """C-ABI error behaviour: non-zero return + message, nothing aborts."""
    import ctypes as C
    from vcvits_b200 import Generator, _lib
    lib = _lib.load()
    bad = Generator(**dict(O.TINY_CFG, upsample_initial_channel=24))  # 24 -> 12 -> 6 channels: not multiples of 8
    with pytest.raises(RuntimeError, match="multiple of"):
        bad.cuda()(torch.zeros(1, 16, 4, device="cuda"))
    m = Generator(**O.TINY_CFG, mode="fp32").cuda()
    plan = m._plan_for(torch.device("cuda", 0))
    # forward before fold
    y = torch.empty(1, 1, 32, device="cuda")
    x = torch.zeros(1, 16, 4, device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    rc = lib.vcd_forward(plan, 0, x.data_ptr(), 64, 4, 1, None, y.data_ptr(), ws.data_ptr(), ws.numel(), 1, 4, 0, None)
    assert rc != 0 and b"vcd_fold_weights" in lib.vcd_last_error()
    m(x)  # folds
    rc = lib.vcd_forward(plan, 0, x.data_ptr(), 64, 4, 1, None, y.data_ptr(), ws.data_ptr(), 16, 1, 4, 0, None)
    assert rc != 0 and b"workspace too small" in lib.vcd_last_error()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 7, 4, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 16, 4))


def test_user_side_cuda_graph_capture_of_inference():
    """The library must cooperate when the caller captures its own CUDA graph around the decoder call."""
    from vcvits_b200 import Generator
    sd = O.seeded_state_dict(O.SMALL_CFG, 3, gain=1.3)
    m = Generator(**O.SMALL_CFG, mode="bf16")
    m.load_state_dict(sd)
    m = m.cuda()
    x = torch.randn(2, 64, 20, device="cuda")
    g = torch.randn(2, 16, 1, device="cuda")
    with torch.no_grad():
        y_eager = m(x, g).clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m(x, g)  # warm-up on the side stream
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            y_static = m(x, g)
        graph.replay()
        torch.cuda.synchronize()
    assert torch.equal(y_static, y_eager)


@pytest.mark.parametrize("B,T", [(1, 1), (3, 7), (1, 130), (2, 257)])
def test_tensor_core_path_ragged_shapes(B, T):
    """tcgen05 kernels on lengths that do not fill a 128-row tile / cross tile and pad boundaries."""
    sd = O.seeded_state_dict(O.SMALL_CFG, 5, gain=1.3)
    torch.manual_seed(B * 1000 + T)
    x, g = torch.randn(B, 64, T), torch.randn(B, 16, 1)
    dy = torch.randn(B, 1, T * 16)
    y16, g16, _ = b200_run(O.SMALL_CFG, sd, x, g, dy, mode="bf16")
    y_ref, gref = oracle_run(O.SMALL_CFG, sd, x, g, dy)
    assert float((y16.double() - y_ref).abs().max()) <= 1e-2
    num = sum(float((g16[n].double() - gref[n]).pow(2).sum()) for n in gref)
    den = sum(float(gref[n].pow(2).sum()) for n in gref)
    assert (num / den) ** 0.5 <= 0.12
    assert all(torch.isfinite(v).all() for v in g16.values())


def test_many_distinct_shapes_recycle_the_graph_cache():
    """Variable-length inference: more distinct (B, T) shapes than the captured-graph cache holds (64).  The cache is
    dropped and rebuilt; results must not change and the plan must stay usable."""
    from vcvits_b200 import Generator
    sd = O.seeded_state_dict(O.SMALL_CFG, 7, gain=1.1)
    m = Generator(**O.SMALL_CFG, mode="bf16")
    m.load_state_dict(sd)
    m = m.cuda()
    torch.manual_seed(3)
    g = torch.randn(1, O.SMALL_CFG["gin_channels"], 1).cuda()
    xs = {T: torch.randn(1, O.SMALL_CFG["initial_channel"], T).cuda() for T in range(4, 74)}
    first = {}
    with torch.no_grad():
        for T, x in xs.items():          # 70 shapes: evicts the cache at least once
            first[T] = m(x, g).clone()
        for T in (4, 40, 73):            # recapture after eviction
            again = m(xs[T], g)
            assert torch.equal(again, first[T]), T
    torch.cuda.synchronize()


def test_batch_split_invariance_base_config_bf16():
    """The decoder has no cross-sample operation (modules.py:203-216): one B=6 step must equal two B=3 steps on the same
    utterances -- the waveform and dz bit for bit (no kernel's arithmetic may depend on the batch position of a tile),
    parameter gradients up to the fp32 summation order of the weight-gradient partials."""
    from vcvits_b200 import Generator
    torch.manual_seed(11)
    m = Generator(**O.BASE_CFG, mode="bf16").cuda()
    gen = torch.Generator().manual_seed(3)
    xs = torch.randn(6, 256, 20, generator=gen).cuda()
    gs = torch.randn(6, 256, 1, generator=gen).cuda()
    dys = torch.randn(6, 1, 20 * m.hop, generator=gen).cuda()

    def step(sl):
        m.zero_grad(set_to_none=True)
        x = xs[sl].clone().requires_grad_(True)
        y = m(x, gs[sl])
        y.backward(dys[sl])
        return y.detach().clone(), x.grad.clone(), [p.grad.double().clone() for p in m.parameters()]

    yf, dxf, gf = step(slice(0, 6))
    ya, dxa, ga = step(slice(0, 3))
    yb, dxb, gb = step(slice(3, 6))
    assert torch.equal(yf, torch.cat([ya, yb]))
    assert torch.equal(dxf, torch.cat([dxa, dxb]))
    for f, a, b in zip(gf, ga, gb):
        assert float((a + b - f).norm()) <= 2e-5 * float(f.norm()) + 1e-12
