"""GPU: the drop-in boundary -- AMP / autocast, remove_weight_norm, DDP-style gradient sync,
stale-state protections of the Python module."""
import os
import socket

import pytest
import torch

from oracle import hifigan_oracle as O
from tests.helpers import b200_run, oracle_run, rel_l2

pytestmark = pytest.mark.gpu


def _small(mode="fp32", seed=3, gain=1.3):
    from vcvits_b200 import Generator
    sd = O.seeded_state_dict(O.SMALL_CFG, seed, gain=gain)
    m = Generator(**O.SMALL_CFG, mode=mode)
    m.load_state_dict(sd)
    return m.cuda(), sd


def test_fp16_autocast_with_grad_scaler():
    """Lightning AMP (train.py:104-106, precision 16): the decoder is called inside autocast with fp16 latents and the
    loss is scaled by a GradScaler.  Output is fp32, gradients come back in the inputs' dtypes, unscaled gradients
    match the fp64 oracle (fp32 mode: the only error is the fp16 rounding of the inputs, which the oracle sees too)."""
    m, sd = _small("fp32")
    torch.manual_seed(0)
    x16 = torch.randn(2, 64, 24, device="cuda").half().requires_grad_(True)
    g16 = torch.randn(2, 16, 1, device="cuda").half().requires_grad_(True)
    dy = torch.randn(2, 1, 24 * 16, device="cuda")
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    opt = torch.optim.SGD(m.parameters(), lr=0.0)
    with torch.autocast("cuda", dtype=torch.float16):
        y = m(x16, g16)
        loss = (y.float() * dy).sum()
    assert y.dtype == torch.float32
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    assert x16.grad.dtype == torch.float16 and g16.grad.dtype == torch.float16
    y_ref, gref = oracle_run(O.SMALL_CFG, sd, x16.detach().float().cpu(), g16.detach().float().cpu(), dy.cpu())
    assert float((y.detach().cpu().double() - y_ref).abs().max()) <= 1e-4
    for n, p in m.named_parameters():   # fp32 mode vs fp64: PyTorch fp32 itself reaches ~2e-3 on a few tensors
        assert rel_l2(p.grad.cpu(), gref[n]) <= 1e-2, n
    num = sum(float((p.grad.cpu().double() - gref[n]).pow(2).sum()) for n, p in m.named_parameters())
    den = sum(float(gref[n].pow(2).sum()) for n, _ in m.named_parameters())
    assert (num / den) ** 0.5 <= 1e-3
    assert rel_l2(x16.grad.float().cpu() / 1024.0, gref["__x__"]) <= 2e-3     # fp16 storage of the (scaled) gradient
    # bf16 mode under the same autocast region runs too (its own arithmetic mode; autocast is disabled inside)
    m16, _ = _small("bf16")
    with torch.autocast("cuda", dtype=torch.float16):
        y16 = m16(x16.detach(), g16.detach())
    assert y16.dtype == torch.float32 and float((y16.cpu().double() - y_ref).abs().max()) <= 1e-2


def test_remove_weight_norm_matches_reference_convention():
    """modules.py:218-222: after remove_weight_norm the module holds `weight` (+ `bias`) and no weight_g / weight_v; the
    waveform is unchanged and the baked model is still differentiable (plain weight gradients)."""
    import warnings
    m, sd = _small("fp32")
    torch.manual_seed(1)
    x, g = torch.randn(2, 64, 20, device="cuda"), torch.randn(2, 16, 1, device="cuda")
    with torch.no_grad():
        y_before = m(x, g).clone()
    m.remove_weight_norm()
    keys = list(m.state_dict().keys())
    assert not any(k.endswith(("weight_g", "weight_v")) for k in keys)
    assert "ups.0.weight" in keys and "resblocks.0.convs1.0.weight" in keys and "conv_pre.weight" in keys
    # the same keys, in the same order, as torch's remove_weight_norm leaves on the oracle
    ref = O.build(O.SMALL_CFG, sd)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for mod in list(ref.ups) + [c for rb in ref.resblocks for c in list(rb.convs1) + list(rb.convs2)]:
            torch.nn.utils.remove_weight_norm(mod)
    assert keys == list(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert torch.allclose(m.state_dict()[k].cpu(), v, rtol=1e-6, atol=1e-8), k
    with torch.no_grad():
        y_after = m(x, g)
    assert float((y_after - y_before).abs().max()) <= 1e-6
    # still trainable: gradients w.r.t. the baked weights equal autograd through the baked oracle
    dy = torch.randn_like(y_after)
    m(x, g).backward(dy)
    ref64 = ref.double()
    ref64(x.cpu().double(), g.cpu().double()).backward(dy.cpu().double())
    for (n, p), (_, q) in zip(m.named_parameters(), ref64.named_parameters()):
        assert rel_l2(p.grad.cpu(), q.grad) <= 1e-3, n
    # a checkpoint saved before the call no longer loads (as in the reference)
    with pytest.raises(RuntimeError):
        m.load_state_dict(sd)


def test_in_place_parameter_edits_are_seen():
    """`p.data` edits do not bump `_version`.  Training-mode forwards always re-fold; inference needs
    invalidate_weights()."""
    m, sd = _small("fp32")
    x, g = torch.randn(1, 64, 12, device="cuda"), torch.randn(1, 16, 1, device="cuda")
    y0 = m(x, g).detach().clone()                    # training-mode forward (parameters require grad)
    m.conv_post.weight.data.mul_(0.5)                # invisible to the version counter
    y1 = m(x, g).detach()
    assert float((y1 - y0).abs().max()) > 1e-4
    with torch.no_grad():
        y2 = m(x, g).clone()
        m.resblocks[0].convs1[0].weight_g.data.mul_(2.0)
        m.invalidate_weights()
        y3 = m(x, g)
    assert float((y3 - y2).abs().max()) > 1e-6
    # replacing a Parameter object without _apply is picked up
    new_w = torch.nn.Parameter(m.conv_pre.weight.detach() * 0.0)
    m.conv_pre.weight = new_w
    y4 = m(x, g)
    y4.sum().backward()
    assert new_w.grad is not None and float(new_w.grad.abs().sum()) > 0


def test_refold_between_forward_and_backward_raises():
    m, _ = _small("fp32")
    x, g = torch.randn(1, 64, 12, device="cuda"), torch.randn(1, 16, 1, device="cuda")
    y = m(x, g)
    m(x, g)                                          # second training-mode forward re-folds the plan's weights
    with pytest.raises(RuntimeError, match="re-folded"):
        y.sum().backward()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    from vcvits_b200 import Generator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # both ranks share cuda:0 (NCCL needs one GPU per rank)
    try:
        torch.cuda.set_device(0)
        sd = O.seeded_state_dict(O.SMALL_CFG, 9, gain=1.3)
        m = Generator(**O.SMALL_CFG, mode="fp32")
        m.load_state_dict(sd)
        m = m.cuda()
        m.set_gradient_sync(dist.group.WORLD)
        torch.manual_seed(5)
        x, g, dy = torch.randn(4, 64, 20), torch.randn(4, 16, 1), torch.randn(4, 1, 320)
        lo, hi = rank * 2, rank * 2 + 2
        m(x[lo:hi].cuda(), g[lo:hi].cuda()).backward(dy[lo:hi].cuda())
        out[rank] = {n: p.grad.cpu() for n, p in m.named_parameters()}
    finally:
        dist.destroy_process_group()


def test_gradient_sync_equals_single_process_full_batch():
    """b5 / train.py:99-100: two ranks, half the batch each, segment-wise all-reduce inside backward == the full-batch
    gradient / world (DDP averaging)."""
    import torch.multiprocessing as mp
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    m, sd = _small("fp32", seed=9)
    torch.manual_seed(5)
    x, g, dy = torch.randn(4, 64, 20), torch.randn(4, 16, 1), torch.randn(4, 1, 320)
    m(x.cuda(), g.cuda()).backward(dy.cuda())
    full = {n: p.grad.cpu() / 2 for n, p in m.named_parameters()}
    for r in (0, 1):
        for n in full:
            assert rel_l2(out[r][n], full[n]) <= 1e-5, (r, n)
    for n in full:
        assert torch.equal(out[0][n], out[1][n]), n


def test_sliced_forward_equals_slice_segments_then_decode():
    """rand_slice_segments' gather folded into the decoder input load (commons.py:48-64).  Same
    waveform bit for bit, same parameter gradients, and dz = the scatter autograd derives from slice_segments."""
    from oracle import synth_parts as S
    m, _ = _small("fp32")
    torch.manual_seed(2)
    z = torch.randn(3, 64, 50, device="cuda", requires_grad=True)
    g = torch.randn(3, 16, 1, device="cuda")
    ids = torch.tensor([0, 17, 38], device="cuda")
    dy = torch.randn(3, 1, 12 * 16, device="cuda")
    y_ref = m(S.slice_segments(z, ids, 12), g)
    y_ref.backward(dy)
    dz_ref, grads_ref = z.grad.clone(), [p.grad.clone() for p in m.parameters()]
    z.grad = None
    m.zero_grad(set_to_none=True)
    y = m.forward_sliced(z, ids, 12, g)
    assert torch.equal(y, y_ref)
    y.backward(dy)
    assert z.grad.shape == z.shape and torch.equal(z.grad, dz_ref)
    for p, r in zip(m.parameters(), grads_ref):      # split partials combine with fp32 atomics: equal up to summation order
        assert rel_l2(p.grad.cpu(), r.cpu()) <= 1e-5
    with pytest.raises(ValueError):
        m.forward_sliced(z, ids[:2], 12, g)


def test_fused_parameter_gradients_match_autograd_leaves():
    """b5: the default backward writes every parameter's ``.grad`` itself (one flat buffer) instead of returning 233
    tensors to autograd; same values as the classic leaf-per-parameter path, ``+=`` when a gradient already exists,
    nothing for frozen parameters, and ``torch.autograd.grad`` w.r.t. parameters through the classic path."""
    m, _ = _small("fp32")
    torch.manual_seed(4)
    x = torch.randn(2, 64, 20, device="cuda", requires_grad=True)
    g = torch.randn(2, 16, 1, device="cuda", requires_grad=True)
    dy = torch.randn(2, 1, 320, device="cuda")
    assert m.fused_param_grads
    m(x, g).backward(dy)
    fused = [p.grad.clone() for p in m.parameters()]
    dx_f, dg_f = x.grad.clone(), g.grad.clone()
    assert all(p.grad.shape == p.shape and p.grad.is_contiguous() for p in m.parameters())
    # accumulate into existing gradients (no zero_grad in between): exactly twice the gradient
    m(x, g).backward(dy)
    for p, r in zip(m.parameters(), fused):
        assert rel_l2(p.grad.cpu(), (2 * r).cpu()) <= 1e-5
    # classic path
    m.zero_grad(set_to_none=True)
    x.grad = g.grad = None
    m.fused_param_grads = False
    m(x, g).backward(dy)
    for p, r in zip(m.parameters(), fused):
        assert rel_l2(p.grad.cpu(), r.cpu()) <= 1e-5
    assert rel_l2(x.grad.cpu(), dx_f.cpu()) <= 1e-5 and rel_l2(g.grad.cpu(), dg_f.cpu()) <= 1e-5
    some = [m.conv_pre.weight, m.ups[0].weight_g]
    got = torch.autograd.grad(m(x, g), some, dy)
    assert rel_l2(got[0].cpu(), m.conv_pre.weight.grad.cpu()) <= 1e-5
    # frozen parameters: no gradient, the rest unchanged; latents only: no parameter gradient at all
    m.fused_param_grads = True
    m.zero_grad(set_to_none=True)
    m.conv_pre.weight.requires_grad_(False)
    m(x, g).backward(dy)
    assert m.conv_pre.weight.grad is None and m.conv_pre.bias.grad is not None
    for p in m.parameters():
        p.requires_grad_(False)
    m.zero_grad(set_to_none=True)
    x.grad = None
    m(x, g).backward(dy)
    assert all(p.grad is None for p in m.parameters())
    assert rel_l2(x.grad.cpu(), dx_f.cpu()) <= 1e-5
    # inference under no_grad still works with the fused default
    with torch.no_grad():
        y = m(x, g)
    assert not y.requires_grad


@pytest.mark.parametrize("mode,cfg_name,B,T", [("bf16", "BASE_CFG", 16, 32), ("bf16", "BASE48K_CFG", 4, 32), ("fp32", "SMALL_CFG", 3, 40)])
def test_deterministic_mode_is_bit_identical_run_to_run(mode, cfg_name, B, T):
    """``Generator.deterministic = True`` (vcd_set_deterministic): every gradient element receives at most two atomic
    contributions, so two runs of the same step agree BIT FOR BIT in all 233 parameter gradients, dz and dg -- at the
    benchmarked configuration.  The default (arrival-order reductions) agrees with it to the fp32 summation order."""
    from oracle import hifigan_oracle as O
    from vcvits_b200 import Generator
    cfg = getattr(O, cfg_name)
    torch.manual_seed(3)
    m = Generator(**cfg, mode=mode).cuda()
    x = torch.randn(B, cfg["initial_channel"], T, device="cuda")
    g = torch.randn(B, cfg["gin_channels"], 1, device="cuda")
    dy = torch.randn(B, 1, T * m.hop, device="cuda")

    def step():
        for p in m.parameters():
            p.grad = None
        xx, gg = x.clone().requires_grad_(True), g.clone().requires_grad_(True)
        m(xx, gg).backward(dy)
        torch.cuda.synchronize()
        out = {n: p.grad.clone() for n, p in m.named_parameters()}
        out["__dx"], out["__dg"] = xx.grad.clone(), gg.grad.clone()
        return out

    default = step()
    m.deterministic = True
    runs = [step() for _ in range(3)]
    m.deterministic = False
    for r in runs[1:]:
        for k in runs[0]:
            assert torch.equal(r[k], runs[0][k]), k
    for k, v in default.items():
        a, b = v.double(), runs[0][k].double()
        assert float((a - b).norm()) <= 2e-4 * float(b.norm()) + 1e-12, k   # weight_g gradients cancel heavily
