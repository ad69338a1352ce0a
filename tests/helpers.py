"""Shared test helpers (oracle = checker only)."""
import numpy as np
import torch

from oracle import hifigan_oracle as O


def load_golden(path):
    d = np.load(path)
    sd = {k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("sd::")}
    grads = {k[6:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("grad::")}
    rest = {k: d[k] for k in d.files if "::" not in k}
    return sd, grads, rest


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def oracle_run(cfg, sd, x, g, dy=None, dtype=torch.float64):
    """Run the CPU oracle in `dtype`; returns y and (if dy given) grads of x, g and parameters."""
    m = O.build(cfg, {k: v.to(dtype) for k, v in sd.items()}, dtype=dtype)
    xx = x.detach().cpu().to(dtype).requires_grad_(dy is not None)
    gg = g.detach().cpu().to(dtype).requires_grad_(dy is not None) if g is not None else None
    y = m(xx, gg)
    if dy is None:
        return y.detach(), None
    y.backward(dy.detach().cpu().to(dtype))
    grads = {n: p.grad for n, p in m.named_parameters()}
    grads["__x__"] = xx.grad
    if gg is not None:
        grads["__g__"] = gg.grad
    return y.detach(), grads


def b200_run(cfg, sd, x, g, dy=None, mode="fp32"):
    from vcvits_b200 import Generator
    m = Generator(**cfg, mode=mode)
    m.load_state_dict(sd)
    m = m.cuda()
    xx = x.detach().cuda().float().requires_grad_(dy is not None)
    gg = g.detach().cuda().float().requires_grad_(dy is not None) if g is not None else None
    y = m(xx, gg)
    if dy is None:
        return y.detach().cpu(), None, m
    y.backward(dy.detach().cuda().float())
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters()}
    grads["__x__"] = xx.grad.detach().cpu()
    if gg is not None:
        grads["__g__"] = gg.grad.detach().cpu()
    return y.detach().cpu(), grads, m


# -------------------------------------------------------------------------------------------------------------------
# Teacher-forced local parity (oracle/bf16_emulation.py): one bf16-mode training step through the C ABI with every
# tensor the CUDA path stores read back from its workspace (forward activations after vcd_forward, gradient tensors
# after each backward segment, whose buffers are shared between stages).
# -------------------------------------------------------------------------------------------------------------------
def _ws_tensor(lib, plan, mode, ws, B, T, name):
    import ctypes as C
    off, ch, ln, pl, pr = C.c_size_t(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = lib.vcd_debug_ws_tensor(plan, mode, B, T, 1, name.encode(), C.byref(off), C.byref(ch), C.byref(ln), C.byref(pl),
                                 C.byref(pr))
    assert rc == 0, lib.vcd_last_error()
    Cn, L, rows = ch.value, ln.value, pl.value + ln.value + pr.value
    nbytes = B * Cn * rows * 2
    t = ws[off.value: off.value + nbytes].view(torch.bfloat16).view(B, Cn // 8, rows, 8)
    return t[:, :, pl.value: pl.value + L, :].permute(0, 1, 3, 2).reshape(B, Cn, L).cpu()


def b200_step_with_stored(cfg, sd, x, g, dy):
    """bf16-mode fold + forward + segment-by-segment backward through the raw C ABI.
    Returns (y, grads, stored): grads keyed like oracle_run, stored = {storage point name: bf16 [B, C, L] CPU tensor}."""
    import ctypes as C
    from vcvits_b200 import Generator, _lib
    lib = _lib.load()
    m = Generator(**cfg, mode="bf16")
    m.load_state_dict(sd)
    m = m.cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    plan, mode = m._plan_for(dev), m._mode
    params = m._ordered_params()
    m._fold_if_needed(params, force=True)
    B, _, T = x.shape
    S, NB = m.num_upsamples, m.num_kernels
    npairs = 3 if m.resblock == "1" else 2
    xd, dyd = x.cuda().float().contiguous(), dy.cuda().float().contiguous()
    gd = g.cuda().float().reshape(B, -1).contiguous() if g is not None else None
    ws_bytes = lib.vcd_workspace_bytes(plan, mode, B, T, 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    y = torch.empty(B, 1, T * m.hop, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.vcd_forward(plan, mode, xd.data_ptr(), xd.stride(0), xd.stride(1), xd.stride(2),
                               gd.data_ptr() if gd is not None else None, y.data_ptr(), ws.data_ptr(), ws_bytes, B, T, 1,
                               stream), "vcd_forward")
    torch.cuda.synchronize()
    stored = {}
    grab = lambda name: stored.__setitem__(name, _ws_tensor(lib, plan, mode, ws, B, T, name))
    grab("xin")
    for i in range(S + 1):
        grab(f"a{i}")
    for i in range(S):
        grab(f"ua{i}")
        for j in range(NB):
            for q in range(npairs):
                if m.resblock == "1":
                    grab(f"ma{i}.{j}.{q}")
                if q < npairs - 1:
                    grab(f"xa{i}.{j}.{q}")
    flat = torch.empty(m._flat_numel, device=dev)
    views = m._grad_views(flat)
    ptrs = (C.c_void_p * len(views))(*[v.data_ptr() for v in views])
    dx = torch.empty(B, m.initial_channel, T, device=dev)
    dg = torch.empty(B, m.gin_channels, device=dev) if gd is not None else None
    for seg in range(S + 1):
        _lib.check(lib.vcd_backward(plan, mode, dyd.data_ptr(), y.data_ptr(), gd.data_ptr() if gd is not None else None,
                                    dx.data_ptr(), dg.data_ptr() if dg is not None else None, ptrs, ws.data_ptr(), ws_bytes,
                                    B, T, 1 << seg, stream), "vcd_backward")
        torch.cuda.synchronize()
        if seg == S:
            break
        i = S - 1 - seg
        grab(f"Gi{i}")
        for j in range(NB):
            for q in range(npairs):
                if q > 0:
                    grab(f"Gt{i}.{j}.{q}")
                if m.resblock == "1":
                    grab(f"dm{i}.{j}.{q}")
        # the phase-packed gradient w.r.t. the upsample output: Z[b][r*cout + ch][q] = d[b][ch][q*u + r - p]
        u, k = m.upsample_rates[i], m.upsample_kernel_sizes[i]
        z = _ws_tensor(lib, plan, mode, ws, B, T, f"duz{i}")
        cout, lz = z.shape[1] // u, z.shape[2]
        p = (k - u) // 2
        lout = T * u
        for uu in m.upsample_rates[:i]:
            lout *= uu
        stored[f"duz{i}"] = z.view(B, u, cout, lz).permute(0, 2, 3, 1).reshape(B, cout, lz * u)[:, :, p: p + lout].contiguous()
        if i == 0:
            grab("d0")
    grads = {n: v.detach().cpu() for n, v in zip(m._names, views)}
    grads["__x__"] = dx.cpu()
    if dg is not None:
        grads["__g__"] = dg.cpu().reshape(g.shape)
    return y.cpu(), grads, stored, m
