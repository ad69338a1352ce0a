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
