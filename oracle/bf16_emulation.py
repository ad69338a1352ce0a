"""bf16-emulating CPU oracle for the decoder's ``mode="bf16"`` path.

TEST INFRASTRUCTURE ONLY (see the header of ``oracle/hifigan_oracle.py``): nothing under ``vcvits_b200/`` may
import this file.

The plain oracle (``hifigan_oracle.OracleGenerator``) computes the reference arithmetic in fp32 / fp64.  The
CUDA path's bf16 mode rounds to bf16 at a small number of *storage points*  and accumulates in
fp32 everywhere else.  This module restates the same decoder (ResBlock bodies: vits/model/modules.py:186-247,
glue: SURVEY.md Appendix A) with a rounding to bf16 inserted at exactly those points, so that the bf16-mode
waveform and EVERY parameter gradient can be compared per tensor at bf16-independent tolerances:

forward storage points (``round_ste``: round in forward, identity in backward)
  * the latent ``x`` and every tensor-core weight ``w = v * (g / ||v||)`` (fold in fp32, then bf16);
    ``conv_post.weight``, ``cond.*`` and all biases stay fp32;
  * every stored activation is ``bf16(leaky_relu(.))``: conv_pre output, upsample outputs, the ResBlock
    mid tensors, the residual stream between pairs, the stage outputs (mean over branches fused with the next
    leaky_relu, slope 0.1 -- 0.01 before conv_post);
  * the raw residual stream is NOT stored: ``x = xt + x`` (modules.py:213) uses
    ``inv_lrelu(bf16(lrelu(x)))`` (``StoreAct`` below);
  * running sums over ResBlock branches stay fp32.
backward storage points (``grad_round``: identity in forward, rounds the incoming gradient to bf16)
  * gradient w.r.t. the conv_pre output, every upsample output (the phase-packed tensor), every ResBlock mid
    pre-activation, the residual stream between pairs, and every stage's branch sum;
  * weight / bias gradients and the weight-norm backward accumulate in fp32 from those bf16 operands.

Everything between storage points is evaluated in ``dtype`` (fp64 by default; fp32 for the largest shapes).

Two ways to use it:
  * free running (``run``): an end-to-end bf16-mode reference.  Rounding is chaotic -- a single flipped bf16
    rounding upstream re-rolls thousands of roundings downstream -- so two correct implementations that differ
    only in fp32 summation order still differ by the network's intrinsic bf16 noise (a few 1e-3 of the waveform
    amplitude, a few percent on small gradient tensors).  Good for noise-level bounds only.
  * teacher forced (``run(..., stored=...)``): at every storage point the value the CUDA path actually stored
    (read back from its workspace) is compared with the value computed here FROM THE CUDA PATH'S OWN STORED
    OPERANDS and then substituted.  Each comparison then isolates one kernel launch (one convolution + fused
    epilogue, forward or data gradient), and the parameter gradients autograd derives from the substituted
    tensors isolate each weight-gradient launch plus the weight-norm backward.  Those per-tensor comparisons
    hold at ~1e-3, far below bf16 noise.  Storage point names: ``xin``, ``a<i>``, ``ua<i>``, ``ma<i>.<j>.<q>``,
    ``xa<i>.<j>.<q>`` (values) and ``d0``, ``duz<i>``, ``dm<i>.<j>.<q>``, ``Gt<i>.<j>.<q>``, ``Gi<i>`` (gradients),
    i = stage, j = ResBlock branch, q = pair -- the names of ``vcd_debug_ws_tensor``.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
from torch.nn import functional as F

SLOPE = float(np.float32(0.1))          # kSlope in vcd_api.cu (LRELU_SLOPE, modules.py:16) as the GPU's fp32 constant
INV_SLOPE = 10.0                        # kInvSlope
FINAL_SLOPE = float(np.float32(0.01))   # kFinalSlope (F.leaky_relu default before conv_post)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    # the GPU rounds fp32 values (RNE); go through fp32 first so an fp64 oracle rounds the same way
    return t.float().to(torch.bfloat16).to(t.dtype)


class Stored:
    """Teacher forcing: tensors the CUDA path stored (name -> [B, C, L] tensor) and the per-point comparison report."""

    def __init__(self, tensors: Optional[Dict[str, torch.Tensor]] = None, record: bool = False):
        self.tensors = tensors if tensors is not None else {}
        self.record = record          # True: capture this run's own storage points instead of substituting
        self.report: Dict[str, dict] = {}

    def check_and_substitute(self, name: Optional[str], computed: torch.Tensor) -> torch.Tensor:
        if name is None:
            return computed
        if self.record:
            self.tensors[name] = computed.detach().to(torch.bfloat16)
            return computed
        if name not in self.tensors:
            return computed
        got = self.tensors[name].to(computed.dtype)
        assert got.shape == computed.shape, (name, tuple(got.shape), tuple(computed.shape))
        diff = (got - computed).double()
        den = float(computed.double().norm())
        # one bf16 ulp of the element's own magnitude (a flipped rounding is exactly one ulp)
        ulp = torch.maximum(computed.abs(), got.abs()).double() * 2.0 ** -7
        self.report[name] = {"rel_l2": float(diff.norm()) / (den if den > 0 else 1.0),
                             "max_abs": float(diff.abs().max()),
                             "max_ulps": float((diff.abs() / ulp.clamp_min(1e-30)).max()),
                             "frac_differ": float((diff != 0).double().mean())}
        return got


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stored, name):
        r = _bf16(x)
        return stored.check_and_substitute(name, r) if stored is not None else r

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class _GradRound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stored, name):
        ctx.stored, ctx.name = stored, name
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        r = _bf16(g)
        if ctx.stored is not None:
            r = ctx.stored.check_and_substitute(ctx.name, r)
        return r, None, None


class _StoreAct(torch.autograd.Function):
    """Stored activation t = bf16(lrelu(x)) plus the residual stream recovered from it, x_rec = inv_lrelu(t).
    Backward: d x = mask(t) * d t + d x_rec  (the identity path of ``x = xt + x`` passes the gradient unchanged)."""

    @staticmethod
    def forward(ctx, x, slope, inv, stored, name):
        t = _bf16(torch.where(x > 0, x, x * slope))
        if stored is not None:
            t = stored.check_and_substitute(name, t)
        ctx.save_for_backward(t)
        ctx.slope = slope
        return t, torch.where(t > 0, t, t * inv)

    @staticmethod
    def backward(ctx, gt, gx):
        (t,) = ctx.saved_tensors
        return gt * torch.where(t > 0, 1.0, ctx.slope).to(gt.dtype) + gx, None, None, None, None


class _ActStore(torch.autograd.Function):
    """Stored activation t = bf16(lrelu(x, slope)).  Backward: d x = d t * (t > 0 ? 1 : slope) -- the mask is taken from
    the STORED tensor, exactly like the CUDA epilogues (on a knife-edge |x| ~ 1e-7 the sign of a recomputed x may differ)."""

    @staticmethod
    def forward(ctx, x, slope, stored, name):
        t = _bf16(torch.where(x > 0, x, x * slope))
        if stored is not None:
            t = stored.check_and_substitute(name, t)
        ctx.save_for_backward(t)
        ctx.slope = slope
        return t

    @staticmethod
    def backward(ctx, g):
        (t,) = ctx.saved_tensors
        return g * torch.where(t > 0, 1.0, ctx.slope).to(g.dtype), None, None, None


def act_store(x, slope, stored=None, name=None):
    return _ActStore.apply(x, slope, stored, name)


def round_ste(x, stored=None, name=None):
    return _RoundSTE.apply(x, stored, name)


def grad_round(x, stored=None, name=None):
    return _GradRound.apply(x, stored, name)


def _lrelu(x, slope):
    return torch.where(x > 0, x, x * slope)


def _fold(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """w = v * (g / ||v||) over all dims but 0 (old-style weight_norm, modules.py:10), then bf16."""
    norm = v.flatten(1).norm(dim=1).reshape(g.shape)
    return round_ste(v * (g / norm))


class EmulatedGenerator:
    """Functional bf16-mode emulation of the decoder.  ``params``: state_dict (fp32 tensors, reference keys)."""

    def __init__(self, cfg: dict, state_dict: Dict[str, torch.Tensor], dtype=torch.float64,
                 stored: Optional[Stored] = None):
        self.cfg = cfg
        self.dtype = dtype
        self.stored = stored
        self.params = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in state_dict.items()}
        self.num_kernels = len(cfg["resblock_kernel_sizes"])
        self.rb1 = str(cfg["resblock"]) == "1"

    # -- building blocks -------------------------------------------------------------------------
    def _wn_conv(self, name: str, t: torch.Tensor, dilation: int) -> torch.Tensor:
        p = self.params
        w = _fold(p[name + ".weight_v"], p[name + ".weight_g"])
        k = w.shape[-1]
        return F.conv1d(t, w, p[name + ".bias"], 1, dilation * (k - 1) // 2, dilation)

    def _resblock(self, i: int, j: int, t0: torch.Tensor, x0: torch.Tensor, dilations) -> torch.Tensor:
        """t0 = stored lrelu(x) (bf16 values), x0 = residual stream recovered from it.  Returns the raw block output."""
        name = f"resblocks.{i * self.num_kernels + j}"
        st = self.stored
        t, x = t0, x0
        npairs = 3 if self.rb1 else 2
        for q in range(npairs):
            if self.rb1:
                h1 = grad_round(self._wn_conv(f"{name}.convs1.{q}", t, dilations[q]), st, f"dm{i}.{j}.{q}")
                mid = act_store(h1, SLOPE, st, f"ma{i}.{j}.{q}")
                h2 = self._wn_conv(f"{name}.convs2.{q}", mid, 1)
            else:
                h2 = self._wn_conv(f"{name}.convs.{q}", t, dilations[q])
            x = x + h2
            if q < npairs - 1:
                t, x = _StoreAct.apply(grad_round(x, st, f"Gt{i}.{j}.{q + 1}"), SLOPE, INV_SLOPE, st, f"xa{i}.{j}.{q}")
        return x

    # -- forward ---------------------------------------------------------------------------------
    def __call__(self, x: torch.Tensor, g: Optional[torch.Tensor] = None) -> torch.Tensor:
        cfg, p, st = self.cfg, self.params, self.stored
        h = F.conv1d(round_ste(x, st, "xin"), round_ste(p["conv_pre.weight"]), p["conv_pre.bias"], 1, 3)
        if g is not None:
            h = h + F.conv1d(g, p["cond.weight"], p["cond.bias"])
        a = act_store(grad_round(h, st, "d0"), SLOPE, st, "a0")
        nb = self.num_kernels
        inv_nb = float(np.float32(1.0) / np.float32(nb))
        n_up = len(cfg["upsample_rates"])
        for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
            w = _fold(p[f"ups.{i}.weight_v"], p[f"ups.{i}.weight_g"])
            up = F.conv_transpose1d(a, w, p[f"ups.{i}.bias"], u, (k - u) // 2)
            t0, x0 = _StoreAct.apply(grad_round(up, st, f"duz{i}"), SLOPE, INV_SLOPE, st, f"ua{i}")
            acc = None
            for j in range(nb):
                out = self._resblock(i, j, t0, x0, cfg["resblock_dilation_sizes"][j])
                acc = out if acc is None else acc + out
            acc = grad_round(acc, st, f"Gi{i}")
            a = act_store(acc * inv_nb, FINAL_SLOPE if i == n_up - 1 else SLOPE, st, f"a{i + 1}")
        return torch.tanh(F.conv1d(a, p["conv_post.weight"], None, 1, 3))


def run(cfg: dict, state_dict, x: torch.Tensor, g: Optional[torch.Tensor], dy: Optional[torch.Tensor] = None,
        dtype=torch.float64, stored: Optional[Stored] = None):
    """Returns (y, grads) like tests.helpers.oracle_run: grads keyed by parameter name plus ``__x__`` / ``__g__``.
    With ``stored`` the run is teacher forced and ``stored.report`` holds the per-storage-point comparison."""
    m = EmulatedGenerator(cfg, state_dict, dtype, stored)
    xx = x.detach().cpu().to(dtype).requires_grad_(dy is not None)
    gg = g.detach().cpu().to(dtype).requires_grad_(dy is not None) if g is not None else None
    y = m(xx, gg)
    if dy is None:
        return y.detach(), None
    y.backward(dy.detach().cpu().to(dtype))
    grads = {n: (t.grad if t.grad is not None else torch.zeros_like(t)) for n, t in m.params.items()}
    grads["__x__"] = xx.grad
    if gg is not None:
        grads["__g__"] = gg.grad
    return y.detach(), grads
