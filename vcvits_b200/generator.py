"""Drop-in replacement for the reference's ``Generator`` (the HiFi-GAN decoder behind ``net_g.dec``).

Mirrors the reference interface for this path:
  * constructor arguments  -- vits/model/synthesizers/synthesizer_tts.py:71-78
  * ``forward(x, g=None)``  -- call sites synthesizer_tts.py:140,166,176 and synthesizer_svc.py:87,108,118
  * parameter names/shapes -- old-style ``weight_norm`` layout (``weight_g``/``weight_v``/``bias``) used by
    vits/model/modules.py:10,190-199,229-230, i.e. checkpoints saved under ``net_g.dec.*`` load unchanged
  * ``remove_weight_norm()`` -- convention of modules.py:218-222

All arithmetic runs in the CUDA library behind the C ABI of ``include/vcd.h`` through ONE
``torch.autograd.Function``; PyTorch only owns memory, streams and (optionally) the process group.  There is
no CPU or eager fallback: calling the module on a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
import warnings
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from .dist import SegmentReducer

__all__ = ["Generator", "LRELU_SLOPE"]

LRELU_SLOPE = 0.1  # vits/model/modules.py:16


class _ConvParams(nn.Module):
    """Parameter holder for a plain Conv1d (``weight``[, ``bias``])."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        super().__init__()
        self.weight = nn.Parameter(weight)
        if bias is not None:
            self.bias = nn.Parameter(bias)
        else:
            self.register_parameter("bias", None)


class _WeightNormParams(nn.Module):
    """Parameter holder with the old-style weight_norm layout: ``bias``, ``weight_g``, ``weight_v``."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        self.bias = nn.Parameter(bias)
        norm = weight.reshape(weight.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (weight.dim() - 1)))
        self.weight_g = nn.Parameter(norm)
        self.weight_v = nn.Parameter(weight)

    def bake(self) -> None:
        """``torch.nn.utils.remove_weight_norm`` on this holder: ``weight = g * v / ||v||`` becomes a plain
        parameter (registered after ``bias``, like the reference), ``weight_g`` / ``weight_v`` disappear."""
        if "weight_v" not in self._parameters:
            return
        v, g = self.weight_v.detach(), self.weight_g.detach()
        norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        w = v * (g / norm)
        del self._parameters["weight_g"]
        del self._parameters["weight_v"]
        self.weight = nn.Parameter(w)


class _ResBlock1Params(nn.Module):
    def __init__(self, convs1: Sequence[nn.Module], convs2: Sequence[nn.Module]):
        super().__init__()
        self.convs1 = nn.ModuleList(convs1)
        self.convs2 = nn.ModuleList(convs2)


class _ResBlock2Params(nn.Module):
    def __init__(self, convs: Sequence[nn.Module]):
        super().__init__()
        self.convs = nn.ModuleList(convs)


def _default_conv_init(cout: int, cin: int, k: int, transposed: bool = False):
    """PyTorch's default Conv1d / ConvTranspose1d init (kaiming_uniform(a=sqrt(5)) + uniform bias), drawn in the
    same RNG order as constructing the torch module, so a seeded construction matches the reference's."""
    shape = (cin, cout, k) if transposed else (cout, cin, k)
    w = torch.empty(shape)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    fan_in = shape[1] * k
    bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
    b = torch.empty(cout).uniform_(-bound, bound)
    return w, b


def _consume_init_weights_rng(mods: Sequence[_WeightNormParams]) -> None:
    # vits/commons.py:8-11 init_weights is a no-op on the effective weights of weight-normed convs
    #  but advances the RNG; mirror that so seeded constructions agree with the reference.
    for m in mods:
        torch.empty_like(m.weight_v).normal_(0.0, 0.01)


class _DecoderFunction(torch.autograd.Function):
    """forward = vcd_forward, backward = vcd_backward (+ optional overlapped gradient all-reduce)."""

    @staticmethod
    def forward(ctx, module: "Generator", need_grad: bool, x: torch.Tensor, g: Optional[torch.Tensor],
                starts: Optional[torch.Tensor], seg_frames: int, anchor: Optional[torch.Tensor], *params: torch.Tensor):
        # `anchor` (fused parameter gradients, the default): the parameters are NOT autograd inputs -- one dummy leaf makes
        # the output require grad, backward writes `.grad` of every parameter itself (one flat buffer, one view call)
        # instead of returning 233 tensors through 233 AccumulateGrad nodes (1.3 ms of host time per step on base.json).
        lib = _lib.load()
        fused = len(params) == 0
        if fused:
            params = module._call_params
        B, _, T = x.shape
        T_full = T
        if starts is not None:   # sliced call: x is the full-length latent, the decoder runs on [starts, starts + seg_frames)
            T = int(seg_frames)
        plan = module._plan_for(x.device)
        mode = module._mode
        module._fold_if_needed(params, force=need_grad)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        xf = x if x.dtype == torch.float32 else x.float()
        gf = None
        if g is not None:
            gf = g.reshape(B, -1).float().contiguous()
        ws_bytes = lib.vcd_workspace_bytes(plan, mode, B, T, 1 if need_grad else 0)
        ws = module._take_workspace(ws_bytes, x.device) if need_grad else module._workspace(ws_bytes, x.device, cache=True)
        y = torch.empty((B, 1, T * module.hop), dtype=torch.float32, device=x.device)
        if starts is None:
            _lib.check(lib.vcd_forward(plan, mode, xf.data_ptr(), xf.stride(0), xf.stride(1), xf.stride(2),
                                       gf.data_ptr() if gf is not None else None, y.data_ptr(), ws.data_ptr(),
                                       ws_bytes, B, T, 1 if need_grad else 0, stream), "vcd_forward")
        else:
            _lib.check(lib.vcd_forward_sliced(plan, mode, xf.data_ptr(), xf.stride(0), xf.stride(1), xf.stride(2),
                                              starts.data_ptr(), gf.data_ptr() if gf is not None else None, y.data_ptr(),
                                              ws.data_ptr(), ws_bytes, B, T, 1 if need_grad else 0, stream),
                       "vcd_forward_sliced")
        if need_grad:
            ctx.module = module
            ctx.starts, ctx.T_full = starts, T_full
            ctx.fold_serial = module._fold_serial
            ctx.ws = ws
            ctx.ws_bytes = ws_bytes
            ctx.shape = (B, T)
            ctx.has_g = g is not None
            ctx.fused = fused
            ctx.params = params
            ctx.g_shape = g.shape if g is not None else None
            ctx.x_dtype, ctx.g_dtype = x.dtype, (g.dtype if g is not None else None)
            ctx.save_for_backward(y, gf if gf is not None else y)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        with torch.cuda.device(dy.device):
            return _DecoderFunction._backward(ctx, dy)

    @staticmethod
    def _backward(ctx, dy: torch.Tensor):
        lib = _lib.load()
        module: Generator = ctx.module
        if ctx.ws is None:
            raise RuntimeError("vcvits_b200.Generator: backward through the same forward twice is not supported "
                               "(the activation workspace is recycled after the first backward)")
        if ctx.fold_serial != module._fold_serial:
            # the packed weights / norms live in the plan, not in the autograd graph: a re-fold between this forward and
            # its backward (in-place parameter update, set_mode, a second training forward) would silently mix new
            # weights with old activations
            raise RuntimeError("vcvits_b200.Generator: the decoder weights were re-folded between this forward and its "
                               "backward (parameter update, set_mode() or another training-mode forward in between)")
        y, gf = ctx.saved_tensors
        if not ctx.has_g:
            gf = None
        B, T = ctx.shape
        dev = y.device
        plan = module._plan_for(dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        dy = dy.contiguous().float()
        need_dx, need_dg = ctx.needs_input_grad[2], ctx.has_g and ctx.needs_input_grad[3]
        starts, T_full = ctx.starts, ctx.T_full
        dx = torch.empty((B, module.initial_channel, T_full), dtype=torch.float32, device=dev) if need_dx else None
        dg = torch.empty((B, module.gin_channels), dtype=torch.float32, device=dev) if need_dg else None
        params = ctx.params
        flat = torch.empty(module._flat_numel, dtype=torch.float32, device=dev)
        views = module._grad_views(flat)
        base = flat.data_ptr()
        ptrs = (C.c_void_p * len(views))(*[base + 4 * off if p is not None else None
                                           for off, p in zip(module._flat_offsets, params)])
        def run(mask):
            if starts is None:
                _lib.check(lib.vcd_backward(plan, module._mode, dy.data_ptr(), y.data_ptr(),
                                            gf.data_ptr() if gf is not None else None,
                                            dx.data_ptr() if dx is not None else None,
                                            dg.data_ptr() if dg is not None else None,
                                            ptrs, ctx.ws.data_ptr(), ctx.ws_bytes, B, T, mask, stream), "vcd_backward")
            else:
                _lib.check(lib.vcd_backward_sliced(plan, module._mode, dy.data_ptr(), y.data_ptr(),
                                                   gf.data_ptr() if gf is not None else None,
                                                   dx.data_ptr() if dx is not None else None, T_full, starts.data_ptr(),
                                                   dg.data_ptr() if dg is not None else None,
                                                   ptrs, ctx.ws.data_ptr(), ctx.ws_bytes, B, T, mask, stream),
                           "vcd_backward_sliced")

        world = dist.get_world_size(module._grad_sync_group) if module._grad_sync_group is not None else 1
        # 1/world is applied where the gradients are written (weight-norm backward), so the all-reduce is a plain SUM
        _lib.check(lib.vcd_set_gradient_scale(plan, 1.0 / world), "vcd_set_gradient_scale")
        if module._grad_sync_group is None or world == 1:
            run(0xFFFFFFFF)
        else:
            # ONE library call; the gradients of a segment are final (library event) while later segments still run: its
            # all-reduce is issued behind that event on a side stream and overlaps the rest of backward (replaces the
            # DDP reducer implied by train.py:99-100)
            reducer = SegmentReducer(flat, module._segment_ranges, module._grad_sync_group, prescaled=True)
            run(0xFFFFFFFF)
            if lib.vcd_segment_events_valid(plan):
                comm = module._comm_stream(dev)
                for seg in range(module._num_segments):
                    _lib.check(lib.vcd_stream_wait_segment(plan, seg, comm.cuda_stream), "vcd_stream_wait_segment")
                    with torch.cuda.stream(comm):
                        reducer.segment_done(seg)
            else:   # profiler / serial mode: everything is already enqueued in order on this stream
                for seg in range(module._num_segments):
                    reducer.segment_done(seg)
            reducer.finish()
        module._give_workspace(ctx.ws)
        ctx.ws = None
        gx = dx.to(ctx.x_dtype) if dx is not None else None
        gg = dg.reshape(ctx.g_shape).to(ctx.g_dtype) if dg is not None else None
        if ctx.fused:
            _accumulate_param_grads(params, views)
            return (None, None, gx, gg, None, None, None)
        grads = [v if (p is not None and p.requires_grad) else None for v, p in zip(views, params)]
        return (None, None, gx, gg, None, None, None, *grads)


def _accumulate_param_grads(params: Sequence[Optional[torch.Tensor]], views: Sequence[torch.Tensor]) -> None:
    """What 233 AccumulateGrad nodes would do: ``p.grad = g`` when it is None (the gradient view is handed over, as
    autograd does with a gradient it may steal), ``p.grad += g`` otherwise (one fused foreach add)."""
    have, add = [], []
    for p, v in zip(params, views):
        if p is None or not p.requires_grad:
            continue
        if p.grad is None:
            p.grad = v
        else:
            have.append(p.grad)
            add.append(v)
    if have:
        with torch.no_grad():
            torch._foreach_add_(have, add)


class Generator(nn.Module):
    """B200-native HiFi-GAN decoder with the reference ``Generator`` constructor (synthesizer_tts.py:71-78).

    Extra keyword ``mode``: ``"bf16"`` (default; tcgen05 tensor-core kernels, fp32 accumulation) or ``"fp32"``
    (FFMA parity mode)."""

    def __init__(self, initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes,
                 upsample_rates, upsample_initial_channel, upsample_kernel_sizes, gin_channels=0, mode="bf16"):
        super().__init__()
        self.initial_channel = int(initial_channel)
        self.resblock = "1" if str(resblock) == "1" else "2"
        self.resblock_kernel_sizes = [int(k) for k in resblock_kernel_sizes]
        self.resblock_dilation_sizes = [[int(d) for d in ds] for ds in resblock_dilation_sizes]
        self.upsample_rates = [int(u) for u in upsample_rates]
        self.upsample_initial_channel = int(upsample_initial_channel)
        self.upsample_kernel_sizes = [int(k) for k in upsample_kernel_sizes]
        self.gin_channels = int(gin_channels)
        self.num_kernels = len(self.resblock_kernel_sizes)
        self.num_upsamples = len(self.upsample_rates)
        self.hop = 1
        for u in self.upsample_rates:
            self.hop *= u
        self.set_mode(mode)
        npairs = 3 if self.resblock == "1" else 2
        for ds in self.resblock_dilation_sizes:
            if len(ds) < npairs:
                raise ValueError(f"ResBlock{self.resblock} needs {npairs} dilations per kernel size "
                                 f"(vits/model/modules.py:190-199,229-230), got {ds}")

        c0 = self.upsample_initial_channel
        # construction order mirrors upstream (conv_pre, ups, resblocks, conv_post, init_weights draws, cond)
        self.conv_pre = _ConvParams(*_default_conv_init(c0, self.initial_channel, 7))
        ups = []
        ch = c0
        for u, k in zip(self.upsample_rates, self.upsample_kernel_sizes):
            ups.append(_WeightNormParams(*_default_conv_init(ch // 2, ch, k, transposed=True)))
            ch //= 2
        self.ups = nn.ModuleList(ups)
        blocks = []
        ch = c0
        for _ in self.upsample_rates:
            ch //= 2
            for k in self.resblock_kernel_sizes:
                if self.resblock == "1":
                    c1 = [_WeightNormParams(*_default_conv_init(ch, ch, k)) for _ in range(3)]
                    _consume_init_weights_rng(c1)
                    c2 = [_WeightNormParams(*_default_conv_init(ch, ch, k)) for _ in range(3)]
                    _consume_init_weights_rng(c2)
                    blocks.append(_ResBlock1Params(c1, c2))
                else:
                    cs = [_WeightNormParams(*_default_conv_init(ch, ch, k)) for _ in range(2)]
                    _consume_init_weights_rng(cs)
                    blocks.append(_ResBlock2Params(cs))
        self.resblocks = nn.ModuleList(blocks)
        w_post = torch.empty(1, ch, 7)
        nn.init.kaiming_uniform_(w_post, a=math.sqrt(5))
        self.conv_post = _ConvParams(w_post, None)
        _consume_init_weights_rng(self.ups)
        if self.gin_channels != 0:
            self.cond = _ConvParams(*_default_conv_init(c0, self.gin_channels, 1))

        self._plans = {}
        self._param_cache = None
        self._slots = None
        self._fold_key = None
        self._fold_serial = 0
        self._ws_cache = {}
        self._ws_pool = {}
        self._grad_sync_group = None
        # True (default): backward assigns / accumulates every parameter's ``.grad`` itself (views of one flat buffer)
        # instead of routing 233 tensors through autograd; set False when something must observe the parameters as
        # autograd leaves of this call (``torch.autograd.grad`` w.r.t. parameters, DistributedDataParallel hooks).
        self.fused_param_grads = True
        self._deterministic = False
        self._anchors = {}
        self._comm_streams = {}
        self._grad_templates = None
        self._call_params = None
        self._names: Optional[List[str]] = None

    # ------------------------------------------------------------------ configuration helpers
    def set_mode(self, mode) -> "Generator":
        if mode in ("bf16", torch.bfloat16, _lib.MODE_BF16):
            self._mode = _lib.MODE_BF16
        elif mode in ("fp32", torch.float32, _lib.MODE_FP32):
            self._mode = _lib.MODE_FP32
        else:
            raise ValueError(f"mode must be 'bf16' or 'fp32', got {mode!r}")
        self._fold_key = None
        return self

    @property
    def mode(self) -> str:
        return "bf16" if self._mode == _lib.MODE_BF16 else "fp32"

    def set_gradient_sync(self, group) -> "Generator":
        """Average parameter gradients over ``group`` (a torch.distributed process group, or
        ``torch.distributed.group.WORLD``) inside backward, segment by segment, overlapped with the remaining
        backward kernels.  ``None`` disables it (e.g. when wrapped in DistributedDataParallel instead)."""
        self._grad_sync_group = group
        return self

    def remove_weight_norm(self) -> "Generator":
        """Reference convention (modules.py:218-222,245-247 and upstream ``Generator.remove_weight_norm``): every
        weight-normed convolution (``ups.*``, ``resblocks.*``) is baked, ``weight = g * v / ||v||`` -- afterwards the
        module exposes ``weight``/``bias`` and no ``weight_g``/``weight_v``, exactly like
        ``torch.nn.utils.remove_weight_norm``.  The library is told so by a NULL ``weight_g`` pointer.  As in the
        reference, a checkpoint saved before the call no longer loads afterwards (and vice versa)."""
        for holder in self.modules():
            if isinstance(holder, _WeightNormParams):
                holder.bake()
        self.invalidate_weights()
        self._param_cache = None
        return self

    @property
    def deterministic(self) -> bool:
        """True: parameter gradients are bit-identical from run to run (``vcd_set_deterministic``: at most two atomic
        contributions per gradient element; slower weight-gradient launches).  Default False."""
        return self._deterministic

    @deterministic.setter
    def deterministic(self, on: bool) -> None:
        self._deterministic = bool(on)
        for plan in self._plans.values():
            _lib.check(_lib.load().vcd_set_deterministic(plan, 1 if self._deterministic else 0), "vcd_set_deterministic")

    def invalidate_weights(self) -> "Generator":
        """Forget the cached weight fold.  Training-mode forwards always re-fold; inference (``torch.no_grad``)
        caches the fold keyed on the parameters' (address, version) -- an in-place edit through ``p.data`` does not
        bump the version, so call this after such an edit (EMA copy, manual clipping, ...)."""
        self._fold_key = None
        return self

    def config_struct(self) -> _lib.VcdConfig:
        cfg = _lib.VcdConfig()
        cfg.initial_channel = self.initial_channel
        cfg.resblock = 1 if self.resblock == "1" else 2
        cfg.num_kernels = self.num_kernels
        if self.num_kernels > _lib.MAX_KERNELS or self.num_upsamples > _lib.MAX_UPSAMPLES:
            raise ValueError("too many resblock kernels / upsample stages for the C ABI")
        for i, k in enumerate(self.resblock_kernel_sizes):
            cfg.resblock_kernel_sizes[i] = k
            for j, d in enumerate(self.resblock_dilation_sizes[i][:_lib.MAX_DILATIONS]):
                cfg.resblock_dilation_sizes[i][j] = d
        cfg.num_upsamples = self.num_upsamples
        for i, (u, k) in enumerate(zip(self.upsample_rates, self.upsample_kernel_sizes)):
            cfg.upsample_rates[i] = u
            cfg.upsample_kernel_sizes[i] = k
        cfg.upsample_initial_channel = self.upsample_initial_channel
        cfg.gin_channels = self.gin_channels
        return cfg

    # ------------------------------------------------------------------ plan / parameters
    def _plan_for(self, device: torch.device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        plan = self._plans.get(key)
        if plan is None:
            lib = _lib.load()
            handle = C.c_void_p()
            cfg = self.config_struct()
            with torch.cuda.device(device):
                _lib.check(lib.vcd_plan_create(C.byref(cfg), C.byref(handle)), "vcd_plan_create")
            plan = handle
            self._plans[key] = plan
            if self._deterministic:
                _lib.check(lib.vcd_set_deterministic(plan, 1), "vcd_set_deterministic")
            n = lib.vcd_num_params(plan)
            names, numels = [], []
            for i in range(n):
                name = C.c_char_p()
                shape = (C.c_int64 * 3)()
                ndim = C.c_int()
                _lib.check(lib.vcd_param_info(plan, i, C.byref(name), shape, C.byref(ndim)), "vcd_param_info")
                names.append(name.value.decode())
                numels.append(tuple(shape[: ndim.value]))
            self._names = names
            self._slots = None
            own = {}
            for nme, shp, prm in zip(names, numels, self._resolve_params(names)):
                if prm is not None and tuple(prm.shape) != shp:
                    raise RuntimeError(f"shape mismatch for {nme}: module {tuple(prm.shape)} vs library {shp}")
                own[nme] = shp
            numel_of = {nme: int(math.prod(shp)) for nme, shp in own.items()}
            # flat gradient buffer layout: parameters grouped by backward segment, in completion order
            nseg = lib.vcd_num_backward_segments(plan)
            self._num_segments = nseg
            order, ranges, off = {}, [], 0
            for seg in range(nseg):
                buf = (C.c_int * n)()
                cnt = lib.vcd_segment_params(plan, seg, buf, n)
                lo = off
                for j in range(cnt):
                    idx = buf[j]
                    order[idx] = off
                    off += numel_of[names[idx]]
                ranges.append((lo, off))
            if len(order) != n:
                raise RuntimeError("backward segments do not cover every parameter")
            self._flat_offsets = [order[i] for i in range(n)]
            self._flat_numel = off
            by_off = sorted(range(n), key=lambda i: order[i])
            self._flat_sizes = [numel_of[names[i]] for i in by_off]
            pos = {i: k for k, i in enumerate(by_off)}
            self._flat_index = [(pos[i], own[names[i]]) for i in range(n)]
            self._grad_templates = None
            self._segment_ranges = ranges
        return plan

    def _resolve_params(self, names: Sequence[str]) -> List[Optional[torch.Tensor]]:
        """Library parameter table (reference checkpoint keys) -> this module's Parameters.  After
        ``remove_weight_norm`` a layer's ``weight_v`` slot is its baked ``weight`` and its ``weight_g`` slot is None."""
        slots, out = [], []
        for name in names:
            prefix, leaf = name.rsplit(".", 1)
            holder = self.get_submodule(prefix)
            if leaf in ("weight_g", "weight_v") and leaf not in holder._parameters:
                if "weight" not in holder._parameters:
                    raise RuntimeError(f"parameter table mismatch between module and library: {name}")
                leaf = "weight" if leaf == "weight_v" else None
            elif leaf not in holder._parameters:
                raise RuntimeError(f"parameter table mismatch between module and library: {name}")
            slots.append((holder, leaf))
            out.append(holder._parameters[leaf] if leaf is not None else None)
        n_own = sum(1 for _ in self.parameters())
        if n_own != sum(1 for t in out if t is not None):
            raise RuntimeError("parameter table mismatch between module and library: the module holds parameters the "
                               "library does not know")
        self._slots = slots
        return out

    def _ordered_params(self) -> List[Optional[torch.Tensor]]:
        """Parameters in library order.  The cached list is validated by identity on every call, so a Parameter that was
        replaced without ``_apply`` (``load_state_dict(assign=True)``, ``m.conv_pre.weight = nn.Parameter(..)``,
        pruning / parametrize utilities) is picked up instead of silently using the orphaned tensor."""
        cached, slots = self._param_cache, self._slots
        if cached is not None and slots is not None:
            for (holder, leaf), t in zip(slots, cached):
                if (holder._parameters.get(leaf) if leaf is not None else None) is not t:
                    cached = None
                    break
        if cached is None:
            cached = self._param_cache = self._resolve_params(self._names)
            self._fold_key = None
        return cached

    def _grad_views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        # ONE native call makes every shaped view (flat-buffer order); then parameter-table order
        if self._grad_templates is None:
            by_pos = sorted(self._flat_index, key=lambda ks: ks[0])
            self._grad_templates = [torch.empty(shape, dtype=torch.float32, device="meta") for _, shape in by_pos]
        pieces = torch._C._nn.unflatten_dense_tensors(flat, self._grad_templates)
        return [pieces[k] for k, _ in self._flat_index]

    def _comm_stream(self, device: torch.device) -> torch.cuda.Stream:
        st = self._comm_streams.get(device)
        if st is None:
            st = self._comm_streams[device] = torch.cuda.Stream(device=device)
        return st

    def _anchor_for(self, device: torch.device) -> torch.Tensor:
        a = self._anchors.get(device)
        if a is None:
            a = self._anchors[device] = torch.zeros(1, device=device, requires_grad=True)
        return a

    def _fold_if_needed(self, params: Sequence[Optional[torch.Tensor]], force: bool = False) -> None:
        """Fold the parameters into the packed operand layouts.  Training-mode forwards (``force``) always fold: the
        parameters change every optimizer step and an in-place edit through ``p.data`` is invisible to any version key
        (two launches, part of the timed step in bench.py).  Inference caches on (address, version) per parameter."""
        key = None
        if not force:
            key = (self._mode, tuple((p.data_ptr(), p._version) if p is not None else None for p in params))
            if key == self._fold_key:
                return
        lib = _lib.load()
        f32 = torch.float32
        raw = []
        for p in params:
            if p is None:
                raw.append(None)
                continue
            if p.dtype != f32 or not p.is_contiguous():
                raise RuntimeError("decoder parameters must be contiguous fp32 tensors")
            raw.append(p.data_ptr())
        dev = next(p for p in params if p is not None).device
        ptrs = (C.c_void_p * len(params))(*raw)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.vcd_fold_weights(self._plan_for(dev), self._mode, ptrs, stream), "vcd_fold_weights")
        self._fold_key = key
        self._fold_serial += 1

    def _workspace(self, nbytes: int, device: torch.device, cache: bool) -> torch.Tensor:
        if not cache:
            return torch.empty(nbytes, dtype=torch.uint8, device=device)
        ws = self._ws_cache.get(device)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws_cache = {device: ws}
        return ws

    def _take_workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        """Training workspaces are pooled so that steady-state steps see the SAME device address (the library
        replays CUDA graphs keyed on it).  A workspace is out of the pool between forward and backward."""
        pool = self._ws_pool.setdefault((device, nbytes), [])
        if pool:
            return pool.pop()
        return torch.empty(nbytes, dtype=torch.uint8, device=device)

    def _give_workspace(self, ws: torch.Tensor) -> None:
        pool = self._ws_pool.setdefault((ws.device, ws.numel()), [])
        if len(pool) < 4:
            pool.append(ws)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._fold_key = None
        self._param_cache = None
        self._ws_cache = {}
        self._ws_pool = {}
        self._anchors = {}
        return out

    # ------------------------------------------------------------------ forward
    def forward_sliced(self, z: torch.Tensor, ids_str: torch.Tensor, segment_size: int,
                       g: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``self(commons.slice_segments(z, ids_str, segment_size), g)`` without materialising the slice: the per-item
        segment gather of ``rand_slice_segments`` (vits/commons.py:48-64; synthesizer_tts.py:138-140,
        synthesizer_svc.py:86-87) happens inside the decoder's input load, and backward scatters the latent gradient
        into a zero-filled full-length tensor.  ``z``: [B, C, T_full]; ``ids_str``: [B] integer start frames."""
        if ids_str.dim() != 1 or ids_str.shape[0] != z.shape[0]:
            raise ValueError(f"expected ids_str of shape [{z.shape[0]}], got {tuple(ids_str.shape)}")
        segment_size = int(segment_size)
        if segment_size < 1 or segment_size > z.shape[2]:
            raise ValueError(f"segment_size {segment_size} does not fit a latent of {z.shape[2]} frames")
        starts = ids_str.to(device=z.device, dtype=torch.int64).contiguous()
        return self.forward(z, g, _starts=starts, _segment=segment_size)

    def forward(self, x: torch.Tensor, g: Optional[torch.Tensor] = None, *, _starts: Optional[torch.Tensor] = None,
                _segment: int = 0) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("vcvits_b200.Generator runs only on CUDA (sm_100a); there is no CPU fallback")
        if x.dim() != 3 or x.shape[1] != self.initial_channel:
            raise ValueError(f"expected x of shape [B, {self.initial_channel}, T], got {tuple(x.shape)}")
        if g is not None:
            if self.gin_channels == 0:
                raise ValueError("g given but the decoder was built with gin_channels=0")
            if g.shape[0] != x.shape[0] or g.numel() != x.shape[0] * self.gin_channels:
                raise ValueError(f"expected g of shape [B, {self.gin_channels}, 1], got {tuple(g.shape)}")
        # the library launches on this device's plan streams: make it current (single-process multi-GPU callers)
        with torch.cuda.device(x.device):
            self._plan_for(x.device)
            params = self._ordered_params()
            grad_on = torch.is_grad_enabled()
            params_need = grad_on and any(p is not None and p.requires_grad for p in params)
            need_grad = grad_on and (params_need or x.requires_grad or (g is not None and g.requires_grad))
            # Lightning AMP (train.py:104-106) calls this inside autocast: the decoder computes in its own mode
            with torch.autocast(device_type="cuda", enabled=False):
                if self.fused_param_grads and (params_need or not need_grad):
                    self._call_params = params
                    anchor = self._anchor_for(x.device) if params_need else None
                    return _DecoderFunction.apply(self, need_grad, x, g, _starts, _segment, anchor)
                return _DecoderFunction.apply(self, need_grad, x, g, _starts, _segment, None, *params)

    def synthesize_host(self, x_host: torch.Tensor, g_host: Optional[torch.Tensor] = None,
                        device: Optional[torch.device] = None) -> torch.Tensor:
        """infer.py-style call on HOST tensors through ``vcd_synthesize_host``: H2D copy, decode, D2H copy."""
        lib = _lib.load()
        device = device or next(self.parameters()).device
        plan = self._plan_for(device)
        with torch.cuda.device(device):
            self._fold_if_needed(self._ordered_params())
        B, _, T = x_host.shape
        xh = x_host.contiguous().float()
        gh = g_host.reshape(B, -1).contiguous().float() if g_host is not None else None
        y = torch.empty((B, 1, T * self.hop), dtype=torch.float32, pin_memory=True)
        nbytes = lib.vcd_workspace_bytes(plan, self._mode, B, T, 0)
        nbytes = (nbytes + 255) // 256 * 256 + lib.vcd_host_call_extra_bytes(plan, B, T)
        ws = self._workspace(nbytes, device, cache=True)
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            _lib.check(lib.vcd_synthesize_host(plan, self._mode, xh.data_ptr(),
                                               gh.data_ptr() if gh is not None else None, y.data_ptr(),
                                               ws.data_ptr(), ws.numel(), B, T, stream), "vcd_synthesize_host")
        return y

    def layer_paths(self) -> List[str]:
        lib = _lib.load()
        plan = self._plan_for(next(self.parameters()).device)
        out, i = [], 0
        while True:
            s = lib.vcd_layer_path(plan, self._mode, i)
            if s is None:
                return out
            out.append(s.decode())
            i += 1

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_plans"] = {}        # library handles are per process / per device; rebuilt lazily
        state["_param_cache"] = None
        state["_slots"] = None
        state["_ws_cache"] = {}
        state["_ws_pool"] = {}
        state["_anchors"] = {}
        state["_comm_streams"] = {}
        state["_call_params"] = None
        state["_grad_templates"] = None
        state["_fold_key"] = None
        state["_grad_sync_group"] = None
        return state

    def __del__(self):
        try:
            lib = _lib.load()
            for plan in self._plans.values():
                lib.vcd_plan_destroy(plan)
        except Exception:
            pass
