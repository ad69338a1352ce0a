"""CPU oracle for the VCVITS HiFi-GAN waveform decoder (`dec`).

TEST INFRASTRUCTURE ONLY.  Nothing under ``vcvits_b200/`` may import this file; it is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the checker and the timed CPU baseline, never as the product path.

What it restates (all citations are into the reference tree, /root/reference):

* ``OracleResBlock1``  follows ``vits/model/modules.py:186-222`` (three pairs of
  [lrelu(0.1) -> dilated weight-normed Conv1d -> lrelu(0.1) -> Conv1d(dil=1) -> +x]).
* ``OracleResBlock2``  follows ``vits/model/modules.py:225-247`` (two [lrelu -> dilated conv -> +x]).
* ``same_padding``     follows ``vits/commons.py:14-15`` (``get_padding``).
* ``LRELU_SLOPE``      is ``vits/model/modules.py:16``.
* ``OracleGenerator``  is the decoder behind ``SynthesizerTTS.dec`` / ``SynthesizerSVC.dec``.  The class
  itself is ABSENT from the reference tree (``vits/model/synthesizers/synthesizer_tts.py:22`` imports a
  missing ``vits/model/vocoder``; ``synthesizer_svc.py:59`` pulls the un-vendored torch.hub dependency
  ``vtuber-plan/hifi-gan`` tag ``v0.3.1``).  Its constructor contract is ``synthesizer_tts.py:71-78`` and
  its forward contract ``synthesizer_tts.py:140`` / ``synthesizer_svc.py:87,108``.  The glue
  (conv_pre k=7 -> +cond(g) -> [lrelu(0.1) -> weight-normed ConvTranspose1d -> mean of ResBlocks] x N ->
  lrelu(0.01 default slope) -> conv_post(k=7, no bias) -> tanh) is the published upstream VITS
  ``models.py::Generator`` algorithm that this repo declares itself a derivative of (``README.md:2``);
  SURVEY.md Appendix A records that decision.

Pinning status: the two ResBlock restatements are pinned bit-for-bit against the reference's own classes
(``oracle/make_golden.py`` imports ``vits.model.modules`` from /root/reference and writes
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them).  The Generator glue has no reference
code, test or golden vector to pin against: for that part **parity is unpinned** and anchored on the
reference's call sites and configs only.
"""
from __future__ import annotations

import warnings
from typing import Optional, Sequence

import torch
from torch import nn
from torch.nn import functional as F

LRELU_SLOPE = 0.1  # vits/model/modules.py:16


def _weight_norm(m: nn.Module) -> nn.Module:
    # old-style weight_norm (weight_g / weight_v parameters), as imported at vits/model/modules.py:10
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.nn.utils.weight_norm(m)


def same_padding(kernel_size: int, dilation: int = 1) -> int:
    """vits/commons.py:14-15."""
    return int((kernel_size * dilation - dilation) / 2)


def _mirror_init_weights_rng(convs) -> None:
    """vits/commons.py:8-11 ``init_weights`` draws N(0, 0.01) into the derived ``.weight`` of each conv
    (modules.py:195,201,232).  Under old-style weight_norm that tensor is recomputed from weight_g/weight_v
    by the forward pre-hook, so the draw never reaches the effective weights; it does advance
    the RNG, which we mirror so seeded default inits agree with the reference classes."""
    for conv in convs:
        torch.empty_like(conv.weight_v).normal_(0.0, 0.01)


class OracleResBlock1(nn.Module):
    """vits/model/modules.py:186-222 (x_mask branches are dead on the decoder path and omitted)."""

    def __init__(self, channels: int, kernel_size: int = 3, dilation: Sequence[int] = (1, 3, 5)):
        super().__init__()
        self.convs1 = nn.ModuleList(
            _weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                   padding=same_padding(kernel_size, d)))
            for d in dilation[:3])  # the reference indexes dilation[0..2] only
        _mirror_init_weights_rng(self.convs1)
        self.convs2 = nn.ModuleList(
            _weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=1,
                                   padding=same_padding(kernel_size, 1)))
            for _ in dilation[:3])
        _mirror_init_weights_rng(self.convs2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for first, second in zip(self.convs1, self.convs2):
            h = first(F.leaky_relu(x, LRELU_SLOPE))
            h = second(F.leaky_relu(h, LRELU_SLOPE))
            x = h + x
        return x


class OracleResBlock2(nn.Module):
    """vits/model/modules.py:225-247."""

    def __init__(self, channels: int, kernel_size: int = 3, dilation: Sequence[int] = (1, 3)):
        super().__init__()
        self.convs = nn.ModuleList(
            _weight_norm(nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                   padding=same_padding(kernel_size, d)))
            for d in dilation[:2])  # the reference indexes dilation[0..1] only
        _mirror_init_weights_rng(self.convs)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for conv in self.convs:
            x = conv(F.leaky_relu(x, LRELU_SLOPE)) + x
        return x


class OracleGenerator(nn.Module):
    """Decoder with the constructor of synthesizer_tts.py:71-78 and forward of synthesizer_tts.py:140."""

    def __init__(self, initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes,
                 upsample_rates, upsample_initial_channel, upsample_kernel_sizes, gin_channels=0):
        super().__init__()
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        self.conv_pre = nn.Conv1d(initial_channel, upsample_initial_channel, 7, 1, padding=3)
        block = OracleResBlock1 if str(resblock) == "1" else OracleResBlock2
        # Construction order (ups, resblocks, conv_post, init_weights draws, cond) matches upstream so that a
        # seeded default init reproduces the same weights as the reference-class composition in
        # oracle/make_golden.py.
        self.ups = nn.ModuleList()
        ch = upsample_initial_channel
        for u, k in zip(upsample_rates, upsample_kernel_sizes):
            self.ups.append(_weight_norm(nn.ConvTranspose1d(ch, ch // 2, k, u, padding=(k - u) // 2)))
            ch //= 2
        self.resblocks = nn.ModuleList()
        ch = upsample_initial_channel
        for _ in upsample_rates:
            ch //= 2
            for rk, rd in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(block(ch, rk, rd))
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3, bias=False)
        _mirror_init_weights_rng(self.ups)  # upstream: self.ups.apply(init_weights)
        if gin_channels != 0:
            self.cond = nn.Conv1d(gin_channels, upsample_initial_channel, 1)

    def forward(self, x: torch.Tensor, g: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self.conv_pre(x)
        if g is not None:
            x = x + self.cond(g)
        for i, up in enumerate(self.ups):
            x = up(F.leaky_relu(x, LRELU_SLOPE))
            branches = self.resblocks[i * self.num_kernels:(i + 1) * self.num_kernels]
            acc = None
            for rb in branches:
                acc = rb(x) if acc is None else acc + rb(x)
            x = acc / self.num_kernels
        x = F.leaky_relu(x)  # default slope 0.01 (upstream), NOT LRELU_SLOPE
        return torch.tanh(self.conv_post(x))


# ---------------------------------------------------------------------------------------------
# Shared helpers for tests / bench (configs follow configs/base.json:55-67, configs/48k_base.json)
# ---------------------------------------------------------------------------------------------
BASE_CFG = dict(initial_channel=256, resblock="1", resblock_kernel_sizes=[3, 7, 11],
                resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], upsample_rates=[8, 8, 4, 2],
                upsample_initial_channel=512, upsample_kernel_sizes=[16, 16, 4, 4], gin_channels=256)
BASE48K_CFG = dict(BASE_CFG, initial_channel=128)
TINY_CFG = dict(initial_channel=16, resblock="1", resblock_kernel_sizes=[3, 5],
                resblock_dilation_sizes=[[1, 3, 5], [1, 2, 3]], upsample_rates=[4, 2],
                upsample_initial_channel=32, upsample_kernel_sizes=[8, 4], gin_channels=8)
TINY2_CFG = dict(TINY_CFG, resblock="2")
# smallest config whose every stage is wide enough for the tcgen05 tiles (64- and 32-channel stages)
SMALL_CFG = dict(initial_channel=64, resblock="1", resblock_kernel_sizes=[3, 7, 11],
                 resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], upsample_rates=[4, 4],
                 upsample_initial_channel=128, upsample_kernel_sizes=[8, 8], gin_channels=16)
SMALL2_CFG = dict(SMALL_CFG, resblock="2")


def seeded_state_dict(cfg: dict, seed: int = 1234, gain: float = 1.0) -> "dict[str, torch.Tensor]":
    """Default-init weights under a fixed seed (configs/base.json:12 uses 1234).  ``gain`` scales every
    ``weight_g`` so the output reaches tanh's non-linear range."""
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        model = OracleGenerator(**cfg)
    finally:
        torch.random.set_rng_state(gen_state)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if gain != 1.0:
        for k in sd:
            if k.endswith("weight_g"):
                sd[k] = sd[k] * gain
    return sd


def build(cfg: dict, state_dict=None, dtype=torch.float32) -> OracleGenerator:
    model = OracleGenerator(**cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model.to(dtype)


def audio_seconds(batch: int, frames: int, hop: int = 512, sr: int = 48000) -> float:
    """configs/base.json:31,33 (target_sampling_rate, hop_length)."""
    return batch * frames * hop / sr


def forward_flops(cfg: dict, batch: int, frames: int) -> int:
    """Algorithmic forward FLOPs, 2*Cin*Cout*k*L per conv."""
    c0 = cfg["upsample_initial_channel"]
    total = 2 * cfg["initial_channel"] * c0 * 7 * frames
    if cfg.get("gin_channels", 0):
        total += 2 * cfg["gin_channels"] * c0
    ch, length = c0, frames
    for u, k in zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"]):
        total += 2 * ch * (ch // 2) * k * length
        ch //= 2
        length *= u
        per_block = 2 if str(cfg["resblock"]) == "1" else 1
        for rk, rd in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]):
            total += len(rd) * per_block * 2 * ch * ch * rk * length
    total += 2 * ch * 1 * 7 * length
    return total * batch
