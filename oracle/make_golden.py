"""Generate tests/golden/*.npz from the REFERENCE's own classes (run in the build container only).

TEST INFRASTRUCTURE ONLY.  Imports ``vits.model.modules`` (ResBlock1/ResBlock2, LRELU_SLOPE) and
``vits.commons`` from /root/reference, which exists only in the build container; the produced fixtures are
small and committed so that the GPU box never needs the reference tree.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Fixtures:
  resblock1_k{K}.npz / resblock2_k{K}.npz  reference ResBlock state_dict + input + fp32 output + fp64 output
  generator_tiny{,2}.npz   Appendix-A glue composed from the REFERENCE ResBlock classes (tiny config):
                           state_dict, x, g, y (fp32), y/grad_x/grad_g/param grads in fp64
  generator_base_probe.npz base.json config, weights re-derivable from seed 1234: weight checksum, input,
                           a strided probe of the fp32 output (no 58 MB state_dict in git)
"""
import os
import sys
import types
import warnings

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
warnings.simplefilter("ignore")

from vits.model import modules as ref_modules  # noqa: E402  (the reference's own code)
from vits import commons as ref_commons  # noqa: E402

from oracle import hifigan_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


class RefGlueGenerator(nn.Module):
    """SURVEY.md Appendix A glue around the reference's own ResBlock1/ResBlock2 classes."""

    def __init__(self, initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes,
                 upsample_rates, upsample_initial_channel, upsample_kernel_sizes, gin_channels=0):
        super().__init__()
        self.num_kernels = len(resblock_kernel_sizes)
        self.num_upsamples = len(upsample_rates)
        self.conv_pre = nn.Conv1d(initial_channel, upsample_initial_channel, 7, 1, padding=3)
        block = ref_modules.ResBlock1 if resblock == "1" else ref_modules.ResBlock2
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
            self.ups.append(torch.nn.utils.weight_norm(nn.ConvTranspose1d(
                upsample_initial_channel // (2 ** i), upsample_initial_channel // (2 ** (i + 1)),
                k, u, padding=(k - u) // 2)))
        self.resblocks = nn.ModuleList()
        for i in range(len(self.ups)):
            ch = upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(block(ch, k, d))
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3, bias=False)
        self.ups.apply(ref_commons.init_weights)
        if gin_channels != 0:
            self.cond = nn.Conv1d(gin_channels, upsample_initial_channel, 1)

    def forward(self, x, g=None):
        x = self.conv_pre(x)
        if g is not None:
            x = x + self.cond(g)
        for i in range(self.num_upsamples):
            x = F.leaky_relu(x, ref_modules.LRELU_SLOPE)
            x = self.ups[i](x)
            xs = None
            for j in range(self.num_kernels):
                r = self.resblocks[i * self.num_kernels + j](x)
                xs = r if xs is None else xs + r
            x = xs / self.num_kernels
        x = F.leaky_relu(x)
        x = self.conv_post(x)
        return torch.tanh(x)


def npify(sd):
    return {"sd::" + k: v.detach().cpu().numpy() for k, v in sd.items()}


def golden_resblocks():
    for kind, cls, dil in (("1", ref_modules.ResBlock1, (1, 3, 5)), ("2", ref_modules.ResBlock2, (1, 3))):
        for k in (3, 7, 11):
            torch.manual_seed(100 + k)
            m = cls(16, k, dil)
            # randomise weight_g so g != ||v|| and the reparameterisation is exercised
            with torch.no_grad():
                for n, p in m.named_parameters():
                    if n.endswith("weight_g"):
                        p.mul_(0.5 + torch.rand_like(p))
            x = torch.randn(2, 16, 96)
            y32 = m(x).detach()
            sd = {k_: v.detach().clone() for k_, v in m.state_dict().items()}
            m64 = cls(16, k, dil).double()
            m64.load_state_dict({k_: v.double() for k_, v in sd.items()})
            y64 = m64(x.double()).detach()
            np.savez_compressed(os.path.join(OUT, f"resblock{kind}_k{k}.npz"), x=x.numpy(), y32=y32.numpy(),
                                y64=y64.numpy(), kernel_size=k, dilation=np.array(dil), **npify(sd))


def golden_generator(name, cfg, B, T, gain):
    torch.manual_seed(4321)
    m = RefGlueGenerator(**cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("weight_g"):
                p.mul_(gain * (0.75 + 0.5 * torch.rand_like(p)))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(B, cfg["initial_channel"], T)
    g = torch.randn(B, cfg["gin_channels"], 1)
    y32 = m(x, g).detach()
    m64 = RefGlueGenerator(**cfg).double()
    m64.load_state_dict({k: v.double() for k, v in sd.items()})
    x64 = x.double().requires_grad_(True)
    g64 = g.double().requires_grad_(True)
    y64 = m64(x64, g64)
    dy = torch.randn(y64.shape, dtype=torch.float32)
    y64.backward(dy.double())
    grads = {"grad::" + n: p.grad.numpy() for n, p in m64.named_parameters()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), g=g.numpy(), y32=y32.numpy(),
                        y64=y64.detach().numpy(), dy=dy.numpy(), grad_x=x64.grad.numpy(),
                        grad_g=g64.grad.numpy(), **npify(sd), **grads)
    print(name, "absmax", float(y32.abs().max()))


def golden_base_probe():
    cfg = O.BASE_CFG
    torch.manual_seed(1234)
    m = RefGlueGenerator(**cfg)
    sd = m.state_dict()
    checksum = float(sum(v.double().abs().sum() for v in sd.values()))
    torch.manual_seed(99)
    x = torch.randn(1, 256, 32)
    g = torch.randn(1, 256, 1)
    with torch.no_grad():
        y = m(x, g)
    np.savez_compressed(os.path.join(OUT, "generator_base_probe.npz"), x=x.numpy(), g=g.numpy(),
                        y_probe=y[0, 0, ::16].numpy(), checksum=checksum, torch_version=torch.__version__)
    print("base probe absmax", float(y.abs().max()), "checksum", checksum)


def golden_mel():
    """Reference mel_spectrogram_torch (vits/mel_processing.py:115-142) run with a librosa stub."""
    import torchaudio
    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    filt = types.ModuleType("librosa.filters")
    util.normalize = util.pad_center = util.tiny = None
    def mel(sr, n_fft, n_mels, fmin, fmax):
        return torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax if fmax else sr / 2), n_mels, sr,
                                                     norm="slaney", mel_scale="slaney").T.contiguous().numpy()
    filt.mel = mel
    lib.util, lib.filters = util, filt
    sys.modules["librosa"], sys.modules["librosa.util"], sys.modules["librosa.filters"] = lib, util, filt
    from vits import mel_processing as ref_mel
    torch.manual_seed(5)
    t = torch.arange(16384) / 48000.0
    y = 0.3 * torch.sin(2 * torch.pi * 440 * t) + 0.05 * torch.randn(2, 16384)
    y = y.clamp(-1, 1)
    m = ref_mel.mel_spectrogram_torch(y, 2048, 256, 48000, 512, 2048, 0.0, None)
    np.savez_compressed(os.path.join(OUT, "mel_probe.npz"), y=y.numpy(), logmel=m.numpy())
    print("mel probe", tuple(m.shape))


if __name__ == "__main__":
    golden_mel()
    golden_resblocks()
    golden_generator("generator_tiny", O.TINY_CFG, 2, 12, gain=1.5)
    golden_generator("generator_tiny2", O.TINY2_CFG, 2, 12, gain=1.5)
    golden_base_probe()
    print("wrote", sorted(os.listdir(OUT)))
