"""A few calls of the mel loss tail at the configs[1] segment shape; used under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vcvits_b200 import mel as V
kw = dict(n_fft=2048, num_mels=256, sampling_rate=48000, hop_size=512, win_size=2048, fmin=0.0, fmax=None)
g = torch.Generator().manual_seed(1)
y = (0.3 * torch.randn(16, 1, 16384, generator=g)).clamp(-1, 1).cuda()
tgt = V.mel_spectrogram_torch((0.3 * torch.randn(16, 16384, generator=g)).clamp(-1, 1).cuda(), **kw)
plan = V._plan(device="cuda", **kw)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    plan.loss_and_grad(y, tgt, 45.0)
torch.cuda.synchronize()
print("done")
