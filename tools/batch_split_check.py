"""Debug: gradients of one B=8 step vs the sum of two B=4 steps (same samples).  The decoder has no cross-sample
operation, so the two must agree to fp32 summation order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator

cfg = getattr(O, sys.argv[1] if len(sys.argv) > 1 else "BASE_CFG")
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
torch.manual_seed(1234)
m = Generator(**cfg, mode=mode).cuda()
g = torch.Generator().manual_seed(7)
xs = torch.randn(8, cfg["initial_channel"], T, generator=g).cuda()
gs = torch.randn(8, cfg["gin_channels"], 1, generator=g).cuda()
dys = torch.randn(8, 1, T * m.hop, generator=g).cuda()

def grads(sl):
    m.zero_grad(set_to_none=True)
    x = xs[sl].clone().requires_grad_(True)
    y = m(x, gs[sl])
    y.backward(dys[sl])
    torch.cuda.synchronize()
    return {n: p.grad.double().clone() for n, p in m.named_parameters()}, y.detach().clone(), x.grad.clone()

full, yf, dxf = grads(slice(0, 8))
a, ya, dxa = grads(slice(0, 4))
b, yb, dxb = grads(slice(4, 8))
print("forward max-abs diff:", float((yf - torch.cat([ya, yb])).abs().max()), " dx rel:", float((dxf - torch.cat([dxa, dxb])).norm() / dxf.norm()))
full2, _, _ = grads(slice(0, 8))
num = sum(float((full[n] - full2[n]).pow(2).sum()) for n in full); den = sum(float(full[n].pow(2).sum()) for n in full)
print("run-to-run (same batch) rel-l2:", (num / den) ** 0.5)
num = sum(float((a[n] + b[n] - full[n]).pow(2).sum()) for n in full)
print("split vs full rel-l2:", (num / den) ** 0.5)
worst = sorted(((float((a[n] + b[n] - full[n]).norm() / (full[n].norm() + 1e-30)), n) for n in full), reverse=True)[:6]
for e, n in worst:
    print(f"  {n:40s} {e:.3e}")
