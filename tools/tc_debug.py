"""Debug helper (GPU): compare bf16 mode against fp32 mode and the fp64 oracle, per tensor."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import hifigan_oracle as O
from tests.helpers import b200_run, oracle_run, rel_l2

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="SMALL_CFG")
ap.add_argument("--B", type=int, default=2)
ap.add_argument("--T", type=int, default=40)
ap.add_argument("--gain", type=float, default=1.3)
ap.add_argument("--nograd", action="store_true")
ap.add_argument("--verbose", action="store_true")
a = ap.parse_args()
cfg = getattr(O, a.cfg)
sd = O.seeded_state_dict(cfg, 1234, gain=a.gain)
torch.manual_seed(0)
x = torch.randn(a.B, cfg["initial_channel"], a.T)
g = torch.randn(a.B, cfg["gin_channels"], 1)
hop = 1
for u in cfg["upsample_rates"]:
    hop *= u
dy = None if a.nograd else torch.randn(a.B, 1, a.T * hop)
y_ref, g_ref = oracle_run(cfg, sd, x, g, dy)
print("oracle absmax", float(y_ref.abs().max()))
for mode in ("fp32", "bf16"):
    y, grads, m = b200_run(cfg, sd, x, g, dy, mode=mode)
    torch.cuda.synchronize()
    if a.verbose and mode == "bf16":
        print("\n".join(m.layer_paths()))
    print(f"[{mode}] fwd max-abs {float((y.double() - y_ref).abs().max()):.3e} rel-l2 {rel_l2(y, y_ref):.3e}")
    if grads:
        rows = sorted(((rel_l2(grads[n], g_ref[n]), n) for n in g_ref), reverse=True)
        num = sum(float((grads[n].double() - g_ref[n]).pow(2).sum()) for n in g_ref)
        den = sum(float(g_ref[n].pow(2).sum()) for n in g_ref)
        print(f"[{mode}] grads global rel-l2 {(num / den) ** 0.5:.3e}; worst: " + ", ".join(f"{n}={r:.2e}" for r, n in rows[:6]))
        if a.verbose:
            for r, n in rows:
                print(f"   {n:40s} {r:.3e}")
