"""One (or a few) training steps of a bench workload; used under ncu."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="BASE_CFG"); ap.add_argument("--B", type=int, default=16)
ap.add_argument("--T", type=int, default=32); ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--mode", default="bf16"); ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--range-last", action="store_true", help="cudaProfilerStart/Stop around the last step (ncu --replay-mode range)")
a = ap.parse_args()
cfg = getattr(O, a.cfg)
torch.manual_seed(1234)
m = Generator(**cfg, mode=a.mode).cuda()
x = torch.randn(a.B, cfg["initial_channel"], a.T, device="cuda")
g = torch.randn(a.B, cfg["gin_channels"], 1, device="cuda")
dy = torch.randn(a.B, 1, a.T * m.hop, device="cuda")
for it in range(a.steps):
    if a.range_last and it == a.steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    if a.fwd_only:
        with torch.no_grad():
            m(x, g)
    else:
        m.zero_grad(set_to_none=True)
        m._fold_key = None
        m(x.requires_grad_(True), g.requires_grad_(True)).backward(dy)
torch.cuda.synchronize()
if a.range_last:
    torch.cuda.profiler.stop()
from vcvits_b200 import _lib
import os
if os.environ.get("VCD_PHASES"):
    _lib.load().vcd_phase_dump(1)
print("done")
