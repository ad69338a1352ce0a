import os, sys, collections, csv
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator, _lib
lib = _lib.load()
torch.manual_seed(0)
m = Generator(**O.BASE_CFG, mode="bf16").cuda()
x = torch.randn(16, 256, 32, device="cuda"); g = torch.randn(16, 256, 1, device="cuda"); dy = torch.randn(16, 1, 32*512, device="cuda")
def step():
    for p in m.parameters(): p.grad = None
    xx, gg = x.clone().requires_grad_(True), g.clone().requires_grad_(True)
    m(xx, gg).backward(dy)
for det in (False, True):
    m.deterministic = det
    for _ in range(3): step()
    torch.cuda.synchronize()
    lib.vcd_profile_enable(1)
    step(); torch.cuda.synchronize()
    path = f"gpurun_out/det_prof_{int(det)}.csv"
    lib.vcd_profile_dump(path.encode())
    lib.vcd_profile_enable(0)
    rows = [l.rstrip("\n").split(",") for l in open(path)]
    print("det", det, "rows", len(rows), rows[0])
    rows = rows[1:]
    tot = sum(float(r[-2]) for r in rows)
    print(" total serial ms", tot)
    for r in sorted(rows, key=lambda r: -float(r[-2]))[:12]:
        print("  ", r)
