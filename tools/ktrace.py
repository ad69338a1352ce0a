"""Debug: in-kernel timeline (CTA 0) of one tcgen05 conv launch.  VCD_KTRACE=<layer>:<fwd|dgrad> python tools/ktrace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VCD_GRAPHS", "0")
import torch
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator, _lib
cfg = O.BASE_CFG
torch.manual_seed(1234)
m = Generator(**cfg, mode="bf16").cuda()
x = torch.randn(16, 256, 32, device="cuda"); g = torch.randn(16, 256, 1, device="cuda"); dy = torch.randn(16, 1, 16384, device="cuda")
for _ in range(3):
    m.zero_grad(set_to_none=True)
    m(x.requires_grad_(True), g.requires_grad_(True)).backward(dy)
torch.cuda.synchronize()
buf = (C.c_uint64 * 64)()
_lib.check(_lib.load().vcd_debug_read_trace(m._plan_for(x.device), buf), "trace")
t0 = buf[0]
names = {0: "entry", 1: "setup done", 2: "first A issued", 3: "W resident landed", 4: "first A landed", 5: "all MMAs issued", 6: "exit",
         8: "mma issued t0", 9: "mma issued t1", 10: "mma issued t2", 11: "mma issued t3", 12: "acc done t0", 13: "acc done t1",
         20: "u0 start", 21: "u0 tmem ld done", 24: "u2 start", 25: "u2 tmem ld done", 28: "u4 start", 29: "u4 tmem ld done",
         14: "acc done t2", 15: "acc done t3", 16: "epi done t0", 17: "epi done t1", 18: "epi done t2", 19: "epi done t3"}
if os.environ.get("VCD_KTRACE", "").endswith(":wgrad"):
    names = {0: "entry", 1: "setup done", 5: "all MMAs issued", 6: "exit", 12: "acc done"}
    names.update({8 + i: f"stage {i} landed" for i in range(4)})
    names.update({21: "accumulators written"})
for k in sorted(names, key=lambda k: buf[k] if buf[k] else 1 << 62):
    if buf[k]:
        print(f"{names[k]:22s} +{(buf[k] - t0) / 1000:8.2f} us")
