"""Host-side cost of one training step (bench.py's step_device): cProfile + coarse timers.  Run on the GPU box."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator, _lib

cfg = O.BASE_CFG
torch.manual_seed(1234)
m = Generator(**cfg, mode="bf16").cuda()
B, T = 16, 32
x = torch.randn(B, cfg["initial_channel"], T, device="cuda")
g = torch.randn(B, cfg["gin_channels"], 1, device="cuda")
dy = torch.randn(B, 1, T * m.hop, device="cuda")
params = list(m.parameters())

def step():
    for p in params:
        p.grad = None
    xx = x.detach().requires_grad_(True)
    gg = g.detach().requires_grad_(True)
    y = m(xx, gg)
    y.backward(dy)

for _ in range(5):
    step()
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/N:.3f} ms/step; with final sync {1e3*(t2-t0)/N:.3f} ms/step")
# forward only / backward only host time
tf = tb = 0.0
for _ in range(N):
    for p in params:
        p.grad = None
    xx = x.detach().requires_grad_(True); gg = g.detach().requires_grad_(True)
    a = time.perf_counter(); y = m(xx, gg); b = time.perf_counter(); y.backward(dy); c = time.perf_counter()
    tf += b - a; tb += c - b
torch.cuda.synchronize()
print(f"host forward {1e3*tf/N:.3f} ms, backward {1e3*tb/N:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
print(s.getvalue()[:5000])
