"""Debug: repeated forward/backward on one module (exercises CUDA-graph replay) checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hifigan_oracle as O
from tests.helpers import oracle_run, rel_l2
from vcvits_b200 import Generator
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
cfg = O.TINY_CFG if mode == "fp32" else O.SMALL_CFG
sd = O.seeded_state_dict(cfg, 2, gain=1.5)
m = Generator(**cfg, mode=mode); m.load_state_dict(sd); m = m.cuda()
hop = m.hop
for it in range(4):
    torch.manual_seed(it)
    x = torch.randn(2, cfg["initial_channel"], 64); g = torch.randn(2, cfg["gin_channels"], 1); dy = torch.randn(2, 1, 64 * hop)
    m.zero_grad(set_to_none=True)
    y = m(x.cuda(), g.cuda()); y.backward(dy.cuda())
    yr, gr = oracle_run(cfg, sd, x, g, dy)
    grads = {n: p.grad.cpu() for n, p in m.named_parameters()}
    num = sum(float((grads[n].double() - gr[n]).pow(2).sum()) for n in grads); den = sum(float(gr[n].pow(2).sum()) for n in grads)
    worst = max((rel_l2(grads[n], gr[n]), n) for n in grads)
    print(f"iter {it}: fwd rel {rel_l2(y.detach().cpu(), yr):.2e} grads rel {(num/den)**0.5:.2e} worst {worst}")
