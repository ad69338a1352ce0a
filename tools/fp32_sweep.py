import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import hifigan_oracle as O
from tests.helpers import b200_run, oracle_run, rel_l2
cfg = O.SMALL_CFG
sd = O.seeded_state_dict(cfg, 1234, gain=1.3)
for T in [int(a) for a in sys.argv[1:]] or [8, 16, 24, 25, 32, 40, 48]:
    torch.manual_seed(0)
    x = torch.randn(2, 64, T); g = torch.randn(2, 16, 1); dy = torch.randn(2, 1, T * 16)
    yr, gr = oracle_run(cfg, sd, x, g, dy)
    y, grads, m = b200_run(cfg, sd, x, g, dy, mode="fp32")
    rows = sorted(((rel_l2(grads[n], gr[n]), n) for n in gr), reverse=True)
    print(T, f"fwd {rel_l2(y, yr):.1e}", " ".join(f"{n}={r:.1e}" for r, n in rows[:3]), flush=True)
