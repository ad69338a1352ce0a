import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import hifigan_oracle as O
from tests.helpers import b200_step_with_stored
cfg = O.BASE_CFG
sd = O.seeded_state_dict(cfg, 1234, gain=1.2)
torch.manual_seed(3)
B, T = 16, 32
x = torch.randn(B, 256, T); g = torch.randn(B, 256, 1); dy = torch.randn(B, 1, T * 512)
y, grads, stored, m = b200_step_with_stored(cfg, sd, x, g, dy)
out = {k: v.float().numpy() for k, v in stored.items()}
out["__order"] = np.array(list(stored.keys()))
np.savez(sys.argv[1], **out)
print("saved", len(stored))
