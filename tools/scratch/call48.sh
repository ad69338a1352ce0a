cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_boundary_gpu.py -q -x -k deterministic 2>&1 | tail -15
timeout 300 python - <<'PY' 2>&1 | tail -8
import torch, time
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator
torch.manual_seed(0)
m = Generator(**O.BASE_CFG, mode="bf16").cuda()
x = torch.randn(16, 256, 32, device="cuda"); g = torch.randn(16, 256, 1, device="cuda"); dy = torch.randn(16, 1, 32*512, device="cuda")
def step():
    for p in m.parameters(): p.grad = None
    xx, gg = x.clone().requires_grad_(True), g.clone().requires_grad_(True)
    m(xx, gg).backward(dy)
for det in (False, True, False):
    m.deterministic = det
    for _ in range(5): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): step()
    b.record(); torch.cuda.synchronize()
    print("deterministic", det, a.elapsed_time(b) / 20, "ms/step")
PY
