"""Debug / smoke: N-rank gradient-synchronised decoder step (torchrun).  Prints progress to stderr."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator

def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.strftime('%H:%M:%S')}", *a, file=sys.stderr, flush=True)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
log("init")
dist.init_process_group("nccl", device_id=dev)
t = torch.ones(4, device=dev) * (rank + 1)
dist.all_reduce(t)
torch.cuda.synchronize()
log("allreduce ok", t.tolist())
cfg = O.SMALL_CFG if len(sys.argv) < 2 else getattr(O, sys.argv[1])
torch.manual_seed(1234)
m = Generator(**cfg, mode="bf16").to(dev)
single = Generator(**cfg, mode="bf16").to(dev)
single.load_state_dict(m.state_dict())
m.set_gradient_sync(dist.group.WORLD)
B, T = 4, 32
g = torch.Generator().manual_seed(7)
xs = torch.randn(world * B, cfg["initial_channel"], T, generator=g).to(dev)
gs = torch.randn(world * B, cfg["gin_channels"], 1, generator=g).to(dev)
dys = torch.randn(world * B, 1, T * m.hop, generator=g).to(dev)
for it in range(3):
    m.zero_grad(set_to_none=True)
    sl = slice(rank * B, (rank + 1) * B)
    y = m(xs[sl], gs[sl])
    log("fwd enqueued", it)
    y.backward(dys[sl])
    log("bwd enqueued", it)
    torch.cuda.synchronize()
    log("step done", it)
# reference: the same global batch on one GPU; DDP averages per-rank gradients of per-rank sums -> sum/world
single.zero_grad(set_to_none=True)
single(xs, gs).backward(dys)
torch.cuda.synchronize()
num = den = 0.0
for (n, p), (_, q) in zip(m.named_parameters(), single.named_parameters()):
    num += float((p.grad * world - q.grad).double().pow(2).sum()); den += float(q.grad.double().pow(2).sum())
log("DDP vs single-GPU gradient rel-l2:", (num / den) ** 0.5)
worst = sorted(((float((p.grad * world - q.grad).double().norm() / (q.grad.double().norm() + 1e-30)), n)
                for (n, p), (_, q) in zip(m.named_parameters(), single.named_parameters())), reverse=True)[:8]
if rank == 0:
    for e, n in worst:
        log(f"  {n:40s} rel-l2 {e:.3e}")
dist.destroy_process_group()
