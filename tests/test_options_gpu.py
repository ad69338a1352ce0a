"""GPU: library options that are read once per process (environment) -- each runs in its own interpreter and is compared
with the default path on the same seeded inputs."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_STEP = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator
cfg = O.SMALL_CFG
m = Generator(**cfg, mode="bf16")
m.load_state_dict(O.seeded_state_dict(cfg, 21, gain=1.3))
m = m.cuda()
torch.manual_seed(8)
x = torch.randn(3, 64, 40, device="cuda", requires_grad=True)
g = torch.randn(3, 16, 1, device="cuda", requires_grad=True)
dy = torch.randn(3, 1, 40 * m.hop, device="cuda")
y = m(x, g)
y.backward(dy)
torch.cuda.synchronize()
out = {{n: p.grad.cpu().numpy() for n, p in m.named_parameters()}}
out["__y"] = y.detach().cpu().numpy(); out["__dx"] = x.grad.cpu().numpy(); out["__dg"] = g.grad.cpu().numpy()
np.savez(sys.argv[1], **out)
"""


def _run(env_extra, path):
    env = dict(os.environ)
    env.update(env_extra)
    subprocess.run([sys.executable, "-c", _STEP.format(root=ROOT), path], check=True, env=env, cwd=ROOT, timeout=300)
    return dict(np.load(path))


@pytest.mark.parametrize("option", [{"VCD_WG_FUSE": "1", "VCD_PAIR_BWD": "0"}, {"VCD_PAIR": "0"}, {"VCD_PAIR_BWD": "0"}, {"VCD_PAIR_MT": "1"},
                                    {"VCD_BWD_WHOLE": "0"}, {"VCD_GRAPHS": "0"}])
def test_option_matches_default_path(option):
    """VCD_WG_FUSE=1: weight gradients accumulated inside the (unfused) data-gradient launches (off by default); VCD_PAIR=0 /
    VCD_PAIR_BWD=0: unfused ResBlock pairs (forward and backward / backward only); VCD_PAIR_MT=1: one row tile per CTA
    tile; VCD_BWD_WHOLE=0: per-segment backward graphs; VCD_GRAPHS=0: no CUDA-graph replay.  Same results as
    the default path: waveform and dz bit for bit (same arithmetic), weight gradients to the fp32 summation order."""
    with tempfile.TemporaryDirectory() as d:
        ref = _run({}, os.path.join(d, "ref.npz"))
        alt = _run(option, os.path.join(d, "alt.npz"))
    assert np.array_equal(ref["__y"], alt["__y"])
    assert np.array_equal(ref["__dx"], alt["__dx"])
    for k in ref:
        a, b = ref[k].astype(np.float64), alt[k].astype(np.float64)
        assert np.linalg.norm(a - b) <= 1e-5 * np.linalg.norm(a) + 1e-12, k


_STEP_BASE = r"""
import sys, numpy as np, torch
sys.path.insert(0, {root!r})
from oracle import hifigan_oracle as O
from vcvits_b200 import Generator
cfg = O.BASE_CFG
torch.manual_seed(11)
m = Generator(**cfg, mode="bf16").cuda()
x = torch.randn(4, 256, 64, device="cuda", requires_grad=True)
g = torch.randn(4, 256, 1, device="cuda", requires_grad=True)
dy = torch.randn(4, 1, 64 * m.hop, device="cuda")
y = m(x, g)
y.backward(dy)
with torch.no_grad():
    y_inf = m(x.detach()[:, :, :48].contiguous(), g.detach())
torch.cuda.synchronize()
out = {{n: p.grad.cpu().numpy() for n, p in m.named_parameters()}}
out["__y"] = y.detach().cpu().numpy(); out["__dx"] = x.grad.cpu().numpy(); out["__dg"] = g.grad.cpu().numpy()
out["__y_inf"] = y_inf.cpu().numpy()
np.savez(sys.argv[1], **out)
"""


def test_cluster_multicast_weight_stages_are_bit_identical():
    """VCD_CONV_CLUSTER=1: every streamed-weight (>= 128-channel) forward / data-gradient launch runs as 2-CTA clusters
    whose CTAs each fetch half of a weight stage and multicast it to both (default: only the long inference-size
    launches).  Same arithmetic in the same order: waveform, dz, dg bit for bit at the base configuration (B = 4 x 64
    frames, training step + an inference call); weight gradients come from the unchanged weight-gradient kernel."""
    def run(env_extra, path):
        env = dict(os.environ)
        env.update(env_extra)
        subprocess.run([sys.executable, "-c", _STEP_BASE.format(root=ROOT), path], check=True, env=env, cwd=ROOT, timeout=300)
        return dict(np.load(path))
    with tempfile.TemporaryDirectory() as d:
        ref = run({"VCD_CONV_CLUSTER": "0"}, os.path.join(d, "ref.npz"))
        alt = run({"VCD_CONV_CLUSTER": "2"}, os.path.join(d, "alt.npz"))
    for k in ("__y", "__y_inf", "__dx", "__dg"):
        assert np.isfinite(ref[k]).all()
    assert np.array_equal(ref["__y"], alt["__y"]) and np.array_equal(ref["__y_inf"], alt["__y_inf"])
    # VCD_CONV_CLUSTER=2 may pick other rows per CTA tile (same per-element arithmetic: tiles do not interact)
    assert np.array_equal(ref["__dx"], alt["__dx"]) and np.array_equal(ref["__dg"], alt["__dg"])
    for k in ref:
        a, b = ref[k].astype(np.float64), alt[k].astype(np.float64)
        assert np.linalg.norm(a - b) <= 1e-5 * np.linalg.norm(a) + 1e-12, k
