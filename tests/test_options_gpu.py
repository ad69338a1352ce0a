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
