"""CPU: oracle/synth_parts.py (the non-decoder parts used by the configs[4] harness) replayed against the fixture that
oracle/make_golden_synth.py generated from the reference's own classes."""
import os

import numpy as np
import torch

from oracle import synth_parts as S


def _load(golden_dir, prefix):
    d = np.load(os.path.join(golden_dir, "synth_parts_tiny.npz"))
    sd = {k[len(prefix) + 2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith(prefix + "::")}
    return d, sd


def test_enc_p_matches_reference(golden_dir):
    d, sd = _load(golden_dir, "enc_p")
    m = S.PreloadHubertContentEncoder(8, 16, 32, 2, 2, 3, 0.0, 24, 10).eval()
    m.load_state_dict(sd)
    lens = torch.from_numpy(d["enc_p_lens"])
    with torch.no_grad():
        x, mean, logs, _ = m(torch.from_numpy(d["enc_p_feats"]), lens, torch.from_numpy(d["enc_p_pitch"]), lens)
    for got, key in ((x, "enc_p_x"), (mean, "enc_p_m"), (logs, "enc_p_logs")):
        assert torch.allclose(got, torch.from_numpy(d[key]), atol=1e-5), key


def test_enc_q_and_flow_match_reference(golden_dir):
    d, sd = _load(golden_dir, "enc_q")
    m = S.PosteriorEncoder(12, 8, 16, 5, 1, 3, gin_channels=6).eval()
    m.load_state_dict(sd)
    lens, g = torch.from_numpy(d["enc_p_lens"]), torch.from_numpy(d["enc_q_g"])
    with torch.no_grad():
        _, mean, logs, _ = m(torch.from_numpy(d["enc_q_spec"]), lens, g=g)
    assert torch.allclose(mean, torch.from_numpy(d["enc_q_m"]), atol=1e-5)
    assert torch.allclose(logs, torch.from_numpy(d["enc_q_logs"]), atol=1e-5)
    d, sd = _load(golden_dir, "flow")
    f = S.ResidualCouplingBlock(8, 16, 5, 1, 2, gin_channels=6).eval()
    f.load_state_dict(sd)
    z, mask = torch.from_numpy(d["flow_z"]), torch.from_numpy(d["flow_mask"])
    with torch.no_grad():
        out = f(z, mask, g=g)
        back = f(out, mask, g=g, reverse=True)
    assert torch.allclose(out, torch.from_numpy(d["flow_out"]), atol=1e-5)
    assert torch.allclose(back * mask, z * mask, atol=1e-4)       # the flow is invertible


def test_slice_segments():
    z = torch.arange(2 * 3 * 10, dtype=torch.float32).view(2, 3, 10)
    out = S.slice_segments(z, torch.tensor([2, 5]), 4)
    assert torch.equal(out[0], z[0, :, 2:6]) and torch.equal(out[1], z[1, :, 5:9])
