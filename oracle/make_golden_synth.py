"""Pins oracle/synth_parts.py against the reference's own classes and writes tests/golden/synth_parts_tiny.npz.

Runs only where /root/reference exists (the build container).  For every restated module the reference class is
built on a tiny configuration, its state_dict is loaded into the restatement (same parameter names) and both are run
on the same seeded inputs; the outputs must agree to 1e-6 (same PyTorch kernels, same order of operations).  The
fixture stores the weights, inputs and reference outputs so tests/test_synth_parts_cpu.py can replay the comparison
on a box without the reference tree.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("fairseq", types.ModuleType("fairseq"))   # content_encoder.py imports it at module level only

from oracle import synth_parts as S  # noqa: E402
from vits.model.encoders.content_encoder import PreloadHubertContentEncoder as RefEncP  # noqa: E402
from vits.model.encoders.posterior_encoder import PosteriorEncoder as RefEncQ  # noqa: E402
from vits.model.flow import ResidualCouplingBlock as RefFlow  # noqa: E402
import vits.commons as ref_commons  # noqa: E402


def main():
    out = {}
    torch.manual_seed(7)
    B, T, H, C = 2, 23, 16, 8
    # enc_p
    ref = RefEncP(C, H, 32, 2, 2, 3, 0.0, 24, 10).eval()
    mine = S.PreloadHubertContentEncoder(C, H, 32, 2, 2, 3, 0.0, 24, 10).eval()
    mine.load_state_dict(ref.state_dict())
    feats, lens = torch.randn(B, 24, T), torch.tensor([T, T - 6])
    pitch = torch.randint(1, 10, (B, T))
    with torch.no_grad():
        r, m = ref(feats, lens, pitch, lens), mine(feats, lens, pitch, lens)
    for a, b in zip(r, m):
        assert torch.allclose(a, b, atol=1e-6), float((a - b).abs().max())
    out.update({f"enc_p::{k}": v.numpy() for k, v in ref.state_dict().items()})
    out.update(enc_p_feats=feats.numpy(), enc_p_lens=lens.numpy(), enc_p_pitch=pitch.numpy(), enc_p_x=r[0].numpy(),
               enc_p_m=r[1].numpy(), enc_p_logs=r[2].numpy())
    # enc_q (the sampled z uses randn_like: compare the deterministic outputs m, logs and z under a fixed seed)
    ref = RefEncQ(12, C, H, 5, 1, 3, gin_channels=6).eval()
    mine = S.PosteriorEncoder(12, C, H, 5, 1, 3, gin_channels=6).eval()
    mine.load_state_dict(ref.state_dict())
    spec, g = torch.rand(B, 12, T), torch.randn(B, 6, 1)
    with torch.no_grad():
        torch.manual_seed(1); r = ref(spec, lens, g=g)
        torch.manual_seed(1); m = mine(spec, lens, g=g)
    for a, b in zip(r, m):
        assert torch.allclose(a, b, atol=1e-6), float((a - b).abs().max())
    out.update({f"enc_q::{k}": v.numpy() for k, v in ref.state_dict().items()})
    out.update(enc_q_spec=spec.numpy(), enc_q_g=g.numpy(), enc_q_m=r[1].numpy(), enc_q_logs=r[2].numpy())
    # flow (post convs are zero-initialised in the reference: randomise them so the coupling is exercised)
    ref = RefFlow(C, H, 5, 1, 2, gin_channels=6).eval()
    for p in ref.parameters():
        if float(p.abs().sum()) == 0:
            p.data.normal_(0, 0.1)
    mine = S.ResidualCouplingBlock(C, H, 5, 1, 2, gin_channels=6).eval()
    mine.load_state_dict(ref.state_dict())
    z = torch.randn(B, C, T)
    mask = torch.unsqueeze(ref_commons.sequence_mask(lens, T), 1).float()
    with torch.no_grad():
        r, m = ref(z, mask, g=g), mine(z, mask, g=g)
        rr, mr = ref(r, mask, g=g, reverse=True), mine(m, mask, g=g, reverse=True)
    assert torch.allclose(r, m, atol=1e-6) and torch.allclose(rr, mr, atol=1e-6)
    out.update({f"flow::{k}": v.numpy() for k, v in ref.state_dict().items()})
    out.update(flow_z=z.numpy(), flow_mask=mask.numpy(), flow_out=r.numpy())
    # slicing helpers
    ids = torch.tensor([3, 0])
    assert torch.equal(ref_commons.slice_segments(z, ids, 5), S.slice_segments(z, ids, 5))
    torch.manual_seed(3); a = ref_commons.rand_slice_segments(z, lens, 4)
    torch.manual_seed(3); b = S.rand_slice_segments(z, lens, 4)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    path = os.path.join(ROOT, "tests", "golden", "synth_parts_tiny.npz")
    np.savez_compressed(path, **out)
    print("pinned enc_p / enc_q / flow / slicing against the reference;", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
