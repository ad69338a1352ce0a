"""CPU: the bf16-emulating oracle (oracle/bf16_emulation.py) -- consistency with the plain oracle, the chaos of free-running
bf16 rounding, and the teacher-forcing mechanism the GPU local-parity tests rely on."""
import torch

from oracle import bf16_emulation as E
from oracle import hifigan_oracle as O
from tests.helpers import oracle_run, rel_l2


def _case(cfg, B=2, T=19, seed=3):
    sd = O.seeded_state_dict(cfg, 77, gain=1.3)
    torch.manual_seed(seed)
    x, g = torch.randn(B, cfg["initial_channel"], T), torch.randn(B, cfg["gin_channels"], 1)
    hop = 1
    for u in cfg["upsample_rates"]:
        hop *= u
    return sd, x, g, torch.randn(B, 1, T * hop)


def test_emulation_is_the_oracle_plus_bf16_noise():
    for cfg in (O.SMALL_CFG, O.TINY2_CFG):
        sd, x, g, dy = _case(cfg)
        y_ref, gref = oracle_run(cfg, sd, x, g, dy)
        y, grads = E.run(cfg, sd, x, g, dy)
        assert float((y - y_ref).abs().max()) <= 1e-2                      # the north star's bf16 waveform bound
        num = sum(float((grads[n] - gref[n]).pow(2).sum()) for n in gref)
        den = sum(float(gref[n].pow(2).sum()) for n in gref)
        assert (num / den) ** 0.5 <= 0.12
        assert set(grads) == set(gref)


def test_teacher_forcing_isolates_every_storage_point():
    """An fp32 run records its own storage points; an fp64 run forced with them must agree at every point to fp32
    accumulation noise, and so must all parameter gradients -- although the same two runs, free running, differ by
    percents on some tensors (rounding chaos)."""
    cfg = O.SMALL_CFG
    sd, x, g, dy = _case(cfg)
    rec = E.Stored(record=True)
    y32, g32 = E.run(cfg, sd, x, g, dy, dtype=torch.float32, stored=rec)
    assert {"xin", "a0", "a2", "ua1", "ma1.2.2", "xa0.0.1", "d0", "duz0", "dm0.1.0", "Gt1.2.2", "Gi1"} <= set(rec.tensors)
    tf = E.Stored(rec.tensors)
    y64, g64 = E.run(cfg, sd, x, g, dy, dtype=torch.float64, stored=tf)
    assert set(tf.report) == set(rec.tensors)
    assert max(r["rel_l2"] for r in tf.report.values()) <= 1e-3
    assert max(rel_l2(g32[n], g64[n]) for n in g64) <= 1e-5
    _, free = E.run(cfg, sd, x, g, dy, dtype=torch.float64)
    assert max(rel_l2(g32[n], free[n]) for n in free) >= 1e-3    # chaos: free-running runs decorrelate
