"""Host-side behaviour of vcvits_b200.Generator that does not need a GPU."""
import copy
import pickle

import pytest
import torch

from oracle import hifigan_oracle as O
from vcvits_b200 import Generator


def _inputs(cfg, B=1, T=4):
    return torch.randn(B, cfg["initial_channel"], T), torch.randn(B, cfg["gin_channels"], 1)


def test_cpu_tensors_are_refused_not_computed():
    """No CPU fallback: the product path fails loudly instead of routing through PyTorch or the oracle."""
    m = Generator(**O.TINY_CFG, mode="fp32")
    x, g = _inputs(O.TINY_CFG)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, g)


def test_unknown_mode_is_rejected():
    with pytest.raises((ValueError, RuntimeError)):
        Generator(**O.TINY_CFG, mode="fp8")


@pytest.mark.parametrize("cfg", [O.TINY_CFG, O.TINY2_CFG])
def test_copy_and_pickle_keep_parameters_and_drop_device_state(cfg):
    m = Generator(**cfg, mode="bf16")
    m.load_state_dict(O.seeded_state_dict(cfg, 5, gain=1.1))
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert list(clone.state_dict().keys()) == list(m.state_dict().keys())
        for a, b in zip(m.state_dict().values(), clone.state_dict().values()):
            assert torch.equal(a, b)
        assert clone.mode == m.mode


def test_state_dict_round_trip_with_reference_key_order():
    """Checkpoints of the reference layout load strictly (weight_g / weight_v pairs, conv_post without bias)."""
    sd = O.seeded_state_dict(O.SMALL_CFG, 11)
    m = Generator(**O.SMALL_CFG, mode="fp32")
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    out = m.state_dict()
    assert list(out.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(out[k], sd[k]), k
    assert "conv_post.bias" not in out and "conv_post.weight" in out
    assert any(k.endswith("weight_g") for k in out) and any(k.endswith("weight_v") for k in out)


def test_remove_weight_norm_keys_and_slot_resolution():
    """modules.py:218-222 convention on the host side: baked `weight` replaces weight_g / weight_v (key order of
    torch.nn.utils.remove_weight_norm), and the library's parameter table (reference checkpoint keys) resolves to the
    baked weight in the weight_v slot and to None in the weight_g slot."""
    import warnings
    sd = O.seeded_state_dict(O.TINY_CFG, 4, gain=1.2)
    m = Generator(**O.TINY_CFG, mode="fp32")
    m.load_state_dict(sd)
    names = list(sd.keys())                      # == the library's parameter table (tests/test_cabi_symbols.py)
    before = m._resolve_params(names)
    assert all(t is not None for t in before)
    m.remove_weight_norm()
    ref = O.build(O.TINY_CFG, sd)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for mod in list(ref.ups) + [c for rb in ref.resblocks for c in list(rb.convs1) + list(rb.convs2)]:
            torch.nn.utils.remove_weight_norm(mod)
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert torch.allclose(m.state_dict()[k], v, rtol=1e-6, atol=1e-8), k
    after = m._resolve_params(names)
    for n, t in zip(names, after):
        if n.endswith("weight_g"):
            assert t is None, n
        elif n.endswith("weight_v"):
            assert t is m.get_submodule(n.rsplit(".", 1)[0]).weight, n
        else:
            assert t is not None, n
    with pytest.raises(RuntimeError):
        m.load_state_dict(sd)
