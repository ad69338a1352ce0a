"""CPU: the C-ABI shared library loads and exports every symbol include/vcd.h declares; no compute."""
import ctypes
import os
import re

import pytest

from vcvits_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vcd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vcd_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built():
    if _lib.needs_build():
        _lib.build()
    assert os.path.exists(_lib.LIB_PATH)


def test_every_declared_symbol_is_exported():
    if _lib.needs_build():
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vcd.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)


def test_version_and_error_strings():
    lib = _lib.load()
    assert b"sm_100a" in lib.vcd_version()
    assert isinstance(lib.vcd_last_error(), bytes)


def test_plan_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vcvits_b200 import Generator
    from oracle import hifigan_oracle as O
    m = Generator(**O.TINY_CFG)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 16, 4))  # CPU tensor: no fallback
    lib = _lib.load()
    handle = ctypes.c_void_p()
    cfg = m.config_struct()
    assert lib.vcd_plan_create(ctypes.byref(cfg), ctypes.byref(handle)) != 0
    assert b"no CUDA device" in lib.vcd_last_error()


def test_module_state_dict_matches_reference_layout():
    import torch
    from vcvits_b200 import Generator
    from oracle import hifigan_oracle as O
    for cfg in (O.TINY_CFG, O.TINY2_CFG, O.BASE_CFG):
        torch.manual_seed(1234)
        m = Generator(**cfg)
        sd = O.seeded_state_dict(cfg, 1234)
        msd = m.state_dict()
        assert list(msd.keys()) == list(sd.keys())
        for k in sd:
            assert msd[k].shape == sd[k].shape, k
            assert torch.equal(msd[k], sd[k]), k  # seeded default init reproduces the reference's
        m.load_state_dict(sd)  # both directions
        O.build(cfg).load_state_dict(msd)
