"""vcvits_b200 -- B200-native (sm_100a) HiFi-GAN waveform decoder for VCVITS.

Only the hot path named by BASELINE.json lives here: the `Generator` behind `net_g.dec`
(reference: vits/model/synthesizers/synthesizer_tts.py:71-78,140), its CUDA kernels (``csrc/``) and the C-ABI
binding (``_lib``).  See DESIGN.md and INTEGRATION.md.
"""
from .generator import Generator, LRELU_SLOPE  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["Generator", "LRELU_SLOPE"]
__version__ = "0.1.0"
