"""ctypes binding of the C ABI in ``include/vcd.h`` (the library is ``vcvits_b200/libvcd.so``).

There is deliberately no fallback: if the shared library is missing or fails to load, every compute entry
point raises.  ``build()`` compiles it in-tree with nvcc for sm_100a (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvcd.so")
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["vcd_api.cu"]
HEADERS = ["common.cuh", "simt_kernels.cuh", "fold.cuh", "fold_fast.cuh", "plan.h", "tc_conv.cuh", "tc_kernels.cuh", "tc_pair.cuh",
           "mel_loss.cuh"]

MODE_FP32 = 0
MODE_BF16 = 1

MAX_KERNELS, MAX_DILATIONS, MAX_UPSAMPLES = 8, 3, 8


class VcdConfig(C.Structure):
    _fields_ = [
        ("initial_channel", C.c_int32),
        ("resblock", C.c_int32),
        ("num_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * MAX_KERNELS),
        ("resblock_dilation_sizes", (C.c_int32 * MAX_DILATIONS) * MAX_KERNELS),
        ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * MAX_UPSAMPLES),
        ("upsample_kernel_sizes", C.c_int32 * MAX_UPSAMPLES),
        ("upsample_initial_channel", C.c_int32),
        ("gin_channels", C.c_int32),
    ]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build vcvits_b200/libvcd.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    hdr = os.path.join(os.path.dirname(_HERE), "include", "vcd.h")
    return os.path.exists(hdr) and os.path.getmtime(hdr) > t


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a.  Returns the path of the .so."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lock = threading.Lock()
_lib = None

_FLOATPP = C.POINTER(C.c_void_p)

_SIGNATURES = {
    "vcd_version": (C.c_char_p, []),
    "vcd_last_error": (C.c_char_p, []),
    "vcd_plan_create": (C.c_int, [C.POINTER(VcdConfig), C.POINTER(C.c_void_p)]),
    "vcd_plan_destroy": (None, [C.c_void_p]),
    "vcd_num_params": (C.c_int, [C.c_void_p]),
    "vcd_param_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "vcd_total_param_elems": (C.c_int64, [C.c_void_p]),
    "vcd_hop": (C.c_int, [C.c_void_p]),
    "vcd_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vcd_fold_weights": (C.c_int, [C.c_void_p, C.c_int, _FLOATPP, C.c_void_p]),
    "vcd_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vcd_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               _FLOATPP, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_void_p]),
    "vcd_forward_sliced": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vcd_backward_sliced": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, _FLOATPP, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32,
                                      C.c_void_p]),
    "vcd_set_gradient_scale": (C.c_int, [C.c_void_p, C.c_float]),
    "vcd_set_deterministic": (C.c_int, [C.c_void_p, C.c_int]),
    "vcd_segment_events_valid": (C.c_int, [C.c_void_p]),
    "vcd_stream_wait_segment": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vcd_num_backward_segments": (C.c_int, [C.c_void_p]),
    "vcd_segment_params": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]),
    "vcd_host_call_extra_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "vcd_synthesize_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
    "vcd_launch_count": (C.c_uint64, [C.c_int]),
    "vcd_profile_enable": (C.c_int, [C.c_int]),
    "vcd_profile_num_classes": (C.c_int, []),
    "vcd_profile_class_name": (C.c_char_p, [C.c_int]),
    "vcd_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double)]),
    "vcd_profile_dump": (C.c_int, [C.c_char_p]),
    "vcd_phase_dump": (C.c_int, [C.c_int]),
    "vcd_debug_read_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "vcd_layer_path": (C.c_char_p, [C.c_void_p, C.c_int, C.c_int]),
    "vcd_debug_ws_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vcd_debug_tc_paths": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    # mel / STFT loss tail (vcvits_b200/mel.py)
    "vcd_mel_plan_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "vcd_mel_plan_destroy": (None, [C.c_void_p]),
    "vcd_mel_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "vcd_mel_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "vcd_mel_spectrogram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
    "vcd_mel_debug_path": (C.c_int, [C.c_void_p, C.c_int]),
    "vcd_mel_loss_sliced": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
    "vcd_mel_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                               C.c_int, C.c_int, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load() -> C.CDLL:
    """Load libvcd.so (never builds implicitly on a GPU box: the prebuilt in-tree .so is what ships)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(vcvits_b200 has no CPU or PyTorch fallback for the decoder path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vcd_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed: {msg}")
