"""ctypes binding of the C ABI in include/b200mel.h (libb200mel.so).

There is no fallback: if the shared library is missing or a call fails, this
module raises.  ``__graft_entry__.build()`` (or ``python asr-ttl-mtl_b200/build.py``)
produces the library in-tree at ``asr-ttl-mtl_b200/lib/libb200mel.so``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MEL_LIB") or os.path.join(HERE, "lib", "libb200mel.so")   # (B200MEL_LIB: bring-up builds)

# enums of include/b200mel.h
OK = 0
ERR_NULL_POINTER, ERR_BAD_N_MELS, ERR_TOO_SHORT, ERR_BAD_ARGUMENT, ERR_BAD_FILTERS, ERR_CUDA, ERR_NO_DEVICE = range(1, 8)
DTYPE_F32, DTYPE_S16 = 0, 1
VARIANT_AUTO, VARIANT_FFT, VARIANT_TCGEN05 = 0, 1, 2
FLAG_GLOBAL_MAX = 1
FLAG_TILE_KEYS = 2
FLAG_OUT_F16 = 4
FLAG_DEFER_CLAMP = 8
ABI_VERSION = 2

#: every symbol include/b200mel.h declares: (restype, argtypes)
SYMBOLS = {
    "b200mel_abi_version": (c_int, []),
    "b200mel_status_string": (c_char_p, [c_int]),
    "b200mel_last_cuda_error": (c_char_p, []),
    "b200mel_frames": (c_int, [c_int64, c_int64, POINTER(c_int64)]),
    "b200mel_plan_create": (c_int, [c_int, POINTER(c_float), POINTER(c_void_p)]),
    "b200mel_plan_destroy": (c_int, [c_void_p]),
    "b200mel_plan_n_mels": (c_int, [c_void_p]),
    "b200mel_workspace_bytes": (c_size_t, [c_int64]),
    "b200mel_workspace_bytes_tiles": (c_size_t, [c_int64, c_int64]),
    "b200mel_logmel_device": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                      c_void_p, c_void_p, c_uint, c_int, c_void_p]),
    "b200mel_normalise_device": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_uint, c_void_p]),
    "b200mel_stem_conv1_gelu_device": (c_int, [c_void_p, c_void_p, c_uint, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int,
                                               c_void_p, c_void_p]),
    "b200mel_stem_conv1_gelu_fm16_device": (c_int, [c_void_p, c_void_p, c_uint, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int,
                                                    c_void_p, c_void_p]),
    "b200mel_stem_conv2_gelu_device": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_uint, c_void_p]),
    "b200mel_mel_windows_device": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_uint, c_void_p]),
    "b200mel_logmel_host": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                    c_void_p, c_uint, c_int]),
    "b200mel_launch_count": (c_uint64, []),
    "b200mel_kernel_fault": (c_uint, [POINTER(c_uint)]),
    "b200mel_profile_enable": (c_int, [c_int]),
    "b200mel_profile_collect": (c_int, [POINTER(ctypes.c_double), POINTER(c_uint64)]),
}

_lib = None
_lock = threading.Lock()


class B200MelError(RuntimeError):
    """A b200mel call returned a non-zero status."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


def load() -> ctypes.CDLL:
    """Load libb200mel.so and bind every declared symbol; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"b200mel CUDA library not found at {LIB_PATH}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `python asr-ttl-mtl_b200/build.py`. "
                "There is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        got = lib.b200mel_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libb200mel.so ABI version {got}, binding expects {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def status_message(status: int) -> str:
    lib = load()
    msg = lib.b200mel_status_string(status).decode()
    if status == ERR_CUDA:
        msg += ": " + lib.b200mel_last_cuda_error().decode()
    return msg


def check(status: int) -> None:
    """Map a C status to the exception type the reference raises for the same condition."""
    if status == OK:
        return
    msg = status_message(status)
    if status == ERR_BAD_N_MELS:
        raise AssertionError(msg)  # whisper/audio.py:103 is an assert
    raise B200MelError(status, f"b200mel: {msg}")


def frames(n_samples: int, padding: int = 0) -> int:
    out = c_int64(0)
    check(load().b200mel_frames(int(n_samples), int(padding), ctypes.byref(out)))
    return out.value


def launch_count() -> int:
    return int(load().b200mel_launch_count())


def kernel_fault() -> tuple[int, int]:
    """(code, CTA) of a timed-out hand-over inside the tcgen05 kernel, (0, 0) if there never was one; synchronises."""
    cta = c_uint(0)
    code = int(load().b200mel_kernel_fault(ctypes.byref(cta)))
    return code, int(cta.value)


PROFILE_KINDS = ("fft_pass", "normalise", "tcgen05_pass", "other")


def profile_enable(on: bool) -> None:
    check(load().b200mel_profile_enable(1 if on else 0))


def profile_collect() -> dict:
    """{kind: (total_ms, launches)} of the launches bracketed since the last collect."""
    ms = (ctypes.c_double * len(PROFILE_KINDS))()
    n = (c_uint64 * len(PROFILE_KINDS))()
    check(load().b200mel_profile_collect(ms, n))
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(PROFILE_KINDS)}
