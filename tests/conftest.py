"""Shared fixtures.  `-m "not gpu"` runs here without a GPU; `-m gpu` runs on a B200."""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "logmel_golden.npz")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


class Golden:
    def __init__(self):
        self.z = np.load(GOLDEN_PATH, allow_pickle=False)
        self.meta = json.loads(bytes(self.z["manifest_json"]).decode())
        self.cases = self.meta["cases"]

    def __getitem__(self, key):
        return self.z[key]

    def signal(self, case):
        from oracle import signals

        x = signals.make_signal(case["kind"], case["n"], case["seed"])
        assert signals.digest(x) == case["input_sha256"], "synthetic input drifted from the golden run"
        return x

    def out(self, case):
        return self.z[f"out_{case['idx']}"]


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def emul():
    """The CPU choreography emulator (tests/emul/emul_fft.cpp), built with g++ on demand."""
    src = os.path.join(EMUL_DIR, "emul_fft.cpp")
    lib = os.path.join(EMUL_DIR, "libemul_fft.so")
    csrc = os.path.join(ROOT, "asr-ttl-mtl_b200", "csrc")
    deps = [src, os.path.join(csrc, "logmel_core.cuh"), os.path.join(csrc, "tables.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", lib, src],
                       check=True)
    h = ctypes.CDLL(lib)
    fp = ctypes.POINTER(ctypes.c_float)
    h.emul_fft_logmel.argtypes = [fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, fp, fp, ctypes.c_int]
    h.emul_fft_logmel.restype = ctypes.c_int
    h.emul_dft20.argtypes = [fp, fp]
    h.emul_key_encode.argtypes = [ctypes.c_float]
    h.emul_key_encode.restype = ctypes.c_uint32
    h.emul_key_decode.argtypes = [ctypes.c_uint32]
    h.emul_key_decode.restype = ctypes.c_float

    def run(x, n_mels, filters, padding=0, valid=None, normalise=True):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        frames = (n + max(padding, 0)) // 160
        out = np.zeros((n_mels, frames), np.float32)
        f = np.ascontiguousarray(filters, dtype=np.float32)
        st = h.emul_fft_logmel(x.ctypes.data_as(fp), n, n if valid is None else valid, padding, n_mels,
                               f.ctypes.data_as(fp), out.ctypes.data_as(fp), int(normalise))
        assert st == 0, f"emulator status {st}"
        return out

    h.run = run
    return h


@pytest.fixture(scope="session")
def emul_tc():
    """CPU emulator of the tcgen05 variant (tests/emul/emul_tc.cpp)."""
    src = os.path.join(EMUL_DIR, "emul_tc.cpp")
    lib = os.path.join(EMUL_DIR, "libemul_tc.so")
    csrc = os.path.join(ROOT, "asr-ttl-mtl_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("logmel_core.cuh", "tables.h", "tc_core.cuh", "tc_tables.h", "mel_bands.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-Wno-strict-aliasing",
                        "-DB200_HOST_HAS_CUDA_HEADERS", "-I/usr/local/cuda/include", "-o", lib, src], check=True)
    h = ctypes.CDLL(lib)
    fp = ctypes.POINTER(ctypes.c_float)
    h.emul_tc_logmel.argtypes = [fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, fp, fp, ctypes.c_int]
    h.emul_tc_logmel.restype = ctypes.c_int
    dp = ctypes.POINTER(ctypes.c_double)
    h.emul_tc_frame_spectrum.argtypes = [fp, dp, dp]
    h.emul_tc_frame_spectrum.restype = ctypes.c_int

    def run(x, n_mels, filters, padding=0, valid=None, normalise=True):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        frames = (n + max(padding, 0)) // 160
        out = np.zeros((n_mels, frames), np.float32)
        f = np.ascontiguousarray(filters, dtype=np.float32)
        st = h.emul_tc_logmel(x.ctypes.data_as(fp), n, n if valid is None else valid, padding, n_mels,
                              f.ctypes.data_as(fp), out.ctypes.data_as(fp), int(normalise))
        assert st == 0, f"emulator status {st}"
        return out

    h.run = run
    return h


@pytest.fixture(scope="session")
def native_lib():
    """libb200mel.so, built in-tree by nvcc (cross-compiles without a GPU)."""
    import __graft_entry__ as entry

    entry.build()
    from asr_ttl_mtl_b200 import _native

    return _native.load()


@pytest.fixture(scope="session")
def b200():
    """The product package, on a box with a GPU and a built library."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as entry

    entry.build()
    import asr_ttl_mtl_b200

    return asr_ttl_mtl_b200
