"""In-tree build of the CUDA shared libraries (sm_100a only; nvcc cross-compiles without a GPU).

    python asr-ttl-mtl_b200/build.py [--force] [--verbose]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    asr-ttl-mtl_b200/lib/libb200mel.so      the product: kernels + C ABI of include/b200mel.h
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb200mel.so")

SOURCES = ["b200mel_api.cu", "logmel_fft.cu", "logmel_tc.cu", "stem_conv.cu", "stem_conv2.cu", "mel_windows.cu"]
HEADERS = ["kernels.h", "logmel_core.cuh", "tables.h", "tc_core.cuh", "tc_tables.h", "mel_bands.h", os.path.join(ROOT, "include", "b200mel.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "--expt-relaxed-constexpr",
    "-shared",
]


def nvcc_path() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the b200mel CUDA library cannot be built")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True builds lib/libb200mel_trace.so instead: the same library with the tcgen05 kernel's timeline stamps
    compiled in (-DB200MEL_TC_TRACE; tools/tc_trace.py loads it through B200MEL_LIB)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if trace:
        _build(os.path.join(LIBDIR, "libb200mel_switches.so"), srcs, force, verbose, ["-DB200MEL_TC_SWITCHES"])
        return _build(os.path.join(LIBDIR, "libb200mel_trace.so"), srcs, force, verbose, ["-DB200MEL_TC_TRACE"])
    return _build(LIB, srcs, force, verbose, [])


def _build(LIB: str, srcs: list[str], force: bool, verbose: bool, extra: list[str]) -> str:
    deps = srcs + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB, *srcs]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}) building {LIB}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, trace="--trace" in sys.argv))
