"""The C-ABI library loads and exports every symbol include/b200mel.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200mel.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mel_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(native_lib):
    from asr_ttl_mtl_b200 import _native

    declared = _declared_symbols()
    assert len(declared) >= 12
    assert sorted(_native.SYMBOLS) == declared
    for name in declared:
        assert getattr(native_lib, name) is not None


def test_frames_rule_through_the_abi(native_lib, golden):
    from asr_ttl_mtl_b200 import _native

    for n, t in zip(golden["frames_len"], golden["frames_T"]):
        assert _native.frames(int(n)) == int(t)
    assert _native.frames(48000, 480000) == 3300 and _native.frames(1000, -5) == 6
    out = ctypes.c_int64(-1)
    assert native_lib.b200mel_frames(200, 0, ctypes.byref(out)) == _native.ERR_TOO_SHORT
    assert native_lib.b200mel_frames(100, 100, ctypes.byref(out)) == _native.ERR_TOO_SHORT
    assert native_lib.b200mel_frames(100, 101, ctypes.byref(out)) == _native.OK and out.value == 1
    assert native_lib.b200mel_frames(16000, 0, None) == _native.ERR_NULL_POINTER
    assert native_lib.b200mel_frames(-1, 0, ctypes.byref(out)) == _native.ERR_BAD_ARGUMENT


def test_status_strings_and_versions(native_lib):
    from asr_ttl_mtl_b200 import _native

    assert native_lib.b200mel_abi_version() == _native.ABI_VERSION
    assert native_lib.b200mel_status_string(0) == b"ok"
    assert b"n_mels" in native_lib.b200mel_status_string(_native.ERR_BAD_N_MELS)
    assert native_lib.b200mel_workspace_bytes(256) >= 1024 and native_lib.b200mel_workspace_bytes(256) % 256 == 0
    with pytest.raises(AssertionError):
        _native.check(_native.ERR_BAD_N_MELS)
    with pytest.raises(_native.B200MelError):
        _native.check(_native.ERR_TOO_SHORT)


def test_plan_argument_errors_need_no_gpu(native_lib):
    from asr_ttl_mtl_b200 import _native
    from asr_ttl_mtl_b200.filterbank import slaney_mel_filterbank

    handle = ctypes.c_void_p()
    bank = slaney_mel_filterbank(80)
    ptr = bank.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    assert native_lib.b200mel_plan_create(80, None, ctypes.byref(handle)) == _native.ERR_NULL_POINTER
    assert native_lib.b200mel_plan_create(64, ptr, ctypes.byref(handle)) == _native.ERR_BAD_N_MELS
    if not torch.cuda.is_available():
        # the product path fails loudly without a device: no CPU fallback
        assert native_lib.b200mel_plan_create(80, ptr, ctypes.byref(handle)) == _native.ERR_NO_DEVICE
    assert native_lib.b200mel_plan_destroy(None) == _native.OK


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "asr-ttl-mtl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle-side", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in text or f.endswith(".py"), f
