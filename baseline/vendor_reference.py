#!/usr/bin/env python
"""Vendor the UNMODIFIED reference front-end into baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

    python baseline/vendor_reference.py [--reference /root/reference]

Copies whisper/audio.py, whisper/utils.py and whisper/assets/mel_filters.npz byte for byte into
baseline/_ref/refwhisper/ next to an empty __init__.py, so `bench.py --impl reference` and the `cpu_baseline`
leg can time the reference's own `log_mel_spectrogram` (whisper/audio.py:110-157) where /root/reference does not
exist; and the reference's two Python packages (whisper/, speech_disorder/), unmodified, into baseline/_ref/tree/ so
that tests/test_consumers.py can run the REAL consumers of the path - MultiTaskSpeechDataset + collate_fn
(speech_disorder/dataset.py) and transcribe() (whisper/transcribe.py) - on the GPU box as well.  Nothing here enters the
git history, and no product code imports it.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "refwhisper")
FILES = ["audio.py", "utils.py", os.path.join("assets", "mel_filters.npz")]


def vendor(reference: str = "/root/reference") -> bool:
    src_root = os.path.join(reference, "whisper")
    if not os.path.isdir(src_root):
        return False
    os.makedirs(os.path.join(DEST, "assets"), exist_ok=True)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(src_root, rel), os.path.join(DEST, rel)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "__init__.py"), "w") as f:
        f.write('"""Package shell around the vendored, unmodified reference files (see baseline/vendor_reference.py)."""\n')
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src_root, "sha256": manifest}, f, indent=1)
    tree = os.path.join(HERE, "_ref", "tree")
    for package in ("whisper", "speech_disorder"):
        dst = os.path.join(tree, package)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(reference, package), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


def tree_path():
    """Directory to put on sys.path to import the reference's `whisper` and `speech_disorder` packages: the reference
    checkout itself where it exists, else the vendored copy, else None."""
    if os.path.isdir("/root/reference/whisper"):
        return "/root/reference"
    tree = os.path.join(HERE, "_ref", "tree")
    return tree if os.path.isdir(os.path.join(tree, "whisper")) else None


def load():
    """The vendored reference `audio` module, or None if baseline/_ref was never built."""
    if not os.path.exists(os.path.join(DEST, "audio.py")):
        return None
    parent = os.path.dirname(DEST)
    if parent not in sys.path:
        sys.path.insert(0, parent)
    import importlib

    return importlib.import_module("refwhisper.audio")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ok = vendor(ap.parse_args().reference)
    print("vendored into", DEST if ok else "(reference tree not found)")
