"""Rebind the reference's audio API to the B200 front-end, leaving its consumers unchanged.

The reference has no plugin registry: its consumers bind the three functions by name at
import time — ``from whisper.audio import load_audio, pad_or_trim, log_mel_spectrogram``
(speech_disorder/dataset.py:7), ``from .audio import ...`` (whisper/transcribe.py:11-19,
whisper/__init__.py:11).  ``install()`` therefore rewrites those module namespaces.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

from . import audio as _audio

REBOUND_NAMES = ("log_mel_spectrogram", "pad_or_trim", "mel_filters")
#: modules of the reference that hold by-name copies of the audio API
CONSUMER_MODULES = ("whisper.audio", "whisper", "whisper.transcribe", "speech_disorder.dataset")

_saved: List[Tuple[object, str, object]] = []


def install(import_missing: bool = False) -> Dict[str, List[str]]:
    """Point every already-imported consumer module at this package's functions.

    With ``import_missing=True`` the consumer modules are imported first (they must be
    importable, i.e. the reference tree is on ``sys.path``).  Returns what was rebound.
    Forked DataLoader workers cannot initialise CUDA: run the reference's loaders with
    ``num_workers=0`` or ``multiprocessing_context="spawn"``.
    """
    rebound: Dict[str, List[str]] = {}
    for mod_name in CONSUMER_MODULES:
        module = sys.modules.get(mod_name)
        if module is None and import_missing:
            try:
                module = importlib.import_module(mod_name)
            except Exception:  # a consumer that cannot be imported here has nothing to rebind
                module = None
        if module is None:
            continue
        for name in REBOUND_NAMES:
            if hasattr(module, name):
                _saved.append((module, name, getattr(module, name)))
                setattr(module, name, getattr(_audio, name))
                rebound.setdefault(mod_name, []).append(name)
    return rebound


def uninstall() -> None:
    """Undo :func:`install`."""
    while _saved:
        module, name, original = _saved.pop()
        setattr(module, name, original)
