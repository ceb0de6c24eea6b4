"""Seeded synthetic 16 kHz waveforms shared by the golden-vector generator, the
parity tests and bench.py (SURVEY.md §8d "Extra parity distributions").

TEST INFRASTRUCTURE — nothing under oracle/ is on the product path.

All generators use ``numpy.random.default_rng(seed)`` (PCG64), float64
arithmetic and a final cast to float32, so the same (kind, n, seed) gives the
same bytes in this container and on the GPU box (same image, same numpy).
"""
from __future__ import annotations

import hashlib

import numpy as np

SAMPLE_RATE = 16000

KINDS = (
    "gauss",        # 0.1 * randn                      (BASELINE configs 1-5)
    "uniform",      # U[-1, 1]
    "sine440",      # 0.5 * sin(2 pi 440 t)
    "sine1k_noise", # 0.8 * sin(2 pi 1000 t) + 1e-5 * randn
    "chirp",        # linear chirp 50 Hz -> 7.95 kHz over the clip
    "two_tone",     # 1 kHz at 0 dB + 3.7 kHz at -70 dB
    "pcm16",        # round(3000 * randn) / 32768  (the grid load_audio produces, audio.py:62)
    "burst",        # first 1/6 of the clip gauss, rest exact zeros (5 s burst + 25 s silence at 30 s)
    "zeros",        # silence
    "impulse",      # single unit sample at n // 3
)


def make_signal(kind: str, n: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    if kind == "gauss":
        x = 0.1 * rng.standard_normal(n)
    elif kind == "uniform":
        x = rng.uniform(-1.0, 1.0, n)
    elif kind == "sine440":
        x = 0.5 * np.sin(2 * np.pi * 440.0 * t)
    elif kind == "sine1k_noise":
        x = 0.8 * np.sin(2 * np.pi * 1000.0 * t) + 1e-5 * rng.standard_normal(n)
    elif kind == "chirp":
        dur = max(n, 1) / SAMPLE_RATE
        f0, f1 = 50.0, 7950.0
        x = 0.9 * np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t))
    elif kind == "two_tone":
        x = 0.5 * np.sin(2 * np.pi * 1000.0 * t) + 0.5 * 10 ** (-70 / 20) * np.sin(2 * np.pi * 3700.0 * t)
    elif kind == "pcm16":
        x = np.clip(np.round(3000.0 * rng.standard_normal(n)), -32768, 32767) / 32768.0
    elif kind == "burst":
        x = np.zeros(n)
        m = n // 6
        x[:m] = 0.1 * rng.standard_normal(m)
    elif kind == "zeros":
        x = np.zeros(n)
    elif kind == "impulse":
        x = np.zeros(n)
        x[n // 3] = 1.0
    else:
        raise ValueError(f"unknown signal kind {kind!r}")
    return x.astype(np.float32)


def make_pcm16(n: int, seed: int = 0) -> np.ndarray:
    """int16 samples whose ``/32768`` float image is ``make_signal('pcm16', n, seed)``."""
    rng = np.random.default_rng(seed)
    return np.clip(np.round(3000.0 * rng.standard_normal(n)), -32768, 32767).astype(np.int16)


def variable_lengths(count: int, seed: int = 4321, lo: int = 16000, hi: int = 480000) -> np.ndarray:
    """BASELINE config 4: clip lengths ~ U{lo..hi} samples (1-30 s)."""
    return np.random.default_rng(seed).integers(lo, hi + 1, size=count).astype(np.int64)


def digest(x: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
