"""Slaney mel filterbank, generated instead of shipped.

The reference loads its filterbank from ``whisper/assets/mel_filters.npz``
(reference ``whisper/audio.py:91-107``), which its docstring says was written by
``librosa.filters.mel(sr=16000, n_fft=400, n_mels=80|128)`` (``audio.py:95-100``).
librosa is not a dependency of this repo and the asset is not copied: the
published Slaney construction below reproduces both arrays bit for bit
(float32), which ``tests/test_filterbank.py`` pins with SHA-256 digests taken
from the reference asset.

All arithmetic is float64 until the final ramp is stored into a float32 array,
which is what gives bit-equality.
"""
from __future__ import annotations

import hashlib

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
N_BINS = N_FFT // 2 + 1  # 201

#: sha256 of ``ndarray.tobytes()`` (float32, C order) of the reference asset's
#: ``mel_80`` / ``mel_128`` arrays; recorded by tests/golden/make_golden.py.
FILTER_SHA256 = {
    80: "4f2701b1d287d74a0dc9871026e9519d98cb76426615f2539b0d151a0ae4ec2e",
    128: "2a5f9822897750e047c85dea37cc268d3be0ecfd23c28a5f10da129d99d05afe",
}

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def _hz_to_mel(hz: float) -> float:
    if hz >= _MIN_LOG_HZ:
        return _MIN_LOG_MEL + np.log(hz / _MIN_LOG_HZ) / _LOGSTEP
    return hz / _F_SP


def _mel_to_hz(mels: np.ndarray) -> np.ndarray:
    mels = np.asarray(mels, dtype=np.float64)
    hz = _F_SP * mels
    log_region = mels >= _MIN_LOG_MEL
    hz[log_region] = _MIN_LOG_HZ * np.exp(_LOGSTEP * (mels[log_region] - _MIN_LOG_MEL))
    return hz


def slaney_mel_filterbank(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """float32 ``[n_mels, n_fft//2+1]`` area-normalised triangular filterbank."""
    n_bins = n_fft // 2 + 1
    bin_hz = np.linspace(0.0, sr / 2.0, n_bins, endpoint=True)
    edges_hz = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    widths = np.diff(edges_hz)
    ramps = np.subtract.outer(edges_hz, bin_hz)

    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    for m in range(n_mels):
        rising = -ramps[m] / widths[m]
        falling = ramps[m + 2] / widths[m + 1]
        weights[m] = np.maximum(0.0, np.minimum(rising, falling))
    area_norm = 2.0 / (edges_hz[2 : n_mels + 2] - edges_hz[:n_mels])
    weights *= area_norm[:, np.newaxis]
    return weights


def filter_digest(weights: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(weights, dtype=np.float32).tobytes()).hexdigest()
