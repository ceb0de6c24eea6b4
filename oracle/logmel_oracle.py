"""CPU restatement of the reference log-mel front-end (``whisper/audio.py``).

TEST INFRASTRUCTURE. Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker or the CPU arm — never as the product path.

Where the arithmetic lives: the reference has no kernel of its own; its
``log_mel_spectrogram`` (reference ``whisper/audio.py:110-157``) lowers to
PyTorch (``requirements.txt:3``: ``torch>=2.7.1``; installed here 2.11.0+cu128,
CPU FFT = MKL DFTI).  Two restatements are kept:

* :func:`logmel_f64` — the published semantics written out in float64 numpy
  (``torch.stft`` with ``center=True, pad_mode="reflect", onesided``; periodic
  Hann; |X|^2; mel projection; log10 clamp; max-8; (x+4)/4).  It is the
  high-precision spec.
* :func:`logmel_f32_port` — the same steps through the same fp32 PyTorch
  operators the reference calls (``audio.py:146-156``), i.e. a port with the
  reference's own cost profile.  It is the CPU baseline arm of ``bench.py``.

PARITY PIN: the reference ships no tests or golden vectors, so
both restatements are pinned against outputs of the reference itself, run in the
build container by ``tests/golden/make_golden.py`` and committed as
``tests/golden/logmel_golden.npz`` (``tests/test_oracle.py`` checks them).  The
filterbank used here is the reference asset's values as captured in that
fixture, not the product's regenerated one.
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_SAMPLES = 480000
N_FRAMES = 3000
N_BINS = N_FFT // 2 + 1

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "logmel_golden.npz")


@lru_cache(maxsize=None)
def reference_filters(n_mels: int) -> np.ndarray:
    """The reference asset's ``mel_{n_mels}`` array (audio.py:105-107), from the golden fixture."""
    if n_mels not in (80, 128):
        raise AssertionError(f"Unsupported n_mels: {n_mels}")  # audio.py:103
    with np.load(_GOLDEN, allow_pickle=False) as z:
        return np.array(z[f"filters_{n_mels}"], dtype=np.float32)


def n_frames_of(n_samples: int, padding: int = 0) -> int:
    """Frames kept by audio.py:148-149: stft yields 1 + L'//160, the last one is dropped."""
    total = n_samples + max(int(padding), 0)
    if total <= N_FFT // 2:
        # torch.stft's reflect pad needs pad < length (functional.py:675-680)
        raise RuntimeError(f"audio too short for reflect padding: {total} samples (need > {N_FFT // 2})")
    return total // HOP_LENGTH


def pad_or_trim_oracle(array: np.ndarray, length: int = N_SAMPLES, axis: int = -1) -> np.ndarray:
    """audio.py:65-88, numpy branch: keep the first ``length`` entries or right-pad with zeros."""
    n = array.shape[axis]
    if n > length:
        index = [slice(None)] * array.ndim
        index[axis] = slice(0, length)
        return array[tuple(index)].copy()
    if n < length:
        widths = [(0, 0)] * array.ndim
        widths[axis] = (0, length - n)
        return np.pad(array, widths)
    return array


def reflect_index(j: np.ndarray, n: int) -> np.ndarray:
    """Index map of ``F.pad(..., mode='reflect')`` as used by torch.stft(center=True)."""
    j = np.where(j < 0, -j, j)
    return np.where(j >= n, 2 * (n - 1) - j, j)


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    """``torch.hann_window(400)`` (audio.py:147): periodic Hann, float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def power_spectrum_f64(audio: np.ndarray, padding: int = 0) -> np.ndarray:
    """float64 ``[201, T]`` power spectrum of audio.py:146-149 for a 1-D waveform."""
    x = np.asarray(audio, dtype=np.float64)
    if padding > 0:
        x = np.concatenate([x, np.zeros(padding)])  # audio.py:145-146
    n = x.shape[0]
    frames = n_frames_of(n)
    pos = HOP_LENGTH * np.arange(frames)[:, None] + np.arange(N_FFT)[None, :] - N_FFT // 2
    framed = x[reflect_index(pos, n)] * hann_periodic()[None, :]
    spec = np.fft.rfft(framed, n=N_FFT, axis=1)  # [T, 201]
    return (spec.real**2 + spec.imag**2).T


def logmel_f64(audio: np.ndarray, n_mels: int = 80, padding: int = 0) -> np.ndarray:
    """High-precision spec of ``log_mel_spectrogram`` for ONE utterance; float64 ``[n_mels, T]``."""
    filters = reference_filters(n_mels).astype(np.float64)
    mel = filters @ power_spectrum_f64(audio, padding)  # audio.py:151-152
    log_spec = np.log10(np.maximum(mel, 1e-10))  # audio.py:154
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)  # audio.py:155
    return (log_spec + 4.0) / 4.0  # audio.py:156


def logmel_f32_port(audio, n_mels: int = 80, padding: int = 0):
    """fp32 PyTorch port: the operators audio.py:146-156 calls, in the same order.

    Accepts a 1-D or 2-D float32 tensor/ndarray.  Like the reference, a 2-D
    input shares ONE max over the whole call (audio.py:155).
    """
    import torch
    import torch.nn.functional as F

    x = torch.from_numpy(audio) if isinstance(audio, np.ndarray) else audio
    if padding > 0:
        x = F.pad(x, (0, padding))
    stft = torch.stft(x, N_FFT, HOP_LENGTH, window=torch.hann_window(N_FFT), return_complex=True)
    power = stft[..., :-1].abs() ** 2
    mel = torch.from_numpy(reference_filters(n_mels)) @ power
    log_spec = torch.clamp(mel, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0


def logmel_f32_port_per_utterance(batch, n_mels: int = 80, padding: int = 0):
    """The batch oracle: stack of per-clip calls, one max per utterance."""
    import torch

    return torch.stack([logmel_f32_port(row, n_mels, padding) for row in batch])
