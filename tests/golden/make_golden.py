"""Generate tests/golden/logmel_golden.npz by RUNNING THE REFERENCE in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference (``/root/reference/whisper/audio.py``) is imported unmodified and
called on the seeded waveforms of ``oracle/signals.py``; its outputs, the two
filterbank arrays of its asset, and a few known answers for ``pad_or_trim`` and
the frame-count rule are stored.  The reference is Python and cannot travel to
the GPU box, so these fixtures are what pins the oracle (and through it the
CUDA path) to the reference's behaviour.  Nothing at test/bench time reads
/root/reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import signals  # noqa: E402
import whisper.audio as ref_audio  # noqa: E402  (the reference, unmodified)

SHORT = 16000

# (kind, n_samples, seed, n_mels, padding)
CASES = []
for i, kind in enumerate(signals.KINDS):
    for n_mels in (80, 128):
        CASES.append((kind, SHORT, 100 + i, n_mels, 0))
# ragged lengths and the padding argument (frame-count rule T = (L + padding) // 160)
for n, pad in ((201, 0), (399, 0), (400, 0), (16001, 0), (5000, 1234), (4321, 480), (31999, 0), (32160, 160)):
    CASES.append(("gauss", n, 7, 80, pad))
CASES.append(("chirp", 23456, 8, 128, 777))
# full 30 s clips (BASELINE configs 1-3 and the hard tonal case)
CASES.append(("gauss", 480000, 0, 80, 0))
CASES.append(("gauss", 480000, 0, 128, 0))
CASES.append(("chirp", 480000, 1, 80, 0))
CASES.append(("burst", 480000, 2, 80, 0))
# transcribe-style call: padding = N_SAMPLES (transcribe.py:139)
CASES.append(("sine1k_noise", 48000, 3, 80, 480000))


def main() -> None:
    torch.manual_seed(0)
    arrays = {}
    manifest = []
    for idx, (kind, n, seed, n_mels, pad) in enumerate(CASES):
        x = signals.make_signal(kind, n, seed)
        out = ref_audio.log_mel_spectrogram(x, n_mels=n_mels, padding=pad)
        assert out.dtype == torch.float32
        arrays[f"out_{idx}"] = out.numpy()
        manifest.append(
            dict(idx=idx, kind=kind, n=n, seed=seed, n_mels=n_mels, padding=pad,
                 input_sha256=signals.digest(x), shape=list(out.shape))
        )

    # the reference's literal 2-D behaviour: ONE max over the whole call (audio.py:155)
    batch = np.stack([signals.make_signal("gauss", SHORT, 50) * s for s in (1.0, 1e-4, 0.3)])
    arrays["batch2d_in_scale"] = np.array([1.0, 1e-4, 0.3], dtype=np.float32)
    arrays["batch2d_out"] = ref_audio.log_mel_spectrogram(batch, n_mels=80).numpy()

    # filterbank asset values (audio.py:105-107)
    for n_mels in (80, 128):
        arrays[f"filters_{n_mels}"] = ref_audio.mel_filters("cpu", n_mels).numpy()

    # pad_or_trim known answers (audio.py:65-88), numpy and torch branches
    arrays["pot_pad_np"] = ref_audio.pad_or_trim(np.arange(5, dtype=np.float32), 8)
    arrays["pot_trim_np"] = ref_audio.pad_or_trim(np.arange(10, dtype=np.float32), 4)
    m = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    arrays["pot_axis1_pad_t"] = ref_audio.pad_or_trim(m, 5, axis=1).numpy()
    arrays["pot_axis0_trim_t"] = ref_audio.pad_or_trim(m, 1, axis=0).numpy()
    arrays["pot_last_pad_t"] = ref_audio.pad_or_trim(m, 6).numpy()

    # frame counts straight from the reference for a sweep of lengths
    lengths = [201, 202, 319, 320, 321, 399, 400, 401, 479, 480, 16000, 16001, 479999, 480000]
    arrays["frames_len"] = np.array(lengths, dtype=np.int64)
    arrays["frames_T"] = np.array(
        [ref_audio.log_mel_spectrogram(np.zeros(n, np.float32)).shape[-1] for n in lengths], dtype=np.int64
    )
    consts = {k: int(getattr(ref_audio, k)) for k in (
        "SAMPLE_RATE", "N_FFT", "HOP_LENGTH", "CHUNK_LENGTH", "N_SAMPLES", "N_FRAMES",
        "N_SAMPLES_PER_TOKEN", "FRAMES_PER_SECOND", "TOKENS_PER_SECOND")}

    meta = dict(cases=manifest, constants=consts, torch=torch.__version__, numpy=np.__version__,
                reference="muhkemallgp/asr-ttl-mtl whisper/audio.py (whisper 20240930)")
    arrays["manifest_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "logmel_golden.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {len(CASES)} cases, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
