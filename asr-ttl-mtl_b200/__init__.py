"""asr-ttl-mtl_b200 — B200-native log-mel front-end behind the ``whisper/audio.py`` API.

The directory name follows the project (``asr-ttl-mtl_b200``); a hyphen cannot be
imported, so ``asr_ttl_mtl_b200`` at the repo root is the import alias of this package.
"""
from .audio import (  # noqa: F401
    CHUNK_LENGTH,
    FRAMES_PER_SECOND,
    HOP_LENGTH,
    N_FFT,
    N_FRAMES,
    N_SAMPLES,
    N_SAMPLES_PER_TOKEN,
    SAMPLE_RATE,
    TOKENS_PER_SECOND,
    collate_log_mels,
    gpu_launches,
    load_audio,
    log_mel_spectrogram,
    log_mel_spectrogram_batch,
    mel_filters,
    mel_windows,
    pad_or_trim,
)
from .install import install, uninstall  # noqa: F401
from .sharding import shard_range  # noqa: F401
from .stem import encoder_stem, encoder_stem2, log_mel_encoder_stem, log_mel_encoder_stem2, pack_conv2_weight  # noqa: F401
