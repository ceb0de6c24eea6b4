"""B200-native drop-in for the reference's ``whisper/audio.py``.

Same names, signatures, tensor layouts and error behaviour as the reference module
(``/root/reference/whisper/audio.py``), so ``speech_disorder/dataset.py:7,82-89``,
``whisper/transcribe.py:11-19,139,151,286`` and ``whisper/__init__.py:11`` consume
it unchanged (see ``install.py``).  The compute of ``log_mel_spectrogram``
(``audio.py:145-156``) is one call into the C ABI of ``include/b200mel.h``, which
launches the fused sm_100a kernels.  There is NO CPU fallback: CPU inputs are staged
through the GPU by ``b200mel_logmel_host`` and come back as CPU tensors, and the call
raises if no CUDA device or no built library is present.

Beyond the reference surface, :func:`log_mel_spectrogram_batch` is the batched,
per-utterance entry point the north star asks for (``[B, L] -> [B, n_mels, T]``,
optional ``lengths``, float32 or int16 PCM).
"""
from __future__ import annotations

import ctypes
import os
import threading
from functools import lru_cache
from subprocess import CalledProcessError, run
from typing import Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from . import _native
from .filterbank import slaney_mel_filterbank


def exact_div(x, y):
    """whisper/utils.py:24-26."""
    assert x % y == 0
    return x // y


# hard-coded audio hyperparameters (whisper/audio.py:13-22)
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000 samples in a 30-second chunk
N_FRAMES = exact_div(N_SAMPLES, HOP_LENGTH)  # 3000 frames in a mel spectrogram input

N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2  # the initial convolutions has stride 2
FRAMES_PER_SECOND = exact_div(SAMPLE_RATE, HOP_LENGTH)  # 10ms per audio frame
TOKENS_PER_SECOND = exact_div(SAMPLE_RATE, N_SAMPLES_PER_TOKEN)  # 20ms per audio token

VARIANTS = {"auto": _native.VARIANT_AUTO, "fft": _native.VARIANT_FFT, "tcgen05": _native.VARIANT_TCGEN05}


def load_audio(file: str, sr: int = SAMPLE_RATE):
    """Decode ``file`` to a mono float32 waveform at ``sr`` Hz through the ffmpeg CLI.

    Same contract as whisper/audio.py:25-62 (s16le pipe, ``/ 32768.0``,
    ``RuntimeError("Failed to load audio: ...")`` on decoder failure).  Decoding stays
    on the CPU; it is outside the accelerated path.
    """
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1",
           "-acodec", "pcm_s16le", "-ar", str(sr), "-"]
    try:
        out = run(cmd, capture_output=True, check=True).stdout
    except CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
    return np.frombuffer(out, np.int16).flatten().astype(np.float32) / 32768.0


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Keep the first ``length`` entries along ``axis`` or right-pad with zeros (whisper/audio.py:65-88).

    Torch in -> torch out (same device and dtype), numpy in -> numpy out; an array that
    already has ``length`` entries is returned as is.
    """
    size = array.shape[axis]
    if torch.is_tensor(array):
        if size > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        elif size < length:
            widths = [0, 0] * array.ndim
            widths[2 * (array.ndim - 1 - (axis % array.ndim)) + 1] = length - size
            array = F.pad(array, widths)
    else:
        if size > length:
            array = array.take(indices=range(length), axis=axis)
        elif size < length:
            widths = [(0, 0)] * array.ndim
            widths[axis] = (0, length - size)
            array = np.pad(array, widths)
    return array


@lru_cache(maxsize=None)
def mel_filters(device, n_mels: int) -> torch.Tensor:
    """The ``[n_mels, 201]`` float32 mel filterbank on ``device`` (whisper/audio.py:91-107).

    The reference loads a librosa-generated asset; here the same Slaney bank is
    regenerated (bit-equal, see filterbank.py) so no reference file is shipped.
    """
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    return torch.from_numpy(slaney_mel_filterbank(n_mels)).to(device)


# ---------------------------------------------------------------------------------------------
# plans: one per (CUDA device index, n_mels); they own the kernels' constant tables on that GPU
# ---------------------------------------------------------------------------------------------
_plans = {}
_plans_lock = threading.Lock()


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "asr-ttl-mtl_b200: no CUDA device is visible; the B200 log-mel front-end has no CPU fallback"
        )


def _plan(device_index: int, n_mels: int) -> ctypes.c_void_p:
    key = (device_index, n_mels)
    plan = _plans.get(key)
    if plan is not None:
        return plan
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    lib = _native.load()
    with _plans_lock:
        plan = _plans.get(key)
        if plan is None:
            bank = np.ascontiguousarray(slaney_mel_filterbank(n_mels), dtype=np.float32)
            handle = ctypes.c_void_p()
            with torch.cuda.device(device_index):
                _native.check(lib.b200mel_plan_create(
                    n_mels, bank.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(handle)))
            plan = _plans[key] = handle
    return plan


def _validate_waveform(audio: torch.Tensor, allow_pcm16: bool) -> int:
    if audio.dim() not in (1, 2):
        raise RuntimeError(f"log_mel_spectrogram: expected a 1D or 2D waveform tensor, got {audio.dim()}D")
    if audio.dtype == torch.float32:
        return _native.DTYPE_F32
    if allow_pcm16 and audio.dtype == torch.int16:
        return _native.DTYPE_S16
    raise RuntimeError(
        f"log_mel_spectrogram: expected a float32 waveform"
        f"{' (or int16 PCM)' if allow_pcm16 else ''}, got {audio.dtype}"
    )


# what "auto" means in this process: the library's own choice unless B200MEL_VARIANT names a kernel (tests flip it)
DEFAULT_VARIANT = os.environ.get("B200MEL_VARIANT", "auto")


def _frames_or_raise(n_samples: int, padding: int) -> int:
    try:
        return _native.frames(n_samples, padding)
    except _native.B200MelError as e:
        if e.status == _native.ERR_TOO_SHORT:
            # torch.stft's reflect pad raises RuntimeError for n_fft // 2 >= length
            raise RuntimeError(
                f"log_mel_spectrogram: padding ({N_FFT // 2}, {N_FFT // 2}) needs more than {N_FFT // 2} "
                f"samples, got {n_samples + max(padding, 0)}"
            ) from None
        raise


def _run(audio: torch.Tensor, n_mels: int, padding: int, lengths, flags: int, variant: str,
         out: Optional[torch.Tensor], allow_pcm16: bool,
         out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``audio`` 1-D/2-D on CPU or CUDA -> log-mel on the same device, squeezed like the input."""
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    dtype = _validate_waveform(audio, allow_pcm16)
    _require_cuda()
    lib = _native.load()
    squeeze = audio.dim() == 1
    wave = audio.detach()
    wave = wave.unsqueeze(0) if squeeze else wave
    if wave.stride(-1) != 1 or (wave.shape[0] > 1 and wave.stride(0) < wave.shape[1]):
        wave = wave.contiguous()
    batch, n_samples = wave.shape
    padding = int(padding)
    n_frames = _frames_or_raise(n_samples, padding)
    stride_b = wave.stride(0) if batch > 1 else n_samples
    shape = (batch, n_mels, n_frames)
    if variant == "auto":
        variant = DEFAULT_VARIANT
    try:
        variant_id = VARIANTS[variant]
    except KeyError:
        raise ValueError(f"unknown variant {variant!r}; expected one of {sorted(VARIANTS)}") from None
    if out_dtype not in (torch.float32, torch.float16):
        raise ValueError(f"out_dtype must be torch.float32 or torch.float16, got {out_dtype}")
    if out_dtype == torch.float16:
        flags |= _native.FLAG_OUT_F16

    if wave.is_cuda:
        index = wave.device.index if wave.device.index is not None else torch.cuda.current_device()
        plan = _plan(index, n_mels)
        with torch.cuda.device(index):
            if out is None:
                out = torch.empty(shape, dtype=out_dtype, device=wave.device)
            elif out.shape != shape or out.dtype != out_dtype or out.device != wave.device or not out.is_contiguous():
                raise ValueError(f"out must be a contiguous {out_dtype} {shape} tensor on {wave.device}")
            len_ptr = None
            if lengths is not None:
                lengths = torch.as_tensor(lengths).to(device=wave.device, dtype=torch.int32).contiguous()
                if lengths.shape != (batch,):
                    raise ValueError(f"lengths must have shape ({batch},)")
                len_ptr = lengths.data_ptr()
            workspace = torch.empty(lib.b200mel_workspace_bytes_tiles(batch, n_frames), dtype=torch.uint8, device=wave.device)
            stream = torch.cuda.current_stream(wave.device)
            _native.check(lib.b200mel_logmel_device(
                plan, wave.data_ptr(), dtype, batch, n_samples, stride_b, len_ptr, padding, out.data_ptr(),
                workspace.data_ptr(), flags | _native.FLAG_TILE_KEYS, variant_id, stream.cuda_stream))
            # the caching allocator may hand these blocks to another stream once we return
            workspace.record_stream(stream)
            wave.record_stream(stream)
    else:
        index = torch.cuda.current_device()
        plan = _plan(index, n_mels)
        if out is None:
            out = torch.empty(shape, dtype=out_dtype)
        elif out.shape != shape or out.dtype != out_dtype or out.is_cuda or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous {out_dtype} {shape} CPU tensor")
        len_ptr = None
        if lengths is not None:
            lengths = torch.as_tensor(lengths).to(device="cpu", dtype=torch.int32).contiguous()
            if lengths.shape != (batch,):
                raise ValueError(f"lengths must have shape ({batch},)")
            len_ptr = lengths.data_ptr()
        _native.check(lib.b200mel_logmel_host(
            plan, wave.data_ptr(), dtype, batch, n_samples, stride_b, len_ptr, padding, out.data_ptr(),
            flags, variant_id))
    return out[0] if squeeze else out


def log_mel_spectrogram(
    audio: Union[str, np.ndarray, torch.Tensor],
    n_mels: int = 80,
    padding: int = 0,
    device: Optional[Union[str, torch.device]] = None,
):
    """Log-mel spectrogram of a 16 kHz waveform — the signature of whisper/audio.py:110-157.

    Parameters and return value are the reference's: ``audio`` is a path, a float32
    NumPy array or a float32 tensor of shape ``(L,)`` (or ``(B, L)``); ``padding`` zeros are
    appended on the right; ``device`` moves the waveform first.  Returns a new float32
    tensor ``(n_mels, (L + padding) // 160)`` on the waveform's device.

    As in the reference, a 2-D input shares ONE dynamic-range max over the whole call
    (``log_spec.max()``, audio.py:155); use :func:`log_mel_spectrogram_batch` for one max
    per utterance.  The compute always runs on the GPU, also for CPU inputs.
    """
    if not torch.is_tensor(audio):
        if isinstance(audio, str):
            audio = load_audio(audio)
        audio = torch.from_numpy(audio)
    if device is not None:
        audio = audio.to(device)
    flags = _native.FLAG_GLOBAL_MAX if audio.dim() == 2 else 0
    return _run(audio, n_mels, padding, None, flags, "auto", None, allow_pcm16=False)


def log_mel_spectrogram_batch(
    audio: Union[np.ndarray, torch.Tensor],
    n_mels: int = 80,
    padding: int = 0,
    device: Optional[Union[str, torch.device]] = None,
    *,
    lengths=None,
    out: Optional[torch.Tensor] = None,
    variant: str = "auto",
    out_dtype: torch.dtype = torch.float32,
):
    """Batched front-end with PER-UTTERANCE normalisation: ``[B, L] -> [B, n_mels, T]``.

    Equals ``torch.stack([log_mel_spectrogram(x, n_mels, padding) for x in audio])`` — what
    ``MultiTaskSpeechDataset`` + ``collate_fn`` build one clip at a time
    (speech_disorder/dataset.py:82-89,179) — in a single launch sequence.

    ``audio`` is float32 (any finite range) or int16 PCM (scaled by 1/32768 in-kernel, the
    arithmetic of audio.py:62).  ``lengths`` (optional, ``[B]``) gives the real samples of
    each row; the rest of the row counts as zeros whatever it holds, i.e. the rows
    behave as ``pad_or_trim``-med clips (audio.py:83-86).  ``out_dtype=torch.float16`` stores the float32 result
    rounded to half (what the fp16 model gets after transcribe.py:286's ``.to(dtype)``), halving the bytes written.
    """
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(audio)
    if device is not None:
        audio = audio.to(device)
    if audio.dim() != 2:
        raise RuntimeError(f"log_mel_spectrogram_batch: expected a 2D [B, L] waveform tensor, got {audio.dim()}D")
    return _run(audio, n_mels, padding, lengths, 0, variant, out, allow_pcm16=True, out_dtype=out_dtype)


def collate_log_mels(
    waveforms,
    n_mels: int = 80,
    device: Optional[Union[str, torch.device]] = None,
    *,
    length: int = N_SAMPLES,
    variant: str = "auto",
) -> torch.Tensor:
    """``batch['mels']`` for a list of un-padded waveforms, in one front-end call on the GPU.

    What ``MultiTaskSpeechDataset.load_and_process_audio`` + ``collate_fn`` build one clip at a time on the CPU
    (speech_disorder/dataset.py:82-89,179) — ``torch.stack([log_mel_spectrogram(pad_or_trim(w)) for w in waveforms])``,
    shape ``[B, n_mels, length // 160]`` — built here from the raw clips: each is trimmed to ``length`` samples, only
    its real samples cross PCIe, and the zero tail of ``pad_or_trim`` (audio.py:83-86) is never materialised: the
    kernel gets the clip lengths and treats the rest of each row as zeros.  Float32 or int16 PCM clips (all the same kind).
    A clip that cannot be processed raises — the reference's silent ``zeros((80, 3000))`` fallback is not reproduced.
    """
    _require_cuda()
    if len(waveforms) == 0:
        raise ValueError("collate_log_mels: empty batch")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("collate_log_mels runs on a CUDA device (this front-end has no CPU fallback)")
    clips = [w if torch.is_tensor(w) else torch.from_numpy(np.ascontiguousarray(w)) for w in waveforms]
    dtype = clips[0].dtype
    if dtype not in (torch.float32, torch.int16) or any(c.dtype != dtype or c.dim() != 1 for c in clips):
        raise RuntimeError("collate_log_mels: expected 1-D float32 (or all int16 PCM) waveforms")
    lens = torch.tensor([min(int(c.shape[0]), length) for c in clips], dtype=torch.int32)
    staged = torch.empty((len(clips), length), dtype=dtype, device=dev)   # rows past `lens` stay unwritten and unread
    for i, c in enumerate(clips):
        n = int(lens[i])
        if n > 0:
            staged[i, :n].copy_(c[:n], non_blocking=True)
    return log_mel_spectrogram_batch(staged, n_mels=n_mels, lengths=lens.to(dev, non_blocking=True), variant=variant)


def mel_windows(
    mel: torch.Tensor,
    seeks,
    sizes=None,
    *,
    window_frames: int = N_FRAMES,
    dtype: torch.dtype = torch.float16,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Cut decoding windows out of one utterance's log-mel spectrogram in ONE launch.

    Equals ``torch.stack([pad_or_trim(mel[:, s : s + n], window_frames).to(dtype) for s, n in zip(seeks, sizes)])`` - the
    ``mel_segment`` of whisper/transcribe.py:282-286 (and :150 with ``seeks=[0]``) for many windows at once, written
    straight into zero-padded ``[n_windows, n_mels, window_frames]`` windows in float32 or float16.  ``mel`` is the
    ``[n_mels, T]`` float32 CUDA tensor ``log_mel_spectrogram(audio, padding=N_SAMPLES)`` returned; ``seeks`` (and the
    optional ``sizes``, default: whole windows) are sequences or int32 tensors of window starts / kept frames.
    """
    _require_cuda()
    if not torch.is_tensor(mel) or mel.dim() != 2 or mel.dtype != torch.float32 or not mel.is_cuda:
        raise RuntimeError("mel_windows: expected a 2D float32 CUDA log-mel spectrogram [n_mels, T]")
    if dtype not in (torch.float32, torch.float16):
        raise ValueError(f"dtype must be torch.float32 or torch.float16, got {dtype}")
    mel = mel.detach().contiguous()
    n_mels, n_frames = mel.shape
    with torch.cuda.device(mel.device):
        seeks_t = torch.as_tensor(seeks).to(device=mel.device, dtype=torch.int32).contiguous().reshape(-1)
        n_windows = int(seeks_t.shape[0])
        sizes_t = None
        if sizes is not None:
            sizes_t = torch.as_tensor(sizes).to(device=mel.device, dtype=torch.int32).contiguous().reshape(-1)
            if sizes_t.shape != seeks_t.shape:
                raise ValueError("sizes must have as many entries as seeks")
        shape = (n_windows, n_mels, int(window_frames))
        if out is None:
            out = torch.empty(shape, dtype=dtype, device=mel.device)
        elif out.shape != shape or out.dtype != dtype or out.device != mel.device or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous {dtype} {shape} tensor on {mel.device}")
        stream = torch.cuda.current_stream(mel.device)
        _native.check(_native.load().b200mel_mel_windows_device(
            mel.data_ptr(), n_mels, n_frames, seeks_t.data_ptr(), sizes_t.data_ptr() if sizes_t is not None else None,
            n_windows, int(window_frames), out.data_ptr(), _native.FLAG_OUT_F16 if dtype == torch.float16 else 0, stream.cuda_stream))
        for t in (mel, seeks_t, sizes_t):
            if t is not None:
                t.record_stream(stream)
    return out


def gpu_launches() -> int:
    """Kernels launched by libb200mel.so in this process (bench.py's ``gpu_launches``)."""
    return _native.launch_count()
