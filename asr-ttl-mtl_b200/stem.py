"""Encoder stem right behind the front-end: ``F.gelu(conv1(mel))`` of ``AudioEncoder.forward``
(reference whisper/model.py:179, :193) on the tcgen05 tensor cores, optionally fed straight from the
front-end's un-clamped output so that the clamp at ``max - 8`` (whisper/audio.py:155) happens on load.

``encoder_stem`` is the first layer alone; ``encoder_stem2`` / ``log_mel_encoder_stem2`` are the whole stem of
``AudioEncoder.forward`` (model.py:193-197): conv1 + GELU, conv2 (stride 2) + GELU, the move to ``[B, frames, n_state]``
and the positional embedding.  The transformer blocks stay the model's.
There is no CPU fallback: without the CUDA library (or a GPU) these functions raise.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native
from . import audio as _audio


def _check_stem_params(weight: torch.Tensor, bias: torch.Tensor, n_mels: int, device: torch.device):
    if n_mels != 80:
        raise ValueError(f"encoder stem kernel: n_mels must be 80, got {n_mels}")
    if weight.dim() != 3 or weight.shape[1] != n_mels or weight.shape[2] != 3:
        raise ValueError(f"conv1 weight must be [n_state, {n_mels}, 3], got {tuple(weight.shape)}")
    n_state = weight.shape[0]
    if n_state % 128 != 0:
        raise ValueError(f"encoder stem kernel: n_state must be a multiple of 128, got {n_state}")
    if bias.shape != (n_state,):
        raise ValueError(f"conv1 bias must be [{n_state}], got {tuple(bias.shape)}")
    weight = weight.detach().to(device=device, dtype=torch.float32).contiguous()
    bias = bias.detach().to(device=device, dtype=torch.float32).contiguous()
    return weight, bias, n_state


def encoder_stem(mel: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, *,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``F.gelu(F.conv1d(mel, weight, bias, padding=1))`` for a CUDA float32 ``mel`` of shape ``[B, 80, T]``
    (or ``[80, T]``): model.py:193.  TF32 operands with float32 accumulation (cudnn's default conv arithmetic on this GPU),
    exact GELU."""
    _audio._require_cuda()
    if not mel.is_cuda or mel.dtype != torch.float32:
        raise ValueError("encoder_stem: mel must be a CUDA float32 tensor")
    squeeze = mel.dim() == 2
    x = (mel.unsqueeze(0) if squeeze else mel).contiguous()
    if x.dim() != 3:
        raise ValueError(f"encoder_stem: expected [B, n_mels, T], got {tuple(mel.shape)}")
    batch, n_mels, n_frames = x.shape
    weight, bias, n_state = _check_stem_params(weight, bias, n_mels, x.device)
    lib = _native.load()
    with torch.cuda.device(x.device):
        shape = (batch, n_state, n_frames)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=x.device)
        elif out.shape != shape or out.dtype != torch.float32 or out.device != x.device or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 {shape} tensor on {x.device}")
        stream = torch.cuda.current_stream(x.device)
        _native.check(lib.b200mel_stem_conv1_gelu_device(
            x.data_ptr(), None, 0, batch, n_mels, n_frames, weight.data_ptr(), bias.data_ptr(), n_state,
            out.data_ptr(), stream.cuda_stream))
        for t in (x, weight, bias):
            t.record_stream(stream)
    return out[0] if squeeze else out


def log_mel_encoder_stem(audio: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, *, padding: int = 0,
                         lengths=None, global_max: bool = False) -> torch.Tensor:
    """``F.gelu(conv1(log_mel_spectrogram_batch(audio, 80, padding, lengths=lengths)))`` for a CUDA ``[B, L]`` waveform
    (float32 or int16 PCM) in two launches: the front-end leaves its output BEFORE the clamp (B200MEL_FLAG_DEFER_CLAMP)
    and the stem clamps while loading - the spectrogram is written once and read once, zero-padded tiles are neither
    written nor read.  ``global_max=True``: one max for the whole call (a 2-D ``log_mel_spectrogram`` call, audio.py:155)."""
    _audio._require_cuda()
    if not audio.is_cuda or audio.dim() != 2:
        raise ValueError("log_mel_encoder_stem: audio must be a CUDA [B, L] tensor")
    dtype = _audio._validate_waveform(audio, True)
    wave = audio.detach()
    if wave.stride(-1) != 1 or (wave.shape[0] > 1 and wave.stride(0) < wave.shape[1]):
        wave = wave.contiguous()
    batch, n_samples = wave.shape
    n_mels = 80
    n_frames = _audio._frames_or_raise(n_samples, int(padding))
    weight, bias, n_state = _check_stem_params(weight, bias, n_mels, wave.device)
    lib = _native.load()
    index = wave.device.index if wave.device.index is not None else torch.cuda.current_device()
    plan = _audio._plan(index, n_mels)
    with torch.cuda.device(index):
        len_ptr = None
        if lengths is not None:
            lengths = torch.as_tensor(lengths).to(device=wave.device, dtype=torch.int32).contiguous()
            if lengths.shape != (batch,):
                raise ValueError(f"lengths must have shape ({batch},)")
            len_ptr = lengths.data_ptr()
        mel = torch.empty((batch, n_mels, n_frames), dtype=torch.float32, device=wave.device)
        out = torch.empty((batch, n_state, n_frames), dtype=torch.float32, device=wave.device)
        workspace = torch.empty(lib.b200mel_workspace_bytes_tiles(batch, n_frames), dtype=torch.uint8, device=wave.device)
        stream = torch.cuda.current_stream(wave.device)
        flags = _native.FLAG_TILE_KEYS | (_native.FLAG_GLOBAL_MAX if global_max else 0)
        stride_b = wave.stride(0) if batch > 1 else n_samples
        _native.check(lib.b200mel_logmel_device(
            plan, wave.data_ptr(), dtype, batch, n_samples, stride_b, len_ptr, int(padding), mel.data_ptr(),
            workspace.data_ptr(), flags | _native.FLAG_DEFER_CLAMP, _native.VARIANT_TCGEN05, stream.cuda_stream))
        _native.check(lib.b200mel_stem_conv1_gelu_device(
            mel.data_ptr(), workspace.data_ptr(), flags, batch, n_mels, n_frames, weight.data_ptr(), bias.data_ptr(),
            n_state, out.data_ptr(), stream.cuda_stream))
        for t in (wave, mel, workspace, weight, bias):
            t.record_stream(stream)
    return out


def pack_conv2_weight(weight: torch.Tensor, device=None) -> torch.Tensor:
    """``conv2.weight`` ``[n_state, n_state, 3]`` (model.py:180) as the second layer's operand: IEEE half
    ``[3, n_state, n_state]`` (tap, out channel, in channel).  Do it once per model; ``encoder_stem2`` takes either form."""
    if weight.dim() != 3 or weight.shape[0] != weight.shape[1] or weight.shape[2] != 3:
        raise ValueError(f"conv2 weight must be [n_state, n_state, 3], got {tuple(weight.shape)}")
    w = weight.detach()
    if device is not None:
        w = w.to(device)
    return w.permute(2, 0, 1).to(torch.float16).contiguous()


def _check_conv2_params(weight, bias, pos, n_state, frames_out, device):
    if weight.dtype == torch.float16 and weight.dim() == 3 and weight.shape[0] == 3:
        packed = weight.detach().to(device).contiguous()
    else:
        packed = pack_conv2_weight(weight, device)
    if tuple(packed.shape) != (3, n_state, n_state):
        raise ValueError(f"conv2 weight must be [{n_state}, {n_state}, 3], got {tuple(weight.shape)}")
    if bias.shape != (n_state,):
        raise ValueError(f"conv2 bias must be [{n_state}], got {tuple(bias.shape)}")
    bias = bias.detach().to(device=device, dtype=torch.float32).contiguous()
    if pos is not None:
        # model.py:196: assert x.shape[1:] == self.positional_embedding.shape, "incorrect audio shape"
        assert tuple(pos.shape) == (frames_out, n_state), "incorrect audio shape"
        pos = pos.detach().to(device=device, dtype=torch.float32).contiguous()
    return packed, bias, pos


def _conv2(lib, h1, batch, frames_padded, packed, bias2, pos, n_state, stream, dtype=torch.float32):
    if dtype not in (torch.float32, torch.float16):
        raise ValueError(f"encoder stem: dtype must be torch.float32 or torch.float16, got {dtype}")
    out = torch.empty((batch, frames_padded // 2, n_state), dtype=dtype, device=h1.device)
    _native.check(lib.b200mel_stem_conv2_gelu_device(
        h1.data_ptr(), batch, frames_padded, packed.data_ptr(), bias2.data_ptr(), None if pos is None else pos.data_ptr(),
        n_state, out.data_ptr(), _native.FLAG_OUT_F16 if dtype == torch.float16 else 0, stream.cuda_stream))
    for t in (h1, packed, bias2) + (() if pos is None else (pos,)):
        t.record_stream(stream)
    return out


def _intermediate(batch, n_frames, n_state, device):
    """conv1's result as conv2 reads it: half [B, frames rounded up to even, n_state]; the extra frame is the padding."""
    frames_padded = n_frames + (n_frames & 1)
    h1 = torch.empty((batch, frames_padded, n_state), dtype=torch.float16, device=device)
    if frames_padded != n_frames:
        h1[:, n_frames:].zero_()
    return h1, frames_padded


def encoder_stem2(mel: torch.Tensor, conv1_weight: torch.Tensor, conv1_bias: torch.Tensor, conv2_weight: torch.Tensor,
                  conv2_bias: torch.Tensor, positional_embedding: Optional[torch.Tensor] = None, *,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """The stem of ``AudioEncoder.forward`` (model.py:193-197) for a CUDA float32 ``mel`` ``[B, 80, T]``::

        x = F.gelu(conv1(mel)); x = F.gelu(conv2(x)); x = x.permute(0, 2, 1); x = x + positional_embedding

    -> float32 ``[B, (T + 1) // 2, n_state]`` in two launches.  conv1 as ``encoder_stem`` (TF32 operands); its result goes
    to conv2 as IEEE half ``[B, T, n_state]`` (frames major: the layout the second GEMM's operand copies want) and conv2
    multiplies half operands with float32 accumulation - the 11-bit significand of TF32, i.e. of torch's own convolution.
    ``conv2_weight``: ``conv2.weight`` or ``pack_conv2_weight(conv2.weight)``; ``positional_embedding`` optional;
    ``dtype=torch.float16``: the float32 result rounded once, for a half-precision model (model.py:197 ``.to(x.dtype)``)."""
    _audio._require_cuda()
    if not mel.is_cuda or mel.dtype != torch.float32 or mel.dim() != 3:
        raise ValueError("encoder_stem2: mel must be a CUDA float32 [B, n_mels, T] tensor")
    x = mel.contiguous()
    batch, n_mels, n_frames = x.shape
    w1, b1, n_state = _check_stem_params(conv1_weight, conv1_bias, n_mels, x.device)
    packed, b2, pos = _check_conv2_params(conv2_weight, conv2_bias, positional_embedding, n_state, (n_frames + 1) // 2, x.device)
    lib = _native.load()
    with torch.cuda.device(x.device):
        h1, frames_padded = _intermediate(batch, n_frames, n_state, x.device)
        stream = torch.cuda.current_stream(x.device)
        _native.check(lib.b200mel_stem_conv1_gelu_fm16_device(
            x.data_ptr(), None, 0, batch, n_mels, n_frames, w1.data_ptr(), b1.data_ptr(), n_state, h1.data_ptr(),
            stream.cuda_stream))
        for t in (x, w1, b1):
            t.record_stream(stream)
        return _conv2(lib, h1, batch, frames_padded, packed, b2, pos, n_state, stream, dtype)


def log_mel_encoder_stem2(audio: torch.Tensor, conv1_weight: torch.Tensor, conv1_bias: torch.Tensor, conv2_weight: torch.Tensor,
                          conv2_bias: torch.Tensor, positional_embedding: Optional[torch.Tensor] = None, *, padding: int = 0,
                          lengths=None, global_max: bool = False, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``encoder_stem2(log_mel_spectrogram_batch(audio, 80, padding, lengths=lengths), ...)`` for a CUDA ``[B, L]`` waveform
    (float32 or int16 PCM) in three launches: front-end (un-clamped output, ``B200MEL_FLAG_DEFER_CLAMP``), conv1 (clamps on
    load), conv2 - waveform in, the transformer blocks' input out (dataset.py:80-96 -> model.py:193-197)."""
    _audio._require_cuda()
    if not audio.is_cuda or audio.dim() != 2:
        raise ValueError("log_mel_encoder_stem2: audio must be a CUDA [B, L] tensor")
    out_dtype = dtype
    dtype = _audio._validate_waveform(audio, True)
    wave = audio.detach()
    if wave.stride(-1) != 1 or (wave.shape[0] > 1 and wave.stride(0) < wave.shape[1]):
        wave = wave.contiguous()
    batch, n_samples = wave.shape
    n_mels = 80
    n_frames = _audio._frames_or_raise(n_samples, int(padding))
    w1, b1, n_state = _check_stem_params(conv1_weight, conv1_bias, n_mels, wave.device)
    packed, b2, pos = _check_conv2_params(conv2_weight, conv2_bias, positional_embedding, n_state, (n_frames + 1) // 2, wave.device)
    lib = _native.load()
    index = wave.device.index if wave.device.index is not None else torch.cuda.current_device()
    plan = _audio._plan(index, n_mels)
    with torch.cuda.device(index):
        len_ptr = None
        if lengths is not None:
            lengths = torch.as_tensor(lengths).to(device=wave.device, dtype=torch.int32).contiguous()
            if lengths.shape != (batch,):
                raise ValueError(f"lengths must have shape ({batch},)")
            len_ptr = lengths.data_ptr()
        mel = torch.empty((batch, n_mels, n_frames), dtype=torch.float32, device=wave.device)
        h1, frames_padded = _intermediate(batch, n_frames, n_state, wave.device)
        workspace = torch.empty(lib.b200mel_workspace_bytes_tiles(batch, n_frames), dtype=torch.uint8, device=wave.device)
        stream = torch.cuda.current_stream(wave.device)
        flags = _native.FLAG_TILE_KEYS | (_native.FLAG_GLOBAL_MAX if global_max else 0)
        stride_b = wave.stride(0) if batch > 1 else n_samples
        _native.check(lib.b200mel_logmel_device(
            plan, wave.data_ptr(), dtype, batch, n_samples, stride_b, len_ptr, int(padding), mel.data_ptr(),
            workspace.data_ptr(), flags | _native.FLAG_DEFER_CLAMP, _native.VARIANT_TCGEN05, stream.cuda_stream))
        _native.check(lib.b200mel_stem_conv1_gelu_fm16_device(
            mel.data_ptr(), workspace.data_ptr(), flags, batch, n_mels, n_frames, w1.data_ptr(), b1.data_ptr(),
            n_state, h1.data_ptr(), stream.cuda_stream))
        for t in (wave, mel, workspace, w1, b1):
            t.record_stream(stream)
        return _conv2(lib, h1, batch, frames_padded, packed, b2, pos, n_state, stream, out_dtype)
