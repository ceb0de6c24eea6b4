"""Encoder stem conv1 + GELU (reference whisper/model.py:179, :193) on the tensor cores, alone and fused behind the
front-end with the clamp applied on load.  `-m gpu`.

Oracle: `F.gelu(F.conv1d(x, w, b, padding=1))` in float64 on the CPU.  Tolerance: the kernel reads its float32 operands as
TF32 (10 mantissa bits, fp32 accumulation) - the arithmetic torch's own convolution uses on this GPU when
`torch.backends.cudnn.allow_tf32` is on (torch's default) - so the bar is 3e-3 absolute on outputs of magnitude ~1 with
conv weights of Whisper's scale, and never worse than twice torch's own TF32 convolution on the same input.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import signals

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 3e-3


def _params(n_state, seed, n_mels=80):
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / np.sqrt(n_mels * 3)          # torch's Conv1d initialisation
    w = (torch.rand(n_state, n_mels, 3, generator=g) * 2 - 1) * bound
    b = (torch.rand(n_state, generator=g) * 2 - 1) * bound
    return w, b


def _truth(x, w, b):
    return F.gelu(F.conv1d(x.double().cpu(), w.double(), b.double(), padding=1))


def _err(got, want):
    return float((got.double().cpu() - want).abs().max())


@pytest.mark.parametrize("n_state", [384, 512, 1280])
@pytest.mark.parametrize("batch,n_frames", [(1, 3000), (3, 1000), (2, 128), (2, 131), (1, 5), (5, 257)])
def test_stem_matches_float64_conv_gelu(b200, n_state, batch, n_frames):
    g = torch.Generator().manual_seed(n_frames)
    x = torch.rand(batch, 80, n_frames, generator=g) * 2.5 - 1.0        # the range of normalised log-mels
    w, b = _params(n_state, 7)
    got = b200.encoder_stem(x.to(DEV), w.to(DEV), b.to(DEV))
    assert tuple(got.shape) == (batch, n_state, n_frames) and got.dtype == torch.float32
    want = _truth(x, w, b)
    err = _err(got, want)
    saved = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        torch_tf32 = _err(F.gelu(F.conv1d(x.to(DEV), w.to(DEV), b.to(DEV), padding=1)), want)
    finally:
        torch.backends.cudnn.allow_tf32 = saved
    print(f"n_state {n_state} [{batch}, 80, {n_frames}]: |ours - f64| {err:.2e}, |torch tf32 - f64| {torch_tf32:.2e}")
    assert err <= TOL
    assert err <= max(2 * torch_tf32, 1e-3)


def test_stem_single_utterance_and_errors(b200):
    w, b = _params(384, 3)
    x = torch.rand(80, 300) - 0.5
    got = b200.encoder_stem(x.to(DEV), w.to(DEV), b.to(DEV))
    assert tuple(got.shape) == (384, 300)
    assert _err(got, _truth(x[None], w, b)[0]) <= TOL
    with pytest.raises(ValueError):
        b200.encoder_stem(torch.rand(128, 300, device=DEV), w.to(DEV), b.to(DEV))          # 128 mels: not built
    with pytest.raises(ValueError):
        b200.encoder_stem(x.to(DEV), w[:100].to(DEV), b[:100].to(DEV))                     # n_state % 128
    with pytest.raises(ValueError):
        b200.encoder_stem(x, w, b)                                                         # CPU tensor: no fallback


def test_stem_propagates_nan_like_torch(b200):
    w, b = _params(384, 4)
    x = torch.rand(1, 80, 400) - 0.5
    x[0, 17, 200] = float("nan")
    got = b200.encoder_stem(x.to(DEV), w.to(DEV), b.to(DEV)).cpu()
    want = _truth(x, w, b)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert int(torch.isnan(got).sum()) == 384 * 3


def _fused_case(b200, wave, lengths, n_state, global_max=False, padding=0):
    w, b = _params(n_state, 11)
    dev_wave = wave.to(DEV)
    got = b200.log_mel_encoder_stem(dev_wave, w.to(DEV), b.to(DEV), lengths=lengths, global_max=global_max, padding=padding)
    if global_max:
        assert lengths is None
        mel = b200.log_mel_spectrogram(dev_wave, padding=padding)                          # 2-D call: one max (audio.py:155)
    else:
        mel = b200.log_mel_spectrogram_batch(dev_wave, padding=padding, lengths=lengths)
    want = _truth(mel, w, b)
    two_step = b200.encoder_stem(mel, w.to(DEV), b.to(DEV))
    # the fused path feeds the stem the same float32 values the finish kernel would have written: bit-identical
    assert torch.equal(got, two_step)
    assert _err(got, want) <= TOL
    return got


def test_fused_front_end_and_stem_noise_and_speech_like(b200):
    kinds = list(signals.KINDS)
    wave = torch.from_numpy(np.stack([signals.make_signal(k, 48000, 40 + i) for i, k in enumerate(kinds)]))
    _fused_case(b200, wave, None, 384)
    _fused_case(b200, wave, None, 512, global_max=True)
    _fused_case(b200, wave, None, 384, padding=4000)


def test_fused_with_zero_padded_clips_and_clamp(b200):
    rng = np.random.default_rng(5)
    n = 160 * 128 * 6
    wave = np.zeros((6, n), np.float32)
    lengths = np.array([n, n // 2, 160 * 128 + 77, 0, 333, n - 1], np.int32)
    for i, L in enumerate(lengths):
        wave[i, :L] = rng.standard_normal(L).astype(np.float32) * 0.1
    # a clip whose second half is 100 dB below its first: the clamp at max - 8 is what the stem sees there
    wave[0, n // 2:] *= 1e-5
    # junk behind `lengths` must not be read
    for i, L in enumerate(lengths):
        wave[i, L:] = 7.0
    _fused_case(b200, torch.from_numpy(wave), torch.from_numpy(lengths), 384)
    wave_zero = wave.copy()
    for i, L in enumerate(lengths):
        wave_zero[i, L:] = 0.0
    _fused_case(b200, torch.from_numpy(wave_zero), None, 384)              # real zeros in memory: silent tiles found by the kernel
    _fused_case(b200, torch.from_numpy(wave_zero), None, 384, global_max=True)


def test_fused_int16_pcm(b200):
    rng = np.random.default_rng(9)
    pcm = (rng.standard_normal((3, 32000)) * 3000).astype(np.int16)
    _fused_case(b200, torch.from_numpy(pcm), None, 384)
