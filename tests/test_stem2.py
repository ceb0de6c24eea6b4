"""The whole encoder stem (reference whisper/model.py:193-197): conv1 + GELU, conv2 (stride 2) + GELU, the permute and the
positional embedding, on the tensor cores.  `-m gpu`.

Oracle: the reference's own operator sequence in float64 on the CPU,
    x = F.gelu(conv1(mel)); x = F.gelu(conv2(x)); x = x.permute(0, 2, 1); x = x + positional_embedding.
Tolerance: both layers read their operands with an 11-bit significand (conv1 as TF32, conv2 as IEEE half, float32
accumulation) - the arithmetic torch's own convolutions use on this GPU with `allow_tf32` (cudnn's default) and the
precision of the reference's fp16 inference (transcribe.py:127).  The bar: 4e-3 absolute on outputs of magnitude ~1
with weights of torch's Conv1d initialisation, and never worse than twice torch's own TF32 pair of convolutions.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import signals

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 4e-3


def _params(n_state, seed, n_mels=80):
    g = torch.Generator().manual_seed(seed)
    k1, k2 = 1.0 / np.sqrt(n_mels * 3), 1.0 / np.sqrt(n_state * 3)          # torch's Conv1d initialisation
    w1 = (torch.rand(n_state, n_mels, 3, generator=g) * 2 - 1) * k1
    b1 = (torch.rand(n_state, generator=g) * 2 - 1) * k1
    w2 = (torch.rand(n_state, n_state, 3, generator=g) * 2 - 1) * k2
    b2 = (torch.rand(n_state, generator=g) * 2 - 1) * k2
    return w1, b1, w2, b2


def _sinusoids(length, channels, max_timescale=10000):
    """model.py:62-68 (the positional embedding's values)."""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2))
    t = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1).float()


def _truth(mel, w1, b1, w2, b2, pos=None):
    x = F.gelu(F.conv1d(mel.double().cpu(), w1.double(), b1.double(), padding=1))
    x = F.gelu(F.conv1d(x, w2.double(), b2.double(), stride=2, padding=1)).permute(0, 2, 1)
    return x if pos is None else x + pos.double()


def _torch_tf32(mel, w1, b1, w2, b2):
    saved = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    try:
        x = F.gelu(F.conv1d(mel.to(DEV), w1.to(DEV), b1.to(DEV), padding=1))
        return F.gelu(F.conv1d(x, w2.to(DEV), b2.to(DEV), stride=2, padding=1)).permute(0, 2, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = saved


def _err(got, want):
    return float((got.double().cpu() - want).abs().max())


@pytest.mark.parametrize("n_state", [384, 512, 1280])
@pytest.mark.parametrize("batch,n_frames", [(1, 3000), (3, 1000), (2, 512), (2, 131), (1, 5), (5, 257), (2, 514)])
def test_stem2_matches_float64(b200, n_state, batch, n_frames):
    g = torch.Generator().manual_seed(n_frames)
    mel = torch.rand(batch, 80, n_frames, generator=g) * 2.5 - 1.0          # the range of normalised log-mels
    p = _params(n_state, 7)
    got = b200.encoder_stem2(mel.to(DEV), *[t.to(DEV) for t in p])
    assert tuple(got.shape) == (batch, (n_frames + 1) // 2, n_state) and got.dtype == torch.float32 and got.is_contiguous()
    want = _truth(mel, *p)
    err, ref = _err(got, want), _err(_torch_tf32(mel, *p), want)
    print(f"n_state {n_state} [{batch}, 80, {n_frames}]: |ours - f64| {err:.2e}, |torch tf32 - f64| {ref:.2e}")
    assert err <= TOL
    assert err <= max(2 * ref, 1.5e-3)


def test_stem2_positional_embedding_and_packed_weight(b200):
    n_state, n_frames = 384, 3000
    mel = torch.rand(2, 80, n_frames, generator=torch.Generator().manual_seed(1)) * 2.5 - 1.0
    w1, b1, w2, b2 = _params(n_state, 9)
    pos = _sinusoids(n_frames // 2, n_state)
    packed = b200.pack_conv2_weight(w2, DEV)
    assert tuple(packed.shape) == (3, n_state, n_state) and packed.dtype == torch.float16
    got = b200.encoder_stem2(mel.to(DEV), w1.to(DEV), b1.to(DEV), packed, b2.to(DEV), pos.to(DEV))
    assert _err(got, _truth(mel, w1, b1, w2, b2, pos)) <= TOL
    plain = b200.encoder_stem2(mel.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV))
    assert torch.equal(got, plain + pos.to(DEV))                              # one float32 addition behind the GELU
    with pytest.raises(AssertionError, match="incorrect audio shape"):        # model.py:196
        b200.encoder_stem2(mel.to(DEV)[:, :, :1000], w1.to(DEV), b1.to(DEV), packed, b2.to(DEV), pos.to(DEV))
    with pytest.raises(ValueError):
        b200.encoder_stem2(mel, w1, b1, w2, b2)                               # CPU tensor: no fallback
    with pytest.raises(ValueError):
        b200.encoder_stem2(mel.to(DEV), w1.to(DEV), b1.to(DEV), w2[:, :100].to(DEV), b2.to(DEV))


def test_stem2_first_layer_is_the_stem_kernel_in_half(b200):
    """The intermediate is encoder_stem's result rounded to half: conv2 of that, in float64, is what comes out."""
    n_state = 384
    mel = torch.rand(2, 80, 700, generator=torch.Generator().manual_seed(2)) * 2.5 - 1.0
    w1, b1, w2, b2 = _params(n_state, 5)
    h1 = b200.encoder_stem(mel.to(DEV), w1.to(DEV), b1.to(DEV)).half().double().cpu()
    want = F.gelu(F.conv1d(h1, w2.half().double(), b2.double(), stride=2, padding=1)).permute(0, 2, 1)
    got = b200.encoder_stem2(mel.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV))
    assert _err(got, want) <= 2e-5                                            # float32 accumulation order and the GELU fit only


def test_stem2_propagates_nan_like_torch(b200):
    w1, b1, w2, b2 = _params(384, 4)
    mel = torch.rand(1, 80, 400) - 0.5
    mel[0, 17, 200] = float("nan")
    got = b200.encoder_stem2(mel.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV)).cpu()
    want = _truth(mel, w1, b1, w2, b2)
    assert torch.equal(torch.isnan(got), torch.isnan(want))


def test_fused_front_end_and_stem2(b200):
    kinds = list(signals.KINDS)
    wave = torch.from_numpy(np.stack([signals.make_signal(k, 48000, 40 + i) for i, k in enumerate(kinds)]))
    p = [t.to(DEV) for t in _params(384, 11)]
    for kwargs in ({}, {"padding": 4000}, {"global_max": True}):
        got = b200.log_mel_encoder_stem2(wave.to(DEV), *p, **kwargs)
        if kwargs.get("global_max"):
            mel = b200.log_mel_spectrogram(wave.to(DEV))
        else:
            mel = b200.log_mel_spectrogram_batch(wave.to(DEV), padding=kwargs.get("padding", 0))
        assert torch.equal(got, b200.encoder_stem2(mel, *p))                  # the clamp on load: the same float32 inputs
        assert _err(got, _truth(mel, *[t.cpu() for t in p])) <= TOL


def test_fused_stem2_zero_padded_clips_lengths_and_pcm(b200):
    rng = np.random.default_rng(5)
    n = 160 * 128 * 6
    wave = np.zeros((5, n), np.float32)
    lengths = np.array([n, n // 2, 160 * 128 + 77, 0, n - 1], np.int32)
    for i, L in enumerate(lengths):
        wave[i, :L] = rng.standard_normal(L).astype(np.float32) * 0.1
    p = [t.to(DEV) for t in _params(384, 12)]
    got = b200.log_mel_encoder_stem2(torch.from_numpy(wave).to(DEV), *p, lengths=torch.from_numpy(lengths))
    mel = b200.log_mel_spectrogram_batch(torch.from_numpy(wave).to(DEV), lengths=torch.from_numpy(lengths))
    assert torch.equal(got, b200.encoder_stem2(mel, *p))
    assert _err(got, _truth(mel, *[t.cpu() for t in p])) <= TOL
    pcm = (rng.standard_normal((3, 32000)) * 3000).astype(np.int16)
    got = b200.log_mel_encoder_stem2(torch.from_numpy(pcm).to(DEV), *p)
    mel = b200.log_mel_spectrogram_batch(torch.from_numpy(pcm).to(DEV))
    assert torch.equal(got, b200.encoder_stem2(mel, *p))


def test_stem2_half_output_is_the_rounded_float32_result(b200):
    """`dtype=torch.float16` (model.py:197 `.to(x.dtype)` of a half-precision model): the float32 result rounded once."""
    p = [t.to(DEV) for t in _params(384, 21)]
    pos = _sinusoids(750, 384).to(DEV)
    for n_frames in (1500, 1499, 37):
        mel = (torch.rand(3, 80, n_frames, generator=torch.Generator().manual_seed(n_frames)) * 2.5 - 1.0).to(DEV)
        e = pos[: (n_frames + 1) // 2] if n_frames != 1500 else pos
        full = b200.encoder_stem2(mel, *p, e)
        half = b200.encoder_stem2(mel, *p, e, dtype=torch.float16)
        assert half.dtype == torch.float16 and half.shape == full.shape
        assert torch.equal(half, full.half())
    wave = torch.from_numpy(np.stack([signals.make_signal("gauss", 48000, 3), signals.make_signal("chirp", 48000, 4)])).to(DEV)
    assert torch.equal(b200.log_mel_encoder_stem2(wave, *p, dtype=torch.float16), b200.log_mel_encoder_stem2(wave, *p).half())
    with pytest.raises(ValueError):
        b200.encoder_stem2(mel, *p, dtype=torch.bfloat16)


@pytest.mark.parametrize("n_frames", [3000, 131, 514, 37])
def test_stem2_kernels_write_inside_their_buffers(b200, n_frames):
    """Both stem kernels through the C ABI into buffers with guard bands: nothing outside [batch, frames, n_state] is touched
    (whole pieces leave by TMA tensor store, the partial ones at a clip's end by themselves; the padding frame of an odd
    frame count is the caller's)."""
    from asr_ttl_mtl_b200 import _native

    lib = _native.load()
    batch, n_state, guard = 3, 384, 4096
    w1, b1, w2, b2 = [t.to(DEV) for t in _params(n_state, 31)]
    packed = b200.pack_conv2_weight(w2)
    mel = (torch.rand(batch, 80, n_frames, generator=torch.Generator().manual_seed(5)) * 2.5 - 1.0).to(DEV)
    padded = n_frames + (n_frames & 1)
    h_elems, o_elems = batch * padded * n_state, batch * (padded // 2) * n_state
    h_buf = torch.full((guard + h_elems + guard,), 7.0, dtype=torch.float16, device=DEV)
    o_buf = torch.full((guard + o_elems + guard,), 7.0, dtype=torch.float32, device=DEV)
    h1 = h_buf[guard:guard + h_elems].view(batch, padded, n_state)
    out = o_buf[guard:guard + o_elems].view(batch, padded // 2, n_state)
    stream = torch.cuda.current_stream().cuda_stream
    _native.check(lib.b200mel_stem_conv1_gelu_fm16_device(mel.data_ptr(), None, 0, batch, 80, n_frames, w1.data_ptr(), b1.data_ptr(),
                                                          n_state, h1.data_ptr(), stream))
    if padded != n_frames:
        assert bool((h1[:, n_frames:] == 7.0).all())                       # the padding frame is not the kernel's to write
        h1[:, n_frames:].zero_()
    _native.check(lib.b200mel_stem_conv2_gelu_device(h1.data_ptr(), batch, padded, packed.data_ptr(), b2.data_ptr(), None, n_state,
                                                     out.data_ptr(), 0, stream))
    torch.cuda.synchronize()
    for buf, n in ((h_buf, h_elems), (o_buf, o_elems)):
        assert bool((buf[:guard] == 7.0).all()) and bool((buf[guard + n:] == 7.0).all())
    assert torch.equal(out, b200.encoder_stem2(mel, w1, b1, packed, b2))
    want = F.gelu(F.conv1d(mel, w1, b1, padding=1)).permute(0, 2, 1)
    assert float((h1[:, :n_frames].float() - want).abs().max()) <= 5e-3    # conv1 in half, frames major
    # bad arguments come back as status codes, not launches
    assert lib.b200mel_stem_conv2_gelu_device(h1.data_ptr(), batch, padded + 1, packed.data_ptr(), b2.data_ptr(), None, n_state,
                                              out.data_ptr(), 0, stream) != 0
    assert lib.b200mel_stem_conv2_gelu_device(h1.data_ptr(), batch, padded, packed.data_ptr(), b2.data_ptr(), None, 100,
                                              out.data_ptr(), 0, stream) != 0
    assert lib.b200mel_stem_conv2_gelu_device(h1.data_ptr(), batch, padded, packed.data_ptr(), b2.data_ptr(), None, n_state,
                                              out.data_ptr(), 1, stream) != 0
    assert lib.b200mel_stem_conv2_gelu_device(None, batch, padded, packed.data_ptr(), b2.data_ptr(), None, n_state,
                                              out.data_ptr(), 0, stream) != 0
