"""Parity of the CUDA path (through the C ABI) against the golden vectors and the oracle.  `-m gpu`.

Bar: bit-exact shapes / frame counts / pad-trim indexing; <= 1e-4 max-abs on
the normalised log-mel in fp32.
"""
import numpy as np
import pytest
import torch

from oracle import logmel_oracle as orc
from oracle import signals

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = "cuda:0"
VARIANTS = ["fft", "tcgen05"]


@pytest.fixture(autouse=True, params=VARIANTS)
def default_variant(request, b200):
    """Every test below runs once per STFT kernel: "auto" (what the unchanged consumers get) is pinned to each in turn."""
    from asr_ttl_mtl_b200 import audio as audio_module

    saved = audio_module.DEFAULT_VARIANT
    audio_module.DEFAULT_VARIANT = request.param
    yield request.param
    audio_module.DEFAULT_VARIANT = saved


def _maxerr(a, b):
    return float((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs().max())


def test_library_is_the_thing_that_runs(b200):
    before = b200.gpu_launches()
    b200.log_mel_spectrogram(torch.zeros(16000, device=DEV))
    torch.cuda.synchronize()
    assert b200.gpu_launches() >= before + 1  # the fused persistent kernel


def test_golden_cases_single_utterance_api(b200, golden):
    worst = 0.0
    for c in golden.cases:
        x = golden.signal(c)
        got = b200.log_mel_spectrogram(x, n_mels=c["n_mels"], padding=c["padding"], device=DEV)
        assert got.device.type == "cuda" and got.dtype == torch.float32 and got.is_contiguous()
        assert tuple(got.shape) == tuple(c["shape"])
        err = _maxerr(got, golden.out(c))
        worst = max(worst, err)
        assert err <= TOL, (c, err)
    print(f"worst golden error {worst:.3e}")


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n_mels", [80, 128])
def test_batch_is_per_utterance_stack_of_oracle_calls(b200, n_mels, variant):
    kinds = list(signals.KINDS)
    batch = np.stack([signals.make_signal(k, 40000, 300 + i) for i, k in enumerate(kinds)])
    got = b200.log_mel_spectrogram_batch(torch.from_numpy(batch).to(DEV), n_mels=n_mels, variant=variant)
    want = orc.logmel_f32_port_per_utterance(torch.from_numpy(batch), n_mels)
    assert tuple(got.shape) == (len(kinds), n_mels, 250)
    for i, k in enumerate(kinds):
        assert _maxerr(got[i], want[i]) <= TOL, (k, _maxerr(got[i], want[i]))
    f64 = np.stack([orc.logmel_f64(x, n_mels) for x in batch])
    # adversarial signals: stay as close to the float64 spec as the reference does (+5e-5)
    for i, k in enumerate(kinds):
        assert _maxerr(got[i], f64[i]) <= _maxerr(want[i], f64[i]) + 5e-5, k


def test_tcgen05_variant_golden_cases(b200, golden):
    worst = 0.0
    for c in golden.cases:
        x = torch.from_numpy(golden.signal(c)).to(DEV)[None]
        got = b200.log_mel_spectrogram_batch(x, n_mels=c["n_mels"], padding=c["padding"], variant="tcgen05")[0]
        assert tuple(got.shape) == tuple(c["shape"])
        err = _maxerr(got, golden.out(c))
        worst = max(worst, err)
        assert err <= TOL, (c, err)
    print(f"worst golden error (tcgen05) {worst:.3e}")


def test_tcgen05_variant_lengths_pcm16_and_full_batch(b200):
    lens = np.array([0, 201, 16000, 47999, 48000, 52000], dtype=np.int64)
    rows = np.stack([signals.make_signal("gauss", 48000, 900 + i) for i in range(len(lens))])
    padded = rows.copy()
    for i, n in enumerate(lens):
        padded[i, min(n, 48000):] = 0.0
    a = b200.log_mel_spectrogram_batch(torch.from_numpy(rows).to(DEV), lengths=torch.from_numpy(lens), variant="tcgen05")
    b = b200.log_mel_spectrogram_batch(torch.from_numpy(padded).to(DEV), variant="tcgen05")
    assert torch.equal(a, b) and torch.all(a[0] == -1.5)
    assert _maxerr(a, orc.logmel_f32_port_per_utterance(torch.from_numpy(padded), 80)) <= TOL
    q = np.stack([signals.make_pcm16(32000, 60 + i) for i in range(3)])
    f = q.astype(np.float32) / 32768.0
    assert torch.equal(b200.log_mel_spectrogram_batch(torch.from_numpy(q).to(DEV), n_mels=128, variant="tcgen05"),
                       b200.log_mel_spectrogram_batch(torch.from_numpy(f).to(DEV), n_mels=128, variant="tcgen05"))
    gen = torch.Generator(device=DEV).manual_seed(7)
    audio = (0.1 * torch.randn(64, 480000, generator=gen, device=DEV)).clamp_(-1, 1)
    tc = b200.log_mel_spectrogram_batch(audio, variant="tcgen05")
    fft = b200.log_mel_spectrogram_batch(audio, variant="fft")
    assert _maxerr(tc, fft) <= TOL and torch.isfinite(tc).all()


def test_reference_2d_semantics_one_max_per_call(b200, golden):
    scale = golden["batch2d_in_scale"]
    batch = np.stack([signals.make_signal("gauss", 16000, 50) * s for s in scale])
    got = b200.log_mel_spectrogram(torch.from_numpy(batch), device=DEV)  # drop-in API: global max
    assert _maxerr(got, golden["batch2d_out"]) <= TOL
    per = b200.log_mel_spectrogram_batch(torch.from_numpy(batch).to(DEV))
    assert _maxerr(per, golden["batch2d_out"]) > 0.1


def test_lengths_fast_path_equals_zero_filled_rows(b200):
    lens = np.array([0, 1, 201, 16000, 31999, 47999, 48000, 52000], dtype=np.int64)
    rows = np.stack([signals.make_signal("gauss", 48000, 900 + i) for i in range(len(lens))])
    padded = rows.copy()
    for i, n in enumerate(lens):
        padded[i, min(n, 48000):] = 0.0
        if i % 2:
            rows[i, min(n, 48000):] = np.nan      # what lies behind an utterance's end must not matter, whatever it is
    a = b200.log_mel_spectrogram_batch(torch.from_numpy(rows).to(DEV), lengths=torch.from_numpy(lens))
    b = b200.log_mel_spectrogram_batch(torch.from_numpy(padded).to(DEV))
    assert torch.equal(a, b)  # same arithmetic on the same values
    want = orc.logmel_f32_port_per_utterance(torch.from_numpy(padded), 80)
    assert _maxerr(a, want) <= TOL
    assert torch.all(a[0] == -1.5)  # an all-zero utterance (reference: every value (-10+4)/4)


def test_lengths_path_with_many_tiles_per_sm(b200):
    """Every SM walks a dozen tiles, about half of them wholly behind an utterance's end (not staged, not folded, not
    multiplied) and one per utterance cut by it: the hand-overs between the kernel's warp roles must come out the same as
    for zero-filled rows, call after call, and no wait may ever time out."""
    from asr_ttl_mtl_b200 import _native

    rng = np.random.default_rng(77)
    lens = rng.integers(0, 480001, size=96).astype(np.int64)
    lens[:4] = [0, 480000, 200, 479999]
    gen = torch.Generator(device=DEV).manual_seed(5)
    x = 0.1 * torch.randn(96, 480000, generator=gen, device=DEV)
    padded = x.clone()
    padded.masked_fill_(torch.arange(480000, device=DEV)[None, :] >= torch.from_numpy(lens).to(DEV)[:, None], 0.0)
    want = b200.log_mel_spectrogram_batch(padded)
    lens_dev = torch.from_numpy(lens).to(DEV).to(torch.int32)
    b200.log_mel_spectrogram_batch(x, lengths=lens_dev)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    got = [b200.log_mel_spectrogram_batch(x, lengths=lens_dev) for _ in range(8)]
    stop.record()
    torch.cuda.synchronize()
    for g in got:
        assert torch.equal(g, want)
    assert _native.kernel_fault() == (0, 0)
    # (a call over 96 clips takes well under a millisecond plus the allocator's work; a hand-over that times out, seconds)
    assert start.elapsed_time(stop) < 1500.0
    for i in (0, 1, 2, 3, 50):
        assert _maxerr(want[i], orc.logmel_f32_port(padded[i].cpu().numpy(), 80)) <= TOL


def test_variable_length_clips_like_the_training_set(b200):
    # BASELINE config 4: 1-30 s clips pad_or_trim-med to N_SAMPLES, batches of 16
    lens = signals.variable_lengths(16)
    clips = [signals.make_signal("gauss", int(n), 40 + i) for i, n in enumerate(lens)]
    padded = np.stack([b200.pad_or_trim(c) for c in clips])
    assert padded.shape == (16, 480000)
    got = b200.log_mel_spectrogram_batch(torch.from_numpy(padded).to(DEV), lengths=torch.from_numpy(lens))
    assert tuple(got.shape) == (16, 80, 3000)
    for i in (0, 5, 15):
        want = orc.logmel_f32_port(padded[i], 80)
        assert _maxerr(got[i], want) <= TOL
        tail = got[i][:, int(lens[i]) // 160 + 3:]
        if tail.numel():
            assert torch.all(tail == tail.flatten()[0])  # silence sits exactly on the max-8 clamp
            assert abs(float(got[i].max() - tail.flatten()[0]) - 2.0) < 1e-6


def test_dynamic_range_clamp_paths(b200):
    """The clamp at max - 8 (audio.py:155) when it touches whole tiles (digital silence), parts of tiles (a noise floor
    100 dB down, a loud click in a quiet clip) and nothing at all - the tcgen05 kernel treats each case differently."""
    rng = np.random.default_rng(5)
    n = 16000 * 12
    loud = (0.2 * rng.standard_normal(n)).astype(np.float32)
    clips = np.zeros((5, n), dtype=np.float32)
    clips[0] = loud                                              # nothing below the clamp
    clips[1, : n // 3] = loud[: n // 3]                          # speech, then digital silence (tiles to fill)
    clips[2] = loud
    clips[2, n // 4: n // 2] *= 1e-5                             # a stretch 100 dB down: clamped value by value
    clips[3] = 1e-4 * loud
    clips[3, 70000:70040] = 0.9                                  # one click sets the max; the rest is mostly below max - 8
    clips[4, 1000] = 1.0                                         # an impulse in silence
    got = b200.log_mel_spectrogram_batch(torch.from_numpy(clips).to(DEV))
    want = orc.logmel_f32_port_per_utterance(torch.from_numpy(clips), 80)
    for i in range(len(clips)):
        assert _maxerr(got[i], want[i]) <= TOL, (i, _maxerr(got[i], want[i]))
    # the same clips shuffled inside a larger batch give the same bytes
    big = np.concatenate([clips[::-1], clips, clips[2:3]])
    again = b200.log_mel_spectrogram_batch(torch.from_numpy(big).to(DEV))
    assert torch.equal(again[5:10], got) and torch.equal(again[10], got[2])


def test_collate_log_mels_equals_the_dataset_path(b200):
    """raw variable-length clips -> batch['mels'], equal to dataset.py:82-89 + collate (:179)."""
    rng = np.random.default_rng(8)
    lens = [16000, 123457, 480000, 480000 + 5000, 300]          # short, odd, exact, to be trimmed, very short
    clips = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in lens]
    got = b200.collate_log_mels(clips, device=DEV)
    assert tuple(got.shape) == (len(clips), 80, 3000) and got.device.type == "cuda"
    for i, c in enumerate(clips):
        want = orc.logmel_f32_port(orc.pad_or_trim_oracle(c, 480000), 80)
        assert _maxerr(got[i], want) <= TOL, (i, _maxerr(got[i], want))
    pcm = [np.round(c * 32767).astype(np.int16) for c in clips]
    got16 = b200.collate_log_mels(pcm, device=DEV)
    ref16 = b200.collate_log_mels([p.astype(np.float32) / 32768.0 for p in pcm], device=DEV)
    assert torch.equal(got16, ref16)
    with pytest.raises(ValueError):
        b200.collate_log_mels([], device=DEV)


def test_float16_emission_is_the_rounded_float32_result(b200, default_variant):
    """out_dtype=float16 stores exactly what `.to(torch.float16)` of the float32 result gives
    (transcribe.py:286 feeds the fp16 model that way), clamp paths included; float32 stays the default."""
    rng = np.random.default_rng(21)
    n = 16000 * 9
    clips = (0.1 * rng.standard_normal((4, n))).astype(np.float32)
    clips[1, n // 2:] = 0.0                    # silent tail: tiles filled by the clamp path
    clips[2, 30000:60000] *= 1e-5              # value-by-value clamp
    x = torch.from_numpy(clips).to(DEV)
    if default_variant == "fft":
        with pytest.raises(Exception):
            b200.log_mel_spectrogram_batch(x, out_dtype=torch.float16)
        return
    f32 = b200.log_mel_spectrogram_batch(x)
    f16 = b200.log_mel_spectrogram_batch(x, out_dtype=torch.float16)
    assert f16.dtype == torch.float16 and f16.shape == f32.shape and f16.is_contiguous()
    assert torch.equal(f16, f32.half())
    for n_mels in (80, 128):                   # transcribe's call shape: one utterance, 30 s of right padding
        one32 = b200.log_mel_spectrogram_batch(x[:1], n_mels=n_mels, padding=480000)
        one16 = b200.log_mel_spectrogram_batch(x[:1], n_mels=n_mels, padding=480000, out_dtype=torch.float16)
        assert torch.equal(one16, one32.half())
    host16 = b200.log_mel_spectrogram_batch(torch.from_numpy(clips), out_dtype=torch.float16)   # host buffers in and out
    assert host16.device.type == "cpu" and torch.equal(host16, f16.cpu())
    assert _maxerr(f16.float(), orc.logmel_f32_port_per_utterance(torch.from_numpy(clips), 80)) <= 1e-3   # half: 2^-11 relative


def test_pcm16_ingest_is_bit_equal_to_the_float_path(b200):
    q = np.stack([signals.make_pcm16(32000, 60 + i) for i in range(4)])
    f = q.astype(np.float32) / 32768.0  # audio.py:62
    a = b200.log_mel_spectrogram_batch(torch.from_numpy(q).to(DEV), n_mels=128)
    b = b200.log_mel_spectrogram_batch(torch.from_numpy(f).to(DEV), n_mels=128)
    assert torch.equal(a, b)
    assert _maxerr(a, orc.logmel_f32_port_per_utterance(torch.from_numpy(f), 128)) <= TOL


def test_host_buffer_entry_point_matches_device_path(b200):
    x = np.stack([signals.make_signal("uniform", 480000, 70 + i) for i in range(5)])
    dev = b200.log_mel_spectrogram_batch(torch.from_numpy(x).to(DEV)).cpu()
    host = b200.log_mel_spectrogram_batch(torch.from_numpy(x))  # CPU tensor in -> CPU tensor out
    assert host.device.type == "cpu" and torch.equal(host, dev)
    pinned = torch.from_numpy(x).pin_memory()
    out = torch.empty(5, 80, 3000).pin_memory()
    res = b200.log_mel_spectrogram_batch(pinned, out=out)
    assert res.data_ptr() == out.data_ptr() and torch.equal(out, dev)
    single = b200.log_mel_spectrogram(x[2])  # numpy in, like dataset.py:89
    assert single.device.type == "cpu" and torch.equal(single, dev[2])
    lens = torch.tensor([480000, 1000, 240000, 0, 479999])
    a = b200.log_mel_spectrogram_batch(torch.from_numpy(x), lengths=lens)
    b = b200.log_mel_spectrogram_batch(torch.from_numpy(x).to(DEV), lengths=lens).cpu()
    assert torch.equal(a, b)


def test_non_contiguous_and_strided_rows(b200):
    base = torch.from_numpy(np.stack([signals.make_signal("gauss", 20000, 80 + i) for i in range(6)])).to(DEV)
    rows = base[::2]  # row pitch 2 * L, inner stride 1: consumed in place through stride_b
    assert torch.equal(b200.log_mel_spectrogram_batch(rows), b200.log_mel_spectrogram_batch(rows.contiguous()))
    inter = base.t().contiguous().t()  # inner stride != 1 -> copied by the wrapper
    assert torch.equal(b200.log_mel_spectrogram_batch(inter), b200.log_mel_spectrogram_batch(base))
    odd = base[:, 1:]  # 4-byte aligned rows (no 16-byte alignment)
    want = orc.logmel_f32_port_per_utterance(odd.cpu(), 80)
    assert _maxerr(b200.log_mel_spectrogram_batch(odd), want) <= TOL


def test_transcribe_call_pattern(b200):
    # transcribe.py:139-141: whole file, padding=N_SAMPLES, then pad_or_trim(mel, N_FRAMES) windows
    x = signals.make_signal("sine1k_noise", 16000 * 47 + 123, 9)
    mel = b200.log_mel_spectrogram(x, 80, padding=b200.N_SAMPLES, device=DEV)
    assert tuple(mel.shape) == (80, (x.size + 480000) // 160)
    assert _maxerr(mel, orc.logmel_f32_port(x, 80, padding=480000)) <= TOL
    content_frames = mel.shape[-1] - b200.N_FRAMES
    seg = b200.pad_or_trim(mel[:, 3000:3000 + min(3000, content_frames - 3000)], b200.N_FRAMES)
    assert tuple(seg.shape) == (80, 3000) and seg.device.type == "cuda"


def test_mel_windows_equal_the_decoding_loop_cuts(b200):
    # transcribe.py:282-286: mel_segment = pad_or_trim(mel[:, seek : seek + segment_size], N_FRAMES).to(device).to(dtype),
    # every window of the decoding loop in one launch, float16 (what the fp16 model gets) and float32
    x = signals.make_signal("chirp", 16000 * 95 + 77, 3)
    mel = b200.log_mel_spectrogram(x, 80, padding=b200.N_SAMPLES, device=DEV)
    content_frames = mel.shape[-1] - b200.N_FRAMES
    seeks = [0, 1, 2999, 3000, 4321, 6001, content_frames - 1234, content_frames - 1, mel.shape[-1] - 10]
    sizes = [min(b200.N_FRAMES, max(content_frames - s, 0)) if i % 2 == 0 else b200.N_FRAMES for i, s in enumerate(seeks)]
    for dtype in (torch.float16, torch.float32):
        want = torch.stack([b200.pad_or_trim(mel[:, s:s + n], b200.N_FRAMES).to(dtype) for s, n in zip(seeks, sizes)])
        got = b200.mel_windows(mel, seeks, sizes, dtype=dtype)
        assert got.dtype == dtype and tuple(got.shape) == (len(seeks), 80, 3000)
        assert torch.equal(got, want)
    # whole windows by default; the language-detection window of transcribe.py:150
    first = b200.mel_windows(mel, [0], dtype=torch.float32)
    assert torch.equal(first[0], b200.pad_or_trim(mel, b200.N_FRAMES))
    # a window size that is not a multiple of four and an output the vector path cannot take
    odd = b200.mel_windows(mel, [5, 17], [101, 7], window_frames=1001, dtype=torch.float16)
    want = torch.stack([b200.pad_or_trim(mel[:, 5:106], 1001), b200.pad_or_trim(mel[:, 17:24], 1001)]).half()
    assert torch.equal(odd, want)
    assert b200.mel_windows(mel, [], dtype=torch.float16).shape == (0, 80, 3000)
    with pytest.raises(RuntimeError):
        b200.mel_windows(mel.cpu(), [0])


def test_front_end_next_to_a_kernel_that_holds_the_sms(b200):
    # the trainer's situation: the front-end is launched while another stream keeps the SMs busy (here: a chain of large
    # matmuls), so only some of its CTAs are resident at a time.  No CTA waits for another one, so it must finish with
    # the same bits as a launch on an idle GPU.
    x = torch.from_numpy(np.stack([signals.make_signal("gauss", 480000, 300 + i) for i in range(24)])).to(DEV)
    x[5, 200000:] = 0.0                                         # one clip with a zero tail: the finish kernel's clamp path
    want = b200.log_mel_spectrogram_batch(x)
    torch.cuda.synchronize()
    a = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    side, main = torch.cuda.Stream(), torch.cuda.current_stream()
    with torch.cuda.stream(side):
        for _ in range(40):
            a = (a @ a).clamp_(-1, 1)
    got = [b200.log_mel_spectrogram_batch(x) for _ in range(6)]   # enqueued while the matmuls run
    busy = not side.query()
    torch.cuda.synchronize()
    assert busy, "the side stream had already drained: the test did not overlap anything"
    for g in got:
        assert torch.equal(g, want)
    from asr_ttl_mtl_b200 import _native
    assert _native.kernel_fault() == (0, 0)
    del main


def test_nan_and_inf_poison_only_their_own_utterance(b200):
    x = np.stack([signals.make_signal("gauss", 16000, 20 + i) for i in range(3)])
    x[1, 777] = np.nan
    x[2, 12345] = np.inf
    got = b200.log_mel_spectrogram_batch(torch.from_numpy(x).to(DEV)).cpu()
    assert not torch.isnan(got[0]).any()
    assert torch.isnan(got[1]).all() and torch.isnan(got[2]).all()  # torch.max propagates NaN (audio.py:155)
    assert torch.isnan(orc.logmel_f32_port(x[1], 80)).all()


def test_error_behaviour_matches_the_reference(b200):
    with pytest.raises(AssertionError, match="Unsupported n_mels"):
        b200.log_mel_spectrogram(torch.zeros(16000, device=DEV), n_mels=100)
    with pytest.raises(RuntimeError):
        b200.log_mel_spectrogram(torch.zeros(200, device=DEV))  # reflect pad needs > 200 samples
    with pytest.raises(RuntimeError):
        b200.log_mel_spectrogram(torch.zeros(2, 3, 1600, device=DEV))
    with pytest.raises(RuntimeError):
        b200.log_mel_spectrogram(torch.zeros(16000, device=DEV, dtype=torch.float64))
    assert tuple(b200.log_mel_spectrogram(torch.zeros(201, device=DEV)).shape) == (80, 1)
    assert tuple(b200.log_mel_spectrogram(torch.zeros(100, device=DEV), padding=101).shape) == (80, 1)
    assert tuple(b200.log_mel_spectrogram(torch.zeros(1000, device=DEV), padding=-7).shape) == (80, 6)


def test_input_is_not_aliased_or_modified(b200):
    x = torch.from_numpy(signals.make_signal("gauss", 16000, 1)).to(DEV)
    keep = x.clone()
    out = b200.log_mel_spectrogram(x)
    assert torch.equal(x, keep) and out.data_ptr() != x.data_ptr() and not out.requires_grad


def test_runs_on_the_current_stream(b200):
    x = torch.from_numpy(signals.make_signal("gauss", 48000, 2)).to(DEV)
    ref = b200.log_mel_spectrogram(x)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        y = b200.log_mel_spectrogram(x)
    side.synchronize()
    assert torch.equal(y, ref)


def _full_size_report(tag, got, ref, f64):
    """Worst deviations over EVERY value of a full-size batch; returns (|gpu - ref|, |gpu - f64|, |ref - f64|) maxima and the
    mask of values where the GPU result is more than TOL from the reference."""
    d_ref, d_f64, e_ref = np.abs(got - ref), np.abs(got - f64), np.abs(ref - f64)
    far = d_ref > TOL
    print(f"{tag}: {got.size / 1e6:.1f} M values  |gpu - ref| {d_ref.max():.3e} ({int(far.sum())} values > {TOL:g})  "
          f"|gpu - f64| {d_f64.max():.3e}  |ref - f64| {e_ref.max():.3e}")
    return d_ref, d_f64, e_ref, far


def _assert_full_size_parity(tag, got, ref, f64, shipped):
    """The bar at full size.  Every value of the GPU result is within 1e-4 of the float64 statement of the reference's
    formulas, and all but a handful in 10^8 within 1e-4 of the reference's own fp32 result.  For the kernel that ships
    (tcgen05) the handful is pinned down: it happens only where the reference ITSELF is more than 5e-5 from float64
    (isolated values whose mel power is ~1e-7 of the frame's energy: any two fp32 evaluations differ there), and there
    the GPU value is the closer one to float64."""
    d_ref, d_f64, e_ref, far = _full_size_report(tag, got, ref, f64)
    assert d_f64.max() <= TOL, f"{tag}: {d_f64.max():.3e} from float64"
    assert far.sum() <= 1e-6 * got.size
    assert d_ref.max() <= 1.5 * TOL
    if shipped:
        assert np.all(e_ref[far] > 5e-5), f"{tag}: more than {TOL:g} from a reference value that is itself within 5e-5 of float64"
        assert np.all(d_f64[far] < e_ref[far]), f"{tag}: further from float64 than the reference"


@pytest.mark.parametrize("n_mels", [80, 128])
def test_full_size_batch_every_clip(b200, n_mels, default_variant):
    """BASELINE configs 2 / 3 at full size - 256 clips x 30 s, the batch bench.py times (seed 1234) - EVERY clip against
    the oracle (fp32 port per utterance and float64), plus the size-independent properties."""
    gen = torch.Generator(device=DEV).manual_seed(1234)
    audio = (0.1 * torch.randn(256, 480000, generator=gen, device=DEV)).clamp_(-1, 1)
    out = b200.log_mel_spectrogram_batch(audio, n_mels=n_mels)
    assert tuple(out.shape) == (256, n_mels, 3000)
    host = audio.cpu()
    ref = orc.logmel_f32_port_per_utterance(host, n_mels).numpy()
    f64 = np.stack([orc.logmel_f64(c, n_mels) for c in host.numpy()])
    _assert_full_size_parity(f"config {'2' if n_mels == 80 else '3'} [{default_variant}]", out.cpu().numpy(), ref, f64,
                             shipped=default_variant == "tcgen05")
    # replicas of one clip at other batch positions, and alone: the same bytes
    audio[17] = audio[3]
    audio[200] = audio[3]
    again = b200.log_mel_spectrogram_batch(audio, n_mels=n_mels)
    assert torch.equal(again[17], out[3]) and torch.equal(again[200], out[3]) and torch.equal(again[3], out[3])
    assert torch.equal(b200.log_mel_spectrogram_batch(audio[3:4], n_mels=n_mels)[0], out[3])
    # dynamic range: every utterance spans at most 8 decades => (max - min) <= 2 after (x+4)/4
    mx, mn = out.amax(dim=(1, 2)), out.amin(dim=(1, 2))
    assert torch.all(mx - mn <= 2.0 + 1e-6) and torch.isfinite(out).all()
    # checksum of checksums is reproducible run to run
    third = b200.log_mel_spectrogram_batch(audio, n_mels=n_mels)
    assert float(again.double().sum(dim=(1, 2)).sum()) == float(third.double().sum(dim=(1, 2)).sum())


def test_variable_length_epoch_every_clip(b200, default_variant):
    """BASELINE config 4: a 1,737-clip "epoch" (data/custom_train.csv) of 1-30 s clips in train batches of 16 and val
    batches of 8 (speech_disorder/config.py:15-16), zero-padded to 30 s as dataset.py:85 does - every clip of a sample of
    those batches against the oracle, padded rows and the `lengths` fast path alike."""
    lens = signals.variable_lengths(1737)
    batches = [(16, list(range(0, 96))), (16, list(range(1728, 1737))), (8, list(range(800, 840)))]
    worst = 0.0
    for bs, idx in batches:
        for b0 in range(0, len(idx), bs):
            ids = idx[b0:b0 + bs]
            clips = np.zeros((len(ids), 480000), np.float32)
            for r, i in enumerate(ids):
                clips[r, :lens[i]] = signals.make_signal("gauss", int(lens[i]), 5000 + i)
            ref = orc.logmel_f32_port_per_utterance(torch.from_numpy(clips), 80).numpy()
            x = torch.from_numpy(clips).to(DEV)
            padded = b200.log_mel_spectrogram_batch(x, n_mels=80)
            fast = b200.log_mel_spectrogram_batch(x, n_mels=80, lengths=torch.from_numpy(lens[ids]))
            assert torch.equal(padded, fast)
            err = float(np.abs(padded.cpu().numpy() - ref).max())
            worst = max(worst, err)
            assert err <= TOL, (bs, ids[0], err)
    print(f"config 4 [{default_variant}]: worst |gpu - ref| {worst:.3e} over {sum(len(i) for _, i in batches)} clips")


@pytest.mark.parametrize("scale", [1e-6, 1e-4, 1e-2, 100.0, 32768.0, 1e6])
def test_amplitude_ladder(b200, scale):
    """The reference takes any float32 waveform (e.g. int16-valued samples that were never divided by 32768): so does
    this front-end, with the same accuracy at every amplitude."""
    clips = np.stack([(scale * np.random.default_rng(50 + i).standard_normal(80000)).astype(np.float32) for i in range(4)])
    ref = orc.logmel_f32_port_per_utterance(torch.from_numpy(clips), 80)
    got = b200.log_mel_spectrogram_batch(torch.from_numpy(clips).to(DEV), n_mels=80)
    assert torch.isfinite(got).all()
    assert _maxerr(got, ref) <= TOL


def test_loud_and_quiet_inside_one_utterance(b200):
    """1 s at full scale, then 4 s at -60 dB (inside the 80 dB window): the quiet part keeps its relative precision."""
    rng = np.random.default_rng(7)
    for quiet in (1e-3, 3e-4):
        clip = (0.5 * np.concatenate([rng.standard_normal(16000), quiet * rng.standard_normal(64000)])).astype(np.float32)
        got = b200.log_mel_spectrogram(torch.from_numpy(clip).to(DEV), 80).cpu()
        assert _maxerr(got, orc.logmel_f32_port(clip, 80)) <= TOL
        assert float(np.abs(got.numpy() - orc.logmel_f64(clip, 80)).max()) <= TOL


def test_large_batch_indexing(b200):
    """BASELINE config 5 shape of work on one GPU: thousands of 30 s clips in one call (64-bit tile / row offsets, the
    per-tile workspace, the tensor map's batch dimension) - every clip equals its own single-clip call, bit for bit."""
    n_clips = 3001
    gen = torch.Generator(device=DEV).manual_seed(99)
    base = (0.1 * torch.randn(7, 480000, generator=gen, device=DEV)).clamp_(-1, 1)
    audio = base.repeat(n_clips // 7 + 1, 1)[:n_clips].contiguous()
    audio[1234, 200000:] = 0.0                       # one clip with a silent tail (clamp path inside a big batch)
    out = b200.log_mel_spectrogram_batch(audio)
    assert tuple(out.shape) == (n_clips, 80, 3000) and torch.isfinite(out).all()
    for i in (0, 6, 7, 1234, 1500, n_clips - 1):
        assert torch.equal(out[i], b200.log_mel_spectrogram_batch(audio[i:i + 1])[0]), i
    assert torch.equal(out[7 * 11 + 3], out[3])      # replicas at different batch positions
    del out, audio
    torch.cuda.empty_cache()


def test_long_single_utterance(b200):
    # one hour-scale file in one call (transcribe path): one max over the whole file
    x = signals.make_signal("gauss", 16000 * 600, 3)  # 10 minutes
    x[: 16000 * 5] *= 1e-3
    got = b200.log_mel_spectrogram(x, device=DEV)
    want = orc.logmel_f32_port(x, 80)
    assert tuple(got.shape) == (80, 60000) and _maxerr(got, want) <= TOL


def test_normalise_entry_point(b200, native_lib):
    from asr_ttl_mtl_b200 import _native

    vals = torch.tensor([[-10.0, -3.0, 0.5, 1.0], [2.0, -9.0, -5.9, -6.1]], device=DEV)
    keys = torch.zeros(64, dtype=torch.int32, device=DEV)
    enc = []
    for row in vals.cpu():
        bits = np.float32(row.max().item()).view(np.uint32)
        enc.append(int(bits | 0x80000000) if not (bits & 0x80000000) else int(~bits & 0xFFFFFFFF))
    keys[:2] = torch.tensor(np.array(enc, dtype=np.uint32).view(np.int32))
    out = vals.clone()
    _native.check(native_lib.b200mel_normalise_device(out.data_ptr(), keys.data_ptr(), 2, 4, 0,
                                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    want = (torch.maximum(vals, vals.amax(dim=1, keepdim=True) - 8.0) + 4.0) / 4.0
    assert torch.equal(out, want)


def test_abi_without_tile_keys_equals_the_wrapper(b200, native_lib, default_variant):
    """The binding of INTEGRATION.md section 2 sizes the workspace with b200mel_workspace_bytes and passes no
    B200MEL_FLAG_TILE_KEYS: zero-padded clips (digital silence is then computed and stored like any other tile, the clamp
    decided per utterance) must come out as through the Python mirror, which passes per-tile keys."""
    from asr_ttl_mtl_b200 import _native, audio

    n = 16000 * 9
    clips = np.zeros((6, n), dtype=np.float32)
    for i in range(6):
        k = [n, n // 2, 3000, 0, n - 1, 20000][i]
        clips[i, :k] = signals.make_signal("gauss", k, 700 + i) if k > 0 else 0
    x = torch.from_numpy(clips).to(DEV)
    want = b200.log_mel_spectrogram_batch(x)
    frames = _native.frames(n, 0)
    out = torch.empty(6, 80, frames, device=DEV)
    ws = torch.empty(native_lib.b200mel_workspace_bytes(6), dtype=torch.uint8, device=DEV)
    variant = {"fft": _native.VARIANT_FFT, "tcgen05": _native.VARIANT_TCGEN05}[default_variant]
    _native.check(native_lib.b200mel_logmel_device(audio._plan(0, 80), x.data_ptr(), _native.DTYPE_F32, 6, n, n, None, 0, out.data_ptr(),
                                                   ws.data_ptr(), 0, variant, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    assert torch.all(out[3] == -1.5)


def test_front_end_call_can_be_captured_in_a_cuda_graph(b200):
    """The library only enqueues work on the caller's stream (one memset, the front-end kernel, the finish kernel): the
    whole call can be captured once and replayed on new data - what a training loop with a fixed batch shape does."""
    gen = torch.Generator(device=DEV).manual_seed(11)
    x1 = 0.1 * torch.randn(16, 480000, generator=gen, device=DEV)
    x2 = 0.1 * torch.randn(16, 480000, generator=gen, device=DEV)
    x2[3, 100000:] = 0.0
    static_in = x1.clone()
    out = torch.empty(16, 80, 3000, device=DEV)
    b200.log_mel_spectrogram_batch(static_in, out=out)          # plans, module loading: outside the capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        b200.log_mel_spectrogram_batch(static_in, out=out)
    for x in (x1, x2, x1):
        static_in.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, b200.log_mel_spectrogram_batch(x))


def test_one_max_per_call_with_zero_padded_rows(b200):
    """The reference's 2-D semantics (ONE max over the call, audio.py:155) on a batch whose rows have zero tails and one row of
    nothing but zeros: the tiles of silence (not stored by the front-end, filled by the finish kernel) take the CALL's clamp."""
    n = 16000 * 8
    rows = np.zeros((5, n), dtype=np.float32)
    rows[0] = signals.make_signal("gauss", n, 1)
    rows[1, : n // 2] = 0.01 * signals.make_signal("gauss", n // 2, 2)
    rows[2, :1000] = signals.make_signal("uniform", 1000, 3)
    rows[4, : n - 7] = 3.0 * signals.make_signal("gauss", n - 7, 4)
    got = b200.log_mel_spectrogram(torch.from_numpy(rows), device=DEV)
    want = orc.logmel_f32_port(rows, 80)
    assert tuple(got.shape) == (5, 80, n // 160)
    assert _maxerr(got, want) <= TOL
    assert torch.all(got[3] == got[3].flatten()[0])      # the silent row sits on the call's max - 8 clamp
    assert abs(float(got.max() - got[3].flatten()[0]) - 2.0) < 1e-6
