"""Host-side mirror of whisper/audio.py: names, constants, pad_or_trim, mel_filters, error behaviour."""
import inspect
import sys
import types

import numpy as np
import pytest
import torch

import asr_ttl_mtl_b200 as b200
from asr_ttl_mtl_b200 import audio


def test_constants_match_the_reference(golden):
    for name, value in golden.meta["constants"].items():
        assert getattr(audio, name) == value and getattr(b200, name) == value


def test_signatures_match_the_reference():
    sig = inspect.signature(audio.log_mel_spectrogram)
    assert list(sig.parameters) == ["audio", "n_mels", "padding", "device"]
    assert [p.default for p in sig.parameters.values()][1:] == [80, 0, None]
    sig = inspect.signature(audio.pad_or_trim)
    assert list(sig.parameters) == ["array", "length", "axis"]
    assert sig.parameters["length"].default == 480000 and sig.parameters["axis"].kind is inspect.Parameter.KEYWORD_ONLY
    assert list(inspect.signature(audio.mel_filters.__wrapped__).parameters) == ["device", "n_mels"]
    assert list(inspect.signature(audio.load_audio).parameters) == ["file", "sr"]


def test_pad_or_trim_known_answers(golden):
    assert np.array_equal(audio.pad_or_trim(np.arange(5, dtype=np.float32), 8), golden["pot_pad_np"])
    assert np.array_equal(audio.pad_or_trim(np.arange(10, dtype=np.float32), 4), golden["pot_trim_np"])
    m = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    assert np.array_equal(audio.pad_or_trim(m, 5, axis=1).numpy(), golden["pot_axis1_pad_t"])
    assert np.array_equal(audio.pad_or_trim(m, 1, axis=0).numpy(), golden["pot_axis0_trim_t"])
    assert np.array_equal(audio.pad_or_trim(m, 6).numpy(), golden["pot_last_pad_t"])
    assert np.array_equal(audio.pad_or_trim(m.numpy(), 5, axis=1), golden["pot_axis1_pad_t"])


def test_pad_or_trim_identity_dtype_and_kind():
    x = np.zeros(480000, np.float32)
    assert audio.pad_or_trim(x) is x
    t = torch.zeros(3000, dtype=torch.float16)
    assert audio.pad_or_trim(t, 3000) is t
    assert audio.pad_or_trim(torch.ones(7, dtype=torch.int16), 9).dtype == torch.int16
    assert isinstance(audio.pad_or_trim(np.ones(3), 2), np.ndarray)
    mel = torch.randn(80, 1234)
    out = audio.pad_or_trim(mel, 3000)  # transcribe.py:151,286 use on a mel
    assert out.shape == (80, 3000) and torch.equal(out[:, :1234], mel) and not out[:, 1234:].any()


def test_mel_filters_contract(golden):
    f = audio.mel_filters("cpu", 80)
    assert f.dtype == torch.float32 and f.shape == (80, 201) and f.device.type == "cpu"
    assert np.array_equal(f.numpy(), golden["filters_80"])
    assert audio.mel_filters("cpu", 80) is f  # lru_cache, like the reference
    assert np.array_equal(audio.mel_filters("cpu", 128).numpy(), golden["filters_128"])
    with pytest.raises(AssertionError, match="Unsupported n_mels: 64"):
        audio.mel_filters("cpu", 64)


def test_load_audio_failure_is_a_runtime_error(tmp_path):
    import shutil

    if shutil.which("ffmpeg") is None:
        with pytest.raises((RuntimeError, FileNotFoundError)):
            audio.load_audio(str(tmp_path / "missing.wav"))
    else:
        with pytest.raises(RuntimeError, match="Failed to load audio"):
            audio.load_audio(str(tmp_path / "missing.wav"))


def test_argument_validation_happens_before_any_gpu_work():
    with pytest.raises(AssertionError, match="Unsupported n_mels"):
        audio.log_mel_spectrogram(np.zeros(16000, np.float32), n_mels=64)
    with pytest.raises(RuntimeError):
        audio.log_mel_spectrogram(np.zeros((2, 2, 16000), np.float32))
    with pytest.raises(RuntimeError):
        audio.log_mel_spectrogram(np.zeros(16000, np.float64))
    with pytest.raises(RuntimeError):
        audio.log_mel_spectrogram(np.zeros(16000, np.int16))
    with pytest.raises(RuntimeError):
        audio.log_mel_spectrogram_batch(np.zeros(16000, np.float32))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        audio.log_mel_spectrogram(np.zeros(16000, np.float32))


def test_pack_conv2_weight_is_tap_major_half():
    """The second stem layer's operand (model.py:180): conv2.weight [n, c, 3] -> half [3, n, c]."""
    w = torch.arange(4 * 4 * 3, dtype=torch.float32).reshape(4, 4, 3) / 7
    packed = b200.pack_conv2_weight(w)
    assert packed.dtype == torch.float16 and tuple(packed.shape) == (3, 4, 4) and packed.is_contiguous()
    for k in range(3):
        assert torch.equal(packed[k], w[:, :, k].half())
    with pytest.raises(ValueError):
        b200.pack_conv2_weight(torch.zeros(4, 5, 3))
    with pytest.raises(ValueError):
        b200.pack_conv2_weight(torch.zeros(4, 4, 5))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_encoder_stem_has_no_cpu_fallback():
    w1, b1 = torch.zeros(384, 80, 3), torch.zeros(384)
    w2, b2 = torch.zeros(384, 384, 3), torch.zeros(384)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200.encoder_stem(torch.zeros(1, 80, 100), w1, b1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200.encoder_stem2(torch.zeros(1, 80, 100), w1, b1, w2, b2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200.log_mel_encoder_stem2(torch.zeros(1, 16000), w1, b1, w2, b2)


def test_install_rebinds_consumer_namespaces():
    fake_audio = types.ModuleType("whisper.audio")
    fake_pkg = types.ModuleType("whisper")
    fake_ds = types.ModuleType("speech_disorder.dataset")
    sentinel = object()
    for mod in (fake_audio, fake_pkg, fake_ds):
        mod.log_mel_spectrogram = sentinel
        mod.pad_or_trim = sentinel
    fake_audio.mel_filters = sentinel
    saved = {k: sys.modules.get(k) for k in ("whisper.audio", "whisper", "speech_disorder.dataset")}
    sys.modules.update({"whisper.audio": fake_audio, "whisper": fake_pkg, "speech_disorder.dataset": fake_ds})
    try:
        rebound = b200.install()
        assert fake_ds.log_mel_spectrogram is audio.log_mel_spectrogram
        assert fake_pkg.pad_or_trim is audio.pad_or_trim and fake_audio.mel_filters is audio.mel_filters
        assert set(rebound) == {"whisper.audio", "whisper", "speech_disorder.dataset"}
        b200.uninstall()
        assert fake_ds.log_mel_spectrogram is sentinel and fake_audio.mel_filters is sentinel
    finally:
        b200.uninstall()
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
