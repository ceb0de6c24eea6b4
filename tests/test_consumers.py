"""The REAL consumers of the path, unchanged, after `install()`:

* `MultiTaskSpeechDataset.__getitem__` + `get_collate_fn()` (speech_disorder/dataset.py:75-96,132-219) - what
  scripts/train_disease.py wraps in its DataLoaders - building `batch['mels']`;
* `transcribe()` (whisper/transcribe.py:38-514): whole-file mel with `padding=N_SAMPLES` (:139), the language-id window
  (:151) and every 3000-frame window of the seek loop (:272-286), with a stub model that records what it is handed.

The reference packages are imported from /root/reference where that exists, else from the unmodified copy
baseline/vendor_reference.py put under baseline/_ref/tree (git-ignored; it travels to the GPU box); `jiwer`, which the
image lacks, is stubbed (speech_disorder/__init__.py pulls trainer.py:7).  The same batch / windows are computed first
with the reference's own audio functions, then through the rebinding.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import vendor_reference  # noqa: E402

from oracle import signals  # noqa: E402

TOL = 1e-4


@pytest.fixture(scope="module")
def reference_packages():
    tree = vendor_reference.tree_path()
    if tree is None:
        pytest.skip("no reference tree (neither /root/reference nor baseline/_ref/tree)")
    saved_path, saved_modules = list(sys.path), {k: v for k, v in sys.modules.items() if k == "jiwer" or k.split(".")[0] in ("whisper", "speech_disorder")}
    sys.path.insert(0, tree)
    sys.modules.setdefault("jiwer", types.ModuleType("jiwer"))
    import speech_disorder.dataset as dataset_module
    import whisper
    import whisper.transcribe  # noqa: F401  (whisper/__init__.py rebinds the name `transcribe` to the function)

    transcribe_module = sys.modules["whisper.transcribe"]
    yield types.SimpleNamespace(whisper=whisper, dataset=dataset_module, transcribe=transcribe_module)
    for name in [k for k in sys.modules if k == "jiwer" or k.split(".")[0] in ("whisper", "speech_disorder")]:
        del sys.modules[name]
    sys.modules.update(saved_modules)
    sys.path[:] = saved_path


CLIPS = {  # "path" -> (kind, samples, seed): 1-30 s clips and one that pad_or_trim has to cut
    "a.wav": ("gauss", 16000 * 3 + 17, 1), "b.wav": ("pcm16", 16000 * 11, 2), "c.wav": ("chirp", 16000 * 30, 3),
    "d.wav": ("gauss", 16000 * 33, 4), "e.wav": ("burst", 16000 * 7 + 5, 5), "f.wav": ("two_tone", 16000 * 19, 6),
}


def _fake_load_audio(path, sr=16000):
    kind, n, seed = CLIPS[os.path.basename(path)]
    return signals.make_signal(kind, n, seed)


def _build_batch(ref, tmp_path):
    import pandas as pd

    csv = tmp_path / "train.csv"
    pd.DataFrame({"file": list(CLIPS), "text": [f"utterance {i}" for i in range(len(CLIPS))], "class": [i % 3 for i in range(len(CLIPS))]}).to_csv(csv, index=False)
    config = types.SimpleNamespace(model_size="tiny", class_to_disease={0: "normal", 1: "dysphonia", 2: "dysarthria"})
    ds = ref.dataset.MultiTaskSpeechDataset(str(csv), config)
    batch = ds.get_collate_fn()([ds[i] for i in range(len(ds))])
    return batch


def test_install_rebinds_the_real_consumer_modules(reference_packages):
    ref = reference_packages
    import asr_ttl_mtl_b200 as b200
    from asr_ttl_mtl_b200 import audio as ours

    original = ref.dataset.log_mel_spectrogram
    assert original is ref.whisper.audio.log_mel_spectrogram
    rebound = b200.install()
    try:
        for module in (ref.dataset, ref.transcribe, ref.whisper, ref.whisper.audio):
            assert module.log_mel_spectrogram is ours.log_mel_spectrogram and module.pad_or_trim is ours.pad_or_trim
        assert set(rebound) >= {"whisper.audio", "whisper", "whisper.transcribe", "speech_disorder.dataset"}
        # the constants the consumers import by name are the same numbers
        for name in ("SAMPLE_RATE", "N_FFT", "HOP_LENGTH", "N_SAMPLES", "N_FRAMES", "FRAMES_PER_SECOND", "N_SAMPLES_PER_TOKEN", "TOKENS_PER_SECOND", "CHUNK_LENGTH"):
            assert getattr(ours, name) == getattr(ref.whisper.audio, name)
    finally:
        b200.uninstall()
    assert ref.dataset.log_mel_spectrogram is original and ref.transcribe.log_mel_spectrogram is original


def test_reference_dataset_batch_is_what_the_oracle_says(reference_packages, tmp_path, monkeypatch):
    """The fixture itself: the unmodified consumer with the unmodified front-end gives the oracle's numbers."""
    from oracle import logmel_oracle as orc

    ref = reference_packages
    monkeypatch.setattr(ref.dataset, "load_audio", _fake_load_audio)
    batch = _build_batch(ref, tmp_path)
    assert tuple(batch["mels"].shape) == (len(CLIPS), 80, 3000) and batch["mels"].dtype == torch.float32
    for i, name in enumerate(CLIPS):
        x = orc.pad_or_trim_oracle(_fake_load_audio(name), 480000)
        assert float((batch["mels"][i] - orc.logmel_f32_port(x, 80)).abs().max()) <= 1e-6


@pytest.mark.gpu
def test_dataset_and_collate_through_the_rebinding(reference_packages, tmp_path, monkeypatch, b200):
    ref = reference_packages
    monkeypatch.setattr(ref.dataset, "load_audio", _fake_load_audio)
    want = _build_batch(ref, tmp_path)
    launches = b200.gpu_launches()
    b200.install()
    try:
        got = _build_batch(ref, tmp_path)
    finally:
        b200.uninstall()
    assert b200.gpu_launches() >= launches + len(CLIPS)          # every sample went through the CUDA library
    assert tuple(got["mels"].shape) == (len(CLIPS), 80, 3000) and got["mels"].dtype == torch.float32
    assert got["mels"].device == want["mels"].device                # numpy in -> CPU tensor out, as the reference
    assert float(got["mels"].abs().max()) > 0                       # (dataset.py:93-96 would hide a failure behind zeros)
    assert float((got["mels"] - want["mels"]).abs().max()) <= TOL
    for key in ("input_tokens", "target_tokens", "classes"):
        assert torch.equal(got[key], want[key])


class _StubModel:
    """What transcribe() touches of a Whisper model, with a decoder that hears nothing: every window is skipped as
    silence, so the seek loop walks the whole file 3000 frames at a time and we see every mel window it builds."""

    def __init__(self, whisper, n_mels):
        self.dims = types.SimpleNamespace(n_mels=n_mels, n_audio_ctx=1500, n_text_ctx=448, n_vocab=51865)
        self.device = torch.device("cpu")
        self.is_multilingual = True
        self.num_languages = 99
        self.windows = []
        self._result = whisper.decoding.DecodingResult

    def detect_language(self, mel):
        self.windows.append(("language", mel.clone()))
        return None, {"en": 1.0}

    def decode(self, mel, options):
        self.windows.append(("decode", mel.clone()))
        return self._result(audio_features=None, language="en", tokens=[], text="", avg_logprob=-5.0, no_speech_prob=1.0,
                            temperature=0.0, compression_ratio=1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("n_mels", [80, 128])
def test_transcribe_mel_stage_through_the_rebinding(reference_packages, b200, n_mels):
    ref = reference_packages
    audio = np.concatenate([signals.make_signal("gauss", 16000 * 41 + 123, 31), signals.make_signal("chirp", 16000 * 30, 32)])   # 71 s: 3 windows

    def run():
        model = _StubModel(ref.whisper, n_mels)
        result = ref.transcribe.transcribe(model, audio, verbose=None, fp16=False, condition_on_previous_text=False)
        return model.windows, result

    want, want_result = run()
    b200.install()
    try:
        got, got_result = run()
    finally:
        b200.uninstall()
    assert [k for k, _ in got] == [k for k, _ in want] and len(got) >= 4     # language window + 3 seek windows
    assert got_result["language"] == want_result["language"] == "en"
    for (kind, g), (_, w) in zip(got, want):
        assert g.shape == w.shape == (n_mels, 3000) and g.dtype == w.dtype and g.device == w.device
        assert float((g - w).abs().max()) <= TOL, kind
