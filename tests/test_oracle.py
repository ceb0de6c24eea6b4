"""Pin the oracle to the reference: golden vectors made by running whisper/audio.py (make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import logmel_oracle as orc
from oracle import signals


def _case_ids(golden_cases):
    return [f"{c['idx']}-{c['kind']}-{c['n']}-{c['n_mels']}-p{c['padding']}" for c in golden_cases]


def test_f32_port_reproduces_reference_outputs(golden):
    worst = 0.0
    for c in golden.cases:
        got = orc.logmel_f32_port(golden.signal(c), c["n_mels"], c["padding"]).numpy()
        ref = golden.out(c)
        assert got.shape == ref.shape == tuple(c["shape"])
        # bit-equal in the container that made the fixture; another host CPU may pick another
        # MKL code path, so allow fp32 FFT noise
        worst = max(worst, float(np.abs(got - ref).max()))
    assert worst <= 2e-5, worst


def test_f64_spec_matches_reference_outputs(golden):
    for c in golden.cases:
        got = orc.logmel_f64(golden.signal(c), c["n_mels"], c["padding"])
        err = float(np.abs(got - golden.out(c)).max())
        assert err <= 1e-4, (c, err)


def test_reference_2d_call_shares_one_max(golden):
    scale = golden["batch2d_in_scale"]
    batch = np.stack([signals.make_signal("gauss", 16000, 50) * s for s in scale])
    got = orc.logmel_f32_port(batch, 80).numpy()
    assert np.abs(got - golden["batch2d_out"]).max() <= 2e-5
    per_clip = orc.logmel_f32_port_per_utterance(torch.from_numpy(batch), 80).numpy()
    assert np.abs(per_clip - golden["batch2d_out"]).max() > 0.1  # stacking per-clip calls is NOT the same


def test_frame_count_rule(golden):
    for n, t in zip(golden["frames_len"], golden["frames_T"]):
        assert orc.n_frames_of(int(n)) == int(t)
    for n in (0, 1, 200):
        with pytest.raises(RuntimeError):
            orc.n_frames_of(n)
    assert orc.n_frames_of(100, 480000) == 3000


def test_pad_or_trim_oracle(golden):
    assert np.array_equal(orc.pad_or_trim_oracle(np.arange(5, dtype=np.float32), 8), golden["pot_pad_np"])
    assert np.array_equal(orc.pad_or_trim_oracle(np.arange(10, dtype=np.float32), 4), golden["pot_trim_np"])
    m = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    assert np.array_equal(orc.pad_or_trim_oracle(m, 5, axis=1), golden["pot_axis1_pad_t"])
    assert np.array_equal(orc.pad_or_trim_oracle(m, 1, axis=0), golden["pot_axis0_trim_t"])


def test_silence_is_minus_one_point_five():
    out = orc.logmel_f64(np.zeros(16000, np.float32), 80)
    assert np.all(out == -1.5)


def test_signal_generators_are_deterministic():
    for kind in signals.KINDS:
        a, b = signals.make_signal(kind, 4000, 3), signals.make_signal(kind, 4000, 3)
        assert a.dtype == np.float32 and np.array_equal(a, b)
    q = signals.make_pcm16(4000, 9)
    assert np.array_equal(q.astype(np.float32) / 32768.0, signals.make_signal("pcm16", 4000, 9))
    lens = signals.variable_lengths(64)
    assert lens.min() >= 16000 and lens.max() <= 480000
