"""CPU check of the kernel's math and choreography (tests/emul/emul_fft.cpp runs the phase
functions of csrc/logmel_core.cuh with the kernel's tile geometry and barrier placement)."""
import ctypes

import numpy as np
import pytest

from oracle import logmel_oracle as orc
from oracle import signals

TOL = 1e-4  # max-abs on the normalised log-mel, fp32


def test_dft20_butterfly(emul):
    rng = np.random.default_rng(0)
    fp = ctypes.POINTER(ctypes.c_float)
    for _ in range(20):
        x = (rng.standard_normal(20) + 1j * rng.standard_normal(20)).astype(np.complex64)
        out = np.zeros(20, np.complex64)
        emul.emul_dft20(x.view(np.float32).ctypes.data_as(fp), out.view(np.float32).ctypes.data_as(fp))
        assert np.abs(out - np.fft.fft(x.astype(np.complex128))).max() < 5e-6


def test_emulated_kernel_matches_every_golden_case(emul, golden):
    for c in golden.cases:
        got = emul.run(golden.signal(c), c["n_mels"], golden[f"filters_{c['n_mels']}"], padding=c["padding"])
        err = float(np.abs(got - golden.out(c)).max())
        assert got.shape == tuple(c["shape"]) and err <= TOL, (c, err)


@pytest.mark.parametrize("kind", ["chirp", "two_tone", "sine1k_noise"])
def test_adversarial_signals_stay_close_to_the_f64_spec(emul, golden, kind):
    # |ours - f64| <= |ref - f64| + 5e-5 on high-dynamic-range signals
    x = signals.make_signal(kind, 32000, 77)
    f64 = orc.logmel_f64(x, 80)
    ref = orc.logmel_f32_port(x, 80).numpy()
    got = emul.run(x, 80, golden["filters_80"])
    assert np.abs(got - f64).max() <= np.abs(ref - f64).max() + 5e-5


def test_lengths_semantics_equal_zero_filled_rows(emul, golden):
    x = signals.make_signal("gauss", 20000, 5)
    for valid in (0, 1, 159, 7777, 19999, 20000):
        padded = x.copy()
        padded[valid:] = 0.0
        a = emul.run(x, 80, golden["filters_80"], valid=valid)
        b = emul.run(padded, 80, golden["filters_80"])
        assert np.array_equal(a, b)
        assert np.abs(a - orc.logmel_f32_port(padded, 80).numpy()).max() <= TOL


def test_zero_tail_sits_on_the_clamp(emul, golden):
    x = signals.make_signal("burst", 48000, 1)  # first 8000 samples noise, rest zeros
    out = emul.run(x, 80, golden["filters_80"])
    tail = out[:, 60:]  # frames wholly inside the silence
    assert np.all(tail == tail[0, 0]) and np.isclose(out.max() - tail[0, 0], 2.0, atol=1e-6)


def test_max_key_is_order_preserving_and_nan_wins(emul):
    vals = np.array([-np.inf, -10.0, -1e-3, -0.0, 0.0, 1e-20, 0.5, 3.25, np.inf], dtype=np.float32)
    keys = [emul.emul_key_encode(float(v)) for v in vals]
    assert keys == sorted(keys) and all(k > 0 for k in keys)
    for v, k in zip(vals, keys):
        assert emul.emul_key_decode(k) == v
    nan_key = emul.emul_key_encode(float("nan"))
    assert nan_key == 0xFFFFFFFF and nan_key > max(keys) and np.isnan(emul.emul_key_decode(nan_key))


def test_nan_poisons_the_whole_utterance_like_torch(emul, golden):
    x = signals.make_signal("gauss", 16000, 2)
    x[5000] = np.nan
    assert np.isnan(orc.logmel_f32_port(x, 80).numpy()).all()
    assert np.isnan(emul.run(x, 80, golden["filters_80"])).all()
