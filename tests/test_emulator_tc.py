"""CPU check of the tcgen05 variant's math, operand layouts and bin maps (tests/emul/emul_tc.cpp runs
the stage-1 / epilogue functions of csrc/tc_core.cuh; the MMAs are replaced by exact fp16 products)."""
import ctypes

import numpy as np
import pytest

from oracle import logmel_oracle as orc
from oracle import signals

TOL = 1e-4


def test_real_fft16(emul_tc):
    rng = np.random.default_rng(1)
    fp = ctypes.POINTER(ctypes.c_float)
    for _ in range(20):
        x = rng.standard_normal(16).astype(np.float32)
        out = np.zeros(18, np.float32)
        emul_tc.emul_fft16_real_x2(x.ctypes.data_as(fp), out.ctypes.data_as(fp))
        assert np.abs((out[0::2] + 1j * out[1::2]) - 2 * np.fft.rfft(x.astype(np.float64))).max() < 5e-6


def test_emulated_tc_variant_matches_golden_cases(emul_tc, golden):
    for c in golden.cases:
        if c["n"] > 100000 and c["kind"] != "chirp":
            continue  # the emulated MMAs are slow; one 30 s clip (the hard one) is enough here
        got = emul_tc.run(golden.signal(c), c["n_mels"], golden[f"filters_{c['n_mels']}"], padding=c["padding"])
        err = float(np.abs(got - golden.out(c)).max())
        assert got.shape == tuple(c["shape"]) and err <= TOL, (c, err)


@pytest.mark.parametrize("kind", ["chirp", "two_tone", "sine1k_noise"])
def test_split_precision_stays_close_to_the_f64_spec(emul_tc, golden, kind):
    x = signals.make_signal(kind, 32000, 77)
    f64 = orc.logmel_f64(x, 80)
    ref = orc.logmel_f32_port(x, 80).numpy()
    got = emul_tc.run(x, 80, golden["filters_80"])
    assert np.abs(got - f64).max() <= np.abs(ref - f64).max() + 5e-5


def test_quiet_and_loud_inputs_survive_the_fp16_operands(emul_tc, golden):
    base = signals.make_signal("gauss", 16000, 9)
    for scale in (1e-4, 1e-2, 1.0, 8.0):
        x = (base * scale).astype(np.float32)
        got = emul_tc.run(x, 80, golden["filters_80"])
        assert np.abs(got - orc.logmel_f32_port(x, 80).numpy()).max() <= TOL, scale
