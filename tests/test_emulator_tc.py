"""CPU check of the tcgen05 variant's math, operand layouts and bin maps (tests/emul/emul_tc.cpp runs
the stage-1 / epilogue functions of csrc/tc_core.cuh; the MMAs are replaced by exact fp16 products)."""
import ctypes

import numpy as np
import pytest

from oracle import logmel_oracle as orc
from oracle import signals

TOL = 1e-4


def test_folded_dft_of_one_frame(emul_tc):
    """Folds, slot maps and matrices: the four products give Re X[k] and |Im X[k]| of the Hann-windowed frame."""
    rng = np.random.default_rng(1)
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    n = np.arange(400)
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * n / 400)
    for trial in range(6):
        x = rng.standard_normal(400).astype(np.float32)
        if trial == 1:
            x[:] = 0; x[200] = 1.0          # centre tap only
        if trial == 2:
            x[:] = 0; x[100] = 1.0; x[300] = -0.5
        re, im = np.zeros(200), np.zeros(200)
        assert emul_tc.emul_tc_frame_spectrum(x.ctypes.data_as(fp), re.ctypes.data_as(dp), im.ctypes.data_as(dp)) == 0
        X = np.fft.rfft(hann * x.astype(np.float64))[:200]
        scale = np.abs(X).max() + 1e-30
        assert np.abs(re - X.real).max() / scale < 2e-6, trial
        assert np.abs(im - np.abs(X.imag)).max() / scale < 2e-6, trial


def test_mel_band_header_is_current():
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "tools", "gen_mel_bands.py"), "--check"]).returncode == 0


def test_emulated_tc_variant_matches_golden_cases(emul_tc, golden):
    for c in golden.cases:
        if c["n"] > 100000 and c["kind"] != "chirp":
            continue  # the emulated MMAs are slow; one 30 s clip (the hard one) is enough here
        got = emul_tc.run(golden.signal(c), c["n_mels"], golden[f"filters_{c['n_mels']}"], padding=c["padding"])
        err = float(np.abs(got - golden.out(c)).max())
        assert got.shape == tuple(c["shape"]) and err <= TOL, (c, err)


def test_compile_time_fold_equals_table_driven_fold(emul_tc, golden):
    """The unrolled chunk the kernel runs (tc_sweep_chunk_ct) and the table-driven one agree bit for bit."""
    x = signals.make_signal("gauss", 16000, 5)
    emul_tc.run(x, 80, golden["filters_80"])
    emul_tc.run(signals.make_signal("chirp", 16000, 6), 128, golden["filters_128"])
    assert emul_tc.emul_tc_chunk_mismatch() == 0


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
