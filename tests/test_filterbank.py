"""The regenerated Slaney filterbank is bit-equal to the reference asset (audio.py:91-107)."""
import numpy as np
import pytest

from asr_ttl_mtl_b200 import filterbank


@pytest.mark.parametrize("n_mels", [80, 128])
def test_filterbank_bit_equal_to_reference_asset(golden, n_mels):
    ours = filterbank.slaney_mel_filterbank(n_mels)
    ref = golden[f"filters_{n_mels}"]
    assert ours.dtype == np.float32 and ours.shape == (n_mels, 201)
    assert np.array_equal(ours, ref)
    assert filterbank.filter_digest(ours) == filterbank.FILTER_SHA256[n_mels]


@pytest.mark.parametrize("n_mels", [80, 128])
def test_filterbank_structure_the_kernel_relies_on(n_mels):
    w = filterbank.slaney_mel_filterbank(n_mels)
    assert not w[:, 0].any() and not w[:, 200].any()  # DC and Nyquist carry no weight
    total = 0
    for row in w:
        nz = np.flatnonzero(row)
        assert nz.size > 0 and np.array_equal(nz, np.arange(nz[0], nz[-1] + 1))  # one contiguous band
        total += nz.size
    assert total <= 512 and (w != 0).sum(axis=0).max() <= 2
