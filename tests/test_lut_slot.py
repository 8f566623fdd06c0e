"""LutSlot (csrc/m3t_b200_device.cuh): the posterior lookup table stores bin `idx` at idx ^ ((idx >> 4) & 15) ^ ((idx >> 8) & 15).
The property every writer / reader relies on: a bijection of [0, n_bins^3) for every power-of-two bin count the
reference allows (color_histograms.cpp:131-159: 2 .. 64), and - the reason it exists - neighbouring colours no longer
share a shared-memory bank pair (entry = 8 bytes, bank pair = slot & 15)."""
import numpy as np
import pytest


def lut_slot(idx):
    return idx ^ ((idx >> 4) & 15) ^ ((idx >> 8) & 15)


@pytest.mark.parametrize("n_bins", [2, 4, 8, 16, 32, 64])
def test_lut_slot_is_a_bijection(n_bins):
    idx = np.arange(n_bins ** 3, dtype=np.int64)
    slot = lut_slot(idx)
    assert slot.min() == 0 and slot.max() == n_bins ** 3 - 1
    assert np.array_equal(np.sort(slot), idx)


def test_lut_slot_spreads_a_colour_blob_over_the_bank_pairs():
    n = 16
    b, g, r = np.meshgrid(np.arange(5, 8), np.arange(6, 9), np.arange(9, 12), indexing="ij")   # a 3 x 3 x 3 blob of bins
    idx = ((b * n + g) * n + r).ravel()
    assert len(np.unique(idx & 15)) == 3                 # index order: the red bin alone picks the bank pair
    assert len(np.unique(lut_slot(idx) & 15)) >= 8       # slots: green and blue take part
