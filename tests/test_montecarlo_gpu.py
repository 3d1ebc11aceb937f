"""the Monte-Carlo base-power table regenerated on the GPU must reproduce entries computed
by the unmodified reference procedure (monteCarlo/monteCarlo.py:133-201; fixtures from
oracle/make_mc_golden.py), and the grid-point identity of monteCarlo/unit_tests_interp.py:74-98 must hold
for a table built this way."""
import os

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


def test_regenerated_entries_match_the_reference_procedure():
    import mdr_b200
    z = np.load(os.path.join(gu.GOLDEN_DIR, "mc_table_samples.npz"))
    idx, ref = z["idx"].astype(np.int64), z["values"]
    got = mdr_b200.regenerate_entries(idx, precision="fp64")
    # entries are means of integer-valued powers: agreement is exact unless a bang-bang comparison sits
    # within 1e-13 K of the threshold (none of the sampled combinations does)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-9)
    assert (got > 0).sum() == (ref > 0).sum()


def test_fp32_entries_are_close_and_table_slice_is_consistent():
    import mdr_b200
    from oracle import mdr_oracle as orc
    z = np.load(os.path.join(gu.GOLDEN_DIR, "mc_table_samples.npz"))
    idx, ref = z["idx"].astype(np.int64), z["values"]
    got32 = mdr_b200.regenerate_entries(idx, precision="fp32")
    # a flipped bang-bang decision changes an entry by a multiple of cap/COP/75; allow a few of them
    close = np.isclose(got32, ref, rtol=0, atol=1e-6)
    assert close.mean() > 0.9
    assert np.abs(got32 - ref).max() < 6000.0 * 3 / 75 + 1e-6
    # a full (thermal, HVAC, hour, date) slice: 9 x 5 x 8 entries; monotone in outdoor temperature on average
    shape = mdr_b200.montecarlo.grid_shape()
    sub = np.stack(np.meshgrid([1], [1], [1], [1], range(9), range(5), range(8), [1], [0], [0], indexing="ij"), -1).reshape(-1, 10)
    vals = mdr_b200.regenerate_entries(sub, precision="fp64").reshape(9, 5, 8)
    assert vals.min() >= 0 and vals.max() <= 6000.0
    assert vals.mean(axis=(0, 1))[-1] > vals.mean(axis=(0, 1))[0]  # hotter outside -> more cooling power
    assert int(np.prod(shape)) == gu.TABLE_SIZE
