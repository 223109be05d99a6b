"""CPU check of the fixed-point phase -> bin fast path of pdm_hist_kernel (pdm_bin_fast_g in
periodicity_b200/csrc/pdm.cu), emulated with exact rational arithmetic (a correctly rounded FMA) and compared with the
reference's binning of phase.py:131,138-140: phi = (t / P) % 1, bin k if k/m0 <= phi < (k+1)/m0.

Claim under test (pdm.cu): the fast bin can differ from the reference's only for samples whose position
inside the bin, shifted up by the guard, is below 2 * guard -- exactly the samples the kernel re-bins with its exact
path.  Everything else must agree, including samples very close to (but not flagged at) a bin edge.
"""
import struct
from fractions import Fraction

import numpy as np
import pytest

GUARD = 4                                           # PDM_FAST_GUARD, units of 2^-32 turn
MAGIC_G = 1572864.0 + GUARD / 4294967296.0          # PDM_FAST_MAGIC_G = 1.5 * 2^20 + guard * 2^-32
LIMIT = 262144.0                                    # PDM_FAST_LIMIT: |t / P| < 2^18


def fma(a, b, c):
    """Correctly rounded a * b + c (what DFMA computes): exact in rationals, one rounding to double."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def fast_bin(t, rP, m0):
    v = fma(t, rP, MAGIC_G)
    lo = struct.unpack("<Q", struct.pack("<d", v))[0] & 0xFFFFFFFF      # __double2loint
    w = lo * m0                                                          # 32 x 32 -> 64 bit product
    return w >> 32, w & 0xFFFFFFFF                                       # bin, position inside the bin (+ guard)


def reference_bins(t, P, m0):
    phi = (t / P) % 1
    thr = np.arange(m0 + 1) / m0
    return np.minimum(np.searchsorted(thr, phi, side="right") - 1, m0 - 1), phi


def check(t, P, m0):
    rP = 1.0 / P
    assert abs(rP) * np.max(np.abs(t)) < LIMIT
    ref, phi = reference_bins(t, P, m0)
    flagged = 0
    for ti, ki, ph in zip(t, ref, phi):
        kf, pos = fast_bin(float(ti), rP, m0)
        assert 0 <= kf < m0                          # always a valid column: the kernel issues the update unconditionally
        if pos < 2 * GUARD * m0:
            flagged += 1                             # the kernel re-bins these with the exact path
        elif ph < 1.0:                               # phi == 1.0 (tiny negative t / P) is a documented deviation
            assert kf == ki, (ti, P, m0, kf, ki, pos)
    return flagged


@pytest.mark.parametrize("m0", [10, 20, 7, 64])
def test_random_samples_agree_unless_flagged(m0):
    rng = np.random.default_rng(m0)
    t = np.concatenate([rng.uniform(0, 1000, 1500), rng.uniform(-500, 0, 500)])
    for P in (3.7, 1.0, 0.0123, 10.999, -2.5):
        check(t, P, m0)


@pytest.mark.parametrize("m0", [10, 20, 12])
def test_samples_on_and_next_to_bin_edges(m0):
    # integer times with rational periods put samples exactly on edges; offsets down to one ulp probe both sides
    t = np.arange(0.0, 400.0)
    for P in (2.0, 2.5, 4.0, 5.0, 8.0, 12.5, 40.0):
        assert check(t, P, m0) > 0                  # the exact edges are flagged
    rng = np.random.default_rng(3)
    P = 3.7
    k = rng.integers(0, m0, 400)
    cyc = rng.integers(0, 200, 400)
    base = (cyc + k / m0) * P
    for delta in (0.0, 1e-13, -1e-13, 1e-11, -1e-11, 1e-9, -1e-9, 2e-8, -2e-8):
        check(base * (1 + delta) if delta else base, P, m0)
    for steps in (1, -1, 3, -3):
        check(np.nextafter(base, np.inf * steps) if abs(steps) == 1 else base + steps * np.spacing(base), P, m0)


def test_flagged_fraction_matches_the_stated_rate():
    # about 2 * guard * m0 / 2^32 of random samples are flagged: none among 4000 random draws
    rng = np.random.default_rng(9)
    assert check(rng.uniform(0, 1000, 4000), 3.7, 20) == 0
