"""The Gregory-Loredo oracle (oracle/gl_numpy.py) against a direct evaluation of the definition with np.histogram and
exact integer factorials, and its behaviour on a known periodic event series.  PARITY UNPINNED BY THE REFERENCE
(phase.py:14 is a TODO)."""
import math
from fractions import Fraction

import numpy as np

from oracle import gl_numpy


def events(n, period, seed, contrast=0.8, T=200.0):
    """Arrival times of an inhomogeneous Poisson process with rate ~ 1 + contrast sin(2 pi t / period) (thinning)."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, 4 * n))
    keep = rng.uniform(size=t.size) < (1 + contrast * np.sin(2 * np.pi * t / period)) / (1 + contrast)
    return t[keep][:n]


def test_oracle_matches_the_definition_with_exact_factorials():
    t = events(60, 3.1, 1, T=40.0)
    N = t.size
    for period, m, nc in ((3.1, 2, 4), (3.1, 5, 3), (1.7, 3, 1), (7.7, 4, 2)):
        F = m * nc
        fine, _ = np.histogram((t / period) % 1, bins=F, range=(0, 1))
        tot = Fraction(0)
        for c in range(nc):
            nj = np.roll(fine, -c).reshape(m, nc).sum(axis=1)
            prod = 1
            for v in nj:
                prod *= math.factorial(int(v))
            tot += prod
        want = (math.log(math.factorial(m - 1)) - math.log(math.factorial(N + m - 1)) + N * math.log(m)
                + math.log(tot / nc))
        assert abs(gl_numpy.ln_odds_m(t, period, m, nc) - want) <= 1e-10 * max(1.0, abs(want))


def test_peak_at_the_injected_period():
    t = events(1500, 3.3, 2)
    periods = np.linspace(1.0, 8.0, 701)
    lo = gl_numpy.gl(t, periods, m_max=8, nc=5)
    assert abs(periods[np.argmax(lo)] - 3.3) < 0.03
    assert lo.max() > 10 and np.median(lo) < 0          # decisive at the period, against periodicity elsewhere
