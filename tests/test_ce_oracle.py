"""The conditional-entropy oracle (oracle/ce_numpy.py; the CUDA path is planned) against an
independent evaluation with np.histogram2d, and its behaviour on a known signal."""
import numpy as np
import pytest

from oracle import ce_numpy


@pytest.mark.parametrize("nphi,nm", [(10, 5), (8, 4), (16, 3)])
def test_oracle_matches_histogram2d_definition(nphi, nm):
    rng = np.random.default_rng(nphi * nm)
    t = np.sort(rng.uniform(0, 80, 1500))
    x = 12 + 0.4 * np.sin(2 * np.pi * t / 2.9) + 0.1 * rng.standard_normal(t.size)
    for period in (2.9, 1.7, 5.8, 9.31):
        phi = (t / period) % 1
        m = (x - x.min()) / (x.max() - x.min())
        # half-open cells; the sample with the maximum magnitude belongs to the last magnitude bin
        H, _, _ = np.histogram2d(phi, np.minimum(m, np.nextafter(1.0, 0)), bins=[nphi, nm], range=[[0, 1], [0, 1]])
        p = H / H.sum()
        pphi = p.sum(axis=1, keepdims=True) * np.ones_like(p)
        ok = p > 0
        want = np.sum(p[ok] * np.log(pphi[ok] / p[ok]))
        got = ce_numpy.ce_theta(t, ce_numpy.magnitude_bins(x, nm), period, nphi, nm)
        assert got == pytest.approx(want, rel=1e-12)


def test_minimum_at_the_injected_period_and_bounds():
    rng = np.random.default_rng(4)
    t = np.sort(rng.uniform(0, 60, 1200))
    x = np.sin(2 * np.pi * t / 3.3) + 0.15 * rng.standard_normal(t.size)
    periods = np.linspace(1.0, 6.0, 501)
    h = ce_numpy.ce(t, x, periods, 10, 5)
    assert abs(periods[np.argmin(h)] - 3.3) < 0.02
    assert np.all(h >= 0) and np.all(h <= np.log(5) + 1e-12)       # 0 <= H(m | phi) <= ln(nm)
