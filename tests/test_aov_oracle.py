"""The AoV oracle (oracle/aov_numpy.py) against an independent implementation of the statistic:
scipy.stats.f_oneway of the values grouped by phase bin.  The reference has no AoV code (phase.py:11 is a TODO),
so this is the pin; the drop-in class is covered for its host logic with a stub context."""
import numpy as np
import pytest
from scipy import stats

from oracle import aov_numpy


def _groups(t, x, period, nb):
    phi = (t / period) % 1
    k = np.minimum((phi * nb).astype(int), nb - 1)
    return [x[k == b] for b in range(nb) if np.any(k == b)]


@pytest.mark.parametrize("n,nb", [(400, 5), (3000, 10), (57, 8)])
def test_oracle_is_the_one_way_anova_f_ratio(n, nb):
    rng = np.random.default_rng(n + nb)
    t = np.sort(rng.uniform(0, 100, n))
    x = 3 + np.sin(2 * np.pi * t / 3.7) + 0.5 * rng.standard_normal(n)
    for period in (3.7, 1.234, 7.4, 11.0):
        want = stats.f_oneway(*_groups(t, x, period, nb)).statistic
        got = aov_numpy.aov_theta(t, x, period, nb)
        assert got == pytest.approx(want, rel=1e-10)


def test_oracle_peaks_at_the_injected_period_and_handles_empty_bins():
    rng = np.random.default_rng(5)
    t = np.sort(rng.uniform(0, 60, 900))
    x = np.sin(2 * np.pi * t / 2.5) + 0.3 * rng.standard_normal(t.size)
    periods = np.linspace(1.0, 6.0, 501)
    theta = aov_numpy.aov(t, x, periods, 10)
    assert abs(periods[np.nanargmax(theta)] - 2.5) < 0.02
    # integer times, period 4, 8 bins: only bins 0, 2, 4, 6 are populated (r = 4)
    ti = np.arange(64.0)
    xi = rng.standard_normal(64)
    want = stats.f_oneway(*[xi[ti % 4 == k] for k in range(4)]).statistic
    assert aov_numpy.aov_theta(ti, xi, 4.0, 8) == pytest.approx(want, rel=1e-12)
    assert np.isnan(aov_numpy.aov_theta(ti, xi, 1.0, 8))          # every sample in one bin


def test_dropin_class_builds_the_pdm_period_grid(monkeypatch):
    """AOV follows PDM's grid conventions (phase.py:167-180) and wraps the result like PDM (FSeries over 1/P)."""
    from periodicity_b200 import AOV, TSeries, _ffi

    class Stub:
        def aov(self, t, x, periods, nb):
            theta = aov_numpy.aov(t, x, periods, nb)
            return theta, int(np.nanargmax(theta)), float(np.nanmax(theta))

    monkeypatch.setattr(_ffi, "default_context", lambda device=None: Stub())
    rng = np.random.default_rng(9)
    t = np.sort(rng.uniform(0, 50, 500))
    x = np.sin(2 * np.pi * t / 4.2) + 0.2 * rng.standard_normal(500)
    sig = TSeries(t, x)
    aov = AOV(nb=8, n_periods=300)
    out = aov(sig)
    np.testing.assert_array_equal(aov.periods, np.linspace(2 * sig.median_dt, sig.baseline, 300))
    np.testing.assert_array_equal(out.frequency, (1 / aov.periods)[::-1])        # FSeries sorts ascending in f
    np.testing.assert_array_equal(out.values, aov_numpy.aov(t, x, aov.periods, 8)[::-1])
    assert aov.argmax_index == int(np.nanargmax(aov_numpy.aov(t, x, aov.periods, 8)))
    aov2 = AOV(nb=8, p_min=3.0, p_max=6.0, n_periods=400)
    aov2(sig)
    assert abs(aov2.periods[aov2.argmax_index] - 4.2) < 0.05
