"""pdc_gl (Gregory-Loredo periodogram for event arrival times) on the GPU against the numpy oracle (oracle/gl_numpy.py,
pinned to the definition with exact factorials in tests/test_gl_oracle.py).

PARITY UNPINNED BY THE REFERENCE: the method is a TODO there (phase.py:14).  The kernel histograms integer COUNTS; the
rest is lgamma / logsumexp in float64 whose terms reach N ln N, so the comparison is absolute: |d ln O| <= 1e-7 +
1e-10 |ln O|, identical arg-max."""
import numpy as np
import pytest

from oracle import gl_numpy
from test_gl_oracle import events

pytestmark = pytest.mark.gpu


def assert_close(got, ref):
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.max(np.abs(got[ok] - ref[ok]) - 1e-10 * np.abs(ref[ok])) <= 1e-7
    assert np.nanargmax(got) == np.nanargmax(ref)


@pytest.mark.parametrize("n,m_max,nc", [(1500, 8, 5), (300, 12, 10), (6000, 6, 4), (40, 3, 1), (2500, 2, 7)])
def test_against_oracle(gpu_ctx, n, m_max, nc):
    t = events(n, 3.3, n + m_max)
    periods = np.linspace(1.0, 8.0, 301)
    lo, am, mx = gpu_ctx.gl(t, periods, m_max, nc)
    ref = gl_numpy.gl(t, periods, m_max, nc)
    assert_close(lo, ref)
    assert am == np.nanargmax(ref) and mx == lo[am]
    if n >= 1500:
        assert abs(periods[am] - 3.3) < 0.05 and mx > 10


def test_large_event_list_uses_sample_splits_and_is_reproducible(gpu_ctx):
    t = events(60_000, 2.9, 7, T=4000.0)
    periods = np.linspace(1.0, 6.0, 900)
    lo, am, mx = gpu_ctx.gl(t, periods, 6, 4)
    sel = np.unique(np.concatenate([np.arange(0, 900, 45), np.arange(am - 2, am + 3)]))
    ref = gl_numpy.gl(t, periods[sel], 6, 4)
    assert np.max(np.abs(lo[sel] - ref) - 1e-10 * np.abs(ref)) <= 1e-6
    # (with many bins available the first sub-harmonic 2P describes the same light curve and may win)
    assert sel[np.argmax(ref)] == am and (abs(periods[am] - 2.9) < 0.01 or abs(periods[am] - 5.8) < 0.02)
    lo2, am2, mx2 = gpu_ctx.gl(t, periods, 6, 4)
    np.testing.assert_array_equal(lo, lo2)
    assert (am, mx) == (am2, mx2)
    # the count plane is shared with PDM / CE: a PDM call in between must find and leave it clean
    x = np.sin(2 * np.pi * t / 2.9)
    th, _, _ = gpu_ctx.pdm(t, x, periods, 10, 2)
    lo3, _, _ = gpu_ctx.gl(t, periods, 6, 4)
    np.testing.assert_array_equal(lo, lo3)
    th2, _, _ = gpu_ctx.pdm(t, x, periods, 10, 2)
    np.testing.assert_array_equal(th, th2)


def test_class_degenerate_periods_and_invalid_arguments(gpu_ctx):
    from periodicity_b200 import GL, TSeries
    t = events(2000, 4.1, 9)
    gl = GL(m_max=6, nc=5, p_min=1.5, p_max=9.0, n_periods=400)
    pg = gl(TSeries(t, np.ones_like(t)))
    ref = gl_numpy.gl(t, gl.periods, 6, 5)
    assert np.max(np.abs(pg.values - ref[::-1])) <= 1e-7
    assert abs(1 / pg.frequency[pg.argmax()] - 4.1) < 0.05
    pg2 = GL(m_max=6, nc=5, p_min=1.5, p_max=9.0, n_periods=400, devices=[0, 0])(t[::-1])   # plain array, unsorted
    np.testing.assert_array_equal(pg2.values, pg.values)
    lo, am, _ = gpu_ctx.gl(t, np.array([0.0, 2.0, np.inf, np.nan, 4.1]), 4, 3)
    assert np.isnan(lo[[0, 2, 3]]).all() and np.isfinite(lo[[1, 4]]).all() and am == 4
    for bad in ((1, 3), (4, 0), (500, 10)):
        with pytest.raises(ValueError):
            gpu_ctx.gl(t, [2.0], *bad)
