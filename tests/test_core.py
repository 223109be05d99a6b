"""Container semantics the hot path relies on (reference tests/test_core.py:7-34 and core.py:859-955)."""
import numpy as np
import pytest

from periodicity_b200.core import FSeries, TSeries


def test_time_array_is_always_sorted():
    sig = TSeries([3, 2, 1], [3, 5, 7])
    assert all(sig.time == [1, 2, 3])
    assert all(sig.values == [7, 5, 3])


def test_input_arrays_with_different_sizes():
    with pytest.raises(ValueError):
        _ = TSeries([1, 2], [1, 2, 3])


def test_dt_of_nonuniform_samples():
    sig = TSeries([1, 3, 4], [1, 1, 1])
    assert sig.median_dt == 1.5
    with pytest.raises(AttributeError):
        _ = sig.dt


def test_baseline():
    assert TSeries(np.arange(10)).baseline == 9


def test_nonuniform_slice_of_uniform_signal():
    sig = TSeries(np.arange(10))
    assert sig.dt == 1.0
    sig_slice = sig[[2, 5, 6]]
    with pytest.raises(AttributeError):
        _ = sig_slice.dt


def test_defaults_and_len():
    s = TSeries(values=[4.0, 5.0, 6.0])
    assert list(s.time) == [0, 1, 2] and len(s) == 3 and s.size == 3
    s2 = TSeries([0.0, 0.5])
    assert list(s2.values) == [1.0, 1.0]


def test_scalar_arithmetic_keeps_time():
    s = TSeries([0.0, 1.0, 2.5], [1.0, 2.0, 3.0])
    w = 0.0 * s + 1.0                      # GLS.window(), spectral.py:165-167
    assert isinstance(w, TSeries)
    assert list(w.values) == [1.0, 1.0, 1.0] and list(w.time) == [0.0, 1.0, 2.5]
    d = s - s.mean()
    assert abs(d.values.sum()) < 1e-12
    assert np.sum(s) == 6.0


def test_copy_is_independent():
    s = TSeries([0.0, 1.0], [1.0, 2.0])
    c = s.copy()
    c.values = np.array([9.0, 9.0])
    assert list(s.values) == [1.0, 2.0]


def test_fseries_sorts_and_exposes_period():
    f = FSeries([0.5, 0.25, 0.125], [1.0, 2.0, 3.0])        # PDM passes 1/periods, descending
    assert list(f.frequency) == [0.125, 0.25, 0.5]
    assert list(f.values) == [3.0, 2.0, 1.0]
    assert list(f.period) == [8.0, 4.0, 2.0]
    assert f.fmax() == 0.125 and f.pmax() == 8.0


def test_fseries_nan_aware_reductions_and_peaks():
    f = FSeries(np.arange(1, 8) / 10.0, [0.0, 1.0, np.nan, 0.5, 3.0, 0.2, 0.1])
    assert f.argmax() == 4 and f.amax() == 3.0
    g = FSeries(np.arange(1, 8) / 10.0, [0.0, 1.0, 0.3, 0.5, 3.0, 0.2, 0.1])
    assert g.period_at_highest_peak == 1.0 / 0.5
    peaks = g.find_peaks()
    assert list(peaks.attrs["indices"]) == [1, 4]
    with pytest.raises(ValueError):
        FSeries([1.0, 2.0], [1.0])


def test_periods_at_half_max_and_zero_crossings():
    """core.py:341-367,957-972 with a scalar half-maximum level (see the method's docstring)."""
    f = np.linspace(0.01, 2.0, 4000)
    v = np.exp(-0.5 * ((f - 0.7) / 0.01) ** 2) + 0.6 * np.exp(-0.5 * ((f - 1.3) / 0.02) ** 2)
    pg = FSeries(f, v)
    lower, upper = pg.periods_at_half_max()
    fwhm = 2 * np.sqrt(2 * np.log(2)) * 0.01
    assert lower < 1 / 0.7 < upper
    assert abs((1 / lower - 1 / upper) - fwhm) < 3 * (f[1] - f[0])
    lo2, up2 = pg.periods_at_half_max(peak_order=2)
    assert lo2 < 1 / 1.3 < up2
    lo3, up3 = pg.periods_at_half_max(peak_order=1, use_prominence=True)
    assert (lo3, up3) == (lower, upper)                      # isolated peak on a zero floor: prominence == height
    zc = (pg - 0.5).find_zero_crossings()
    assert zc.size == 4 and np.all(np.diff(np.signbit(v - 0.5))[zc])
