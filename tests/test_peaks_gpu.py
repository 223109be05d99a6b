"""GPU parity tests for the device peak picker (pdc_peaks_topk) vs scipy.signal.find_peaks,
the routine behind the reference's FSeries.find_peaks / period_at_highest_peak (core.py:283-317,952-955)."""
import numpy as np
import pytest
from scipy.signal import find_peaks

pytestmark = pytest.mark.gpu


def reference_topk(v, k):
    peaks, _ = find_peaks(v, prominence=0.0)
    order = np.lexsort((peaks, -v[peaks]))            # height descending, then index ascending
    top = peaks[order][:k]
    idx = np.full(k, -1, dtype=np.int64)
    val = np.full(k, np.nan)
    idx[: top.size] = top
    val[: top.size] = v[top]
    return idx, val


@pytest.mark.parametrize("n", [3, 4, 17, 4095, 4096, 4097, 8193, 100_000])
def test_random_rows_match_scipy(gpu_ctx, n):
    rng = np.random.default_rng(n)
    v = rng.standard_normal((3, n))
    idx, val = gpu_ctx.peaks_topk(v, 7)
    for r in range(3):
        ri, rv = reference_topk(v[r], 7)
        np.testing.assert_array_equal(idx[r], ri)
        np.testing.assert_array_equal(val[r], rv)


def test_plateaus_nans_edges_and_ties(gpu_ctx):
    v = np.array([5.0, 1.0, 3.0, 3.0, 3.0, 2.0, 4.0, 4.0, 1.0, np.nan, 2.0, 1.0, 2.0, 2.0, 2.0, 2.0, 0.0, 9.0])
    idx, val = gpu_ctx.peaks_topk(v, 6)
    ri, rv = reference_topk(v, 6)
    np.testing.assert_array_equal(idx[0], ri)
    np.testing.assert_array_equal(val[0], rv)
    assert 0 not in idx[0] and len(v) - 1 not in idx[0]            # edges are never peaks
    # many equal-height peaks: ties resolve to the lower index first
    t = np.tile([0.0, 1.0], 3000)
    idx, val = gpu_ctx.peaks_topk(t, 5)
    np.testing.assert_array_equal(idx[0], [1, 3, 5, 7, 9])
    # monotone and constant rows have no peaks
    idx, val = gpu_ctx.peaks_topk(np.stack([np.arange(50.0), np.ones(50)]), 3)
    assert (idx == -1).all() and np.isnan(val).all()
    # a plateau crossing a block boundary (4096 samples per block)
    p = np.zeros(9000)
    p[4090:4100] = 1.0
    p[100] = 0.5
    idx, val = gpu_ctx.peaks_topk(p, 2)
    np.testing.assert_array_equal(idx[0], reference_topk(p, 2)[0])


def test_invalid_k(gpu_ctx):
    with pytest.raises(ValueError):
        gpu_ctx.peaks_topk(np.zeros(10), 0)
    with pytest.raises(ValueError):
        gpu_ctx.peaks_topk(np.zeros(10), 65)


def test_gls_top_peaks_and_survey_topk(gpu_ctx):
    from periodicity_b200 import GLS, TSeries
    from periodicity_b200.survey import gls_survey
    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0, 60, 1500))
    y = np.sin(2 * np.pi * t / 2.5) + 0.6 * np.sin(2 * np.pi * t / 7.0) + 0.3 * rng.standard_normal(1500)
    gls = GLS(fmax=2.0)
    ls = gls(TSeries(t, y))
    freq, power = gls.top_peaks(4)
    assert 1.0 / freq[0] == ls.period_at_highest_peak           # reference accessor, core.py:952-955
    want = ls.psort_by_peak()[:4]
    np.testing.assert_array_equal(1.0 / freq, want)
    assert abs(1 / freq[0] - 2.5) < 0.05 and abs(1 / freq[1] - 7.0) < 0.3
    out = gls_survey([TSeries(t, y), TSeries(t[:700], y[:700])], nf=3000, top_k=3, want_power=True)
    for b in range(2):
        ri, rv = reference_topk(out["power"][b], 3)
        np.testing.assert_array_equal(out["peak_index"][b], ri)
        np.testing.assert_array_equal(out["peak_power"][b], rv)
        assert out["peak_index"][b][0] == out["argmax"][b] or out["max_power"][b] >= rv[0]


def _halfmax_numpy(v, p, height=None):
    half = v[p] - (v[p] if height is None else height) / 2
    d = v - half
    ch = np.where(np.diff(np.signbit(d)))[0]
    left = ch[ch <= p - 2]
    right = ch[ch >= p]
    return (left[-1] if left.size else -1), (right[0] if right.size else -1)


def test_halfmax_crossings_match_numpy(gpu_ctx):
    rng = np.random.default_rng(41)
    rows, n, k = 5, 30_000, 12
    v = np.abs(rng.standard_normal((rows, n))).cumsum(axis=1)
    v = np.sin(v / 40.0) ** 2 * rng.uniform(0.5, 2.0, (rows, 1)) + 0.01 * rng.standard_normal((rows, n))
    idx, val = gpu_ctx.peaks_topk(v, k)
    left, right = gpu_ctx.peaks_halfmax(v, idx)
    prom = rng.uniform(0.1, 1.0, idx.shape) * val
    left_p, right_p = gpu_ctx.peaks_halfmax(v, idx, prom)
    for r in range(rows):
        for j in range(k):
            assert (left[r, j], right[r, j]) == _halfmax_numpy(v[r], idx[r, j])
            assert (left_p[r, j], right_p[r, j]) == _halfmax_numpy(v[r], idx[r, j], prom[r, j])
    # peaks next to the borders, missing peaks, no crossing at all
    w = np.array([[0.0, 1.0, 0.9, 0.95, 0.2, 0.1, 3.0, 2.9]])
    pk = np.array([[1, 6, -1, 3]])
    l2, r2 = gpu_ctx.peaks_halfmax(w, pk)
    want = [_halfmax_numpy(w[0], p) if p >= 0 else (-1, -1) for p in pk[0]]
    assert [(a, b) for a, b in zip(l2[0], r2[0])] == want


def test_top_peak_widths_match_container_method():
    from periodicity_b200 import GLS, TSeries
    rng = np.random.default_rng(42)
    t = np.sort(rng.uniform(0, 200, 1500))
    y = np.sin(2 * np.pi * t / 7.3) + 0.5 * np.sin(2 * np.pi * t / 2.9) + 0.2 * rng.standard_normal(1500)
    gls = GLS(fmax=1.0)
    pg = gls(TSeries(t, y))
    freq, power, lower, upper = gls.top_peak_widths(4)
    for j in range(4):
        lo, up = pg.periods_at_half_max(peak_order=j + 1)
        assert (lower[j], upper[j]) == (lo, up)
    assert abs(1 / freq[0] - 7.3) < 0.1 and lower[0] < 7.3 < upper[0]
