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


def _emulate_float2_path(t, x, periods, nb):
    """What pdc_aov computes, restated on the CPU: x' = (x - mean) / std as float32, per-bin float32 sums in sample
    order merged into float64 every 8192 samples (pdm_hist_kernel's float2 columns), float64 epilogue with
    q = N - 1 (pdm_epilogue_kernel<PDC_STAT_AOV>)."""
    n = t.size
    mean = x.mean()
    sd = np.sqrt(((x - mean) ** 2).sum() / (n - 1))
    xp = ((x - mean) / sd).astype(np.float32)
    thr = np.arange(nb + 1) / nb
    out = []
    for P in periods:
        k = np.minimum(np.searchsorted(thr, (t / P) % 1, side="right") - 1, nb - 1)
        N = np.bincount(k, minlength=nb).astype(float)
        S = np.zeros(nb)
        for s0 in range(0, n, 8192):
            kk, xx = k[s0:s0 + 8192], xp[s0:s0 + 8192]
            for b in range(nb):
                v = xx[kk == b]
                if v.size:
                    S[b] += float(np.cumsum(v, dtype=np.float32)[-1])
        ok = N >= 1
        sq = (S[ok] ** 2 / N[ok]).sum()
        ntot, stot, r = N[ok].sum(), S[ok].sum(), ok.sum()
        out.append((ntot - r) / (r - 1) * (sq - stot ** 2 / ntot) / ((n - 1) - sq))
    return np.array(out)


@pytest.mark.parametrize("n,nb", [(6000, 10), (4500, 4), (20_000, 16)])
def test_fp32_histogram_columns_keep_aov_inside_the_gpu_tolerance(n, nb):
    """The accuracy argument behind tests/test_aov_gpu.py: FP32 bin sums (error proportional to the sum itself)
    keep Theta within 1e-5 of its peak / 2e-5 relative on values >= 1 % of the peak; the packed fixed-point path PDM
    uses (error 2^-q sigma per sample whatever the sum) would not, which is why pdc_aov never takes it."""
    rng = np.random.default_rng(n + nb)
    t = np.sort(rng.uniform(0, 0.05 * n, n))
    x = 1000 + np.sin(2 * np.pi * t / 3.7) + 0.6 * np.sin(4 * np.pi * t / 3.7) + 0.5 * rng.standard_normal(n)
    periods = np.linspace(1.0, 9.0, 60)
    ref = aov_numpy.aov(t, x, periods, nb)
    got = _emulate_float2_path(t, x, periods, nb)
    peak = ref.max()
    big = ref >= 1e-2 * peak
    assert np.max(np.abs(got - ref)) <= 1e-5 * peak
    assert np.max(np.abs(got - ref)[big] / ref[big]) <= 2e-5
    assert np.argmax(got) == np.argmax(ref)
    # the packed path's quantisation for comparison: q = floor(log2(16000 / max|x'|)), sums of rint(x' 2^q) 2^-q
    xp = (x - x.mean()) / x.std(ddof=1)
    q = int(np.floor(np.log2(16000 / np.abs(xp).max())))
    xq = np.rint(xp * 2.0 ** q) / 2.0 ** q
    thr = np.arange(nb + 1) / nb
    worst = 0.0
    for P, r in zip(periods, ref):
        k = np.minimum(np.searchsorted(thr, (t / P) % 1, side="right") - 1, nb - 1)
        N = np.bincount(k, minlength=nb).astype(float)
        S = np.bincount(k, xq, minlength=nb)
        sq = (S ** 2 / N).sum()
        th = (n - nb) / (nb - 1) * (sq - S.sum() ** 2 / n) / ((n - 1) - sq)
        worst = max(worst, abs(th - r) / max(r, 1e-3))
    assert worst > 2e-5          # i.e. the packed path would not meet the tolerance above
