"""pdc_gls_freqs: GLS power at arbitrary (non-uniform, user-supplied) frequency lists, against the formula oracle
(oracle/oracle.c::orc_gls_exact_freqs -- the same code as the grid oracle the golden vectors pin) and against
pdc_gls on a uniform list.  Same tolerances as tests/test_gls_gpu.py: <= 1e-5 of the peak, <= 1e-5 relative where the
power is >= 1 % of the peak, identical arg-max."""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu

TOL = 1e-5


def assert_power_close(p, ref, tol=TOL):
    peak = np.nanmax(np.abs(ref))
    assert np.nanmax(np.abs(p - ref)) <= tol * peak
    big = np.abs(ref) >= 1e-2 * peak
    assert np.nanmax(np.abs(p[big] - ref[big]) / np.abs(ref[big])) <= tol


def synth(N, T, seed, period=2.75):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    y = 1000 + np.sin(2 * np.pi * t / period + 0.3) + rng.standard_normal(N)
    return t, y, rng


@pytest.mark.parametrize("weighted,fit_mean", [(False, True), (True, True), (False, False), (True, False)])
def test_log_spaced_grid_vs_oracle(gpu_ctx, weighted, fit_mean):
    t, y, rng = synth(5000, 200.0, 81)
    if not fit_mean:
        y = y - y.mean()
    err = rng.uniform(0.5, 1.5, t.size) if weighted else None
    w = None if err is None else err ** -2.0
    freqs = np.geomspace(2e-3, 40.0, 4001)           # spans sub-cycle (f T < 1) to 8000 cycles over the baseline
    p, am, mx = gpu_ctx.gls_freqs(t, y, w, freqs, fit_mean=fit_mean)
    ref = cport.gls_exact_freqs(t, y, err, freqs, fit_mean)
    assert_power_close(p, ref)
    assert am == np.nanargmax(ref) and mx == p[am]
    assert abs(1 / freqs[am] - 2.75) < 0.01


def test_uniform_list_matches_the_grid_entry_point(gpu_ctx):
    t, y, _ = synth(70_000, 1500.0, 82, period=7.3)       # long enough for several sample splits
    df = 1 / (t[-1] - t[0]) / 5
    fmin, nf = 0.5 * df, 20_000
    grid = fmin + df * np.arange(nf)
    pg, ag, mg = gpu_ctx.gls(t, y, None, fmin, df, nf)
    pf, af, mf = gpu_ctx.gls_freqs(t, y, None, grid)
    assert af == ag and np.max(np.abs(pf - pg)) <= 2e-6 * mg
    pf2, af2, mf2 = gpu_ctx.gls_freqs(t, y, None, grid)
    np.testing.assert_array_equal(pf, pf2)               # integer plane: bit-reproducible
    perm = np.random.default_rng(0).permutation(nf)      # any order of the list
    pp, ap, _ = gpu_ctx.gls_freqs(t, y, None, grid[perm])
    np.testing.assert_array_equal(pp, pf[perm])
    assert perm[ap] == af


def test_psd_zero_negative_and_huge_frequencies(gpu_ctx):
    t, y, rng = synth(3000, 80.0, 83)
    err = rng.uniform(0.3, 0.6, t.size)
    w = err ** -2.0
    freqs = np.array([0.0, -0.3636, 0.3636, 1e-4, 0.013, 5000.123, 1e5 + 0.25, 0.37, 2.0, 1e-9])
    p, am, mx = gpu_ctx.gls_freqs(t, y, w, freqs, psd_scale=0.5 * w.sum())
    with np.errstate(all="ignore"):
        ref = cport.gls_exact_freqs(t, y, err, freqs, True, psd=True)
    ok = np.isfinite(ref) & (np.abs(freqs) * (t[-1] - t[0]) > 1e-3)      # f ~ 0 is 0/0-like: same degenerate class only
    assert_power_close(p[ok], ref[ok])
    assert np.isnan(p[0]) or abs(p[0]) < 1e-6 * np.nanmax(ref[ok])
    assert abs(p[1] - p[2]) <= 2e-6 * np.nanmax(p[ok])                   # power is even in f
    assert am == int(np.flatnonzero(ok)[np.nanargmax(ref[ok])])


@pytest.mark.parametrize("N,nf", [(2, 1), (17, 3), (1025, 129), (4000, 50_001)])
def test_ragged_sizes(gpu_ctx, N, nf):
    rng = np.random.default_rng(N + nf)
    t = np.sort(rng.uniform(0, 10, N))
    y = rng.standard_normal(N)
    freqs = np.sort(rng.uniform(0.2, 30.0, nf))
    p, am, mx = gpu_ctx.gls_freqs(t, y, None, freqs)
    ref = cport.gls_exact_freqs(t, y, None, freqs)
    ok = np.isfinite(ref) & (np.abs(ref) < 1e6)
    tol = 1e-3 if N <= 3 else 1e-4
    assert np.nanmax(np.abs(p[ok] - ref[ok])) <= tol * max(1.0, np.nanmax(np.abs(ref[ok])))
    assert p.shape == (nf,)


def test_dropin_class_frequency_keyword_and_multi_device():
    from periodicity_b200 import GLS, TSeries, _ffi
    t, y, rng = synth(20_000, 300.0, 84)
    freqs = np.concatenate([np.linspace(0.01, 1.0, 3000), np.geomspace(1.0, 25.0, 3000)[1:]])
    gls = GLS(frequency=freqs[::-1])                       # any order in, ascending out
    ls = gls(TSeries(t, y))
    np.testing.assert_array_equal(ls.frequency, np.sort(freqs))
    ref = cport.gls_exact_freqs(t, y, None, np.sort(freqs))
    assert_power_close(ls.values, ref)
    assert ls.argmax() == np.nanargmax(ref) == gls.argmax_index and abs(ls.pmax() - 2.75) < 0.01
    err = rng.uniform(0.5, 1.5, t.size)
    y0 = y - y.mean()                                      # fit_mean=False is meant for pre-centred data (window(), spectral.py:165)
    lw = GLS(frequency=freqs, psd=True)(TSeries(t, y0), err=err, fit_mean=False)
    assert_power_close(lw.values, cport.gls_exact_freqs(t, y0, err, np.sort(freqs), False, psd=True))
    import os
    os.environ["PDC_MULTI_MIN_EVALS"] = "1"
    try:
        mctx = _ffi.Context([0, 0, 0])
    finally:
        os.environ.pop("PDC_MULTI_MIN_EVALS", None)
    pm, am, mm = mctx.gls_freqs(t, y, None, np.sort(freqs))
    assert np.max(np.abs(pm - ls.values)) <= 2e-6 * ls.amax()   # a slice may choose another sample split (other FP32 tiles)
    assert am == ls.argmax() and mm == pm[am]
    mctx.close()
    with pytest.raises(ValueError):
        _ffi.default_context(0).gls_freqs(np.arange(5.0), np.arange(4.0), None, [1.0])
