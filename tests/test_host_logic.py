"""Host-side logic of the drop-in classes, with the device call replaced by the oracle.

The Python front end (signal coercion, grid, weights, attribute side effects,
sub-harmonic averaging, FSeries wrap) must reproduce reference
spectral.py:86-108,133-135 and phase.py:160-194; the golden vectors were produced
by the reference itself.  The C-ABI call is monkey-patched by a stand-in that
evaluates the oracle, so these tests run without a GPU.
"""
import numpy as np
import pytest

from conftest import GLS_CASES, PDM_CASES, PDM_KW, SL_CASES, load_golden, opt
from oracle import cport, stringlength_numpy
from periodicity_b200 import GLS, PDM, FSeries, StringLength, TSeries, _ffi


class OracleContext:
    """Same method surface as _ffi.Context, computing with oracle/ (tests only)."""

    def gls(self, t, y, w, fmin, df, nf, fit_mean=True, psd_scale=None, j0=0, want_power=True):
        err = None if w is None else np.asarray(w, dtype=np.float64) ** -0.5
        p = cport.gls_exact(t, y, err, fmin, df, nf, fit_mean, psd_scale is not None, j0=j0)
        if psd_scale is not None:
            assert np.isclose(psd_scale, 0.5 * (np.sum(w) if w is not None else len(t)))
        return p, int(np.nanargmax(p)), float(np.nanmax(p))

    def gls_batch(self, t, y, w, offsets, fmin, df, nf, fit_mean=True, psd_scale=None, want_power=True):
        B = len(offsets) - 1
        out = [self.gls(t[offsets[b]:offsets[b + 1]], y[offsets[b]:offsets[b + 1]],
                        None if w is None else w[offsets[b]:offsets[b + 1]],
                        np.broadcast_to(fmin, (B,))[b], np.broadcast_to(df, (B,))[b], nf, fit_mean,
                        None if psd_scale is None else np.broadcast_to(psd_scale, (B,))[b]) for b in range(B)]
        return (np.stack([o[0] for o in out]) if want_power else None,
                np.array([o[1] for o in out]), np.array([o[2] for o in out]))

    def gls_multi(self, t, Y, w, fmin, df, nf, fit_mean=True, psd_scale=None, j0=0, want_power=True):
        out = [self.gls(t, y, w, fmin, df, nf, fit_mean, psd_scale, j0) for y in np.atleast_2d(Y)]
        return (np.stack([o[0] for o in out]) if want_power else None,
                np.array([o[1] for o in out]), np.array([o[2] for o in out]))

    def peaks_topk(self, values, k):
        from scipy.signal import find_peaks
        v = np.atleast_2d(values)
        idx = np.full((v.shape[0], k), -1, dtype=np.int64)
        val = np.full((v.shape[0], k), np.nan)
        for r, row in enumerate(v):
            pk, _ = find_peaks(row, prominence=0.0)
            top = pk[np.lexsort((pk, -row[pk]))][:k]
            idx[r, :top.size], val[r, :top.size] = top, row[top]
        return idx, val

    def pdm(self, t, x, periods, nb, nc):
        th = cport.pdm(t, x, periods, nb, nc)
        return th, int(np.nanargmin(th)), float(np.nanmin(th))

    def gls_freqs(self, t, y, w, freqs, fit_mean=True, psd_scale=None):
        err = None if w is None else np.asarray(w, dtype=np.float64) ** -0.5
        p = cport.gls_exact_freqs(t, y, err, freqs, fit_mean, psd_scale is not None)
        return p, int(np.nanargmax(p)), float(np.nanmax(p))

    def ce(self, t, x, periods, nphi, nm):
        from oracle import ce_numpy
        h = ce_numpy.ce(np.asarray(t, dtype=np.float64), np.asarray(x, dtype=np.float64), periods, nphi, nm)
        return h, int(np.nanargmin(h)), float(np.nanmin(h))

    def gl(self, t, periods, m_max=12, nc=10):
        from oracle import gl_numpy
        lo = gl_numpy.gl(np.asarray(t, dtype=np.float64), periods, m_max, nc)
        return lo, int(np.nanargmax(lo)), float(np.nanmax(lo))

    def stringlength(self, t, m, periods):
        ell = stringlength_numpy.string_lengths(np.asarray(t, dtype=np.float64), np.asarray(m), periods)
        return ell, int(np.nanargmin(ell)), float(np.nanmin(ell))


@pytest.fixture(autouse=True)
def oracle_backend(monkeypatch):
    monkeypatch.setattr(_ffi, "default_context", lambda device=None: OracleContext())


def _signal(g, key):
    v = g[key]
    t = opt(g["t"])
    return v if t is None else TSeries(t, v)


@pytest.mark.parametrize("case", GLS_CASES)
def test_gls_front_end_matches_reference(case):
    g = load_golden(case)
    gls = GLS(fmin=opt(g["fmin"]), fmax=opt(g["fmax"]), n=g["n"], psd=bool(g["psd"]))
    err = opt(g["err"])
    out = gls(_signal(g, "y"), err=err, fit_mean=bool(g["fit_mean"]))
    assert isinstance(out, FSeries) and out is gls.periodogram
    np.testing.assert_array_equal(out.frequency, g["frequency"])
    np.testing.assert_array_equal(gls.frequency, g["frequency"])
    scale = np.nanmax(np.abs(g["power_exact"]))
    assert np.nanmax(np.abs(out.values - g["power_exact"])) <= 1e-7 * scale
    assert out.argmax() == np.nanargmax(g["power_ref"])
    assert isinstance(gls.signal, TSeries) and gls.signal.size == len(g["y"])
    if err is None:
        np.testing.assert_array_equal(gls.err, np.ones_like(g["y"]))   # spectral.py:99-101
    else:
        np.testing.assert_array_equal(gls.err, err)


def test_gls_reference_known_answers():
    sine = TSeries(values=np.sin((np.arange(100) / 100) * 20 * np.pi))
    assert GLS()(sine).period_at_highest_peak == 10.0            # tests/test_spectral.py:27-31
    t0, ts = 2.5, 0.1
    ls = GLS(n=1)(TSeries(np.arange(0, t0 + ts, ts)))            # tests/test_spectral.py:7-24
    freq = ls.frequency
    assert sorted(freq) == list(freq)
    assert freq[0] == (1 / t0) / 2
    assert np.round(freq[-1], 6) == (1 / ts) / 2
    assert np.max(np.abs(np.diff(freq) - 1 / t0)) < 1e-10


def test_gls_window_bootstrap_model_copy():
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 20, 80))
    y = np.sin(2 * np.pi * t / 2.5) + 0.1 * rng.standard_normal(80)
    err = rng.uniform(0.05, 0.15, 80)
    gls = GLS(fmax=2.0)
    ls = gls(TSeries(t, y), err=err)
    win = gls.window()                                           # spectral.py:165-167
    assert win.size == ls.size and np.nanmax(win.values) <= 1.0 + 1e-9
    assert gls.periodogram is ls                                 # window() works on a copy
    reps = gls.bootstrap(5, random_seed=7, batch=2)
    # the reference loop, re-done by hand with the same generator (spectral.py:140-152)
    rng2 = np.random.default_rng(7)
    want = []
    for _ in range(5):
        bs = rng2.integers(0, 80, 80)
        want.append(GLS(fmax=2.0)(TSeries(t, y[bs]), err=err[bs]).amax())
    np.testing.assert_allclose(reps, want, rtol=1e-12)
    assert 0.0 <= gls.fap(ls.amax()) <= 1.0 and gls.fal(0.5) == np.quantile(reps, 0.5)
    # err=None: uniform weights -> the shared-time path (pdc_gls_multi) must give the reference loop too
    gu = GLS(fmax=2.0)
    gu(TSeries(t, y))
    repsu = gu.bootstrap(4, random_seed=11, batch=3)
    rng3 = np.random.default_rng(11)
    wantu = [GLS(fmax=2.0)(TSeries(t, y[rng3.integers(0, 80, 80)])).amax() for _ in range(4)]
    np.testing.assert_allclose(repsu, wantu, rtol=1e-12)
    fit = gls.model(t, 1 / 2.5)                                  # spectral.py:169-204
    assert np.sqrt(np.mean((fit.values - y) ** 2)) < 0.2
    assert gls.copy() is not gls and gls.copy().fmax == 2.0


def test_gls_rejects_mismatched_err():
    with pytest.raises(ValueError):
        _ffi.Context.gls(OracleContext(), np.arange(5.0), np.arange(4.0), None, 0.1, 0.1, 3)


@pytest.mark.parametrize("case", PDM_CASES)
def test_pdm_front_end_matches_reference(case):
    g = load_golden(case)
    kw = {}
    for k in PDM_KW:
        if k in g and opt(g[k]) is not None:
            v = opt(g[k])
            kw[k] = int(v) if k in ("nb", "nc", "n_periods") else (bool(v) if k == "do_subharmonic" else v)
    pdm = PDM(cores=3, **kw)
    out = pdm(_signal(g, "x"))
    np.testing.assert_array_equal(pdm.periods, g["periods"])
    np.testing.assert_array_equal(out.frequency, g["periodogram_frequency"])   # reversed (core.py:877-881)
    np.testing.assert_allclose(out.values, g["periodogram_values"], rtol=1e-11)
    assert out.argmin() == np.nanargmin(g["periodogram_values"])
    assert pdm.sigma == pytest.approx(g["sigma"], rel=1e-14)
    assert pdm.t is pdm.signal.time and pdm.x is pdm.signal.values
    assert pdm._pdm(pdm.periods[3]) == pytest.approx(
        cport.pdm(pdm.t, pdm.x, pdm.periods[3:4], pdm.nb, pdm.nc)[0], rel=1e-13)


def test_survey_front_end_and_top_peaks():
    from periodicity_b200.survey import gls_survey
    rng = np.random.default_rng(8)
    sigs, errs = [], []
    for b in range(3):
        n = 150 + 40 * b
        t = np.sort(rng.uniform(0, 25, n))
        sigs.append(TSeries(t, np.sin(2 * np.pi * t / (1.5 + b)) + 0.2 * rng.standard_normal(n)))
        errs.append(rng.uniform(0.1, 0.3, n))
    out = gls_survey(sigs, errs=errs, nf=400, want_power=True)
    for b, s in enumerate(sigs):
        df = 1.0 / s.baseline / 5                                   # spectral.py:88
        assert out["df"][b] == df and out["fmin"][b] == 0.5 * df    # spectral.py:89-90
        ref = cport.gls_exact(s.time, s.values, errs[b], 0.5 * df, df, 400)
        np.testing.assert_allclose(out["power"][b], ref, rtol=1e-12)
        assert out["argmax"][b] == np.nanargmax(ref)
        assert out["best_period"][b] == 1.0 / (0.5 * df + df * np.nanargmax(ref))
        assert abs(out["best_period"][b] - (1.5 + b)) < 0.1
    with pytest.raises(ValueError):
        gls_survey([], nf=10)
    with pytest.raises(ValueError):
        gls_survey(sigs, errs=errs[:2], nf=10)
    gls = GLS(fmax=2.0)
    ls = gls(sigs[0])
    f, p = gls.top_peaks(3)
    assert 1.0 / f[0] == ls.period_at_highest_peak
    np.testing.assert_array_equal(1.0 / f, ls.psort_by_peak()[:3])


@pytest.mark.parametrize("case", SL_CASES)
def test_stringlength_front_end_matches_reference(case):
    g = load_golden(case)
    sl = StringLength(dphi=float(g["dphi"]), n_periods=int(g["n_periods"]), cores=2)
    out = sl(_signal(g, "x"))
    np.testing.assert_array_equal(sl.m.values, g["m"])                          # phase.py:64-65
    np.testing.assert_array_equal(out.frequency, g["periodogram_frequency"])    # phase.py:66-67,71
    np.testing.assert_allclose(out.values, g["periodogram_values"], rtol=1e-13)
    assert out is sl.periodogram and sl.signal.size == g["x"].size
    p = 1 / out.frequency[5]
    assert sl._stringlength(p) == pytest.approx(out.values[5], rel=1e-13)


def test_new_method_classes_follow_pdm_grid_conventions():
    """CE / GL (reference TODOs, phase.py:13-14) use PDM's period-grid options and FSeries wrap (phase.py:167-180,194);
    GLS(frequency=...) evaluates on the user's grid, sorted ascending, with the same attribute side effects."""
    from periodicity_b200 import CE, GL
    from test_gl_oracle import events
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 60, 800))
    x = np.sin(2 * np.pi * t / 3.3) + 0.2 * rng.standard_normal(t.size)
    sig = TSeries(t, x)
    pdm = PDM(n_periods=300)
    pdm(sig)
    ce = CE(n_periods=300)
    pg = ce(sig)
    np.testing.assert_array_equal(ce.periods, pdm.periods)                       # same defaults: 2*median_dt .. baseline
    np.testing.assert_array_equal(pg.frequency, pdm.periodogram.frequency)      # FSeries over 1/P, ascending frequency
    assert pg.values[::-1][ce.argmin_index] == ce.min_entropy
    ce2 = CE(nb=8, nm=4, p_min=1.0, p_max=8.0, n_periods=None, oversample=2)
    ce2(sig)
    assert ce2.periods.size == int((1 / 1.0 - 1 / 8.0) * 2 * sig.baseline + 1)   # phase.py:176-179
    assert abs(ce2.periods[ce2.argmin_index] - 3.3) < 0.1
    ev = events(900, 4.1, 3, T=120.0)
    gl = GL(m_max=5, nc=3, p_min=2.0, p_max=9.0, n_periods=200)
    pg = gl(ev[::-1])                                                            # plain, unsorted array of event times
    assert np.all(np.diff(gl.t) >= 0) and gl.periods[0] == 2.0 and gl.periods[-1] == 9.0
    assert abs(gl.periods[gl.argmax_index] - 4.1) < 0.1 and pg.values[::-1][gl.argmax_index] == gl.max_lnodds
    freqs = np.array([0.9, 0.1, 0.303, 0.5, 0.2])
    gls = GLS(frequency=freqs, psd=True)
    err = rng.uniform(0.1, 0.3, t.size)
    out = gls(sig, err=err)
    np.testing.assert_array_equal(out.frequency, np.sort(freqs))
    np.testing.assert_array_equal(gls.frequency, np.sort(freqs))
    assert gls.err is not None and out is gls.periodogram and out.frequency[out.argmax()] == 0.303
    np.testing.assert_allclose(out.values, cport.gls_exact_freqs(t, x, err, np.sort(freqs), True, psd=True), rtol=1e-12)
