"""Pins the CPU oracle (oracle/) against outputs of the unmodified reference.

Fixtures: tests/golden/*.npz, written by tests/golden/make_golden.py from
/root/reference/src/periodicity/{spectral,phase}.py.  No GPU needed.
"""
import numpy as np
import pytest

from conftest import GLS_CASES, PDM_CASES, PDM_KW, SL_CASES, load_golden, opt
from oracle import cport, gls_numpy, pdm_numpy, stringlength_numpy


def _gls_inputs(g):
    y = g["y"]
    t = opt(g["t"])
    if t is None:
        t = np.arange(len(y))                      # TSeries(values=signal), core.py:462-463
    return np.asarray(t, dtype=np.float64), y, opt(g["err"])


@pytest.mark.parametrize("case", GLS_CASES)
def test_gls_numpy_restatement_matches_reference_as_shipped(case):
    g = load_golden(case)
    t, y, err = _gls_inputs(g)
    f, p = gls_numpy.gls(t, y, err, fmin=opt(g["fmin"]), fmax=opt(g["fmax"]), n=g["n"], psd=bool(g["psd"]),
                         fit_mean=bool(g["fit_mean"]), exact=False)
    assert f.shape == g["frequency"].shape
    np.testing.assert_array_equal(f, g["frequency"])          # grid: bit-exact (spectral.py:88-97)
    scale = np.nanmax(np.abs(g["power_ref"]))
    assert np.nanmax(np.abs(p - g["power_ref"])) <= 1e-9 * scale
    assert np.nanargmax(p) == np.nanargmax(g["power_ref"])


@pytest.mark.parametrize("case", GLS_CASES)
def test_gls_formula_oracles_match_reference_with_exact_sums(case):
    g = load_golden(case)
    t, y, err = _gls_inputs(g)
    fmin, df, f = gls_numpy.gls_grid(t, g["n"], opt(g["fmin"]), opt(g["fmax"]))
    pe = g["power_exact"]
    scale = np.nanmax(np.abs(pe))
    big = np.abs(pe) >= 1e-2 * scale
    # numpy direct sums
    p_np = gls_numpy.gls_power(t, y, err, fmin, df, f.size, bool(g["fit_mean"]), bool(g["psd"]),
                               trig_sum=gls_numpy.trig_sum_exact)
    # C restatement (phase reduced mod 1 before sin/cos)
    p_c = cport.gls_exact(t, y, err, fmin, df, f.size, bool(g["fit_mean"]), bool(g["psd"]))
    for p in (p_np, p_c):
        assert np.nanmax(np.abs(p - pe)) <= 1e-7 * scale
        assert np.nanmax(np.abs(p[big] - pe[big]) / np.abs(pe[big])) <= 1e-6
        assert np.nanargmax(p) == np.nanargmax(pe)
    # and the exact formula agrees with the as-shipped reference on the peak index
    assert np.nanargmax(p_c) == np.nanargmax(g["power_ref"])


def test_reference_known_answer_period_is_10():
    """tests/test_spectral.py:27-31 of the reference: period_at_highest_peak == 10.0 exactly."""
    from scipy.signal import find_peaks
    g = load_golden("gls_sine100")
    for key in ("power_ref", "power_exact"):
        peaks, _ = find_peaks(g[key], prominence=0.0)
        best = peaks[np.nanargmax(g[key][peaks])]
        assert best == 49 and g["frequency"].size == 248
        assert 1.0 / g["frequency"][best] == 10.0
    p_c = cport.gls_exact(np.arange(100.0), g["y"], None, g["frequency"][0],
                          g["frequency"][1] - g["frequency"][0], 248)
    peaks, _ = find_peaks(p_c, prominence=0.0)
    assert peaks[np.nanargmax(p_c[peaks])] == 49


def test_reference_known_answer_grid():
    """tests/test_spectral.py:7-24 of the reference."""
    g = load_golden("gls_grid_n1")
    fmin, df, f = gls_numpy.gls_grid(g["t"], n=1)
    np.testing.assert_array_equal(f, g["frequency"])
    f0, fs = 1 / 2.5, 10.0
    assert f[0] == f0 / 2 and np.round(f[-1], 6) == fs / 2
    assert np.max(np.abs(np.diff(f) - f0)) < 1e-10


def _pdm_kwargs(g):
    kw = {}
    for k in PDM_KW:
        if k in g:
            v = opt(g[k])
            if v is not None:
                kw[k] = int(v) if k in ("nb", "nc", "n_periods") else (bool(v) if k == "do_subharmonic" else v)
    return kw


@pytest.mark.parametrize("case", PDM_CASES)
@pytest.mark.parametrize("impl", ["masks", "hist", "c"])
def test_pdm_restatements_match_reference(case, impl):
    g = load_golden(case)
    x = g["x"]
    t = opt(g["t"])
    if t is None:
        t = np.arange(len(x))
    t = np.asarray(t, dtype=np.float64)
    kw = _pdm_kwargs(g)
    sub = kw.pop("do_subharmonic", False)
    nb, nc = kw.get("nb", 5), kw.get("nc", 2)
    p_min, p_max, periods = pdm_numpy.pdm_grid(t, kw.get("p_min"), kw.get("p_max"), kw.get("n_periods", 1000))
    np.testing.assert_array_equal(periods, g["periods"])            # phase.py:180, bit-exact
    if impl == "c":
        theta = cport.pdm(t, x, periods, nb, nc)
    else:
        fn = pdm_numpy.pdm_theta_masks if impl == "masks" else pdm_numpy.pdm_theta_hist
        s2 = np.var(x, ddof=1)
        theta = np.array([fn(t, x, p, nb, nc, s2) for p in periods])
    if sub:
        theta = pdm_numpy.subharmonic_average(theta, periods, p_min, p_max, x.size)
    # the reference wraps FSeries(1/periods, thetas), which re-sorts ascending in frequency (core.py:877-881)
    order = np.argsort(1 / periods, kind="stable")
    np.testing.assert_allclose((1 / periods)[order], g["periodogram_frequency"], rtol=0, atol=0)
    np.testing.assert_allclose(theta[order], g["periodogram_values"], rtol=1e-11, atol=0)
    assert np.nanargmin(theta[order]) == np.nanargmin(g["periodogram_values"])


@pytest.mark.parametrize("case", SL_CASES)
def test_stringlength_restatement_matches_reference_code(case):
    """phase.py:18-72 executed unmodified behind the stand-in core (scalar max()/min(), see oracle/refload.py)."""
    g = load_golden(case)
    x = g["x"]
    t = opt(g["t"])
    t = np.arange(len(x), dtype=np.float64) if t is None else t
    periods, ell = stringlength_numpy.stringlength(t, x, dphi=float(g["dphi"]), n_periods=int(g["n_periods"]))
    np.testing.assert_array_equal(stringlength_numpy.scale(x), g["m"])
    np.testing.assert_array_equal((1 / periods)[::-1], g["periodogram_frequency"])     # FSeries re-sorts (core.py:877-881)
    np.testing.assert_allclose(ell[::-1], g["periodogram_values"], rtol=1e-13)


def test_c_oracle_indexed_and_free_frequency_entry_points_agree_with_the_grid_one():
    """orc_gls_exact_at / orc_gls_exact_freqs (strided checks of huge grids, non-uniform grids) are the same
    code as orc_gls_exact, which the golden vectors pin: same values at the same frequencies."""
    g = load_golden("gls_err")
    t, y, err = _gls_inputs(g)
    fmin, df, f = gls_numpy.gls_grid(t, g["n"], opt(g["fmin"]), opt(g["fmax"]))
    full = cport.gls_exact(t, y, err, fmin, df, f.size, bool(g["fit_mean"]), bool(g["psd"]))
    sel = np.unique(np.concatenate([np.arange(0, f.size, 7), [f.size - 1]]))
    at = cport.gls_exact_at(t, y, err, fmin, df, sel, bool(g["fit_mean"]), bool(g["psd"]))
    np.testing.assert_array_equal(at, full[sel])
    fr = cport.gls_exact_freqs(t, y, err, fmin + sel * df, bool(g["fit_mean"]), bool(g["psd"]))
    np.testing.assert_array_equal(fr, full[sel])
