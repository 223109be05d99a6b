"""Generate the golden vectors that pin ``oracle/`` -- run in the AUTHORING container only.

    python tests/golden/make_golden.py

Runs the UNMODIFIED reference files ``/root/reference/src/periodicity/spectral.py``
and ``phase.py`` (loaded by ``oracle/refload.py`` behind a numpy-only stand-in for
``periodicity.core``; the package itself needs xarray, which is absent) on small
seeded inputs and stores inputs + outputs under ``tests/golden/*.npz``.

For GLS two outputs are stored per case:
  ``power_ref``    the reference as shipped (FFT/extirpolation ``_trig_sum``),
  ``power_exact``  the reference's ``GLS.__call__`` with ``spectral._trig_sum``
                   monkey-patched by the direct float64 sums it approximates
                   (docstring ``spectral.py:12-16``) -- the "formula oracle".
The fixtures travel with the repo; ``/root/reference`` does not exist on the GPU box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import refload  # noqa: E402


def exact_trig_sum(t, w, df, nf, fmin, n=5):
    f = fmin + df * np.arange(nf)
    S = np.empty(nf)
    C = np.empty(nf)
    step = max(1, 2_000_000 // max(1, len(t)))
    for a in range(0, nf, step):
        ph = 2 * np.pi * np.outer(f[a:a + step], t)
        S[a:a + step] = np.sin(ph) @ w
        C[a:a + step] = np.cos(ph) @ w
    return S, C


def run_gls(spectral, t, y, err, kw, fit_mean=True):
    TS = refload.TSeries
    sig = TS(t, y) if t is not None else y
    g = spectral.GLS(**kw)
    ref = g(sig, err=err, fit_mean=fit_mean)
    freq, power_ref = np.array(ref.frequency), np.array(ref.values)
    orig = spectral._trig_sum
    spectral._trig_sum = exact_trig_sum
    try:
        ex = spectral.GLS(**kw)(sig, err=err, fit_mean=fit_mean)
    finally:
        spectral._trig_sum = orig
    return freq, power_ref, np.array(ex.values)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path)} bytes)")


def main():
    spectral, phase = refload.load()
    NONE = np.array([])

    # ---- GLS -----------------------------------------------------------------
    # (1) the reference's own known-answer test, tests/test_spectral.py:27-31
    y = np.sin((np.arange(100) / 100) * 20 * np.pi)
    f, pr, pe = run_gls(spectral, None, y, None, {})
    save("gls_sine100", t=NONE, y=y, err=NONE, fit_mean=True, psd=False, n=5, fmin=np.nan, fmax=np.nan,
         frequency=f, power_ref=pr, power_exact=pe)

    # (2) the reference's grid test, tests/test_spectral.py:7-24 (values are all ones: only the grid is pinned)
    time = np.arange(0, 2.5 + 0.1, 0.1)
    sig = refload.TSeries(time)
    freq = np.array(spectral.GLS(n=1)(sig).frequency)
    save("gls_grid_n1", t=time, frequency=freq)

    # (3) C1-like irregular sinusoid + noise (SURVEY.md §8d recipe), reduced nf
    rng = np.random.default_rng(1)
    N, T, nf = 1000, 100.0, 2000
    t = np.sort(rng.uniform(0, T, N))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    fmax = fmin + (nf - 1.5) * df
    fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + 0.5 * rng.standard_normal(N)
    f, pr, pe = run_gls(spectral, t, y, None, dict(fmin=fmin, fmax=fmax))
    save("gls_c1_small", t=t, y=y, err=NONE, fit_mean=True, psd=False, n=5, fmin=fmin, fmax=fmax,
         frequency=f, power_ref=pr, power_exact=pe)

    # (4) heteroscedastic errors
    err = rng.uniform(0.5, 1.5, N)
    f, pr, pe = run_gls(spectral, t, y, err, dict(fmin=fmin, fmax=fmax))
    save("gls_err", t=t, y=y, err=err, fit_mean=True, psd=False, n=5, fmin=fmin, fmax=fmax,
         frequency=f, power_ref=pr, power_exact=pe)

    # (5) fit_mean=False (the window() branch, spectral.py:114-115,165-167) on a zero-mean signal
    y0 = y - y.mean()
    f, pr, pe = run_gls(spectral, t, y0, None, dict(fmin=fmin, fmax=fmax), fit_mean=False)
    save("gls_nofitmean", t=t, y=y0, err=NONE, fit_mean=False, psd=False, n=5, fmin=fmin, fmax=fmax,
         frequency=f, power_ref=pr, power_exact=pe)

    # (6) psd=True with errors (spectral.py:129-130)
    f, pr, pe = run_gls(spectral, t, y, err, dict(fmin=fmin, fmax=fmax, psd=True))
    save("gls_psd", t=t, y=y, err=err, fit_mean=True, psd=True, n=5, fmin=fmin, fmax=fmax,
         frequency=f, power_ref=pr, power_exact=pe)

    # (7) default grid on irregular times with large time origin (JD-like), n=3
    t7 = 2450000.0 + np.sort(rng.uniform(0, 30.0, 400))
    y7 = 10 + 0.3 * np.sin(2 * np.pi * t7 / 2.345) + 0.1 * rng.standard_normal(400)
    f, pr, pe = run_gls(spectral, t7, y7, None, dict(n=3))
    save("gls_jd_default", t=t7, y=y7, err=NONE, fit_mean=True, psd=False, n=3, fmin=np.nan, fmax=np.nan,
         frequency=f, power_ref=pr, power_exact=pe)

    # (8) real irregular light curve bundled with the reference (t, flux, flux_err), first 600 points
    star = np.load(os.path.join(refload.REFERENCE_ROOT, "src", "periodicity", "data", "spotted_star.npy"))
    ts, ys, es = star[0][:600], star[1][:600], star[2][:600]
    f, pr, pe = run_gls(spectral, ts, ys, es, dict(fmax=2.0))
    save("gls_spotted_star", t=ts, y=ys, err=es, fit_mean=True, psd=False, n=5, fmin=np.nan, fmax=2.0,
         frequency=f, power_ref=pr, power_exact=pe)

    # ---- PDM -----------------------------------------------------------------
    def run_pdm(name, t, x, **kw):
        sig = refload.TSeries(t, x) if t is not None else x
        p = phase.PDM(cores=2, **kw)
        out = p(sig)
        save(name, t=NONE if t is None else t, x=x, periods=np.array(p.periods),
             periodogram_frequency=np.array(out.frequency), periodogram_values=np.array(out.values),
             sigma=p.sigma, **{k: (np.nan if v is None else v) for k, v in kw.items()})

    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0, 100, 1000))
    x = 1000 + np.sin(2 * np.pi * t / 3.7) + 0.8 * np.sin(4 * np.pi * t / 3.7) + rng.standard_normal(1000)
    run_pdm("pdm_basic", t, x, nb=5, nc=2, p_min=1.0, p_max=11.0, n_periods=200)
    run_pdm("pdm_negative_t_nc3", t[:300] - 50.0, x[:300], nb=10, nc=3, p_min=1.0, p_max=11.0, n_periods=150)
    run_pdm("pdm_sparse", t[::25], x[::25], nb=10, nc=2, p_min=1.0, p_max=11.0, n_periods=120)
    xi = np.sin(2 * np.pi * np.arange(240) / 12.0) + 0.3 * rng.standard_normal(240)
    run_pdm("pdm_integer_t_ties", None, xi, nb=5, nc=2, p_min=2.0, p_max=26.0, n_periods=97)
    run_pdm("pdm_subharmonic", t, x, nb=5, nc=2, p_min=1.0, p_max=11.0, n_periods=201, do_subharmonic=True)
    t5 = np.sort(rng.uniform(0, 400, 5000))
    x5 = np.sin(2 * np.pi * t5 / 7.1) + 0.5 * rng.standard_normal(5000)
    run_pdm("pdm_nc1", t5, x5, nb=7, nc=1, p_min=2.0, p_max=20.0, n_periods=90)
    run_pdm("pdm_defaults", t, x, n_periods=64)

    # ---- String Length ----------------------------------------------------------
    # phase.py:18-72 executed unmodified; the stand-in core's max()/min() return scalars (see oracle/refload.py:
    # with the reference's own core.py the scaling line phase.py:65 cannot run)
    def run_sl(name, t, x, **kw):
        sig = refload.TSeries(t, x) if t is not None else x
        s_ = phase.StringLength(cores=2, **kw)
        out = s_(sig)
        save(name, t=NONE if t is None else t, x=x, m=np.array(s_.m.values),
             periodogram_frequency=np.array(out.frequency), periodogram_values=np.array(out.values), **kw)

    rng = np.random.default_rng(6)
    t = np.sort(rng.uniform(0, 60, 300))
    x = 12.0 + 0.4 * np.sin(2 * np.pi * t / 4.3) + 0.05 * rng.standard_normal(300)
    run_sl("sl_basic", t, x, dphi=0.1, n_periods=400)
    run_sl("sl_sparse_negative_t", t[::7] - 30.0, x[::7], dphi=0.25, n_periods=150)
    xi = np.round(np.sin(2 * np.pi * np.arange(120) / 8.0), 3)       # integer times: many equal phases (ties)
    run_sl("sl_integer_t_ties", None, xi, dphi=0.5, n_periods=60)
    t3 = np.sort(rng.uniform(0, 300, 3000))
    x3 = np.sin(2 * np.pi * t3 / 11.7) ** 3 + 0.1 * rng.standard_normal(3000)
    run_sl("sl_3000", t3, x3, dphi=0.1, n_periods=100)


if __name__ == "__main__":
    main()
