"""Ad-hoc GPU sanity run (development aid; the real parity tests live in tests/)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from periodicity_b200 import _ffi
from oracle import gls_numpy, pdm_numpy

ctx = _ffi.Context(0)
print("sm_count", ctx.sm_count)

def gls_case(N, T, nf, sigma, seed, weighted=False, fit_mean=True, check_rows=2000):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + sigma * rng.standard_normal(N)
    err = rng.uniform(0.5, 1.5, N) if weighted else None
    w = None if err is None else err ** -2.0
    t0 = time.perf_counter()
    p, am, mx = ctx.gls(t, y, w, fmin, df, nf, fit_mean=fit_mean)
    t1 = time.perf_counter()
    p, am, mx = ctx.gls(t, y, w, fmin, df, nf, fit_mean=fit_mean)
    t2 = time.perf_counter()
    kms = ctx.last_main_kernel_ms()
    # exact oracle on a subset of frequencies: strided + window around the peak
    sel = np.unique(np.concatenate([np.arange(0, nf, max(1, nf // check_rows)), np.arange(max(0, am - 50), min(nf, am + 50))]))
    pe = np.empty(sel.size)
    # evaluate exact sums per selected frequency via gls_power on single-frequency grids (chunked trick)
    def ts(tt, ww, dff, nff, fm):
        # called with (df, nf, fmin) or (2df, nf, 2fmin): map selection accordingly
        scale = dff / df
        f = (fmin + df * sel) * scale
        ph = 2 * np.pi * np.outer(f, tt)
        return np.sin(ph) @ ww, np.cos(ph) @ ww
    pe = gls_numpy.gls_power(t, y, err, fmin, df, sel.size, fit_mean, False, trig_sum=ts)
    pk = np.nanmax(p)
    e1 = np.nanmax(np.abs(p[sel] - pe)) / pk
    big = pe >= 1e-2 * pk
    e2 = np.nanmax(np.abs(p[sel][big] - pe[big]) / pe[big])
    e3 = np.nanmax(np.abs(p[sel] - pe) / np.abs(pe))
    pf = gls_numpy.gls_power(t, y, err, fmin, df, nf, fit_mean, False)
    print(json.dumps(dict(case=f"gls N={N} nf={nf} w={weighted} fm={fit_mean}", first_ms=(t1 - t0) * 1e3, call_ms=(t2 - t1) * 1e3, kernel_ms=kms,
                          evals_per_s_kernel=N * nf / (kms * 1e-3), peaknorm_err=e1, rel_err_big=e2, rel_err_all=e3,
                          argmax=am, argmax_np=int(np.nanargmax(p)), argmax_exact_sel=int(sel[np.nanargmax(pe)]), argmax_fast=int(np.nanargmax(pf)), max=mx)))

def pdm_case(N, T, npd, nb, nc, seed, check=50):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    x = 1000 + np.sin(2 * np.pi * t / 3.7) + 0.8 * np.sin(4 * np.pi * t / 3.7) + rng.standard_normal(N)
    periods = np.linspace(1.0, 11.0, npd)
    t0 = time.perf_counter()
    th, am, mn = ctx.pdm(t, x, periods, nb, nc)
    t1 = time.perf_counter()
    th, am, mn = ctx.pdm(t, x, periods, nb, nc)
    t2 = time.perf_counter()
    kms = ctx.last_main_kernel_ms()
    sel = np.unique(np.concatenate([np.arange(0, npd, max(1, npd // check)), np.arange(max(0, am - 5), min(npd, am + 5))]))
    s2 = np.var(x, ddof=1)
    ref = np.array([pdm_numpy.pdm_theta_masks(t, x, periods[i], nb, nc, s2) for i in sel])
    print(json.dumps(dict(case=f"pdm N={N} np={npd} nb={nb} nc={nc}", first_ms=(t1 - t0) * 1e3, call_ms=(t2 - t1) * 1e3, kernel_ms=kms,
                          evals_per_s_kernel=N * npd / (kms * 1e-3), max_rel_err=float(np.max(np.abs(th[sel] - ref) / ref)),
                          argmin=am, argmin_np=int(np.nanargmin(th)), argmin_ref_sel=int(sel[np.argmin(ref)]), min=mn)))

gls_case(1000, 100, 10_000, 0.5, 1)
gls_case(1000, 100, 10_000, 0.5, 1, weighted=True)
gls_case(1000, 100, 10_000, 0.5, 1, fit_mean=False)
gls_case(65_000, 1470, 100_000, 1.0, 2, check_rows=300)
gls_case(65_000, 1470, 100_000, 1.0, 2, weighted=True, check_rows=300)
pdm_case(1000, 100, 1000, 5, 2, 3)
pdm_case(100_000, 1000, 2000, 10, 2, 3, check=20)
pdm_case(100_000, 1000, 100_000, 10, 2, 3, check=10)
if len(sys.argv) > 1:
    gls_case(1_000_000, 1000, 1_000_000, 1.0, 5, check_rows=30)
