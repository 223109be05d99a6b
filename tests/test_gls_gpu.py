"""GPU parity tests for GLS: CUDA path (through the C ABI) vs the oracle and the golden vectors.

Tolerances (BASELINE.json north_star): relative power error <= 1e-5 --
asserted as (i) max|dp| / max(p) <= 1e-5 and (ii) elementwise relative error <= 1e-5 on
bins with p >= 1e-2 max(p) -- against the FORMULA oracle (exact sums), and an identical
peak index against both the formula oracle and the reference as shipped.
"""
import numpy as np
import pytest

from conftest import GLS_CASES, load_golden, opt
from oracle import cport, gls_numpy

pytestmark = pytest.mark.gpu

TOL = 1e-5


def assert_power_close(p, ref, tol=TOL):
    peak = np.nanmax(np.abs(ref))
    assert np.nanmax(np.abs(p - ref)) <= tol * peak
    big = np.abs(ref) >= 1e-2 * peak
    assert np.nanmax(np.abs(p[big] - ref[big]) / np.abs(ref[big])) <= tol


def synth(N, T, nf, sigma, seed, fsig=None):
    """SURVEY.md §8d common GLS recipe."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    if fsig is None:
        fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + sigma * rng.standard_normal(N)
    return t, y, fmin, df


@pytest.mark.parametrize("case", GLS_CASES)
def test_golden_cases_through_dropin_class(case):
    from periodicity_b200 import GLS, TSeries
    g = load_golden(case)
    t = opt(g["t"])
    sig = g["y"] if t is None else TSeries(t, g["y"])
    gls = GLS(fmin=opt(g["fmin"]), fmax=opt(g["fmax"]), n=g["n"], psd=bool(g["psd"]))
    out = gls(sig, err=opt(g["err"]), fit_mean=bool(g["fit_mean"]))
    np.testing.assert_array_equal(out.frequency, g["frequency"])
    assert_power_close(out.values, g["power_exact"])
    assert out.argmax() == np.nanargmax(g["power_exact"]) == np.nanargmax(g["power_ref"])
    assert gls.argmax_index == out.argmax() and gls.max_power == out.amax()


def test_reference_known_answer_tests_on_gpu():
    from periodicity_b200 import GLS, TSeries
    sine = TSeries(values=np.sin((np.arange(100) / 100) * 20 * np.pi))
    assert GLS()(sine).period_at_highest_peak == 10.0              # reference tests/test_spectral.py:27-31
    ls = GLS(n=1)(TSeries(np.arange(0, 2.6, 0.1)))                  # reference tests/test_spectral.py:7-24
    freq = ls.frequency
    assert sorted(freq) == list(freq) and freq[0] == 0.4 / 2
    assert np.round(freq[-1], 6) == 5.0 and np.max(np.abs(np.diff(freq) - 0.4)) < 1e-10


@pytest.mark.parametrize("weighted,fit_mean", [(False, True), (True, True), (False, False), (True, False)])
def test_c1_full_grid_vs_c_oracle(gpu_ctx, weighted, fit_mean):
    N, nf = 1000, 10_000
    t, y, fmin, df = synth(N, 100.0, nf, 0.5, 1)
    if not fit_mean:
        y = y - y.mean()
    err = np.random.default_rng(9).uniform(0.5, 1.5, N) if weighted else None
    w = None if err is None else err ** -2.0
    p, am, mx = gpu_ctx.gls(t, y, w, fmin, df, nf, fit_mean=fit_mean)
    ref = cport.gls_exact(t, y, err, fmin, df, nf, fit_mean)
    assert_power_close(p, ref)
    fast = gls_numpy.gls_power(t, y, err, fmin, df, nf, fit_mean, False)
    assert am == np.nanargmax(p) == np.nanargmax(ref) == np.nanargmax(fast)
    assert mx == np.nanmax(p)


def test_psd_normalisation(gpu_ctx):
    t, y, fmin, df = synth(700, 60.0, 3000, 0.5, 4)
    err = np.random.default_rng(5).uniform(0.5, 1.5, 700)
    w = err ** -2.0
    p, _, _ = gpu_ctx.gls(t, y, w, fmin, df, 3000, psd_scale=0.5 * w.sum())
    assert_power_close(p, cport.gls_exact(t, y, err, fmin, df, 3000, True, psd=True))


def test_c2_kepler_like_subset_vs_oracle_and_peak_vs_reference_algorithm(gpu_ctx):
    """BASELINE config C2: 65,000 points x 1e5 frequencies."""
    rng = np.random.default_rng(2)
    slots = np.sort(rng.choice(71_940, 65_000, replace=False))
    t = slots * (29.4244 / 1440.0) + rng.uniform(0, 1 / 1440.0, 65_000)
    nf = 100_000
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + rng.standard_normal(65_000)
    p, am, mx = gpu_ctx.gls(t, y, None, fmin, df, nf)
    sel = np.unique(np.concatenate([np.arange(0, nf, 331), np.arange(am - 40, am + 40), [nf - 1]]))
    ref = cport.gls_exact_at(t, y, None, fmin, df, sel)
    peak = np.nanmax(p)
    assert np.max(np.abs(p[sel] - ref)) <= TOL * peak
    big = ref >= 1e-2 * peak
    assert np.max(np.abs(p[sel][big] - ref[big]) / ref[big]) <= TOL
    assert sel[np.argmax(ref)] == am == np.nanargmax(p)
    fast = gls_numpy.gls_power(t, y, None, fmin, df, nf, True, False)   # the reference's own algorithm
    assert np.nanargmax(fast) == am


def test_grid_shard_offsets_are_consistent(gpu_ctx):
    """Frequency-grid sharding: [0, nf) == [0, a) ++ [a, nf) with j0 offsets."""
    nf = 9001
    t, y, fmin, df = synth(3000, 200.0, nf, 1.0, 6)
    full, am, mx = gpu_ctx.gls(t, y, None, fmin, df, nf)
    a = 4097
    lo, am0, mx0 = gpu_ctx.gls(t, y, None, fmin, df, a, j0=0)
    hi, am1, mx1 = gpu_ctx.gls(t, y, None, fmin, df, nf - a, j0=a)
    cat = np.concatenate([lo, hi])
    assert np.nanmax(np.abs(cat - full)) <= 2e-6 * np.nanmax(full)
    best = (am0, mx0) if mx0 >= mx1 else (am1 + a, mx1)
    assert best[0] == am


def test_invariances_and_determinism(gpu_ctx):
    nf = 5000
    t, y, fmin, df = synth(4000, 300.0, nf, 1.0, 7)
    p1, a1, m1 = gpu_ctx.gls(t, y, None, fmin, df, nf)
    p2, a2, m2 = gpu_ctx.gls(t, y, None, fmin, df, nf)
    np.testing.assert_array_equal(p1, p2)                         # run-to-run bitwise identical
    assert (a1, m1) == (a2, m2)
    p3, a3, _ = gpu_ctx.gls(t, 3.5 * y - 17.0, None, fmin, df, nf)     # affine invariance with fit_mean
    assert a3 == a1 and np.nanmax(np.abs(p3 - p1)) <= 2e-6 * m1
    p4, a4, _ = gpu_ctx.gls(t + 2_450_000.0, y, None, fmin, df, nf)    # time-origin invariance (JD offsets)
    assert a4 == a1 and np.nanmax(np.abs(p4 - p1)) <= 1e-5 * m1
    perm = np.random.default_rng(0).permutation(t.size)                # the C ABI does not need sorted times
    p5, a5, _ = gpu_ctx.gls(t[perm], y[perm], None, fmin, df, nf)
    assert a5 == a1 and np.nanmax(np.abs(p5 - p1)) <= 2e-6 * m1


@pytest.mark.parametrize("N,nf", [(2, 1), (3, 15), (17, 4097), (1025, 33), (2049, 257)])
def test_ragged_sizes(gpu_ctx, N, nf):
    rng = np.random.default_rng(N * 1000 + nf)
    t = np.sort(rng.uniform(0, 10, N))
    y = rng.standard_normal(N)
    df = 1 / (t[-1] - t[0]) / 5
    p, am, mx = gpu_ctx.gls(t, y, None, 0.5 * df, df, nf)
    ref = cport.gls_exact(t, y, None, 0.5 * df, df, nf)
    ok = np.isfinite(ref) & (np.abs(ref) < 1e6)
    # N <= 3 with a floating mean is an exact fit (3 parameters): CC, SS -> 0 and the problem is
    # ill-conditioned by construction, so only a loose bound is meaningful there.
    tol = 1e-3 if N <= 3 else 1e-4
    assert np.nanmax(np.abs(p[ok] - ref[ok])) <= tol * max(1.0, np.nanmax(np.abs(ref[ok])))
    assert p.shape == (nf,)


def test_constant_signal_does_not_crash(gpu_ctx):
    from periodicity_b200 import GLS, TSeries
    ls = GLS(n=1)(TSeries(np.arange(0, 2.6, 0.1)))                   # all-ones values: YY ~ 0, power is junk
    assert ls.size == 13


def test_invalid_arguments_raise_value_error(gpu_ctx):
    with pytest.raises(ValueError):
        gpu_ctx.gls(np.arange(5.0), np.arange(4.0), None, 0.1, 0.1, 3)
    with pytest.raises(ValueError):
        gpu_ctx.gls(np.arange(5.0), np.arange(5.0), None, 0.1, 0.1, 0)
    with pytest.raises(ValueError):
        gpu_ctx.gls_batch(np.arange(5.0), np.arange(5.0), None, [0, 3, 3, 5], 0.1, 0.1, 4)


def test_batch_matches_single_calls_ragged(gpu_ctx):
    """pdc_gls_batch (survey workload / bootstrap): ragged curves, per-curve grids."""
    rng = np.random.default_rng(21)
    sizes = [500, 1300, 37, 2048, 999]
    nf = 777
    ts, ys, ws, fm, dfs = [], [], [], [], []
    for n in sizes:
        t = np.sort(rng.uniform(0, rng.uniform(20, 60), n))
        y = np.sin(2 * np.pi * t / rng.uniform(0.5, 5)) + rng.standard_normal(n)
        ts.append(t); ys.append(y); ws.append(rng.uniform(0.5, 2, n))
        d = 1 / (t[-1] - t[0]) / 5
        dfs.append(d); fm.append(0.5 * d)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    T, Y, W = np.concatenate(ts), np.concatenate(ys), np.concatenate(ws)
    for w_all, w_list in ((None, [None] * 5), (W, ws)):
        P, A, M = gpu_ctx.gls_batch(T, Y, w_all, offsets, fm, dfs, nf)
        _, A2, M2 = gpu_ctx.gls_batch(T, Y, w_all, offsets, fm, dfs, nf, want_power=False)
        np.testing.assert_array_equal(A, A2)
        np.testing.assert_array_equal(M, M2)
        for b in range(5):
            err = None if w_list[b] is None else w_list[b] ** -0.5
            ref = cport.gls_exact(ts[b], ys[b], err, fm[b], dfs[b], nf)
            assert_power_close(P[b], ref)
            assert A[b] == np.nanargmax(ref) and M[b] == np.nanmax(P[b])


def test_torch_device_pointer_entry_matches_host_entry(gpu_ctx):
    import torch
    from periodicity_b200 import dist as pdist
    nf = 3000
    t, y, fmin, df = synth(2500, 100.0, nf, 1.0, 8)
    p, am, mx = gpu_ctx.gls(t, y, None, fmin, df, nf)
    tt, yy = torch.from_numpy(t).cuda(), torch.from_numpy(y).cuda()
    pd, ad, md = pdist.gls_torch(tt, yy, None, fmin, df, nf, ctx=gpu_ctx)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(pd.cpu().numpy(), p)                # same kernels, same decomposition
    assert int(ad.item()) == am and float(md.item()) == mx


def test_bootstrap_on_gpu_matches_looped_calls():
    from periodicity_b200 import GLS, TSeries
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 20, 120))
    y = np.sin(2 * np.pi * t / 2.5) + 0.2 * rng.standard_normal(120)
    err = rng.uniform(0.1, 0.3, 120)
    gls = GLS(fmax=3.0)
    gls(TSeries(t, y), err=err)
    reps = gls.bootstrap(6, random_seed=3, batch=4)
    rng2 = np.random.default_rng(3)
    want = []
    for _ in range(6):
        bs = rng2.integers(0, 120, 120)
        want.append(GLS(fmax=3.0)(TSeries(t, y[bs]), err=err[bs]).amax())
    np.testing.assert_allclose(reps, want, rtol=5e-6)


def test_survey_api_matches_per_curve_class_calls():
    from periodicity_b200 import GLS, TSeries
    from periodicity_b200.survey import gls_survey
    rng = np.random.default_rng(31)
    sigs, errs, periods = [], [], []
    for b in range(6):
        n = int(rng.integers(300, 900))
        t = np.sort(rng.uniform(0, 27.8, n))
        P = rng.uniform(0.5, 10.0)
        sigs.append(TSeries(t, 1000 + np.sin(2 * np.pi * t / P) + 0.5 * rng.standard_normal(n)))
        errs.append(rng.uniform(0.4, 0.6, n))
        periods.append(P)
    for e in (None, errs):
        out = gls_survey(sigs, errs=e, nf=2000, want_power=True)
        for b, s in enumerate(sigs):
            g = GLS(fmax=out["fmin"][b] + 1998.5 * out["df"][b])
            ls = g(s, err=None if e is None else e[b])
            assert ls.size == 2000
            assert np.nanmax(np.abs(ls.values - out["power"][b])) <= 2e-6 * ls.amax()
            assert ls.argmax() == out["argmax"][b]
            assert abs(out["best_period"][b] - periods[b]) / periods[b] < 0.05


def test_sharded_entry_points_single_rank(gpu_ctx):
    """dist.gls_sharded / pdm_sharded / gls_batch_sharded with the real device compute (world size 1)."""
    from periodicity_b200 import GLS, PDM, TSeries
    from periodicity_b200 import dist as pdist
    nf = 2500
    t, y, fmin, df = synth(1500, 80.0, nf, 1.0, 41)
    p, am, mx = gpu_ctx.gls(t, y, None, fmin, df, nf)
    ps, ams, mxs = pdist.gls_sharded(t, y, None, fmin, df, nf, device=0)
    np.testing.assert_array_equal(ps, p)
    assert (ams, mxs) == (am, mx)
    ls = GLS(fmin=fmin, fmax=fmin + (nf - 1.5) * df, shard=True)(TSeries(t, y))
    np.testing.assert_array_equal(ls.values, p)
    periods = np.linspace(0.5, 5.0, 700)
    th, a2, m2 = gpu_ctx.pdm(t, y, periods, 5, 2)
    ths, a2s, m2s = pdist.pdm_sharded(t, y, periods, 5, 2, device=0)
    np.testing.assert_array_equal(ths, th)
    assert (a2s, m2s) == (a2, m2)
    pg = PDM(p_min=0.5, p_max=5.0, n_periods=700, shard=True)(TSeries(t, y))
    np.testing.assert_array_equal(pg.values, th[::-1])
    offsets = np.array([0, 400, 1500])
    P, A, M = gpu_ctx.gls_batch(t, y, None, offsets, [fmin, fmin], [df, 2 * df], 300)
    Ps, As, Ms = pdist.gls_batch_sharded(t, y, None, offsets, [fmin, fmin], [df, 2 * df], 300, want_power=True, device=0)
    np.testing.assert_array_equal(Ps, P)
    np.testing.assert_array_equal(As, A)
    np.testing.assert_array_equal(Ms, M)


def test_zero_frequency_bin_matches_oracle_nan_class(gpu_ctx):
    """fmin = 0 puts f = 0 on the grid: sin == 0 there and the formula degenerates to 0/0;
    the NaN-aware device argmax must skip that bin."""
    rng = np.random.default_rng(51)
    t = np.sort(rng.uniform(0, 40, 500))
    y = np.sin(2 * np.pi * t / 3.1) + 0.2 * rng.standard_normal(500)
    df = 1 / (t[-1] - t[0]) / 5
    p, am, mx = gpu_ctx.gls(t, y, None, 0.0, df, 600)
    ref = cport.gls_exact(t, y, None, 0.0, df, 600)
    # f = 0 is 0/0-like in the formula: rounding decides between NaN and a ~1e-18 value (the C oracle gives
    # -2e-18, exact arithmetic gives NaN); either is the same degenerate class and never a peak.
    assert np.isnan(p[0]) or abs(p[0]) < 1e-9
    assert_power_close(p[1:], ref[1:])
    assert am == 1 + np.nanargmax(ref[1:]) and np.isfinite(mx) and mx == np.nanmax(p)


def test_weighted_batch_with_psd(gpu_ctx):
    rng = np.random.default_rng(52)
    sizes = [700, 333]
    ts = [np.sort(rng.uniform(0, 50, n)) for n in sizes]
    ys = [5 + np.sin(2 * np.pi * t / 2.2) + 0.4 * rng.standard_normal(t.size) for t in ts]
    es = [rng.uniform(0.3, 0.6, n) for n in sizes]
    offsets = np.array([0, 700, 1033])
    dfs = [1 / (t[-1] - t[0]) / 5 for t in ts]
    W = np.concatenate(es) ** -2.0
    psd = [0.5 * (e ** -2.0).sum() for e in es]
    P, A, M = gpu_ctx.gls_batch(np.concatenate(ts), np.concatenate(ys), W, offsets, [0.5 * d for d in dfs], dfs, 900,
                                psd_scale=psd)
    for b in range(2):
        ref = cport.gls_exact(ts[b], ys[b], es[b], 0.5 * dfs[b], dfs[b], 900, True, psd=True)
        assert_power_close(P[b], ref)
        assert A[b] == np.nanargmax(ref)


def test_million_point_curve_window_vs_oracle(gpu_ctx):
    """C5-sized sample axis (1e6 points): 3e4 frequencies around the injected line vs the C oracle on a subset."""
    rng = np.random.default_rng(5)
    n = 1_000_000
    t = np.sort(rng.uniform(0, 1000.0, n))
    y = 1000 + np.sin(2 * np.pi * 17.123 * t + 0.3) + rng.standard_normal(n)
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    j0 = int((17.123 - fmin) / df) - 15_000                 # a shard of the 1e7-point C5 grid
    nf = 30_000
    p, am, mx = gpu_ctx.gls(t, y, None, fmin, df, nf, j0=j0)
    assert abs((fmin + (j0 + am) * df) - 17.123) < df
    sel = np.unique(np.concatenate([np.arange(0, nf, 2999), np.arange(am - 6, am + 7)]))
    ref = cport.gls_exact_at(t, y, None, fmin, df, j0 + sel)
    assert np.max(np.abs(p[sel] - ref)) <= TOL * np.nanmax(p)
    big = ref >= 1e-2 * np.nanmax(p)
    assert np.max(np.abs(p[sel][big] - ref[big]) / ref[big]) <= TOL
    assert sel[np.argmax(ref)] == am


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("S", [1, 5, 8, 19])
def test_shared_time_multi_series_matches_single_calls(gpu_ctx, S, weighted):
    """pdc_gls_multi (series on common timestamps) == pdc_gls per series == formula oracle."""
    rng = np.random.default_rng(100 + S)
    n, nf = 1777, 2100
    t = np.sort(rng.uniform(0, 90, n))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    Y = np.stack([50 * s + np.sin(2 * np.pi * t / rng.uniform(0.7, 9)) * rng.uniform(0.5, 3) +
                  rng.standard_normal(n) for s in range(S)])
    err = rng.uniform(0.5, 1.5, n) if weighted else None
    w = None if err is None else err ** -2.0
    P, A, M = gpu_ctx.gls_multi(t, Y, w, fmin, df, nf)
    _, A2, M2 = gpu_ctx.gls_multi(t, Y, w, fmin, df, nf, want_power=False)
    np.testing.assert_array_equal(A, A2)
    np.testing.assert_array_equal(M, M2)
    for s in range(S):
        ref = cport.gls_exact(t, Y[s], err, fmin, df, nf)
        assert_power_close(P[s], ref)
        assert A[s] == np.nanargmax(ref) and M[s] == np.nanmax(P[s])
        p1, a1, _ = gpu_ctx.gls(t, Y[s], w, fmin, df, nf)
        assert a1 == A[s] and np.nanmax(np.abs(p1 - P[s])) <= 2e-6 * M[s]


def test_shared_time_multi_series_shard_offset_and_psd(gpu_ctx):
    rng = np.random.default_rng(77)
    n, nf = 900, 1500
    t = np.sort(rng.uniform(0, 40, n))
    df = 1 / (t[-1] - t[0]) / 5
    Y = np.stack([np.sin(2 * np.pi * t / 3.3) + 0.5 * rng.standard_normal(n) for _ in range(3)])
    err = rng.uniform(0.2, 0.4, n)
    w = err ** -2.0
    P, A, M = gpu_ctx.gls_multi(t, Y, w, 0.5 * df, df, nf - 700, j0=700, psd_scale=0.5 * w.sum())
    for s in range(3):
        ref = cport.gls_exact(t, Y[s], err, 0.5 * df, df, nf - 700, True, psd=True, j0=700)
        assert_power_close(P[s], ref)
        assert A[s] == np.nanargmax(ref)


def test_bootstrap_uniform_errors_uses_shared_time_kernel():
    from periodicity_b200 import GLS, TSeries
    rng = np.random.default_rng(5)
    t = np.sort(rng.uniform(0, 30, 400))
    y = np.sin(2 * np.pi * t / 2.0) + 0.5 * rng.standard_normal(400)
    gls = GLS(fmax=3.0)
    gls(TSeries(t, y))
    reps = gls.bootstrap(10, random_seed=9, batch=4)
    rng2 = np.random.default_rng(9)
    want = [GLS(fmax=3.0)(TSeries(t, y[rng2.integers(0, 400, 400)])).amax() for _ in range(10)]
    np.testing.assert_allclose(reps, want, rtol=5e-6)


def test_calls_on_different_streams_share_a_ctx_safely(gpu_ctx):
    """A ctx's scratch is shared: calls issued on different CUDA streams must still be ordered."""
    import torch
    from periodicity_b200 import dist as pdist
    nf = 20_000
    t, y, fmin, df = synth(20_000, 400.0, nf, 1.0, 61)
    t2, y2, fmin2, df2 = synth(15_000, 300.0, nf, 1.0, 62)
    want1 = gpu_ctx.gls(t, y, None, fmin, df, nf)[0]
    want2 = gpu_ctx.gls(t2, y2, None, fmin2, df2, nf)[0]
    d = [torch.from_numpy(a).cuda() for a in (t, y, t2, y2)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        with torch.cuda.stream(s1):
            p1 = pdist.gls_torch(d[0], d[1], None, fmin, df, nf, ctx=gpu_ctx)[0]
        with torch.cuda.stream(s2):
            p2 = pdist.gls_torch(d[2], d[3], None, fmin2, df2, nf, ctx=gpu_ctx)[0]
        outs.append((p1, p2))
    torch.cuda.synchronize()
    for p1, p2 in outs:
        np.testing.assert_array_equal(p1.cpu().numpy(), want1)
        np.testing.assert_array_equal(p2.cpu().numpy(), want2)


@pytest.mark.parametrize("n_per_peak,weighted", [(1.0, False), (2.0, True), (2.9, False), (3.0, True), (3.1, False),
                                                 (10.0, True), (40.0, False)])
def test_grid_density_selects_strip_step_form(gpu_ctx, n_per_peak, weighted):
    """df * baseline <= 0.34 turns uses the three-term recurrence along the frequency axis, coarser grids
    (n < 3 samples per peak, spectral.py:88) the plain rotation: both against the formula oracle."""
    N, nf = 3000, 6000
    rng = np.random.default_rng(21)
    t = np.sort(rng.uniform(0, 50.0, N))
    df = 1 / (t[-1] - t[0]) / n_per_peak
    fmin = 0.5 * df
    y = 3 + np.sin(2 * np.pi * (fmin + 0.41 * nf * df) * t) + 0.7 * rng.standard_normal(N)
    err = rng.uniform(0.5, 1.5, N) if weighted else None
    p, am, _ = gpu_ctx.gls(t, y, None if err is None else err ** -2.0, fmin, df, nf)
    ref = cport.gls_exact(t, y, err, fmin, df, nf, True)
    assert_power_close(p, ref)
    assert am == np.nanargmax(ref)


def test_three_term_and_rotation_forms_agree(monkeypatch):
    """PDC_GLS_THREE_TERM=0 (read at ctx creation) forces the rotation form: same periodogram within tolerance."""
    from periodicity_b200 import _ffi
    t, y, fmin, df = synth(20_000, 400.0, 30_000, 1.0, 22)
    w = np.random.default_rng(23).uniform(0.5, 2.0, t.size)
    a = _ffi.Context(0)
    monkeypatch.setenv("PDC_GLS_THREE_TERM", "0")
    b = _ffi.Context(0)
    for ww in (None, w):
        pa, ia, _ = a.gls(t, y, ww, fmin, df, 30_000)
        pb, ib, _ = b.gls(t, y, ww, fmin, df, 30_000)
        assert ia == ib
        assert np.max(np.abs(pa - pb)) <= 2e-6 * np.max(pb)
    lo = ia - 200
    ref = cport.gls_exact(t, y, None if ww is None else ww ** -0.5, fmin + lo * df, df, 400, True)
    assert_power_close(pa[lo:lo + 400], ref)


def test_grid_crossing_zero_frequency(gpu_ctx):
    """Negative and sub-cycle frequencies in one grid (three-term form, FP64 low-frequency bins)."""
    t, y, fmin, df = synth(2000, 80.0, 4000, 0.5, 24)
    p2, am2, _ = gpu_ctx.gls(t, y, None, -1000.25 * df, df, 4000)
    ref2 = cport.gls_exact(t, y, None, -1000.25 * df, df, 4000, True)
    assert_power_close(p2, ref2)
    assert am2 == np.nanargmax(ref2)


def test_batch_mixes_step_forms_per_curve(gpu_ctx):
    """One batch whose curves have different grid densities: the strip kernel picks the three-term or the
    rotation step per curve (block-uniform), weighted and unweighted."""
    rng = np.random.default_rng(31)
    sizes = [900, 1500, 700, 1200]
    dens = [5.0, 1.5, 3.0, 2.0]          # samples per peak: 5 and 3 -> three-term, 1.5 and 2 -> rotation
    nf = 1500
    ts, ys, ws, fm, dfs = [], [], [], [], []
    for n, d_ in zip(sizes, dens):
        t = np.sort(rng.uniform(0, 40.0, n))
        d = 1 / (t[-1] - t[0]) / d_
        ts.append(t)
        ys.append(2 + np.sin(2 * np.pi * (0.5 * d + 0.37 * nf * d) * t) + rng.standard_normal(n))
        ws.append(rng.uniform(0.5, 2, n))
        dfs.append(d)
        fm.append(0.5 * d)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    T, Y, W = np.concatenate(ts), np.concatenate(ys), np.concatenate(ws)
    for w_all, w_list in ((None, [None] * 4), (W, ws)):
        P, A, M = gpu_ctx.gls_batch(T, Y, w_all, offsets, fm, dfs, nf)
        for b in range(4):
            err = None if w_list[b] is None else w_list[b] ** -0.5
            ref = cport.gls_exact(ts[b], ys[b], err, fm[b], dfs[b], nf)
            assert_power_close(P[b], ref)
            assert A[b] == np.nanargmax(ref)
            single, a1, _ = gpu_ctx.gls(ts[b], ys[b], w_list[b], fm[b], dfs[b], nf)
            assert np.max(np.abs(single - P[b])) <= 2e-6 * np.max(single)   # same kernel, different sample split
            assert a1 == A[b]


def test_shared_time_multi_series_coarse_grid_uses_rotation(gpu_ctx):
    """pdc_gls_multi with n = 2 samples per peak (rotation form) and 20 series (three groups: window sums come
    from the first group only)."""
    rng = np.random.default_rng(32)
    n, S, nf = 1500, 20, 2000
    t = np.sort(rng.uniform(0, 30.0, n))
    for dens in (2.0, 5.0):
        df = 1 / (t[-1] - t[0]) / dens
        fmin = 0.5 * df
        Y = 1 + np.sin(2 * np.pi * t[None, :] / rng.uniform(0.3, 3.0, S)[:, None]) + rng.standard_normal((S, n))
        w = rng.uniform(0.5, 2.0, n)
        for ww in (None, w):
            P, A, M = gpu_ctx.gls_multi(t, Y, ww, fmin, df, nf)
            for s in (0, 7, 8, 19):
                ref = cport.gls_exact(t, Y[s], None if ww is None else ww ** -0.5, fmin, df, nf)
                assert_power_close(P[s], ref)
                assert A[s] == np.nanargmax(ref)


@pytest.mark.gpu
def test_large_batch_is_uploaded_in_pipelined_runs(monkeypatch):
    """Host entry pdc_gls_batch cuts survey-sized batches into runs of whole curves whose uploads overlap the
    kernels of the previous run (PDC_BATCH_PIPE_BYTES, read at ctx creation, sets the size where that starts):
    same peaks and periodograms as the one-shot path, with weights, PSD scaling and a non-zero first offset."""
    from periodicity_b200 import _ffi
    rng = np.random.default_rng(77)
    B, nf = 45, 640
    sizes = rng.integers(30, 1500, B)
    sizes[7] = 5000                                  # one curve much longer than the rest
    ts, ys, ws, fm, dfs = [], [], [], [], []
    for n in sizes:
        t = np.sort(rng.uniform(0, rng.uniform(20, 60), n))
        ts.append(t)
        ys.append(np.sin(2 * np.pi * t / rng.uniform(0.5, 5)) + rng.standard_normal(n))
        ws.append(rng.uniform(0.5, 2, n))
        d = 1 / (t[-1] - t[0]) / 5
        dfs.append(d); fm.append(0.5 * d)
    lead = 123                                       # offsets[0] != 0: the batch starts inside the caller's arrays
    T = np.concatenate([np.zeros(lead)] + ts)
    Y = np.concatenate([np.zeros(lead)] + ys)
    W = np.concatenate([np.ones(lead)] + ws)
    offsets = lead + np.concatenate([[0], np.cumsum(sizes)])
    one_shot = _ffi.Context(0)
    monkeypatch.setenv("PDC_BATCH_PIPE_BYTES", "1")
    piped = _ffi.Context(0)
    for w_all, psd in ((None, False), (W, False), (W, True)):
        scale = rng.uniform(0.5, 2.0, B) if psd else None
        P1, A1, M1 = one_shot.gls_batch(T, Y, w_all, offsets, fm, dfs, nf, psd_scale=scale)
        P2, A2, M2 = piped.gls_batch(T, Y, w_all, offsets, fm, dfs, nf, psd_scale=scale)
        _, A3, M3 = piped.gls_batch(T, Y, w_all, offsets, fm, dfs, nf, psd_scale=scale, want_power=False)
        np.testing.assert_array_equal(A1, A2)
        np.testing.assert_array_equal(A2, A3)
        np.testing.assert_array_equal(M2, M3)
        for b in range(B):
            assert np.max(np.abs(P1[b] - P2[b])) <= 2e-6 * np.max(np.abs(P1[b]))
    for b in (0, 7, B - 1):                           # and against the oracle
        ref = cport.gls_exact(ts[b], ys[b], None, fm[b], dfs[b], nf)
        P, A, _ = piped.gls_batch(T, Y, None, offsets, fm, dfs, nf)
        assert_power_close(P[b], ref)
        assert A[b] == np.nanargmax(ref)


def test_c4_tess_recipe_batch_vs_oracle_and_reference_peaks(gpu_ctx):
    """BASELINE config C4 at the bench recipe (bench.make_gls_c4: TESS-like 2-min cadence, 20,000 points x 1e4
    frequencies per curve, per-curve grid): 32 curves through pdc_gls_batch.  Power vs the formula oracle on three
    curves, peak index vs the reference's own algorithm (spectral.py:11-40 as shipped) on all of them."""
    import bench
    wl = bench.make_gls_c4(32)
    B, nf, off = 32, wl["nf"], wl["offsets"]
    P, A, M = gpu_ctx.gls_batch(wl["t"], wl["y"], None, off, wl["fmin"], wl["df"], nf)
    _, A2, M2 = gpu_ctx.gls_batch(wl["t"], wl["y"], None, off, wl["fmin"], wl["df"], nf, want_power=False)
    np.testing.assert_array_equal(A, A2)
    np.testing.assert_array_equal(M, M2)
    for b in range(B):
        tb, yb = wl["t"][off[b]:off[b + 1]], wl["y"][off[b]:off[b + 1]]
        fast = gls_numpy.gls_power(tb, yb, None, wl["fmin"][b], wl["df"][b], nf, True, False)
        assert A[b] == np.nanargmax(P[b]) == np.nanargmax(fast), f"curve {b}"
        assert M[b] == np.nanmax(P[b])
        if b in (0, 13, 31):
            ref = cport.gls_exact(tb, yb, None, wl["fmin"][b], wl["df"][b], nf)
            assert_power_close(P[b], ref)
            assert A[b] == np.nanargmax(ref)


@pytest.mark.slow
def test_c5_full_grid_peak_index_vs_reference_algorithm_and_strided_oracle(gpu_ctx):
    """BASELINE config C5 in full: 1e6 points x 1e7 frequencies on one GPU (~3 s).  north_star: "matching the
    reference's peak index on every config" -- the arg-max over the WHOLE grid equals the arg-max of the reference's
    own FFT-extirpolation algorithm (nfft = 2^26, ~4 GB, ~20 s on the host), and the power agrees with the formula
    oracle on a strided subset of 1e4 frequencies plus a window round the peak."""
    import bench
    wl = bench.make_gls_c5(10_000_000)
    t, y, fmin, df, nf = wl["t"], wl["y"], wl["fmin"], wl["df"], wl["nf"]
    p, am, mx = gpu_ctx.gls(t, y, None, fmin, df, nf)
    assert am == np.nanargmax(p) and mx == p[am]
    assert abs((fmin + am * df) - 17.123) < df
    sel = np.unique(np.concatenate([np.arange(0, nf, 1000), np.arange(am - 8, am + 9), [nf - 1]]))
    ref = cport.gls_exact_at(t, y, None, fmin, df, sel)
    peak = np.nanmax(p)
    assert np.max(np.abs(p[sel] - ref)) <= TOL * peak
    big = ref >= 1e-2 * peak
    assert np.max(np.abs(p[sel][big] - ref[big]) / ref[big]) <= TOL
    assert sel[np.argmax(ref)] == am
    fast = gls_numpy.gls_power(t, y, None, fmin, df, nf, True, False)      # the reference's algorithm as shipped
    assert int(np.nanargmax(fast)) == am


@pytest.mark.parametrize("n_per_peak,weighted", [(100.0, False), (100.0, True), (30.0, False)])
def test_dense_grid_has_many_sub_cycle_bins(gpu_ctx, n_per_peak, weighted):
    """GLS(n=100) puts ~100 bins below one cycle over the baseline, where CC - C^2 cancels ~300x and FP32
    sums are not enough (gls_common.cuh).  The FP64 range is sized from the actual count (round 1 stopped at 16)."""
    N, nf = 4000, 3000
    rng = np.random.default_rng(33)
    t = np.sort(rng.uniform(0, 80.0, N))
    df = 1 / (t[-1] - t[0]) / n_per_peak
    fmin = 0.5 * df
    y = 3 + np.sin(2 * np.pi * (fmin + 0.61 * nf * df) * t) + 0.7 * rng.standard_normal(N)
    err = rng.uniform(0.5, 1.5, N) if weighted else None
    p, am, _ = gpu_ctx.gls(t, y, None if err is None else err ** -2.0, fmin, df, nf)
    ref = cport.gls_exact(t, y, err, fmin, df, nf, True)
    assert_power_close(p, ref)
    low = int(n_per_peak)                                     # the sub-cycle bins themselves, elementwise
    assert np.max(np.abs(p[:low] - ref[:low]) / np.maximum(np.abs(ref[:low]), 1e-3 * np.max(ref))) <= 1e-5
    assert am == np.nanargmax(ref)
