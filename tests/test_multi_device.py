"""Multi-device contexts behind the C ABI (pdc_ctx_create_multi, csrc/multi.cu): the host entry points shard the grid /
the batch over the ctx's devices inside the library -- one process, no torchrun, no torch.distributed.

The sharded path (slicing, one worker thread per device, results written straight into the caller's buffer, host
reduction of the arg-extremum candidates) is exercised on ANY box: an ordinal may repeat in the device list, so
`[0, 0, 0]` runs three concurrent workers on one GPU.  With >= 2 GPUs visible the same tests also run on distinct
devices.  Checks: each slice is bit-identical to the single-device call for that slice, the whole result agrees with the
single-device full-grid call within its run-to-run decomposition tolerance, and with the C oracle at 1e-5.
"""
import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def device_lists():
    lists = [[0, 0], [0, 0, 0]]
    n = _ngpu()
    if n >= 2:
        lists.append(list(range(min(n, 8))))
        lists.append([1, 0])
    return lists


@pytest.fixture(scope="module", params=device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def mctx(request):
    import os
    from periodicity_b200 import _ffi
    old = os.environ.get("PDC_MULTI_MIN_EVALS")
    os.environ["PDC_MULTI_MIN_EVALS"] = "1"        # read at ctx creation: shard even the small test problems
    try:
        ctx = _ffi.Context(request.param)
    finally:
        if old is None:
            os.environ.pop("PDC_MULTI_MIN_EVALS", None)
        else:
            os.environ["PDC_MULTI_MIN_EVALS"] = old
    assert ctx.device_count == len(request.param)
    yield ctx
    ctx.close()


def slices(n, k):
    L = -(-n // k)
    return [(min(n, d * L), min(n, d * L + L)) for d in range(k)]


def synth(N, T, nf, seed):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    y = 1000 + np.sin(2 * np.pi * (fmin + 0.3137 * nf * df) * t + 0.3) + rng.standard_normal(N)
    return t, y, fmin, df


@pytest.mark.parametrize("weighted", [False, True])
def test_gls_grid_is_sharded_and_matches_single_device_slices(gpu_ctx, mctx, weighted):
    nf = 30_011                                     # not a multiple of the device count
    t, y, fmin, df = synth(20_000, 400.0, nf, 71)
    w = np.random.default_rng(72).uniform(0.5, 2.0, t.size) if weighted else None
    p, am, mx = mctx.gls(t, y, w, fmin, df, nf)
    k = mctx.device_count
    for a, b in slices(nf, k):                      # every slice: the very same single-device call
        ps, _, _ = gpu_ctx.gls(t, y, w, fmin, df, b - a, j0=a)
        np.testing.assert_array_equal(p[a:b], ps)
    full, am1, mx1 = gpu_ctx.gls(t, y, w, fmin, df, nf)
    assert np.max(np.abs(p - full)) <= 2e-6 * mx1 and am == am1 == np.nanargmax(p) and mx == p[am]
    sel = np.unique(np.concatenate([np.arange(0, nf, 997), np.arange(am - 3, am + 4)]))
    ref = cport.gls_exact_at(t, y, None if w is None else w ** -0.5, fmin, df, sel)
    assert np.max(np.abs(p[sel] - ref)) <= TOL * np.nanmax(ref)
    assert sel[np.argmax(ref)] == am
    # shard offset on top of the caller's own j0, peaks only
    _, am2, mx2 = mctx.gls(t, y, w, fmin, df, nf - 5000, j0=5000, want_power=False)
    assert am2 == int(np.nanargmax(p[5000:])) and abs(mx2 - np.nanmax(p[5000:])) <= 2e-6 * mx


def test_gls_nan_bins_and_first_occurrence_in_the_host_reduction(mctx):
    """fmin = 0 puts a NaN (0/0) bin into the first device's slice; the host reduction must skip it."""
    t, y, _, df = synth(3000, 50.0, 4000, 73)
    p, am, mx = mctx.gls(t, y, None, 0.0, df, 4000)
    assert am == np.nanargmax(p) and mx == np.nanmax(p) and np.isfinite(mx)


def test_pdm_aov_stringlength_period_grids_are_sharded(gpu_ctx, mctx):
    rng = np.random.default_rng(74)
    n = 12_000
    t = np.sort(rng.uniform(0, 300.0, n))
    x = 10 + np.sin(2 * np.pi * t / 3.7) + 0.8 * np.sin(4 * np.pi * t / 3.7) + rng.standard_normal(n)
    periods = np.linspace(1.0, 11.0, 5003)
    k = mctx.device_count
    th, ai, av = mctx.pdm(t, x, periods, 10, 2)
    for a, b in slices(periods.size, k):
        np.testing.assert_array_equal(th[a:b], gpu_ctx.pdm(t, x, periods[a:b], 10, 2)[0])
    assert ai == np.nanargmin(th) and av == th[ai]
    ref = cport.pdm(t, x, periods[::25], 10, 2)
    np.testing.assert_allclose(th[::25], ref, rtol=TOL)
    ao, aoi, aov = mctx.aov(t, x, periods, 8)
    np.testing.assert_array_equal(ao, np.concatenate([gpu_ctx.aov(t, x, periods[a:b], 8)[0] for a, b in slices(periods.size, k)]))
    assert aoi == np.nanargmax(ao) and aov == ao[aoi]
    m = (x - x.max()) / (2 * (x.max() - x.min())) + 0.25
    sl_periods = periods[:700]
    ell, li, lv = mctx.stringlength(t[:1500], m[:1500], sl_periods)
    np.testing.assert_array_equal(ell, gpu_ctx.stringlength(t[:1500], m[:1500], sl_periods)[0])   # per-period work: identical
    assert li == np.nanargmin(ell) and lv == ell[li]


def test_batch_and_shared_time_series_are_split_by_curve(gpu_ctx, mctx):
    rng = np.random.default_rng(75)
    sizes = rng.integers(200, 1500, 23)
    nf = 800
    ts, ys, fm, dfs = [], [], [], []
    for n in sizes:
        t = np.sort(rng.uniform(0, rng.uniform(20, 60), n))
        ts.append(t)
        ys.append(np.sin(2 * np.pi * t / rng.uniform(0.5, 5)) + rng.standard_normal(n))
        d = 1 / (t[-1] - t[0]) / 5
        dfs.append(d)
        fm.append(0.5 * d)
    off = np.concatenate([[0], np.cumsum(sizes)])
    T, Y = np.concatenate(ts), np.concatenate(ys)
    W = rng.uniform(0.5, 2.0, T.size)
    for w in (None, W):
        P, A, M = mctx.gls_batch(T, Y, w, off, fm, dfs, nf)
        P1, A1, M1 = gpu_ctx.gls_batch(T, Y, w, off, fm, dfs, nf)
        np.testing.assert_array_equal(A, A1)
        assert np.max(np.abs(P - P1)) <= 2e-6 * np.max(P1)     # a curve's sample split may differ with the batch size
        _, A2, M2 = mctx.gls_batch(T, Y, w, off, fm, dfs, nf, want_power=False)
        np.testing.assert_array_equal(A2, A)
        np.testing.assert_array_equal(M2, M)
    for b in (0, 11, 22):
        ref = cport.gls_exact(ts[b], ys[b], None, fm[b], dfs[b], nf)
        Pn = mctx.gls_batch(T, Y, None, off, fm, dfs, nf)[0]
        assert np.max(np.abs(Pn[b] - ref)) <= TOL * np.max(ref)
    # series on shared timestamps: whole groups of 8 series per device
    n, S = 1500, 21
    t = np.sort(rng.uniform(0, 40.0, n))
    df = 1 / (t[-1] - t[0]) / 5
    Ym = 1 + np.sin(2 * np.pi * t[None, :] / rng.uniform(0.5, 5.0, S)[:, None]) + rng.standard_normal((S, n))
    Pm, Am, Mm = mctx.gls_multi(t, Ym, None, 0.5 * df, df, 1200)
    P1, A1, M1 = gpu_ctx.gls_multi(t, Ym, None, 0.5 * df, df, 1200)
    np.testing.assert_array_equal(Am, A1)
    assert np.max(np.abs(Pm - P1)) <= 2e-6 * np.max(P1)


def test_small_problems_stay_on_one_device_and_errors_propagate(gpu_ctx):
    from periodicity_b200 import _ffi
    ctx = _ffi.Context([0, 0])                       # default threshold: ~5e8 evaluations per device
    t, y, fmin, df = synth(1000, 100.0, 10_000, 1)   # configs[0]: 1e7 evaluations -> must not be cut
    p, am, mx = ctx.gls(t, y, None, fmin, df, 10_000)
    np.testing.assert_array_equal(p, gpu_ctx.gls(t, y, None, fmin, df, 10_000)[0])
    with pytest.raises(ValueError):
        ctx.pdm(np.arange(5.0), np.arange(5.0), [1.0], 0, 2)
    with pytest.raises(ValueError):
        ctx.gls(np.arange(5.0), np.arange(5.0), None, 0.1, 0.1, 0)
    ctx.close()
    with pytest.raises(ValueError):
        _ffi.Context([])
    with pytest.raises((ValueError, RuntimeError)):
        _ffi.Context([0, 4096])


def test_dropin_classes_take_a_device_list():
    """`GLS(devices=[...])(signal)` / `PDM(devices=[...])(signal)` in an ordinary Python process."""
    from periodicity_b200 import GLS, PDM, TSeries
    devs = list(range(max(2, min(_ngpu(), 8)))) if _ngpu() >= 2 else [0, 0]
    rng = np.random.default_rng(76)
    t = np.sort(rng.uniform(0, 200.0, 40_000))
    y = 5 + np.sin(2 * np.pi * t / 2.75) + rng.standard_normal(t.size)
    one = GLS(fmax=20.0, device=0)(TSeries(t, y))
    many = GLS(fmax=20.0, devices=devs)(TSeries(t, y))
    np.testing.assert_array_equal(many.frequency, one.frequency)
    assert many.argmax() == one.argmax() and np.max(np.abs(many.values - one.values)) <= 2e-6 * one.amax()
    assert abs(many.pmax() - 2.75) < 0.01
    p1 = PDM(nb=10, nc=2, p_min=1.0, p_max=8.0, n_periods=30_000, device=0)(TSeries(t, y))
    pm = PDM(nb=10, nc=2, p_min=1.0, p_max=8.0, n_periods=30_000, devices=devs)(TSeries(t, y))
    assert pm.argmin() == p1.argmin() and np.max(np.abs(pm.values - p1.values) / p1.values) <= 2e-6
