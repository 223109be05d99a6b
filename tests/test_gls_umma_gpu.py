"""GPU parity tests of the tensor-core formulation of the GLS sums (csrc/gls_umma.cu, tcgen05): angle addition across
blocks of the frequency grid as fp16 hi/lo GEMMs.  Same oracle and tolerance as tests/test_gls_gpu.py (BASELINE.json
north_star: relative power error <= 1e-5): (i) max|dp| / max p <= 1e-5, (ii) elementwise relative error <= 1e-5 on
bins with p >= 1e-2 max p, identical peak index -- against the formula oracle (reference spectral.py:113-132 with exact
sums, oracle/oracle.c).  The kernel choice is a ctx property (environment at ctx creation): PDC_GLS_UMMA=1 forces the
tensor path whenever a call is eligible, 0 forces gls_strip_kernel; the default is automatic by problem size.

Size-independent properties checked here: the two kernels agree with each other far inside the tolerance; the result
does not depend on whether the fine operand is computed in the kernel or precomputed per curve (bit-identical: same
integer phase arithmetic, same instruction order); repeated calls are bit-identical; a frequency shard (j0 > 0) equals
the slice of the full grid within the decomposition tolerance.
"""
import os

import numpy as np
import pytest

from oracle import cport

pytestmark = pytest.mark.gpu

TOL = 1e-5


def make_ctx(**env):
    from periodicity_b200 import _ffi
    keys = ["PDC_GLS_UMMA", "PDC_GLS_UMMA_FINE", "PDC_GLS_UMMA_CHUNK", "PDC_GLS_UMMA_NSPLIT", "PDC_GLS_UMMA_RZCOMP",
            "PDC_GLS_UMMA_CG2"]
    saved = {k: os.environ.get(k) for k in keys}
    try:
        for k in keys:
            os.environ.pop(k, None)
        for k, v in env.items():
            os.environ[k] = str(v)
        return _ffi.Context(0)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def ctx_t():
    return make_ctx(PDC_GLS_UMMA=1)


@pytest.fixture(scope="module")
def ctx_s():
    return make_ctx(PDC_GLS_UMMA=0)


def synth(N, T, nf, sigma, seed, weighted=False, nper=5):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    df = 1 / (t[-1] - t[0]) / nper
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + sigma * rng.standard_normal(N)
    err = rng.uniform(0.5, 2.0, N) if weighted else None
    return t, y, err, fmin, df


def assert_power_close(p, ref, tol=TOL):
    peak = np.nanmax(np.abs(ref))
    assert np.nanmax(np.abs(p - ref)) <= tol * peak
    big = np.abs(ref) >= 1e-2 * peak
    assert np.nanmax(np.abs(p[big] - ref[big]) / np.abs(ref[big])) <= tol


# (N, nf, weighted, fit_mean, sigma): partial tiles (nf not a multiple of 128 / of a tile), sample counts that are not
# multiples of 8, 16, 64 or 256, one-sample tails, a single coarse tile, several tiles of both types, weak and strong peaks
CASES = [
    (3000, 1600, False, True, 1.0),
    (3000, 1600, True, True, 1.0),
    (5000, 20000, False, True, 1.0),
    (777, 300, False, True, 1.0),
    (20000, 10000, True, True, 1.0),
    (4097, 8321, False, True, 1.0),
    (4096, 8192, False, False, 1.0),
    (1025, 257, True, False, 0.1),
    (257, 4099, False, True, 3.0),
    (40, 2000, False, True, 0.5),
    (9, 1000, True, True, 0.5),
    (16385, 3000, False, True, 5.0),
]


@pytest.mark.parametrize("N,nf,weighted,fit_mean,sigma", CASES)
def test_tensor_core_kernel_vs_oracle_and_strip_kernel(ctx_t, ctx_s, N, nf, weighted, fit_mean, sigma):
    t, y, err, fmin, df = synth(N, 100.0, nf, sigma, N + nf, weighted)
    w = None if err is None else err ** -2.0
    ref = cport.gls_exact(t, y, err, fmin, df, nf, fit_mean=fit_mean)
    p, am, mx = ctx_t.gls(t, y, w, fmin, df, nf, fit_mean=fit_mean)
    assert ctx_t.last_gls_path() in (1, 2, 3)              # the tensor-core kernel really ran
    assert not np.isnan(p).any()
    assert_power_close(p, ref)
    assert am == int(np.nanargmax(ref)) and mx == p[am]
    ps, ams, _ = ctx_s.gls(t, y, w, fmin, df, nf, fit_mean=fit_mean)
    assert ctx_s.last_gls_path() == 0
    assert ams == am
    assert np.nanmax(np.abs(p - ps)) <= 4e-6 * np.nanmax(ref)


def test_small_or_backward_grids_stay_on_the_strip_kernel(ctx_t):
    t, y, err, fmin, df = synth(500, 50.0, 200, 1.0, 3)
    p, am, _ = ctx_t.gls(t, y, None, fmin, df, 200)               # fewer than 256 frequencies: not eligible
    assert ctx_t.last_gls_path() == 0
    assert_power_close(p, cport.gls_exact(t, y, None, fmin, df, 200))
    p, am, _ = ctx_t.gls(t, y, None, fmin + 999 * df, -df, 1000)  # df < 0: the strip kernel's rotation form
    assert ctx_t.last_gls_path() == 0
    ref = cport.gls_exact(t, y, None, fmin, df, 1000)[::-1]
    assert_power_close(p, ref)


def test_automatic_choice_by_problem_size(gpu_ctx):
    t, y, err, fmin, df = synth(1000, 50.0, 10000, 1.0, 5)          # C1 size: 1e7 evaluations -> strip kernel
    gpu_ctx.gls(t, y, None, fmin, df, 10000)
    assert gpu_ctx.last_gls_path() == 0
    t, y, err, fmin, df = synth(8000, 50.0, 30000, 1.0, 6)          # 2.4e8 evaluations -> tensor cores
    p, am, _ = gpu_ctx.gls(t, y, None, fmin, df, 30000)
    assert gpu_ctx.last_gls_path() in (1, 2, 3)
    assert_power_close(p, cport.gls_exact(t, y, None, fmin, df, 30000))


def test_precomputed_fine_operand_is_bit_identical_to_the_in_kernel_one():
    a = make_ctx(PDC_GLS_UMMA=1, PDC_GLS_UMMA_FINE=1, PDC_GLS_UMMA_CG2=0)
    b = make_ctx(PDC_GLS_UMMA=1, PDC_GLS_UMMA_FINE=0, PDC_GLS_UMMA_CG2=0)
    for (N, nf, weighted) in [(5000, 20000, False), (3001, 1500, True), (70000, 3000, False)]:
        t, y, err, fmin, df = synth(N, 100.0, nf, 1.0, 7 + N, weighted)
        w = None if err is None else err ** -2.0
        pa, ama, _ = a.gls(t, y, w, fmin, df, nf)
        assert a.last_gls_path() == 2
        pb, amb, _ = b.gls(t, y, w, fmin, df, nf)
        assert b.last_gls_path() == 1
        np.testing.assert_array_equal(pa, pb)
        assert ama == amb
        pa2, _, _ = a.gls(t, y, w, fmin, df, nf)            # and reproducible call to call
        np.testing.assert_array_equal(pa, pa2)


# (N, nf, weighted): one to several tiles of both types, partial tiles, tile halves of odd size, short and long jobs
CG2_CASES = [(3000, 4096, False), (5000, 20000, True), (4097, 8321, False), (2000, 70001, False), (30000, 16384, True),
             (70000, 33000, False), (130, 50000, False)]


@pytest.mark.parametrize("N,nf,weighted", CG2_CASES)
def test_pair_of_ctas_kernel_vs_oracle(N, nf, weighted):
    """gls_umma2_kernel (tcgen05 cta_group::2: a pair of CTAs per tile of 256 fine indices) against the oracle and against
    the one-CTA kernel; automatic for one curve from 16384 frequencies on, forced here from 4096."""
    ctx2, ctx1 = make_ctx(PDC_GLS_UMMA=1, PDC_GLS_UMMA_CG2=1), make_ctx(PDC_GLS_UMMA=1, PDC_GLS_UMMA_CG2=0)
    t, y, err, fmin, df = synth(N, 100.0, nf, 1.0, 3 * N + nf, weighted)
    w = None if err is None else err ** -2.0
    p2, am2, mx2 = ctx2.gls(t, y, w, fmin, df, nf)
    assert ctx2.last_gls_path() == 3
    assert not np.isnan(p2).any()
    idx = np.unique(np.concatenate([np.arange(0, nf, max(1, nf // 3000)), np.arange(max(am2 - 64, 0), min(am2 + 64, nf)),
                                    np.arange(nf - 300, nf)]))
    ref = cport.gls_exact_at(t, y, err, fmin, df, idx)
    assert_power_close(p2[idx], ref)
    assert idx[int(np.nanargmax(ref))] == am2 and mx2 == p2[am2]
    p1, am1, _ = ctx1.gls(t, y, w, fmin, df, nf)
    assert ctx1.last_gls_path() in (1, 2)
    assert am1 == am2
    assert np.nanmax(np.abs(p1 - p2)) <= 2e-6 * np.nanmax(p1)
    p2b, _, _ = ctx2.gls(t, y, w, fmin, df, nf)
    np.testing.assert_array_equal(p2, p2b)                     # bit-reproducible
    # a shard of the grid (as the multi-GPU paths cut it) through the same kernel
    j0, cnt = nf // 3, nf - nf // 3
    if cnt >= 4096:
        ps, a_, _ = ctx2.gls(t, y, w, fmin, df, cnt, j0=j0)
        assert ctx2.last_gls_path() == 3
        assert np.nanmax(np.abs(ps - p2[j0:])) <= 2e-6 * np.nanmax(p2)


def test_frequency_shards_and_psd(ctx_t):
    N, nf = 6000, 12000
    t, y, err, fmin, df = synth(N, 100.0, nf, 1.0, 11, True)
    w = err ** -2.0
    ref = cport.gls_exact(t, y, err, fmin, df, nf)
    full, am, mx = ctx_t.gls(t, y, w, fmin, df, nf)
    assert_power_close(full, ref)
    peak = np.nanmax(ref)
    for j0, cnt in [(0, 5000), (5000, 4096), (9096, 2904)]:       # shards of the grid, as the multi-GPU paths cut it
        p, a_, _ = ctx_t.gls(t, y, w, fmin, df, cnt, j0=j0)
        assert ctx_t.last_gls_path() in (1, 2, 3)
        assert np.nanmax(np.abs(p - ref[j0:j0 + cnt])) <= TOL * peak
        assert np.nanmax(np.abs(p - full[j0:j0 + cnt])) <= 2e-6 * peak
        assert a_ == int(np.nanargmax(ref[j0:j0 + cnt]))
    # PSD normalisation (spectral.py:130) goes through the same epilogue
    psd, _, _ = ctx_t.gls(t, y, w, fmin, df, nf, psd_scale=0.5 * w.sum())
    assert_power_close(psd, cport.gls_exact(t, y, err, fmin, df, nf, psd=True))


def test_batch_of_ragged_curves_on_tensor_cores(ctx_t, ctx_s):
    rng = np.random.default_rng(21)
    lens = [3000, 4097, 1500, 8192, 2500, 6001]
    nf = 4000
    ts, ys, offs, fmins, dfs = [], [], [0], [], []
    for b, n in enumerate(lens):
        t = np.sort(rng.uniform(0, 30.0 + b, n))
        df = 1 / (t[-1] - t[0]) / 5
        ts.append(t)
        ys.append(np.sin(2 * np.pi * (2.0 + 0.3 * b) * t) + rng.standard_normal(n))
        offs.append(offs[-1] + n)
        fmins.append(0.5 * df)
        dfs.append(df)
    t, y = np.concatenate(ts), np.concatenate(ys)
    offs = np.array(offs, dtype=np.int64)
    p, am, mx = ctx_t.gls_batch(t, y, None, offs, np.array(fmins), np.array(dfs), nf)
    assert ctx_t.last_gls_path() == 1                      # batches compute the fine operand in the kernel
    p = p.reshape(len(lens), nf)
    ps, ams, _ = ctx_s.gls_batch(t, y, None, offs, np.array(fmins), np.array(dfs), nf)
    for b in range(len(lens)):
        ref = cport.gls_exact(ts[b], ys[b], None, fmins[b], dfs[b], nf)
        assert_power_close(p[b], ref)
        assert am[b] == int(np.nanargmax(ref)) == ams[b]


def test_non_finite_input_gives_nan_and_leaves_the_planes_clean(ctx_t):
    t, y, err, fmin, df = synth(4000, 100.0, 6000, 1.0, 31)
    yb = y.copy()
    yb[123] = np.nan
    p, am, mx = ctx_t.gls(t, yb, None, fmin, df, 6000)
    assert np.isnan(p).all() and am == -1
    p, am, mx = ctx_t.gls(t, y, None, fmin, df, 6000)       # the next call on the same ctx is unaffected
    assert_power_close(p, cport.gls_exact(t, y, None, fmin, df, 6000))


def test_dense_grid_with_many_sub_cycle_bins(ctx_t):
    # n = 60 samples per peak: ~60 bins with f T < 1 take their FP64 sums from the small plane, the rest from the tensor path
    t, y, err, fmin, df = synth(3000, 100.0, 9000, 1.0, 41, nper=60)
    p, am, _ = ctx_t.gls(t, y, None, fmin, df, 9000)
    assert ctx_t.last_gls_path() in (1, 2, 3)
    assert_power_close(p, cport.gls_exact(t, y, None, fmin, df, 9000))


@pytest.mark.parametrize("chunk", [2, 4, 8, 16, 32])
def test_accumulation_run_length_and_truncation_compensation(chunk):
    """The TMEM accumulator truncates; the drain compensates the expected loss.  Any run length stays inside the tolerance
    with the compensation; without it the error grows with the run length (and is still inside at the default)."""
    t, y, err, fmin, df = synth(30000, 400.0, 2000, 0.2, 51)
    ref = cport.gls_exact(t, y, None, fmin, df, 2000)
    p, am, _ = make_ctx(PDC_GLS_UMMA=1, PDC_GLS_UMMA_CHUNK=chunk).gls(t, y, None, fmin, df, 2000)
    assert_power_close(p, ref)
    e_comp = np.nanmax(np.abs(p - ref)) / np.nanmax(ref)
    p0, _, _ = make_ctx(PDC_GLS_UMMA=1, PDC_GLS_UMMA_CHUNK=chunk, PDC_GLS_UMMA_RZCOMP=0).gls(t, y, None, fmin, df, 2000)
    e_raw = np.nanmax(np.abs(p0 - ref)) / np.nanmax(ref)
    assert e_comp <= 3e-6
    if chunk >= 16:
        assert e_raw > 1.5 * e_comp                          # the bias is real and the compensation removes most of it


def test_multi_device_ctx_on_the_tensor_path():
    from periodicity_b200 import _ffi
    saved = os.environ.get("PDC_GLS_UMMA")
    os.environ["PDC_GLS_UMMA"] = "1"
    try:
        m = _ffi.Context([0, 0])
        one = _ffi.Context(0)
    finally:
        if saved is None:
            os.environ.pop("PDC_GLS_UMMA", None)
        else:
            os.environ["PDC_GLS_UMMA"] = saved
    t, y, err, fmin, df = synth(20000, 300.0, 60000, 1.0, 61)       # 1.2e9 evaluations: cut over both workers
    pm, amm, mxm = m.gls(t, y, None, fmin, df, 60000)
    p1, am1, mx1 = one.gls(t, y, None, fmin, df, 60000)
    assert amm == am1
    assert np.nanmax(np.abs(pm - p1)) <= 2e-6 * np.nanmax(p1)
    idx = np.arange(0, 60000, 97)
    assert_power_close(pm[idx], cport.gls_exact_at(t, y, None, fmin, df, idx))
