"""GPU parity tests for PDM: CUDA path (through the C ABI) vs the oracle and the golden vectors.

Tolerance: relative theta error <= 1e-5 and identical argmin.
"""
import numpy as np
import pytest

from conftest import PDM_CASES, PDM_KW, load_golden, opt
from oracle import cport

pytestmark = pytest.mark.gpu

TOL = 1e-5


def synth(N, T, seed):
    """SURVEY.md §8d C3 recipe."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    x = 1000 + np.sin(2 * np.pi * t / 3.7) + 0.8 * np.sin(4 * np.pi * t / 3.7) + rng.standard_normal(N)
    return t, x


@pytest.mark.parametrize("case", PDM_CASES)
def test_golden_cases_through_dropin_class(case):
    from periodicity_b200 import PDM, TSeries
    g = load_golden(case)
    kw = {}
    for k in PDM_KW:
        if k in g and opt(g[k]) is not None:
            v = opt(g[k])
            kw[k] = int(v) if k in ("nb", "nc", "n_periods") else (bool(v) if k == "do_subharmonic" else v)
    t = opt(g["t"])
    sig = g["x"] if t is None else TSeries(t, g["x"])
    pdm = PDM(**kw)
    out = pdm(sig)
    np.testing.assert_array_equal(pdm.periods, g["periods"])
    np.testing.assert_array_equal(out.frequency, g["periodogram_frequency"])
    np.testing.assert_allclose(out.values, g["periodogram_values"], rtol=TOL)
    assert out.argmin() == np.nanargmin(g["periodogram_values"])


def test_integer_times_rational_periods_bin_edges_exact(gpu_ctx):
    """Samples exactly on bin edges (t integer, P 'nice'): must bin like the reference's >= / < tests."""
    rng = np.random.default_rng(5)
    t = np.arange(600.0)
    x = np.sin(2 * np.pi * t / 12.0) + 0.3 * rng.standard_normal(600)
    periods = np.array([2.0, 2.5, 4.0, 5.0, 8.0, 10.0, 12.0, 12.5, 16.0, 20.0, 25.0, 40.0, 3.0, 6.0, 7.0, 9.6])
    for nb, nc in ((5, 2), (10, 2), (4, 1), (10, 3), (8, 4)):
        th, am, mn = gpu_ctx.pdm(t, x, periods, nb, nc)
        ref = cport.pdm(t, x, periods, nb, nc)
        np.testing.assert_allclose(th, ref, rtol=2e-6)
        assert am == np.nanargmin(ref)


def test_c3_reduced_vs_c_oracle(gpu_ctx):
    """BASELINE config C3 with 2,000 of the 1e5 trial periods (same N, nb, nc)."""
    t, x = synth(100_000, 1000.0, 3)
    periods = np.linspace(1.0, 11.0, 100_000)[::50]
    th, am, mn = gpu_ctx.pdm(t, x, periods, 10, 2)
    ref = cport.pdm(t, x, periods, 10, 2)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref) and mn == th[am]


def test_c3_full_size_properties(gpu_ctx):
    """Full C3 (1e5 x 1e5): determinism, affine invariance, argmin vs oracle window."""
    t, x = synth(100_000, 1000.0, 3)
    periods = np.linspace(1.0, 11.0, 100_000)
    th, am, mn = gpu_ctx.pdm(t, x, periods, 10, 2)
    th2, am2, mn2 = gpu_ctx.pdm(t, x, periods, 10, 2)
    np.testing.assert_array_equal(th, th2)
    assert (am, mn) == (am2, mn2) and am == np.nanargmin(th)
    th3, am3, _ = gpu_ctx.pdm(t, 2.5 * x - 40.0, periods, 10, 2)
    assert am3 == am and np.max(np.abs(th3 - th) / th) <= 2e-6
    lo, hi = max(0, am - 100), min(periods.size, am + 100)
    ref = cport.pdm(t, x, periods[lo:hi], 10, 2)
    np.testing.assert_allclose(th[lo:hi], ref, rtol=TOL)
    assert lo + np.argmin(ref) == am
    # strided oracle pass over the WHOLE grid (every 50th period + the window above): a wrong global minimum
    # anywhere else would show here, not only next to the GPU's own argmin
    sel = np.arange(0, periods.size, 50)
    ref_s = cport.pdm(t, x, periods[sel], 10, 2)
    np.testing.assert_allclose(th[sel], ref_s, rtol=TOL)
    assert ref_s.min() >= ref.min()                      # no strided period beats the window's minimum
    assert np.nanmin(th) == th[am] and th[am] <= ref_s.min() * (1 + TOL)
    assert abs(periods[am] - 3.7) < 0.01 or abs(periods[am] - 7.4) < 0.02


@pytest.mark.parametrize("nb,nc", [(1, 1), (2, 1), (50, 3), (100, 2), (13, 7)])
def test_bin_geometries(gpu_ctx, nb, nc):
    t, x = synth(5000, 200.0, 12)
    periods = np.linspace(0.7, 9.0, 333)
    th, am, _ = gpu_ctx.pdm(t, x, periods, nb, nc)
    ref = cport.pdm(t, x, periods, nb, nc)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref)


def test_sparse_bins_and_negative_times(gpu_ctx):
    t, x = synth(60, 30.0, 13)
    t = t - 15.0
    periods = np.linspace(0.9, 7.0, 101)
    th, am, _ = gpu_ctx.pdm(t, x, periods, 10, 2)        # many bins have <= 1 sample and are dropped
    ref = cport.pdm(t, x, periods, 10, 2)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref)


def test_few_periods_many_samples_uses_sample_split(gpu_ctx):
    t, x = synth(300_000, 3000.0, 14)
    periods = np.linspace(2.0, 9.0, 40)
    th, am, _ = gpu_ctx.pdm(t, x, periods, 5, 2)
    ref = cport.pdm(t, x, periods, 5, 2)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref)


def test_invalid_arguments_raise_value_error(gpu_ctx):
    with pytest.raises(ValueError):
        gpu_ctx.pdm(np.arange(5.0), np.arange(4.0), [1.0, 2.0], 5, 2)
    with pytest.raises(ValueError):
        gpu_ctx.pdm(np.arange(5.0), np.arange(5.0), [1.0], 0, 2)
    with pytest.raises(ValueError):
        gpu_ctx.pdm(np.arange(5.0), np.arange(5.0), [1.0], 10_000, 10)


def test_torch_device_pointer_entry_matches_host_entry(gpu_ctx):
    import torch
    from periodicity_b200 import dist as pdist
    t, x = synth(20_000, 500.0, 15)
    periods = np.linspace(1.0, 11.0, 3000)
    th, am, mn = gpu_ctx.pdm(t, x, periods, 10, 2)
    td, ad, md = pdist.pdm_torch(torch.from_numpy(t).cuda(), torch.from_numpy(x).cuda(),
                                 torch.from_numpy(periods).cuda(), 10, 2, ctx=gpu_ctx)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(td.cpu().numpy(), th)
    assert int(ad.item()) == am and float(md.item()) == mn


def test_million_samples_crosses_flush_boundaries(gpu_ctx):
    """1e6 samples per period thread: FP32 columns are merged into FP64 every 8192 samples."""
    t, x = synth(1_000_000, 10_000.0, 16)
    periods = np.linspace(2.0, 9.0, 600)
    th, am, _ = gpu_ctx.pdm(t, x, periods, 10, 2)
    sel = np.unique(np.concatenate([np.arange(0, 600, 60), np.arange(max(0, am - 3), min(600, am + 4))]))
    ref = cport.pdm(t, x, periods[sel], 10, 2)
    np.testing.assert_allclose(th[sel], ref, rtol=TOL)
    assert sel[np.argmin(ref)] == am


def test_outlier_heavy_values(gpu_ctx):
    """Values with 1e3-sigma outliers and a large offset: centring/scaling before the FP32 cast must hold."""
    t, x = synth(20_000, 300.0, 17)
    x = x * 1e-3 + 5e6
    x[::997] += 50.0
    periods = np.linspace(1.0, 11.0, 400)
    th, am, _ = gpu_ctx.pdm(t, x, periods, 10, 2)
    ref = cport.pdm(t, x, periods, 10, 2)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref)


def test_degenerate_trial_periods_follow_the_reference_classes(gpu_ctx):
    """P = 0 / denormal / NaN give NaN (phases are inf or NaN, phase.py:131), P = inf gives 1, P < 0 works."""
    t, x = synth(3000, 100.0, 18)
    periods = np.array([0.0, 1.0, np.inf, np.nan, 2.0, 1e-320, -3.0, 3.7])
    th, am, mn = gpu_ctx.pdm(t, x, periods, 5, 2)
    with np.errstate(all="ignore"):
        ref = cport.pdm(t, x, periods, 5, 2)
    np.testing.assert_array_equal(np.isnan(th), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(th[ok], ref[ok], rtol=TOL)
    assert am == np.nanargmin(ref) and mn == th[am]


def test_non_finite_samples_do_not_corrupt_memory(gpu_ctx):
    """NaN / inf time stamps are in no bin (every mask of phase.py:138-140 is false); NaN values poison
    sigma and hence every theta (phase.py:165).  Either way the guarded kernel variant runs."""
    t, x = synth(20_000, 300.0, 19)
    periods = np.linspace(1.0, 11.0, 700)
    base, _, _ = gpu_ctx.pdm(t, x, periods, 10, 2)
    tb = t.copy()
    tb[[5, 777, 19_999]] = [np.nan, np.inf, -np.inf]
    th, am, _ = gpu_ctx.pdm(tb, x, periods, 10, 2)
    with np.errstate(all="ignore"):
        ref = cport.pdm(tb, x, periods, 10, 2)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref)
    xb = x.copy()
    xb[123] = np.nan
    th, am, _ = gpu_ctx.pdm(t, xb, periods, 10, 2)
    assert np.isnan(th).all()
    again, _, _ = gpu_ctx.pdm(t, x, periods, 10, 2)      # the ctx is still healthy and deterministic
    np.testing.assert_array_equal(again, base)


def test_every_bin_dropped_gives_nan_like_the_reference(gpu_ctx):
    """phase.py:142-147: when no coarse bin holds more than one sample the reference divides 0.0 by 0.0 -> NaN
    (nan-aware reductions then skip the period).  The FP32 residue of the numerator must not become +-inf."""
    t = np.array([0.0, 0.31, 0.77, 1.13, 1.52, 2.2, 2.9])
    x = np.array([1.0, 2.5, -0.3, 0.9, 1.7, 0.2, 1.1])
    periods = np.array([2.0, 3.0, 1.9, 0.5, 7.0])
    for nb, nc in ((50, 1), (64, 1), (12, 1), (6, 1)):
        th, am, mn = gpu_ctx.pdm(t, x, periods, nb, nc)
        with np.errstate(all="ignore"):
            ref = cport.pdm(t, x, periods, nb, nc)
        np.testing.assert_array_equal(np.isnan(th), np.isnan(ref))
        assert not np.isinf(th).any()
        ok = ~np.isnan(ref)
        if ok.any():
            np.testing.assert_allclose(th[ok], ref[ok], rtol=TOL)
            assert am == np.nanargmin(ref)
        else:
            assert am == -1 and np.isnan(mn)


def test_outlier_burst_is_fed_in_guaranteed_pieces(gpu_ctx):
    """The packed path feeds 512-sample windows whole only after checking sum |increment| < 2^21; a flare (hundreds of
    consecutive samples several sigma high) fails that check and must be fed in 128-sample pieces -- still exact."""
    t, x = synth(50_000, 600.0, 21)
    x[20_000:21_500] += 6.0              # 1,500 consecutive samples ~4 sigma (of the flared curve) above the rest
    x[40_000:40_700] -= 5.0
    periods = np.linspace(1.0, 11.0, 900)
    th, am, _ = gpu_ctx.pdm(t, x, periods, 10, 2)
    ref = cport.pdm(t, x, periods, 10, 2)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref)


@pytest.mark.parametrize("offset", [2_450_000.0, 2_457_000.5, -1.0e6])
def test_julian_date_stamps_use_the_shifted_fast_path(gpu_ctx, offset):
    """Absolute JD / BJD stamps (|t / P| ~ 2^21 .. 2^24) run the fixed-point phase on t - t0 with a guard band sized
    by |t / P|: samples inside it are re-binned exactly from the ORIGINAL stamp, so the result must equal the
    reference's (t / P) % 1 binning -- including its rounding of the quotient at that magnitude -- on irregular stamps
    and on a regular quarter-day cadence whose phases sit exactly on bin edges."""
    t, x = synth(30_000, 400.0, 23)
    periods = np.concatenate([np.linspace(0.2, 11.0, 1500), [0.25, 0.5, 1.0, 2.0, 2.5, 4.0, 5.0]])
    th, am, mn = gpu_ctx.pdm(t + offset, x, periods, 10, 2)
    ref = cport.pdm(t + offset, x, periods, 10, 2)
    np.testing.assert_allclose(th, ref, rtol=TOL)
    assert am == np.nanargmin(ref)
    tc = offset + 0.25 * np.arange(20_000.0)                            # regular cadence: t / P exactly on bin edges
    xc = np.sin(2 * np.pi * tc / 2.5) + 0.3 * np.random.default_rng(5).standard_normal(tc.size)
    thc, amc, _ = gpu_ctx.pdm(tc, xc, periods[-7:], 10, 2)
    # (a single mis-binned sample of 20,000 would move theta by ~1e-4; the packed path's own quantisation is ~5e-6 here)
    np.testing.assert_allclose(thc, cport.pdm(tc, xc, periods[-7:], 10, 2), rtol=TOL)
