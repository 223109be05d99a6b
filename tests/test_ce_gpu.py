"""pdc_ce on the GPU against the numpy oracle (oracle/ce_numpy.py, itself pinned to np.histogram2d).

PARITY UNPINNED BY THE REFERENCE: conditional entropy is a TODO there (phase.py:13), no code exists.  The oracle uses the
reference's phase and bin-edge conventions (phase.py:131,138-140).  The kernel histograms integer COUNTS, so the only
differences from the oracle are the order of the float64 sum over cells and c ln c vs p ln p: asserted <= 1e-9 relative
(values are O(1)), identical arg-min.
"""
import numpy as np
import pytest

from oracle import ce_numpy

pytestmark = pytest.mark.gpu


def assert_close(h, ref):
    np.testing.assert_array_equal(np.isnan(h), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.max(np.abs(h[ok] - ref[ok])) <= 1e-9 * max(1.0, np.max(np.abs(ref[ok])))
    assert np.nanargmin(h) == np.nanargmin(ref)


def synth(n, T, seed, period=3.3, noise=0.3):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, n))
    x = 12 + np.sin(2 * np.pi * t / period) + 0.4 * np.sin(4 * np.pi * t / period) + noise * rng.standard_normal(n)
    return t, x


@pytest.mark.parametrize("n,nphi,nm", [(1500, 10, 5), (1500, 8, 4), (5000, 16, 3), (20_000, 10, 5), (300, 5, 2),
                                        (7000, 32, 8), (2000, 1, 1), (4000, 100, 12)])
def test_against_oracle(gpu_ctx, n, nphi, nm):
    t, x = synth(n, 200.0, n + nphi)
    periods = np.linspace(0.8, 9.0, 777)
    h, am, mn = gpu_ctx.ce(t, x, periods, nphi, nm)
    ref = ce_numpy.ce(t, x, periods, nphi, nm)
    assert_close(h, ref)
    assert am == np.nanargmin(ref) and mn == h[am]
    if nphi * nm > 1:
        assert abs(periods[am] - 3.3) < 0.05 or abs(periods[am] - 6.6) < 0.1 or n <= 300
    assert np.all(h[~np.isnan(h)] >= -1e-12) and np.all(h[~np.isnan(h)] <= np.log(nm) + 1e-12)   # 0 <= H(m | phi) <= ln nm


def test_c3_shape_reduced_uses_sample_splits(gpu_ctx):
    """1e5 samples x 2,000 periods: few period blocks, so the sample axis is split across blocks that all add into
    the same count plane."""
    t, x = synth(100_000, 1000.0, 3, period=3.7, noise=1.0)
    periods = np.linspace(1.0, 11.0, 100_000)[::50]
    h, am, mn = gpu_ctx.ce(t, x, periods, 10, 5)
    ref = ce_numpy.ce(t, x, periods, 10, 5)
    assert_close(h, ref)
    h2, am2, mn2 = gpu_ctx.ce(t, x, periods, 10, 5)              # integer counts: bit-reproducible, plane left clean
    np.testing.assert_array_equal(h, h2)
    assert (am, mn) == (am2, mn2)


def test_integer_times_rational_periods_bin_edges_exact(gpu_ctx):
    """Samples exactly on phase-bin edges must land where the reference's >= / < thresholds put them."""
    rng = np.random.default_rng(5)
    t = np.arange(6000.0)
    x = np.sin(2 * np.pi * t / 12.0) + 0.3 * rng.standard_normal(t.size)
    periods = np.array([2.0, 2.5, 4.0, 5.0, 8.0, 10.0, 12.0, 12.5, 16.0, 20.0, 25.0, 40.0, 3.0, 6.0, 7.0, 9.6])
    for nphi, nm in ((10, 5), (8, 4), (5, 3)):
        h, am, _ = gpu_ctx.ce(t, x, periods, nphi, nm)
        assert_close(h, ce_numpy.ce(t, x, periods, nphi, nm))


def test_negative_and_julian_date_times(gpu_ctx):
    t, x = synth(6000, 300.0, 9)
    periods = np.linspace(1.0, 8.0, 400)
    for shift in (-150.0, 2_450_000.0, 3.0e12):   # JD stamps: shifted fixed-point phase; 3e12: exact FP64 path for every sample
        h, am, _ = gpu_ctx.ce(t + shift, x, periods, 10, 5)
        assert_close(h, ce_numpy.ce(t + shift, x, periods, 10, 5))


def test_degenerate_periods_and_non_finite_samples(gpu_ctx):
    t, x = synth(5000, 100.0, 11)
    periods = np.array([0.0, 1.0, np.inf, np.nan, 2.0, 1e-320, -3.0, 3.3])
    h, am, mn = gpu_ctx.ce(t, x, periods, 10, 5)
    assert np.isnan(h[[0, 2, 3, 5]]).all() and np.isfinite(h[[1, 4, 6, 7]]).all()
    with np.errstate(all="ignore"):
        ref = ce_numpy.ce(t, x, periods[[1, 4, 6, 7]], 10, 5)
    np.testing.assert_allclose(h[[1, 4, 6, 7]], ref, rtol=1e-9)
    assert am == 7 and mn == h[7]
    base, _, _ = gpu_ctx.ce(t, x, periods[[1, 4, 7]], 10, 5)
    tb = t.copy()
    tb[[5, 777]] = [np.nan, np.inf]
    hb, _, _ = gpu_ctx.ce(tb, x, periods[[1, 4, 7]], 10, 5)       # those two samples are in no cell
    keep = np.isfinite(tb)
    np.testing.assert_allclose(hb, ce_numpy.ce(t[keep], x[keep], periods[[1, 4, 7]], 10, 5), rtol=1e-9)
    again, _, _ = gpu_ctx.ce(t, x, periods[[1, 4, 7]], 10, 5)     # the ctx (and its count plane) is still healthy
    np.testing.assert_array_equal(again, base)
    const, _, _ = gpu_ctx.ce(t, np.full_like(x, 3.0), periods[[1, 4]], 10, 5)   # max == min: no magnitude bins
    assert np.isnan(const).all()


def test_dropin_class_torch_entry_and_invalid_arguments(gpu_ctx):
    import torch
    from periodicity_b200 import CE, TSeries
    from periodicity_b200 import dist as pdist
    t, x = synth(3000, 120.0, 13)
    ce = CE(nb=10, nm=5, p_min=1.0, p_max=8.0, n_periods=600)
    pg = ce(TSeries(t, x))
    ref = ce_numpy.ce(t, x, ce.periods, 10, 5)
    np.testing.assert_allclose(pg.values, ref[::-1], rtol=1e-9)           # FSeries sorts by frequency: reversed
    assert abs(1 / pg.frequency[pg.argmin()] - 3.3) < 0.05
    hd, ad, md = pdist.ce_torch(torch.from_numpy(t).cuda(), torch.from_numpy(x).cuda(),
                                torch.from_numpy(ce.periods).cuda(), 10, 5, ctx=gpu_ctx)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(hd.cpu().numpy(), pg.values[::-1])
    assert int(ad.item()) == ce.argmin_index and float(md.item()) == ce.min_entropy
    with pytest.raises(ValueError):
        gpu_ctx.ce(np.arange(5.0), np.arange(4.0), [1.0], 10, 5)
    with pytest.raises(ValueError):
        gpu_ctx.ce(np.arange(5.0), np.arange(5.0), [1.0], 0, 5)
    with pytest.raises(ValueError):
        gpu_ctx.ce(np.arange(5.0), np.arange(5.0), [1.0], 1000, 100)
