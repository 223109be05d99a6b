"""GPU parity tests for String Length: CUDA path (through the C ABI) vs the oracle and the golden vectors.

Arithmetic is float64 throughout and the sort order is exact (stable by phase), so only the order of the
final summation differs from numpy: relative error <= 1e-12 is asserted, and an identical argmin.
"""
import numpy as np
import pytest

from conftest import SL_CASES, load_golden, opt
from oracle import stringlength_numpy as slo

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.mark.parametrize("case", SL_CASES)
def test_golden_cases_through_dropin_class(case):
    from periodicity_b200 import StringLength, TSeries
    g = load_golden(case)
    t = opt(g["t"])
    sig = g["x"] if t is None else TSeries(t, g["x"])
    sl = StringLength(dphi=float(g["dphi"]), n_periods=int(g["n_periods"]))
    out = sl(sig)
    np.testing.assert_array_equal(out.frequency, g["periodogram_frequency"])
    np.testing.assert_allclose(out.values, g["periodogram_values"], rtol=TOL)
    assert out.argmin() == np.nanargmin(g["periodogram_values"])


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 1000, 1024, 1025, 5000])
def test_sizes_around_the_padding_boundaries(gpu_ctx, n):
    rng = np.random.default_rng(n)
    t = np.sort(rng.uniform(-20, 80, n))
    m = slo.scale(np.sin(t / 3.0) + 0.1 * rng.standard_normal(n)) if n > 1 else np.array([0.25])
    periods = np.linspace(0.7, 23.0, 57)
    ell, am, mn = gpu_ctx.stringlength(t, m, periods)
    ref = slo.string_lengths(t, m, periods)
    np.testing.assert_allclose(ell, ref, rtol=TOL, atol=1e-15)
    assert am == np.nanargmin(ref) and mn == ell[am]


def test_equal_phases_keep_time_order(gpu_ctx):
    """Integer times and integer periods: many samples share a phase; the sort must be stable (core.py:477)."""
    rng = np.random.default_rng(3)
    t = np.arange(400.0)
    m = slo.scale(rng.standard_normal(400))
    periods = np.array([1.0, 2.0, 4.0, 5.0, 8.0, 10.0, 16.0, 20.0, 25.0, 50.0, 100.0, 400.0, 3.0, 7.0])
    ell, am, _ = gpu_ctx.stringlength(t, m, periods)
    ref = slo.string_lengths(t, m, periods)
    np.testing.assert_allclose(ell, ref, rtol=TOL)
    assert am == np.nanargmin(ref)


def test_finds_the_period_of_a_sparse_light_curve(gpu_ctx):
    from periodicity_b200 import StringLength, TSeries
    rng = np.random.default_rng(11)
    t = np.sort(rng.uniform(0, 200, 150))
    x = 5 + np.sin(2 * np.pi * t / 6.3) + 0.05 * rng.standard_normal(150)
    out = StringLength(dphi=0.05, n_periods=4000)(TSeries(t, x))
    best = 1 / out.frequency[out.argmin()]
    assert abs(best - 6.3) < 0.05 or abs(best - 12.6) < 0.1 or abs(best - 18.9) < 0.15


@pytest.mark.parametrize("n,ties", [(40_000, False), (17_000, False), (16_385, True), (70_000, True)])
def test_large_curve_uses_global_scratch(gpu_ctx, n, ties):
    """N > 16384 does not fit the shared-memory sort: same results from the hybrid path (chunks of 8192 records sorted /
    merged in shared memory, only the chunk-spanning stages in global scratch), incl. tied phases (integer times), whose
    order must be that of the stable sort."""
    rng = np.random.default_rng(12)
    t = np.sort(rng.integers(0, 4000, n).astype(np.float64)) if ties else np.sort(rng.uniform(0, 1000, n))
    m = slo.scale(np.sin(2 * np.pi * t / 9.1) + 0.3 * rng.standard_normal(n))
    periods = np.linspace(2.0, 30.0, 24)
    ell, am, _ = gpu_ctx.stringlength(t, m, periods)
    ref = slo.string_lengths(t, m, periods)
    np.testing.assert_allclose(ell, ref, rtol=TOL)
    assert am == np.nanargmin(ref)


def test_determinism_and_device_entry(gpu_ctx):
    import torch
    rng = np.random.default_rng(13)
    t = np.sort(rng.uniform(0, 100, 2000))
    m = slo.scale(np.sin(t) + 0.2 * rng.standard_normal(2000))
    periods = np.linspace(1.0, 20.0, 500)
    a, ia, ma = gpu_ctx.stringlength(t, m, periods)
    b, ib, mb = gpu_ctx.stringlength(t, m, periods)
    np.testing.assert_array_equal(a, b)
    assert (ia, ma) == (ib, mb)
    td, md, pd_ = (torch.from_numpy(v).cuda() for v in (t, m, periods))
    out = torch.empty(500, dtype=torch.float64, device="cuda")
    arg = torch.empty(1, dtype=torch.int64, device="cuda")
    mn = torch.empty(1, dtype=torch.float64, device="cuda")
    gpu_ctx.stringlength_dev(td.data_ptr(), md.data_ptr(), 2000, pd_.data_ptr(), 500, out.data_ptr(), arg.data_ptr(),
                             mn.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), a)
    assert int(arg.item()) == ia and float(mn.item()) == ma


def test_invalid_arguments_raise_value_error(gpu_ctx):
    with pytest.raises(ValueError):
        gpu_ctx.stringlength(np.arange(5.0), np.arange(4.0), [1.0, 2.0])
    with pytest.raises(ValueError):
        gpu_ctx.stringlength(np.arange(5.0), np.arange(5.0), [])
