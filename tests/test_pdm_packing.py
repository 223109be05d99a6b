"""CPU check of the integer scheme pdm_hist_kernel relies on (periodicity_b200/csrc/pdm.cu), restated in numpy with the
same 32-bit wrap-around arithmetic: one packed word per bin, (count << 23) + sum of rint(x' 2^q), fed by adds modulo 2^32,
unpacked every PDM_PACK_FLUSH = 256 samples into exact integer (count, sum) planes.

These are the invariants the kernel's comments claim: the unpacking is exact for any sign pattern up to the stated
bound, an update can be undone by adding the negated increment (the deferred bin-edge fix-up), and the exponent chosen
by pdm_stats3_kernel keeps a 256-sample window inside the 23-bit sum field.
"""
import numpy as np
import pytest

PACK_FLUSH = 256          # PDM_PACK_FLUSH
MIN_Q = 11                # PDM_PACK_MIN_Q


def pack_q(xmax):
    """pdm_stats3_kernel: q = floor(log2(16000 / max|x'|)), capped at 20, unusable below MIN_Q."""
    q = int(np.floor(np.log2(16000.0 / xmax)))
    q = min(q, 20)
    return q if q >= MIN_Q else -1


def increments(xs, q):
    """pdm_center_kernel: (1 << 23) + (unsigned)(int)rint(x' * 2^q) as uint32 (two's complement wrap)."""
    fix = np.rint(xs * float(1 << q)).astype(np.int64)
    return ((1 << 23) + fix).astype(np.uint64).astype(np.uint32), fix


def unpack(w):
    """flush32: sum = low 23 bits sign-extended, count = (w - sum) >> 23, all in 32-bit arithmetic."""
    w = np.uint32(w)
    sfix = np.int32(np.uint32(w << np.uint32(9))) >> np.int32(9)
    cnt = np.uint32(w - np.uint32(sfix)) >> np.uint32(23)
    return int(cnt), int(sfix)


@pytest.mark.parametrize("xmax", [1.0, 3.3, 7.8, 0.02])
def test_a_flush_window_unpacks_exactly_for_any_signs(xmax):
    rng = np.random.default_rng(int(xmax * 1000))
    q = pack_q(xmax)
    assert q >= MIN_Q
    with np.errstate(over="ignore"):
        for pattern in ("random", "all_max", "all_min", "alternating"):
            n = PACK_FLUSH if pattern != "random" else int(rng.integers(1, PACK_FLUSH + 1))
            xs = {"random": rng.uniform(-xmax, xmax, n), "all_max": np.full(n, xmax), "all_min": np.full(n, -xmax),
                  "alternating": xmax * (-1.0) ** np.arange(n)}[pattern]
            inc, fix = increments(xs, q)
            w = np.uint32(0)
            for v in inc:                       # ATOMS.ADD: addition modulo 2^32
                w = np.uint32(w + v)
            cnt, sfix = unpack(w)
            assert cnt == n and sfix == int(fix.sum())


def test_worst_case_window_stays_inside_the_sum_field():
    # q is chosen so that 256 * max|x'| * 2^q < 2^22 for every max|x'| (the bound quoted in pdm.cu and DESIGN 4.4)
    for xmax in np.geomspace(1e-3, 7.8, 200):
        q = pack_q(xmax)
        if q >= 0:
            assert PACK_FLUSH * np.rint(xmax * 2.0 ** q) < 2 ** 22
    assert pack_q(16.0) == -1 and pack_q(7.81) == MIN_Q      # outliers beyond ~7.8 sigma: the float2 path is used


def test_an_update_is_undone_by_adding_the_negated_increment():
    # deferred bin-edge fix-up: the increment is moved from the fast bin to the exact bin
    rng = np.random.default_rng(1)
    q = pack_q(4.0)
    inc, fix = increments(rng.uniform(-4, 4, 200), q)
    with np.errstate(over="ignore"):
        a = np.uint32(0)
        b = np.uint32(0)
        for v in inc[:100]:
            a = np.uint32(a + v)
        for v in inc[100:]:
            b = np.uint32(b + v)
        moved = inc[37]                         # sample 37 went to bin a but belongs to bin b
        a = np.uint32(a + np.uint32(0 - moved))
        b = np.uint32(b + moved)
    assert unpack(a) == (99, int(fix[:100].sum() - fix[37]))
    assert unpack(b) == (101, int(fix[100:].sum() + fix[37]))


def test_second_level_integer_planes_hold_a_global_flush_interval():
    # level 2 is int32 fixed point, merged into FP64 every 8192 samples: 8192 * 16000 < 2^31
    assert 8192 * 16000 < 2 ** 31
    # and the quantisation is what DESIGN 4.4 states: |rint(x 2^q) 2^-q - x| <= 2^-(q+1)
    q = pack_q(3.0)
    xs = np.random.default_rng(2).uniform(-3, 3, 10_000)
    _, fix = increments(xs, q)
    assert np.max(np.abs(fix / 2.0 ** q - xs)) <= 2.0 ** -(q + 1)


def test_packed_quantisation_keeps_theta_inside_the_parity_tolerance():
    """The accuracy argument of the packed path: histogramming rint(x' 2^q) 2^-q instead of x' moves
    PDM's theta by ~1e-6 relative on a C3-like curve -- inside the 1e-5 the GPU parity tests assert."""
    from oracle import pdm_numpy
    rng = np.random.default_rng(3)
    n, nb, nc = 100_000, 10, 2
    m0 = nb * nc
    t = np.sort(rng.uniform(0, 1000.0, n))
    x = 1000 + np.sin(2 * np.pi * t / 3.7) + 0.8 * np.sin(4 * np.pi * t / 3.7) + rng.standard_normal(n)
    xp = (x - x.mean()) / x.std(ddof=1)
    q = pack_q(np.abs(xp).max())
    assert q >= MIN_Q
    xq = np.rint(xp * 2.0 ** q) / 2.0 ** q
    idx = (np.arange(m0)[:, None] + np.arange(nc)[None, :]) % m0
    worst = 0.0
    for P in (3.7, 1.85, 7.4, 2.345, 10.1):
        k = pdm_numpy.fine_bin((t / P) % 1, m0)
        N = np.bincount(k, minlength=m0).astype(float)[idx].sum(1)
        S = np.bincount(k, xq, minlength=m0)[idx].sum(1)
        # epilogue of pdm.cu: theta = [nc (N - 1) - sum S_k^2 / n_k] / sum (n_k - 1)
        theta = (nc * (n - 1) - (S ** 2 / N).sum()) / (N - 1).sum()
        ref = pdm_numpy.pdm_theta_hist(t, x, P, nb, nc)
        worst = max(worst, abs(theta - ref) / ref)
    assert worst < 5e-6
