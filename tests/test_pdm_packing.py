"""CPU check of the integer scheme pdm_hist_kernel relies on (periodicity_b200/csrc/pdm.cu), restated in numpy with the
same 32-bit wrap-around arithmetic: one packed word per bin, (count << SUM_BITS) + sum of rint(x' 2^q), fed by adds modulo
2^32, unpacked once per feed window into exact integer (count, sum) planes.  Two layouts exist in pdm.cu: the shipped one,
PDM_PACK_FLUSH = 512 (10 count bits + 22 sum bits; a window is fed whole only when the block has verified
sum |increment| < 2^21 over it, otherwise in guaranteed 128-sample pieces), and the round-1 one, 256 samples (9 + 23 bits).

These are the invariants the kernel's comments claim: the unpacking is exact for any sign pattern up to the stated
bound, an update can be undone by adding the negated increment (the deferred bin-edge fix-up), and the exponent chosen
by pdm_stats3_kernel keeps a 256-sample window inside the 23-bit sum field.
"""
import numpy as np
import pytest

LAYOUTS = [(512, 10), (256, 9)]   # (PDM_PACK_FLUSH, PDM_CNT_BITS); the first is the shipped build
PACK_FLUSH, CNT_BITS = LAYOUTS[0]
PACK_SUB = 128            # PDM_PACK_SUB: guaranteed sub-window of the 512-sample layout
MIN_Q = 11                # PDM_PACK_MIN_Q


def pack_q(xmax):
    """pdm_stats3_kernel: q = floor(log2(16000 / max|x'|)), capped at 20, unusable below MIN_Q."""
    q = int(np.floor(np.log2(16000.0 / xmax)))
    q = min(q, 20)
    return q if q >= MIN_Q else -1


def increments(xs, q, cnt_bits=CNT_BITS):
    """pdm_center_kernel: (1 << SUM_BITS) + (unsigned)(int)rint(x' * 2^q) as uint32 (two's complement wrap)."""
    fix = np.rint(xs * float(1 << q)).astype(np.int64)
    return ((1 << (32 - cnt_bits)) + fix).astype(np.uint64).astype(np.uint32), fix


def unpack(w, cnt_bits=CNT_BITS):
    """flush32: sum = low SUM_BITS bits sign-extended, count = (w - sum) >> SUM_BITS, all in 32-bit arithmetic."""
    w = np.uint32(w)
    sfix = np.int32(np.uint32(w << np.uint32(cnt_bits))) >> np.int32(cnt_bits)
    cnt = np.uint32(w - np.uint32(sfix)) >> np.uint32(32 - cnt_bits)
    return int(cnt), int(sfix)


def feed_pieces(fix, flush, cnt_bits):
    """Piece length pdm_hist_kernel feeds a window of increments `fix` in: the whole window if it is the 256-sample
    layout or if sum |increment| < 2^(SUM_BITS - 1) (checked by the block while staging the tile), else PACK_SUB."""
    if flush <= 256 or np.abs(fix).sum() < (1 << (32 - cnt_bits - 1)):
        return flush
    return PACK_SUB


@pytest.mark.parametrize("flush,cnt_bits", LAYOUTS)
@pytest.mark.parametrize("xmax", [1.0, 3.3, 7.8, 0.02])
def test_a_flush_window_unpacks_exactly_for_any_signs(xmax, flush, cnt_bits):
    """Whatever the sign pattern -- including a whole window of extreme values, which the 512-sample layout must
    detect and feed in 128-sample pieces -- every piece unpacks to the exact count and the exact integer sum."""
    rng = np.random.default_rng(int(xmax * 1000))
    q = pack_q(xmax)
    assert q >= MIN_Q
    with np.errstate(over="ignore"):
        for pattern in ("random", "gaussian", "all_max", "all_min", "alternating"):
            n = flush if pattern != "random" else int(rng.integers(1, flush + 1))
            xs = {"random": rng.uniform(-xmax, xmax, n), "all_max": np.full(n, xmax), "all_min": np.full(n, -xmax),
                  "gaussian": np.clip(rng.standard_normal(n) * xmax / 4.5, -xmax, xmax),
                  "alternating": xmax * (-1.0) ** np.arange(n)}[pattern]
            inc, fix = increments(xs, q, cnt_bits)
            piece = feed_pieces(fix, flush, cnt_bits)
            if pattern == "gaussian":
                assert piece == flush                       # ordinary data is fed whole: half as many feeds as round 1
            if pattern in ("all_max", "all_min") and flush > 256:
                assert piece == PACK_SUB                    # a burst of extremes is caught by the window check
            for a in range(0, n, piece):
                w = np.uint32(0)
                for v in inc[a:a + piece]:      # ATOMS.ADD: addition modulo 2^32
                    w = np.uint32(w + v)
                cnt, sfix = unpack(w, cnt_bits)
                assert cnt == len(inc[a:a + piece]) and sfix == int(fix[a:a + piece].sum())


def test_worst_case_window_stays_inside_the_sum_field():
    # q is chosen so that 256 * max|x'| * 2^q < 2^22 (9 + 23 bits) == 128 * max|x'| * 2^q < 2^21 (10 + 22 bits, guaranteed
    # sub-window) for every max|x'| (the bound quoted in pdm.cu and DESIGN 4.4)
    for xmax in np.geomspace(1e-3, 7.8, 200):
        q = pack_q(xmax)
        if q >= 0:
            assert 256 * np.rint(xmax * 2.0 ** q) < 2 ** 22
            assert PACK_SUB * np.rint(xmax * 2.0 ** q) < 2 ** 21
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
