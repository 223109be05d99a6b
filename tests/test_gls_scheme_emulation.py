"""CPU emulation of the numerical scheme of gls_strip_kernel (periodicity_b200/csrc/gls.cu) in float32 numpy, against
the exact float64 sums: the accuracy argument behind the kernel's design, checkable without a GPU.

Emulated (unweighted three-term form): per-sample step angle 2 pi (df (t - tmin) + gamma) with the per-curve phase origin
gamma = 1/4 - span/2 that centres the step angles on a quarter turn; per strip of K = 16 frequencies an exact float64
seed quantised to 2^-32 turn and evaluated in float32, index 1 by one float32 rotation, then the three-term recurrence
c[k+1] = 2 cos(d) c[k] - c[k-1] in float32; six float32 sums {C, S, YC, YS, CC, CS} accumulated sequentially over
tiles of 1024 samples and merged in float64; the reference's tau-offset algebra (spectral.py:113-132) in float64.
Sub-cycle frequencies (|f| (tmax - tmin) < 1) are evaluated in float64 by gls_lowfreq_kernel and are left out here.
"""
import numpy as np
import pytest

from oracle import gls_numpy

K, TILE = 16, 1024
f32 = np.float32


def emulate_strip_sums(t, y, fmin, df, nf):
    tt = t - t.min()
    span = df * tt.max()
    assert span <= 0.34                                    # GLS_TT_MAX_SPAN: the three-term form is selected
    gamma = 0.25 - 0.5 * span
    b = (df * tt) % 1 + gamma
    cr, sr = np.cos(2 * np.pi * b).astype(f32), np.sin(2 * np.pi * b).astype(f32)
    tc = (cr + cr).astype(f32)
    yc = y - y.mean()
    yv = (yc / np.sqrt(np.mean(yc * yc))).astype(f32)      # unit RMS before the float32 cast
    nstrip = -(-nf // K)
    js = np.arange(nstrip) * K
    sums = np.zeros((6, nstrip, K))
    for t0 in range(0, t.size, TILE):
        sl = slice(t0, t0 + TILE)
        # exact seed at the strip's first frequency: phase in turns, quantised to 2^-32, evaluated in float32
        ph = (np.outer(fmin + js * df, tt[sl]) % 1 + (js * gamma % 1)[:, None]) % 1
        fx = np.rint(((ph + 0.5) % 1 - 0.5) * 2.0 ** 32)
        x = (fx.astype(f32) * f32(2 * np.pi / 2.0 ** 32)).astype(np.float64)
        cp, sp = np.cos(x).astype(f32), np.sin(x).astype(f32)
        c = (cp.astype(np.float64) * cr[sl] - (sp * sr[sl]).astype(f32)).astype(f32)       # index 1: one rotation
        s = (sp.astype(np.float64) * cr[sl] + (cp * sr[sl]).astype(f32)).astype(f32)
        for k in range(K):
            cc, ss = (cp, sp) if k == 0 else (c, s)
            terms = (cc, ss, cc * yv[sl], ss * yv[sl], cc * cc, ss * cc)
            for q, term in enumerate(terms):               # sequential float32 accumulation over the tile
                sums[q, :, k] += np.cumsum(term.astype(f32), axis=1, dtype=f32)[:, -1]
            if 1 <= k < K - 1:
                cn = (c.astype(np.float64) * tc[sl] - cp).astype(f32)
                sn = (s.astype(np.float64) * tc[sl] - sp).astype(f32)
                cp, sp, c, s = c, s, cn, sn
    return sums.reshape(6, -1)[:, :nf], yv.astype(np.float64)


@pytest.mark.parametrize("n_per_peak,seed", [(5, 1), (3, 2), (8, 3)])
def test_three_term_float32_strip_meets_the_parity_tolerance(n_per_peak, seed):
    rng = np.random.default_rng(seed)
    n, nf = 3000, 1600
    t = np.sort(rng.uniform(0, 100.0, n))
    df = 1 / (t[-1] - t[0]) / n_per_peak
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + rng.standard_normal(n)
    (C, S, YC, YS, CC, CS), yv = emulate_strip_sums(t, y, fmin, df, nf)
    # the reference's epilogue on the emulated sums: (Sh, Ch), (S2, C2) = (2 CS, 2 CC - 1), (S, C), weights 1/n
    feed = iter([(YS / n, YC / n), (2 * CS / n, 2 * CC / n - 1), (S / n, C / n)])
    got = gls_numpy.gls_power(t, yv, None, fmin, df, nf, True, False, trig_sum=lambda *a: next(feed))
    ref = gls_numpy.gls_power(t, y, None, fmin, df, nf, True, False, trig_sum=gls_numpy.trig_sum_exact)
    keep = np.abs(fmin + np.arange(nf) * df) * (t[-1] - t[0]) >= 1.0       # float64 bins of gls_lowfreq_kernel left out
    peak = ref[keep].max()
    assert np.argmax(np.where(keep, got, -np.inf)) == np.argmax(np.where(keep, ref, -np.inf))
    assert np.max(np.abs(got - ref)[keep]) <= 1e-5 * peak
    big = keep & (ref >= 1e-2 * peak)
    assert np.max(np.abs(got - ref)[big] / ref[big]) <= 1e-5
