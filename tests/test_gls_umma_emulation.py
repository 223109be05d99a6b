"""CPU emulation of the numerical scheme of gls_umma_kernel (periodicity_b200/csrc/gls_umma.cu) against the exact
float64 sums: the accuracy argument behind the tensor-core formulation, checkable without a GPU.

Emulated, step by step as the kernel does it:
  * grid index j = 128 cb + k; per sample the phase step b_i as a 64-bit fraction, the coarse base phase as a 32-bit
    fraction; fine phase = (k * b64) >> 32, coarse phase = A32 + ((128 cb) * b64 >> 32), both mod 2^32 (integer arithmetic);
  * (cos, sin) in float32 of the 2^-32-turn phase converted to float32 radians;
  * every operand value x split into fp16 hi = rn(x) and lo = rn(x - hi); products hi hi + hi lo + lo hi, each product
    exact (11-bit significands);
  * one tcgen05.mma = 8 samples (16 K-slots): D <- truncate_toward_zero_float32(D + sum of 16 exact products), three
    instructions per K-step in the kernel's order (lo hi, hi lo, hi hi) -- the truncation is what the hardware probe
    measured (profiles/r02/umma_probe.txt);
  * a run of CS stages (16 samples each) starts from zero; its sum is multiplied by 1 + 2.07e-8 * instructions (the
    expected truncation loss) and added to a float32 master with round-to-nearest;
  * the reference's tau-offset algebra (spectral.py:113-132) in float64 on the six sums, C2 / S2 taken directly at the
    doubled angle (the reference's second `_trig_sum`, spectral.py:110).
"""
import numpy as np
import pytest

from oracle import gls_numpy

f32, f16 = np.float32, np.float16
TWO_PI_32 = f32(2 * np.pi / 2.0 ** 32)


def trunc_f32(x):
    """float64 -> float32 rounding toward zero."""
    r = x.astype(f32)
    over = np.abs(r.astype(np.float64)) > np.abs(x)
    r[over] = np.nextafter(r[over], f32(0))
    return r


def split(x):
    hi = x.astype(f16)
    lo = (x - hi.astype(f32)).astype(f16)
    return hi.astype(np.float64), lo.astype(np.float64)


def sincos_fx(fx):
    """fx: uint32 phase in 2^-32 turn (numpy int64 holding 0 .. 2^32-1) -> float32 (cos, sin) as the kernel's MUFU path."""
    sfx = np.where(fx >= 2 ** 31, fx - 2 ** 32, fx).astype(np.int64)
    x = (sfx.astype(f32) * TWO_PI_32).astype(np.float64)
    return np.cos(x).astype(f32), np.sin(x).astype(f32)


def emulate(t, y, fmin, df, nf, CS, comp_on=True):
    n = t.size
    tt = t - t.min()
    yc = y - y.mean()
    yv = (yc / np.sqrt(np.mean(yc * yc))).astype(f32)
    nC = -(-nf // 128)
    b = (df * tt) % 1.0
    b64 = [int(round(v * 2.0 ** 64)) % (1 << 64) for v in b]                 # python ints: exact
    A32 = np.array([int(round(((fmin * ti) % 1.0) * 2.0 ** 32)) % (1 << 32) for ti in tt], dtype=np.int64)
    sums = np.zeros((6, nC, 128))
    for typ, kmul in ((1, 1), (2, 2)):
        fine_fx = np.array([[(kmul * k * bb >> 32) & 0xffffffff for bb in b64] for k in range(128)], dtype=np.int64)
        coarse_fx = np.array([[((kmul * a) + (kmul * cb * 128 * bb >> 32)) & 0xffffffff for a, bb in zip(A32, b64)]
                              for cb in range(nC)], dtype=np.int64)
        fc, fs = sincos_fx(fine_fx)                                         # [128, n]
        cc, cs = sincos_fx(coarse_fx)                                       # [nC, n]
        # K-slots per sample: fine (cos, sin); coarse rows (w c, -w s) and (w s, w c), plain and y-weighted for type 1
        fine = np.stack([fc, fs], axis=2).reshape(128, 2 * n)
        rows = []
        weights = [np.ones(n, dtype=f32), yv] if typ == 1 else [np.ones(n, dtype=f32)]
        for wv in weights:
            wc, ws = (wv * cc).astype(f32), (wv * cs).astype(f32)
            rows.append(np.stack([wc, -ws], axis=2).reshape(nC, 2 * n))
            rows.append(np.stack([ws, wc], axis=2).reshape(nC, 2 * n))
        coarse = np.concatenate(rows, axis=0)                               # [rows, 2n]
        fh, fl = split(fine)
        ch_, cl = split(coarse)
        pad = (-n) % (CS * 16)
        if pad:
            z = lambda a: np.concatenate([a, np.zeros((a.shape[0], 2 * pad))], axis=1)
            fh, fl, ch_, cl = z(fh), z(fl), z(ch_), z(cl)
        master = np.zeros((128, coarse.shape[0]), dtype=f32)
        nrun = (n + pad) // (CS * 16)
        for run in range(nrun):
            acc = np.zeros((128, coarse.shape[0]), dtype=f32)
            for ks in range(CS * 2):                                        # K-steps of 8 samples = 16 slots
                sl = slice((run * CS * 2 + ks) * 16, (run * CS * 2 + ks + 1) * 16)
                for A, B in ((fl, ch_), (fh, cl), (fh, ch_)):               # lo hi, hi lo, hi hi
                    acc = trunc_f32(acc.astype(np.float64) + A[:, sl] @ B[:, sl].T)
            real = min(CS * 16, n - run * CS * 16)
            comp = f32(1.0 + (2.07e-8 * 3 * ((real + 7) // 8) if comp_on else 0.0))
            master = (master.astype(np.float64) + acc.astype(np.float64) * np.float64(comp)).astype(f32)   # one FFMA
        m = master.astype(np.float64).T                                      # [rows, 128]
        if typ == 1:
            sums[0], sums[1], sums[2], sums[3] = m[0:nC], m[nC:2 * nC], m[2 * nC:3 * nC], m[3 * nC:4 * nC]
        else:
            sums[4], sums[5] = m[0:nC], m[nC:2 * nC]
    return sums.reshape(6, -1)[:, :nf], yv.astype(np.float64)


def powers(t, y, fmin, df, nf, CS, comp_on=True):
    n = t.size
    (C, S, YC, YS, C2, S2), yv = emulate(t, y, fmin, df, nf, CS, comp_on)
    feed = iter([(YS / n, YC / n), (S2 / n, C2 / n), (S / n, C / n)])
    got = gls_numpy.gls_power(t, yv, None, fmin, df, nf, True, False, trig_sum=lambda *a: next(feed))
    ref = gls_numpy.gls_power(t, y, None, fmin, df, nf, True, False, trig_sum=gls_numpy.trig_sum_exact)
    keep = np.abs(fmin + np.arange(nf) * df) * (t[-1] - t[0]) >= 1.0          # sub-cycle bins are FP64 on the GPU
    return got, ref, keep


def make(n, nf, sigma, seed):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, 100.0, n))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + sigma * rng.standard_normal(n)
    return t, y, fmin, df


@pytest.mark.parametrize("n,nf,sigma,CS,seed", [(1200, 250, 1.0, 4, 1), (1530, 256, 0.3, 8, 2), (1024, 200, 1.0, 16, 3)])
def test_fp16_split_gemm_with_truncating_accumulator_meets_the_parity_tolerance(n, nf, sigma, CS, seed):
    t, y, fmin, df = make(n, nf, sigma, seed)
    got, ref, keep = powers(t, y, fmin, df, nf, CS)
    peak = ref[keep].max()
    assert np.argmax(np.where(keep, got, -np.inf)) == np.argmax(np.where(keep, ref, -np.inf))
    assert np.max(np.abs(got - ref)[keep]) <= 1e-5 * peak
    big = keep & (ref >= 1e-2 * peak)
    assert np.max(np.abs(got - ref)[big] / ref[big]) <= 1e-5


def test_truncation_bias_grows_with_the_run_length_and_the_compensation_removes_it():
    t, y, fmin, df = make(2048, 128, 0.05, 7)                                # strong, coherent peak
    errs = {}
    for CS, comp in ((4, False), (16, False), (16, True)):
        got, ref, keep = powers(t, y, fmin, df, 128, CS, comp)
        j = np.argmax(np.where(keep, ref, -np.inf))
        errs[(CS, comp)] = (got[j] - ref[j]) / ref[j]
    assert errs[(16, False)] < errs[(4, False)] < 0                          # truncation loses power, more for longer runs
    assert abs(errs[(16, True)]) < 0.35 * abs(errs[(16, False)])             # most of it is compensated
    assert abs(errs[(16, True)]) <= 1e-6
