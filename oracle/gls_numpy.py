"""numpy restatement of the reference GLS path -- TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/src/periodicity/spectral.py``:

* ``trig_sum_fast``   restates ``_trig_sum`` (``spectral.py:11-40``): the
  Press & Rybicki extirpolation of the weighted samples onto an ``nfft``-point
  grid followed by one inverse FFT.  It is an *approximation* of the sums
  below (error ~ oversampling**-4).  This is oracle O2 ("as shipped") and the
  reference CPU algorithm timed by ``bench.py --impl reference``.
* ``trig_sum_exact``  the sums the reference approximates,
  ``S_j = sum_i w_i sin(2 pi f_j t_i)``, ``C_j = sum_i w_i cos(2 pi f_j t_i)``,
  ``f_j = fmin + j df`` (docstring ``spectral.py:12-16``), evaluated directly in
  float64.  Plugged into the same epilogue it gives oracle O1 (the "formula
  oracle"), the <=1e-5 tolerance target of the CUDA kernel.
* ``gls_grid`` / ``gls_weights`` / ``gls_power``  restate ``GLS.__call__``
  (``spectral.py:86-132``): grid, weights, mean removal, tau-offset algebra and
  normalisation.

Pinned against outputs of the unmodified reference in
``tests/test_oracle_golden.py`` (fixtures: ``tests/golden/gls_*.npz``).
"""
import numpy as np

TWO_PI = 2.0 * np.pi


def trig_sum_fast(t, w, df, nf, fmin, oversampling=5):
    """(S, C) by 4-point Lagrange extirpolation + one inverse FFT (``spectral.py:11-40``).

    ``oversampling`` is the reference's ``n=5`` default of ``_trig_sum``; note the
    reference never forwards ``GLS.n`` to it (``spectral.py:109-112``).
    """
    t = np.asarray(t, dtype=np.float64)
    w = np.asarray(w)
    nfft = 1 << int(nf * oversampling - 1).bit_length()          # spectral.py:18
    t0 = t.min()
    cw = w * np.exp(1j * TWO_PI * fmin * (t - t0))              # shift grid start to fmin (:20)
    pos = ((t - t0) * nfft * df) % nfft                          # fractional FFT-grid position (:21)
    re = np.zeros(nfft)
    im = np.zeros(nfft)
    on_grid = (pos % 1) == 0                                     # exact hits go in unsplit (:23-24)
    if on_grid.any():
        idx = pos[on_grid].astype(np.int64)
        re += np.bincount(idx, cw.real[on_grid], nfft)
        im += np.bincount(idx, cw.imag[on_grid], nfft)
    pos, cw = pos[~on_grid], cw[~on_grid]
    lo = np.clip((pos - 2).astype(np.int64), 0, nfft - 4)        # first of 4 neighbours (:26)
    d = pos[None, :] - (lo[None, :] + np.arange(4)[:, None])     # distances to the 4 nodes
    full = cw * (d[0] * d[1] * d[2] * d[3])                      # common numerator (:27)
    # Lagrange denominators prod_{l != k}(k - l) for k = 0..3 are -6, 2, -2, 6 (:28-33)
    for k, den in enumerate((-6.0, 2.0, -2.0, 6.0)):
        contrib = full / (den * d[k])
        re += np.bincount(lo + k, contrib.real, nfft)
        im += np.bincount(lo + k, contrib.imag, nfft)
    spec = np.fft.ifft(re + 1j * im)[:nf]                        # (:34)
    if t0 != 0:
        spec = spec * np.exp(1j * TWO_PI * t0 * (fmin + df * np.arange(nf)))   # (:35-37)
    return nfft * spec.imag, nfft * spec.real                    # (S, C)  (:38-40)


def trig_sum_exact(t, w, df, nf, fmin, chunk=None):
    """(S, C) as direct float64 sums over samples, chunked over frequency."""
    t = np.asarray(t, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    S = np.empty(nf)
    C = np.empty(nf)
    if chunk is None:
        chunk = max(1, int(4_000_000 // max(1, t.size)))
    for a in range(0, nf, chunk):
        b = min(nf, a + chunk)
        f = fmin + df * np.arange(a, b)
        ph = TWO_PI * np.outer(f, t)
        S[a:b] = np.sin(ph) @ w
        C[a:b] = np.cos(ph) @ w
    return S, C


def gls_grid(time, n=5, fmin=None, fmax=None):
    """(fmin, df, frequency) exactly as ``spectral.py:88-98``."""
    time = np.asarray(time)
    df = 1.0 / (time[-1] - time[0]) / n
    if fmin is None:
        fmin = 0.5 * df
    if fmax is None:
        fmax = 0.5 / np.median(np.diff(time))
    frequency = np.arange(fmin, fmax + df, df)
    return fmin, df, frequency


def gls_power(t, values, err=None, fmin=None, df=None, nf=None, fit_mean=True, psd=False,
              trig_sum=trig_sum_fast):
    """Periodogram values on ``fmin + j*df`` (``spectral.py:99-132``)."""
    t = np.asarray(t, dtype=np.float64)
    values = np.asarray(values)
    if err is None:
        err = np.ones_like(values)
    w = err ** -2.0
    w = w / w.sum()
    y = values - np.dot(w, values) if fit_mean else values
    Sh, Ch = trig_sum(t, w * y, df, nf, fmin)
    S2, C2 = trig_sum(t, w, 2 * df, nf, 2 * fmin)
    with np.errstate(divide="ignore", invalid="ignore"):
        if fit_mean:
            S, C = trig_sum(t, w, df, nf, fmin)
            tan2 = (S2 - 2 * S * C) / (C2 - (C * C - S * S))
        else:
            tan2 = S2 / C2
        hyp = np.sqrt(1 + tan2 * tan2)
        S2w = tan2 / hyp
        C2w = 1 / hyp
        Cw = np.sqrt(0.5) * np.sqrt(1 + C2w)
        Sw = np.sqrt(0.5) * np.sign(S2w) * np.sqrt(1 - C2w)
        YY = np.dot(w, y ** 2)
        YC = Ch * Cw + Sh * Sw
        YS = Sh * Cw - Ch * Sw
        CC = 0.5 * (1 + C2 * C2w + S2 * S2w)
        SS = 0.5 * (1 - C2 * C2w - S2 * S2w)
        if fit_mean:
            CC = CC - (C * Cw + S * Sw) ** 2
            SS = SS - (S * Cw - C * Sw) ** 2
        power = YC * YC / CC + YS * YS / SS
        if psd:
            power = power * (0.5 * (err ** -2.0).sum())
        else:
            power = power / YY
    return power


def gls(time, values, err=None, fmin=None, fmax=None, n=5, psd=False, fit_mean=True, exact=False):
    """Full call: returns (frequency, power)."""
    fmin_, df, frequency = gls_grid(time, n=n, fmin=fmin, fmax=fmax)
    power = gls_power(time, values, err, fmin_, df, frequency.size, fit_mean, psd,
                      trig_sum=trig_sum_exact if exact else trig_sum_fast)
    return frequency, power
