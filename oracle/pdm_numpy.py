"""numpy restatement of the reference PDM path -- TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/src/periodicity/phase.py``:

* ``pdm_theta_masks``  restates ``PDM._pdm`` (``phase.py:128-149``) step by step:
  phase ``(t/P) % 1``, ``m0 = nb*nc`` overlapping bins selected with the same
  three comparisons against ``k/m0`` (``:137-141``), bins with <= 1 sample
  dropped (``:142-144``), pooled variance (``:145-149``).  ``sort=True`` also
  performs the reference's argsort by phase (``:132-134``; it cannot change the
  result, only the cost) and is what the CPU timing baseline uses.
* ``pdm_theta_hist``   the fine-bin form the CUDA kernel implements: every
  sample falls in one of ``m0`` fine bins delimited by the SAME float64
  thresholds ``k/m0``; coarse bin k is the circular union of fine bins
  k..k+nc-1; per bin only (n, sum x', sum x'^2) with x' = x - mean(x) is kept.
* ``pdm_grid`` / ``pdm`` restate ``PDM.__call__`` (``phase.py:160-194``) including the
  optional sub-harmonic averaging (``:188-193``).

Pinned against outputs of the unmodified reference in
``tests/test_oracle_golden.py`` (fixtures: ``tests/golden/pdm_*.npz``).  The
reference has no PDM test (``tests/test_phase.py`` is empty).
"""
import numpy as np


def pdm_theta_masks(t, x, period, nb, nc, sigma2=None, sort=False):
    m0 = nb * nc
    phi = (t / period) % 1
    if sort:
        order = np.argsort(phi)
        phi = phi[order]
        x = x[order]
    if sigma2 is None:
        sigma2 = np.var(x, ddof=1)
    num = 0.0
    ntot = 0
    good = 0
    for k in range(m0):
        sel = phi >= k / m0
        sel &= phi < (k + nc) / m0
        sel |= phi < (k - (m0 - nc)) / m0
        xs = x[sel]
        if xs.size > 1:
            num += (xs.size - 1) * np.var(xs, ddof=1)
            ntot += xs.size
            good += 1
    return (num / (ntot - good)) / sigma2


def fine_bin(phi, m0):
    """Index k with k/m0 <= phi < (k+1)/m0 using the reference's own thresholds."""
    thr = np.arange(m0 + 1) / m0
    q = np.searchsorted(thr, phi, side="right") - 1
    return np.minimum(q, m0 - 1)


def pdm_theta_hist(t, x, period, nb, nc, sigma2=None):
    m0 = nb * nc
    phi = (t / period) % 1
    q = fine_bin(phi, m0)
    xc = x - x.mean()
    n = np.bincount(q, minlength=m0).astype(np.float64)
    s1 = np.bincount(q, xc, minlength=m0)
    s2 = np.bincount(q, xc * xc, minlength=m0)
    idx = (np.arange(m0)[:, None] + np.arange(nc)[None, :]) % m0
    N = n[idx].sum(1)
    S1 = s1[idx].sum(1)
    S2 = s2[idx].sum(1)
    ok = N > 1
    if sigma2 is None:
        sigma2 = np.var(x, ddof=1)
    num = np.sum(S2[ok] - S1[ok] ** 2 / N[ok])
    return (num / (N[ok].sum() - ok.sum())) / sigma2


def pdm_grid(time, p_min=None, p_max=None, n_periods=1000, oversample=1):
    """Trial periods exactly as ``phase.py:166-180``."""
    time = np.asarray(time)
    t0 = time[-1] - time[0]
    if p_min is None:
        p_min = 2 * np.median(np.diff(time))
    if p_max is None:
        p_max = oversample * t0
    if n_periods is None:
        n_periods = int((1 / p_min - 1 / p_max) * oversample * t0 + 1)
    return p_min, p_max, np.linspace(p_min, p_max, n_periods)


def subharmonic_average(thetas, periods, p_min, p_max, size):
    """``phase.py:166,188-193`` (applied in period order, before the FSeries wrap)."""
    thetas = np.array(thetas, dtype=np.float64)
    theta_crit = 1.0 - 11.0 / size ** 0.8
    dp = periods[1] - periods[0]
    (can,) = np.where((thetas < theta_crit) & (periods <= p_max / 2))
    sub = np.round(2 * can + p_min / dp).astype(int)
    thetas[can] = (thetas[can] + thetas[sub]) / 2
    return thetas


def pdm(time, values, nb=5, nc=2, p_min=None, p_max=None, n_periods=1000, oversample=1,
        do_subharmonic=False, kernel=pdm_theta_masks, **kw):
    """Full call: returns (periods ascending, thetas in period order)."""
    time = np.asarray(time, dtype=np.float64)
    values = np.asarray(values)
    p_min_, p_max_, periods = pdm_grid(time, p_min, p_max, n_periods, oversample)
    sigma2 = np.var(values, ddof=1)
    thetas = np.array([kernel(time, values, p, nb, nc, sigma2, **kw) for p in periods])
    if do_subharmonic:
        thetas = subharmonic_average(thetas, periods, p_min_, p_max_, values.size)
    return periods, thetas


def _pool_init(time, values, nb, nc, sigma2, sort):
    global _G
    _G = (time, values, nb, nc, sigma2, sort)


def _pool_eval(period):
    time, values, nb, nc, sigma2, sort = _G
    return pdm_theta_masks(time, values, period, nb, nc, sigma2, sort=sort)


def pdm_pool(time, values, periods, nb, nc, cores, sort=True):
    """CPU baseline: fan the trial periods out over ``cores`` processes (``phase.py:185-186``)."""
    from multiprocessing import Pool
    sigma2 = np.var(values, ddof=1)
    with Pool(cores, initializer=_pool_init, initargs=(time, values, nb, nc, sigma2, sort)) as pool:
        return np.array(pool.map(_pool_eval, periods))
