"""numpy statement of the Analysis-of-Variance periodogram (Schwarzenberg-Czerny 1989) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY THE REFERENCE: ``/root/reference/src/periodicity/phase.py:11`` lists the method as a
TODO and ships no code, tests or golden vectors for it.  The oracle follows the conventions of the
reference's ``PDM._pdm`` (``phase.py:128-141``) for everything they share -- phase ``(t / P) % 1``
(``:131``), bins selected by the comparisons ``phi >= k/nb`` and ``phi < (k+1)/nb`` (``:138-140`` with
``nc = 1``) -- and is pinned instead to an independent implementation of the statistic itself:
``scipy.stats.f_oneway`` over the values grouped by phase bin (``tests/test_aov_oracle.py``).
"""
import numpy as np


def aov_theta(t, x, period, nb):
    """[(N - r)/(r - 1)] * s1 / s2 over the r populated bins (float64, two-pass means)."""
    phi = (t / period) % 1
    groups = []
    for k in range(nb):
        sel = phi >= k / nb
        sel &= phi < (k + 1) / nb
        if sel.any():
            groups.append(x[sel])
    r = len(groups)
    n = sum(g.size for g in groups)
    if r < 2 or n <= r:
        return np.nan
    allx = np.concatenate(groups)
    mean = allx.mean()
    s1 = sum(g.size * (g.mean() - mean) ** 2 for g in groups)
    s2 = sum(((g - g.mean()) ** 2).sum() for g in groups)
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((n - r) / (r - 1)) * (s1 / s2)


def aov(t, x, periods, nb):
    t = np.asarray(t, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return np.array([aov_theta(t, x, p, nb) for p in np.asarray(periods, dtype=np.float64)])
