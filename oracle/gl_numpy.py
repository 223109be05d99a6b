"""numpy statement of the Gregory-Loredo (1992) periodogram for event arrival times -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY THE REFERENCE: ``/root/reference/src/periodicity/phase.py:14`` lists the method as a TODO and ships
no code for it.  Conventions follow the reference's ``PDM._pdm`` where they overlap: phase ``(t / P) % 1``
(``phase.py:131``), fine phase bin k selected by ``phi >= k/F`` and ``phi < (k+1)/F`` (``phase.py:138-140``).

For N events and a stepwise model with m phase bins, the odds against a constant rate, marginalised over the bin rates
and over the phase offset (Gregory & Loredo 1992, section 5):

    O_m(P) = [(m - 1)! / (N + m - 1)!]  m^N  < prod_j n_j(P, phi)! >_phi

The offset average runs over ``nc`` offsets per bin: events are counted in ``F = m * nc`` fine bins and the m bins at
offset c are circular unions of nc consecutive fine bins starting at c.  ``ln O(P) = ln mean_{m=2..m_max} O_m(P)``.
Pinned in ``tests/test_gl_oracle.py`` to a direct evaluation with ``np.histogram`` + ``math.factorial`` on small cases.
"""
import numpy as np
from scipy.special import gammaln, logsumexp


def fine_counts(t, period, F):
    phi = (t / period) % 1
    thr = np.arange(F + 1) / F
    k = np.minimum(np.searchsorted(thr, phi, side="right") - 1, F - 1)
    return np.bincount(k, minlength=F)


def ln_odds_m(t, period, m, nc):
    N = t.size
    F = m * nc
    fine = fine_counts(t, period, F)
    L = np.empty(nc)
    for c in range(nc):
        nj = np.roll(fine, -c).reshape(m, nc).sum(axis=1)
        L[c] = gammaln(nj + 1.0).sum()
    return gammaln(m) - gammaln(N + m) + N * np.log(m) + logsumexp(L) - np.log(nc)


def gl(t, periods, m_max=12, nc=10):
    t = np.asarray(t, dtype=np.float64)
    out = np.empty(len(periods))
    for i, p in enumerate(np.asarray(periods, dtype=np.float64)):
        out[i] = logsumexp([ln_odds_m(t, p, m, nc) for m in range(2, m_max + 1)]) - np.log(m_max - 1)
    return out
