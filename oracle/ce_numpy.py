"""numpy statement of the conditional-entropy periodogram (Graham et al. 2013) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY THE REFERENCE: ``/root/reference/src/periodicity/phase.py:15`` lists the method as a TODO and ships
no code for it.  Prepared for the CUDA path planned in DESIGN.md section 9 (not built yet); the conventions follow the
reference's ``PDM._pdm`` where they overlap: phase ``(t / P) % 1`` (``phase.py:131``), phase bin k selected by
``phi >= k/nphi`` and ``phi < (k+1)/nphi`` (``phase.py:138-140`` with ``nc = 1``).  Magnitudes are scaled to [0, 1] with
the sample minimum and maximum and cut into ``nm`` equal bins (the maximum goes to the last bin).

    H_c(P) = sum_ij p(m_i, phi_j) ln( p(phi_j) / p(m_i, phi_j) ),   p = cell occupation / N

over the occupied cells; the best period MINIMISES it.  Pinned to an independent evaluation through
``np.histogram2d`` in ``tests/test_ce_oracle.py``.
"""
import numpy as np


def magnitude_bins(x, nm):
    """Per-sample magnitude bin (period independent): floor(nm (x - min) / (max - min)), maximum in the last bin."""
    x = np.asarray(x, dtype=np.float64)
    lo, hi = np.nanmin(x), np.nanmax(x)
    scaled = (x - lo) / (hi - lo)
    return np.minimum((scaled * nm).astype(np.int64), nm - 1)


def ce_theta(t, mbin, period, nphi, nm):
    phi = (t / period) % 1
    thr = np.arange(nphi + 1) / nphi
    pbin = np.minimum(np.searchsorted(thr, phi, side="right") - 1, nphi - 1)
    counts = np.bincount(pbin * nm + mbin, minlength=nphi * nm).reshape(nphi, nm).astype(np.float64)
    p = counts / counts.sum()
    pphi = p.sum(axis=1, keepdims=True) * np.ones_like(p)
    ok = p > 0
    return float(np.sum(p[ok] * np.log(pphi[ok] / p[ok])))


def ce(t, x, periods, nphi=10, nm=5):
    t = np.asarray(t, dtype=np.float64)
    mbin = magnitude_bins(x, nm)
    return np.array([ce_theta(t, mbin, p, nphi, nm) for p in np.asarray(periods, dtype=np.float64)])
