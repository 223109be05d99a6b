"""CPU restatement of the reference String Length method -- TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows ``src/periodicity/phase.py``:
  * ``StringLength.__call__`` ``:53-72``: scale the signal to [-0.25, 0.25] (``:64-65``),
    ``df = dphi / baseline`` (``:66``), ``periods = 1 / linspace(n_periods*df, df, n_periods)`` (``:67``),
    fan-out (``:68-70``), ``FSeries(1 / periods, ell)`` (``:71``; re-sorted to ascending frequency, ``core.py:877-881``);
  * ``StringLength._stringlength`` ``:45-51``: ``fold(period)`` = ``TSeries((t / P) % 1, values)`` (``core.py:543-544``),
    whose constructor re-sorts by phase with a stable sort (``core.py:473-477``), then
    ``hypot(roll(m, -1) - m, roll(phi, -1) - phi).sum()``.

Parity pin: WEAK.  The reference class cannot complete a call with the reference's own ``core.py``
(``signal - signal.max()`` subtracts a one-sample series, ``core.py:217-220,175-178``) and has no test.  The
golden vectors ``tests/golden/sl_*.npz`` come from the UNMODIFIED ``phase.py`` code executed behind the
numpy stand-in core of ``oracle/refload.py`` whose ``max()`` / ``min()`` return scalars -- i.e. they pin this
restatement to the reference's statements, under the one assumption about ``max()`` stated above.
"""
import numpy as np


def scale(values):
    """``phase.py:64-65`` with scalar extrema."""
    values = np.asarray(values, dtype=np.float64)
    vmax, vmin = np.nanmax(values), np.nanmin(values)
    return (values - vmax) / (2 * (vmax - vmin)) + 0.25


def period_grid(time, dphi=0.1, n_periods=1000):
    """``phase.py:66-67``."""
    df = dphi / (time[-1] - time[0])
    return 1 / np.linspace(n_periods * df, df, n_periods)


def string_length(t, m, period):
    """``phase.py:45-51`` for one trial period."""
    phi = (t / period) % 1
    order = np.argsort(phi, kind="stable")
    phi, mm = phi[order], m[order]
    return np.hypot(np.roll(mm, -1) - mm, np.roll(phi, -1) - phi).sum()


def string_lengths(t, m, periods):
    return np.array([string_length(t, m, p) for p in periods])


def stringlength(time, values, dphi=0.1, n_periods=1000):
    """Full call: ``(periods, ell)`` in the order of ``periods`` (the FSeries wrap reverses both)."""
    time = np.asarray(time, dtype=np.float64)
    m = scale(values)
    periods = period_grid(time, dphi, n_periods)
    return periods, string_lengths(time, m, periods)
