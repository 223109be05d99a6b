"""Drop-in ``PDM`` backed by the sm_100a phase-bin histogram kernel.

Same constructor, call signature, attributes and side effects as the reference
class (``src/periodicity/phase.py:75-195``).  The reference fans
``self._pdm(period)`` out over a ``multiprocessing.Pool`` (``phase.py:185-186``);
here all trial periods go to a B200 in one ``pdc_pdm`` call.  ``cores`` is
accepted for signature compatibility and ignored.

There is no CPU fallback: without the CUDA library / a B200 the call raises.
"""
import numpy as np

from . import _ffi
from .core import FSeries, TSeries

__all__ = ["PDM"]


class PDM(object):
    """Phase Dispersion Minimisation (Stellingwerf 1978).

    Parameters (identical to the reference, ``phase.py:78-126``)
    ----------
    nb : int      number of phase bins (default 5)
    nc : int      number of covers per bin (default 2)
    p_min, p_max : float, optional   trial-period range; defaults ``2*median_dt``
                  and ``oversample * baseline``
    n_periods : int   number of trial periods (default 1000; ``None`` derives it
                  from the frequency resolution, ``phase.py:176-179``)
    oversample : scalar
    do_subharmonic : bool   average theta(P) with theta(2P) where significant
    cores : ignored (kept for signature compatibility)

    Extra, keyword-only: ``device`` (CUDA ordinal), ``shard`` (split the period
    grid across ``torch.distributed`` ranks, ``dist.pdm_sharded``; ``shard="p2p"`` fuses the
    all-gather into the epilogue kernel over NVLink peer memory, ``dist.pdm_sharded_p2p``).
    """

    def __init__(self, nb=5, nc=2, p_min=None, p_max=None, n_periods=1000, oversample=1,
                 do_subharmonic=False, cores=None, *, device=None, shard=False):
        self.nb = nb
        self.nc = nc
        self.p_min = p_min
        self.p_max = p_max
        self.n_periods = n_periods
        self.oversample = oversample
        self.do_subharmonic = do_subharmonic
        self.cores = cores
        self.device = device
        self.shard = shard

    def _theta(self, periods):
        """theta for each period, in the order given (replaces the Pool map of ``phase.py:185-187``)."""
        if self.shard:
            from . import dist
            sharded = dist.pdm_sharded_p2p if self.shard == "p2p" else dist.pdm_sharded
            theta, self.argmin_index, self.min_theta = sharded(
                self.t, self.x, periods, self.nb, self.nc, device=self.device)
        else:
            ctx = _ffi.default_context(self.device)
            theta, self.argmin_index, self.min_theta = ctx.pdm(self.t, self.x, periods, self.nb, self.nc)
        return theta

    def _pdm(self, period):
        """theta for a single trial period (``phase.py:128-149``)."""
        return float(self._theta(np.array([period], dtype=np.float64))[0])

    def __call__(self, signal):
        """theta(P) on ``linspace(p_min, p_max, n_periods)`` as an ``FSeries`` over ``1/P``
        (``phase.py:151-195``); sets ``signal, t, x, sigma, periods, periodogram``."""
        if not isinstance(signal, TSeries):
            signal = TSeries(values=signal)
        self.signal = signal
        self.t = signal.time
        self.x = signal.values
        self.sigma = np.var(signal.values, ddof=1)
        theta_crit = 1.0 - 11.0 / signal.size ** 0.8
        t0 = signal.baseline
        p_min = 2 * signal.median_dt if self.p_min is None else self.p_min
        p_max = self.oversample * t0 if self.p_max is None else self.p_max
        if self.n_periods is None:
            n_periods = int((1 / p_min - 1 / p_max) * self.oversample * t0 + 1)
        else:
            n_periods = self.n_periods
        self.periods = np.linspace(p_min, p_max, n_periods)
        dp = self.periods[1] - self.periods[0]
        thetas = self._theta(self.periods)
        if self.do_subharmonic:
            # phase.py:188-193, in period order, before the FSeries wrap re-sorts by frequency
            (can_average,) = np.where((thetas < theta_crit) & (self.periods <= p_max / 2))
            sub_indices = np.round(2 * can_average + p_min / dp).astype(int)
            thetas[can_average] = (thetas[can_average] + thetas[sub_indices]) / 2
        self.periodogram = FSeries(1 / self.periods, thetas)
        return self.periodogram
