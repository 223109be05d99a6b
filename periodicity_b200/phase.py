"""Drop-in ``PDM`` and ``StringLength`` backed by sm_100a kernels (phase-bin histograms; per-period sort).

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

__all__ = ["StringLength", "PDM", "AOV", "CE", "ConditionalEntropy", "GL", "GregoryLoredo"]


class StringLength(object):
    """String Length (Dworetsky 1983): same constructor and call signature as the reference class
    (``src/periodicity/phase.py:18-72``); ``cores`` is accepted and ignored.

    The reference fans ``self._stringlength(period)`` out over a ``multiprocessing.Pool``
    (``phase.py:68-70``); here all trial periods go to a B200 in one ``pdc_stringlength`` call (one
    thread block per period: fold, sort by phase, sum the segment lengths).

    Semantics note.  ``phase.py:65`` scales the signal with ``signal - signal.max()``.  In the reference's
    own ``core.py`` ``Signal.max()`` returns a ONE-SAMPLE series (``core.py:217-220``), and the subtraction
    of two series of different length is refused by xarray's exact join (``core.py:175-178``), so the
    reference class cannot complete a call as shipped and has no test (``tests/test_phase.py`` is empty).
    This class implements what the line states -- scale the values to [-0.25, 0.25] with the scalar
    maximum and minimum (NaN-ignoring, as ``amax``/``amin`` are, ``core.py:212-215``) -- and everything
    after it literally: ``periods = 1 / linspace(n_periods*df, df, n_periods)``, ``df = dphi / baseline``;
    per period ``phi = (t / P) % 1``, stable sort by ``phi``, closed-polygon length with ``np.roll``.
    """

    def __init__(self, dphi=0.1, n_periods=1000, cores=None, *, device=None, shard=False, devices=None):
        self.dphi = dphi
        self.n_periods = n_periods
        self.cores = cores
        # `devices=[0, 1, ...]`: one multi-device context (pdc_ctx_create_multi) -- the library shards the grid over these
        # GPUs of THIS process, no torchrun / torch.distributed needed; `device` = a single ordinal
        self.device = list(devices) if devices is not None else device
        self.shard = shard      # split the period grid across torch.distributed ranks (dist.stringlength_sharded)

    def _ell(self, periods):
        if self.shard:
            from . import dist
            ell, self.argmin_index, self.min_length = dist.stringlength_sharded(
                self.m.time, self.m.values, periods, device=self.device)
        else:
            ctx = _ffi.default_context(self.device)
            ell, self.argmin_index, self.min_length = ctx.stringlength(self.m.time, self.m.values, periods)
        return ell

    def _stringlength(self, period):
        """String length for a single trial period (``phase.py:45-51``)."""
        return float(self._ell(np.array([period], dtype=np.float64))[0])

    def __call__(self, signal):
        """String length on the reference's period grid as an ``FSeries`` over ``1/P`` (``phase.py:53-72``);
        sets ``signal, m, periodogram``."""
        if not isinstance(signal, TSeries):
            signal = TSeries(values=signal)
        self.signal = signal
        # scale signal to range from -0.25 to +0.25 (phase.py:64-65)
        vmax = np.nanmax(signal.values)
        vmin = np.nanmin(signal.values)
        self.m = TSeries(signal.time, (signal.values - vmax) / (2 * (vmax - vmin)) + 0.25, assume_sorted=True)
        df = self.dphi / signal.baseline
        periods = 1 / np.linspace(self.n_periods * df, df, self.n_periods)
        ell = self._ell(periods)
        self.periodogram = FSeries(1 / periods, ell)
        return self.periodogram


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
                 do_subharmonic=False, cores=None, *, device=None, shard=False, devices=None):
        self.nb = nb
        self.nc = nc
        self.p_min = p_min
        self.p_max = p_max
        self.n_periods = n_periods
        self.oversample = oversample
        self.do_subharmonic = do_subharmonic
        self.cores = cores
        # `devices=[0, 1, ...]`: one multi-device context (pdc_ctx_create_multi) -- the library shards the grid over these
        # GPUs of THIS process, no torchrun / torch.distributed needed; `device` = a single ordinal
        self.device = list(devices) if devices is not None else device
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


class AOV(object):
    """Analysis of Variance periodogram (Schwarzenberg-Czerny 1989).

    The reference lists this method as a TODO (``phase.py:11``) and has no implementation; the class
    follows the conventions of its ``PDM`` (``phase.py:75-195``): the same period-grid options with the
    same defaults (``p_min = 2*median_dt``, ``p_max = oversample*baseline``, ``linspace`` in period),
    the same phase definition ``(t / P) % 1`` (``phase.py:131``) and bin edges ``k / nb``
    (``phase.py:138-140`` with ``nc = 1``), an ``FSeries`` over ``1/P`` as result.  The statistic is the
    one-way ANOVA F ratio of the values grouped by phase bin,
    ``[(N - r)/(r - 1)] * sum_b n_b (mean_b - mean)**2 / sum_b sum_i (x_i - mean_b)**2`` over the ``r``
    populated bins; the best period MAXIMISES it.  It is evaluated on a B200 from the same shared-memory
    phase-bin histograms as PDM (``pdc_aov``).  ``cores`` is accepted and ignored.
    """

    def __init__(self, nb=10, p_min=None, p_max=None, n_periods=1000, oversample=1, cores=None, *, device=None, devices=None):
        self.nb = nb
        self.p_min = p_min
        self.p_max = p_max
        self.n_periods = n_periods
        self.oversample = oversample
        self.cores = cores
        # `devices=[0, 1, ...]`: one multi-device context (pdc_ctx_create_multi) -- the library shards the grid over these
        # GPUs of THIS process, no torchrun / torch.distributed needed; `device` = a single ordinal
        self.device = list(devices) if devices is not None else device

    def _theta(self, periods):
        ctx = _ffi.default_context(self.device)
        theta, self.argmax_index, self.max_theta = ctx.aov(self.t, self.x, periods, self.nb)
        return theta

    def _aov(self, period):
        """The statistic for a single trial period."""
        return float(self._theta(np.array([period], dtype=np.float64))[0])

    def __call__(self, signal):
        """Theta_AoV(P) on ``linspace(p_min, p_max, n_periods)`` as an ``FSeries`` over ``1/P``;
        sets ``signal, t, x, periods, periodogram``."""
        if not isinstance(signal, TSeries):
            signal = TSeries(values=signal)
        self.signal = signal
        self.t = signal.time
        self.x = signal.values
        t0 = signal.baseline
        p_min = 2 * signal.median_dt if self.p_min is None else self.p_min
        p_max = self.oversample * t0 if self.p_max is None else self.p_max
        if self.n_periods is None:
            n_periods = int((1 / p_min - 1 / p_max) * self.oversample * t0 + 1)
        else:
            n_periods = self.n_periods
        self.periods = np.linspace(p_min, p_max, n_periods)
        self.periodogram = FSeries(1 / self.periods, self._theta(self.periods))
        return self.periodogram


class CE(object):
    """Conditional-entropy periodogram (Graham et al. 2013).

    The reference lists this method as a TODO (``phase.py:13``) and has no implementation; the class follows the
    conventions of its ``PDM`` (``phase.py:75-195``): the same period-grid options with the same defaults
    (``p_min = 2*median_dt``, ``p_max = oversample*baseline``, ``linspace`` in period), the same phase definition
    ``(t / P) % 1`` (``phase.py:131``) and bin edges ``k / nb`` (``phase.py:138-140`` with ``nc = 1``), an ``FSeries``
    over ``1/P`` as result.  Values are scaled to [0, 1] with their minimum and maximum and cut into ``nm`` equal
    magnitude bins; the statistic is ``H(m | phi) = sum p(phi, m) ln(p(phi) / p(phi, m))`` over the occupied cells of
    the ``nb x nm`` phase-magnitude histogram and the best period MINIMISES it.  It is evaluated on a B200 from
    per-period count histograms in shared memory (``pdc_ce``).  ``cores`` is accepted and ignored.
    """

    def __init__(self, nb=10, nm=5, p_min=None, p_max=None, n_periods=1000, oversample=1, cores=None, *,
                 device=None, devices=None):
        self.nb = nb
        self.nm = nm
        self.p_min = p_min
        self.p_max = p_max
        self.n_periods = n_periods
        self.oversample = oversample
        self.cores = cores
        # `devices=[0, 1, ...]`: one multi-device context (pdc_ctx_create_multi); `device` = a single ordinal
        self.device = list(devices) if devices is not None else device

    def _entropy(self, periods):
        ctx = _ffi.default_context(self.device)
        h, self.argmin_index, self.min_entropy = ctx.ce(self.t, self.x, periods, self.nb, self.nm)
        return h

    def _ce(self, period):
        """The statistic for a single trial period."""
        return float(self._entropy(np.array([period], dtype=np.float64))[0])

    def __call__(self, signal):
        """H_c(P) on ``linspace(p_min, p_max, n_periods)`` as an ``FSeries`` over ``1/P``;
        sets ``signal, t, x, periods, periodogram``."""
        if not isinstance(signal, TSeries):
            signal = TSeries(values=signal)
        self.signal = signal
        self.t = signal.time
        self.x = signal.values
        t0 = signal.baseline
        p_min = 2 * signal.median_dt if self.p_min is None else self.p_min
        p_max = self.oversample * t0 if self.p_max is None else self.p_max
        if self.n_periods is None:
            n_periods = int((1 / p_min - 1 / p_max) * self.oversample * t0 + 1)
        else:
            n_periods = self.n_periods
        self.periods = np.linspace(p_min, p_max, n_periods)
        self.periodogram = FSeries(1 / self.periods, self._entropy(self.periods))
        return self.periodogram


ConditionalEntropy = CE


class GL(object):
    """Gregory-Loredo (1992) periodogram for event arrival times.

    The reference lists this method as a TODO (``phase.py:14``) and has no implementation; the class follows the
    conventions of its ``PDM`` (``phase.py:75-195``): the same period-grid options with the same defaults, the phase
    definition ``(t / P) % 1`` (``phase.py:131``), an ``FSeries`` over ``1/P`` as result.  The input is a series of
    event times (a ``TSeries``' time axis, or a plain array of times; values are ignored).  For every trial period
    the odds of stepwise periodic models with ``m = 2 .. m_max`` phase bins against a constant rate, marginalised over
    the bin rates and the phase offset (``nc`` offsets per bin), are averaged; the best period MAXIMISES ``ln O``.
    Evaluated on a B200 from count histograms in shared memory (``pdc_gl``).  ``cores`` is accepted and ignored.
    """

    def __init__(self, m_max=12, nc=10, p_min=None, p_max=None, n_periods=1000, oversample=1, cores=None, *,
                 device=None, devices=None):
        self.m_max = m_max
        self.nc = nc
        self.p_min = p_min
        self.p_max = p_max
        self.n_periods = n_periods
        self.oversample = oversample
        self.cores = cores
        self.device = list(devices) if devices is not None else device

    def _lnodds(self, periods):
        ctx = _ffi.default_context(self.device)
        out, self.argmax_index, self.max_lnodds = ctx.gl(self.t, periods, self.m_max, self.nc)
        return out

    def __call__(self, events):
        """ln O(P) on ``linspace(p_min, p_max, n_periods)`` as an ``FSeries`` over ``1/P``; sets ``t, periods, periodogram``."""
        t = np.sort(np.asarray(events.time if isinstance(events, TSeries) else events, dtype=np.float64).ravel())
        self.t = t
        t0 = t[-1] - t[0]
        p_min = 2 * np.median(np.diff(t)) if self.p_min is None else self.p_min
        p_max = self.oversample * t0 if self.p_max is None else self.p_max
        if self.n_periods is None:
            n_periods = int((1 / p_min - 1 / p_max) * self.oversample * t0 + 1)
        else:
            n_periods = self.n_periods
        self.periods = np.linspace(p_min, p_max, n_periods)
        self.periodogram = FSeries(1 / self.periods, self._lnodds(self.periods))
        return self.periodogram


GregoryLoredo = GL
