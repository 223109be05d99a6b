"""Drop-in ``GLS`` backed by the sm_100a strip kernel.

Same constructor, call signature, attributes and side effects as the reference
class (``src/periodicity/spectral.py:43-204``); the difference is *how* the
periodogram is evaluated.  The reference approximates the trigonometric sums
with an FFT extirpolation (``_trig_sum``, ``spectral.py:11-40``); here the sums
are evaluated exactly, by brute force, on a B200 through
``libperiodicity_b200.so`` (``pdc_gls``), and the tau-offset algebra of
``spectral.py:113-132`` runs on the device in float64.

There is no CPU fallback: without the CUDA library / a B200 the call raises.
"""
import copy

import numpy as np

from . import _ffi
from .core import FSeries, TSeries

__all__ = ["GLS", "BGLST"]


class GLS(object):
    """Generalised Lomb-Scargle periodogram (Zechmeister & Kuerster 2009).

    Parameters (identical to the reference, ``spectral.py:53-72``)
    ----------
    fmin, fmax : float, optional
        Frequency range; defaults: half a cycle over the baseline, and the
        pseudo-Nyquist frequency ``0.5 / median_dt``.
    n : float, optional
        Samples per peak: ``df = 1 / (baseline * n)`` (default 5).
    psd : bool, optional
        Leave the periodogram unnormalised.

    Extra, keyword-only (not in the reference; defaults keep its behaviour)
    ----------
    device : int, optional
        CUDA device ordinal (default: ``LOCAL_RANK`` or 0).
    shard : bool, optional
        If True and ``torch.distributed`` is initialised, shard the frequency
        grid across ranks and all-gather power + argmax (``dist.gls_sharded``, one NCCL
        all-gather).  ``shard="p2p"`` uses ``dist.gls_sharded_p2p`` instead: the epilogue kernel
        stores its results directly into every rank's buffer over NVLink (no NCCL call).
    """

    def __init__(self, fmin=None, fmax=None, n=5, psd=False, *, device=None, shard=False, devices=None,
                 frequency=None):
        # `frequency`: evaluate on this user-supplied list of frequencies (any spacing) instead of the uniform grid
        # derived from fmin / fmax / n (pdc_gls_freqs); not in the reference, whose FFT needs a uniform grid
        self.user_frequency = None if frequency is None else np.sort(np.asarray(frequency, dtype=np.float64).ravel())
        self.fmin = fmin
        self.fmax = fmax
        self.n = n
        self.psd = psd
        # `devices=[0, 1, ...]`: one multi-device context (pdc_ctx_create_multi) -- the library shards the grid over these
        # GPUs of THIS process, no torchrun / torch.distributed needed; `device` = a single ordinal
        self.device = list(devices) if devices is not None else device
        self.shard = shard

    # -- helpers ---------------------------------------------------------------
    def _grid(self, signal):
        """(fmin, df, frequency) following ``spectral.py:88-98`` to the letter."""
        df = 1.0 / signal.baseline / self.n
        fmin = 0.5 * df if self.fmin is None else self.fmin
        fmax = 0.5 / signal.median_dt if self.fmax is None else self.fmax
        return fmin, df, np.arange(fmin, fmax + df, df)

    def __call__(self, signal, err=None, fit_mean=True):
        """Evaluate the periodogram of ``signal`` (``spectral.py:74-135``).

        ``err`` are per-sample uncertainties (weights ``err**-2``);
        ``fit_mean`` lets the mean float with the fit.
        Returns an ``FSeries`` and sets ``frequency, err, signal, periodogram``.
        """
        if not isinstance(signal, TSeries):
            signal = TSeries(values=signal)
        if self.user_frequency is not None:
            return self._call_on_user_grid(signal, err, fit_mean)
        fmin, df, self.frequency = self._grid(signal)
        nf = self.frequency.size
        if err is None:
            err = np.ones_like(signal.values)
            weights = None                      # uniform weights: unweighted kernel
        else:
            err = np.asarray(err)
            weights = np.asarray(err, dtype=np.float64) ** -2.0
        self.err = err
        psd_scale = 0.5 * (np.asarray(err, dtype=np.float64) ** -2.0).sum() if self.psd else None
        if self.shard:
            from . import dist
            sharded = dist.gls_sharded_p2p if self.shard == "p2p" else dist.gls_sharded
            power, self.argmax_index, self.max_power = sharded(
                signal.time, signal.values, weights, fmin, df, nf, fit_mean, psd_scale, device=self.device)
        else:
            ctx = _ffi.default_context(self.device)
            power, self.argmax_index, self.max_power = ctx.gls(
                signal.time, signal.values, weights, fmin, df, nf, fit_mean=fit_mean, psd_scale=psd_scale)
        self.signal = signal
        self.periodogram = FSeries(self.frequency, power)
        return self.periodogram

    def _call_on_user_grid(self, signal, err, fit_mean):
        """``__call__`` on ``frequency=`` (non-uniform grid): same weights, mean removal, formula and attributes."""
        self.frequency = self.user_frequency
        if err is None:
            err = np.ones_like(signal.values)
            weights = None
        else:
            err = np.asarray(err)
            weights = np.asarray(err, dtype=np.float64) ** -2.0
        self.err = err
        psd_scale = 0.5 * (np.asarray(err, dtype=np.float64) ** -2.0).sum() if self.psd else None
        ctx = _ffi.default_context(self.device)
        power, self.argmax_index, self.max_power = ctx.gls_freqs(
            signal.time, signal.values, weights, self.frequency, fit_mean=fit_mean, psd_scale=psd_scale)
        self.signal = signal
        self.periodogram = FSeries(self.frequency, power, assume_sorted=True)
        return self.periodogram

    def copy(self):
        return copy.deepcopy(self)

    def top_peaks(self, k=5):
        """The ``k`` highest peaks of the last periodogram, found on the GPU (``pdc_peaks_topk``).

        Same peak definition as ``self.periodogram.find_peaks()`` (``core.py:283-317``: local maxima,
        edges excluded), so ``top_peaks(1)`` is ``periodogram.period_at_highest_peak`` without a host
        pass over the whole grid.  Returns ``(frequency[k], power[k])``, NaN-padded if fewer peaks exist."""
        ctx = _ffi.default_context(self.device)
        idx, val = ctx.peaks_topk(self.periodogram.values, k)
        idx, val = idx[0], val[0]
        freq = np.where(idx >= 0, self.periodogram.frequency[np.maximum(idx, 0)], np.nan)
        return freq, val

    def top_peak_widths(self, k=5):
        """Half-maximum period intervals of the ``k`` highest peaks, found on the GPU (``pdc_peaks_topk`` +
        ``pdc_peaks_halfmax``): row ``j`` is what ``periodogram.periods_at_half_max(peak_order=j + 1)`` returns
        (``core.py:957-972``).  Returns ``(frequency[k], power[k], lower_period[k], upper_period[k])``; NaN where a
        peak or a crossing does not exist."""
        ctx = _ffi.default_context(self.device)
        pg = self.periodogram
        idx, val = ctx.peaks_topk(pg.values, k)
        left, right = ctx.peaks_halfmax(pg.values, idx)
        idx, val, left, right = idx[0], val[0], left[0], right[0]
        period = pg.period
        freq = np.where(idx >= 0, pg.frequency[np.maximum(idx, 0)], np.nan)
        lower = np.where(right >= 0, period[np.maximum(right, 0)], np.nan)   # crossing on the high-frequency side
        upper = np.where(left >= 0, period[np.maximum(left, 0)], np.nan)
        return freq, val, lower, upper

    def bootstrap(self, n_bootstraps, random_seed=None, batch=256):
        """Maximum power of ``n_bootstraps`` resamples (``spectral.py:140-152``).

        The reference loops ``gls(bs_sample, err=bs_err).amax()``; here the
        resamples (same times, resampled values and errors, drawn from the same
        generator in the same order) go through the batched kernels keeping only
        each curve's maximum: ``pdc_gls_multi`` (shared times and weights) when the
        errors are uniform, ``pdc_gls_batch`` otherwise.
        """
        rng = np.random.default_rng(random_seed)
        ndata = len(self.signal)
        t = np.asarray(self.signal.time, dtype=np.float64)
        values = np.asarray(self.signal.values, dtype=np.float64)
        err = np.asarray(self.err, dtype=np.float64)
        fmin, df, frequency = self._grid(self.signal)
        nf = frequency.size
        ctx = _ffi.default_context(self.device)
        out = np.empty(n_bootstraps)
        uniform = bool(np.all(err == err.flat[0]))   # err=None in the original call: every replicate has equal weights
        for a in range(0, n_bootstraps, batch):
            b = min(n_bootstraps, a + batch)
            idx = np.stack([rng.integers(0, ndata, ndata) for _ in range(a, b)])
            if uniform:
                # same times, same (uniform) weights: rotation and window sums are shared (pdc_gls_multi)
                psd_scale = 0.5 * (err ** -2.0).sum() if self.psd else None
                _, _, mx = ctx.gls_multi(t, values[idx], None, fmin, df, nf, fit_mean=True, psd_scale=psd_scale,
                                         want_power=False)
            else:
                yb = values[idx].ravel()
                eb = err[idx]
                wb = (eb ** -2.0).ravel()
                offsets = np.arange(b - a + 1, dtype=np.int64) * ndata
                psd_scale = 0.5 * (eb ** -2.0).sum(axis=1) if self.psd else None
                _, _, mx = ctx.gls_batch(np.tile(t, b - a), yb, wb, offsets, fmin, df, nf, fit_mean=True,
                                         psd_scale=psd_scale, want_power=False)
            out[a:b] = mx
        self.bs_replicates = out
        return self.bs_replicates

    def fap(self, power):
        """False-alarm probability of ``power`` from the bootstrap replicates (``spectral.py:154-160``)."""
        return np.mean(power < self.bs_replicates)

    def fal(self, fap):
        """Power level with false-alarm probability ``fap`` (``spectral.py:162-163``)."""
        return np.quantile(self.bs_replicates, 1 - fap)

    def window(self):
        """Spectral window: periodogram of an all-ones signal without mean fit (``spectral.py:165-167``)."""
        gls = self.copy()
        return gls(0.0 * self.signal + 1.0, fit_mean=False)

    def model(self, tf, f0):
        """Best-fit offset + sinusoid at frequency ``f0`` evaluated at ``tf`` (``spectral.py:169-204``)."""
        t = self.signal.time
        sigma = np.asarray(self.err, dtype=np.float64)
        w = sigma ** -2.0
        y_mean = np.dot(self.signal.values, w) / w.sum()
        resid = self.signal.values - y_mean

        def design(times):
            arg = 2 * np.pi * f0 * np.asarray(times)
            return np.vstack([np.ones_like(arg), np.sin(arg), np.cos(arg)])

        A = design(t) / sigma
        theta = np.linalg.solve(A @ A.T, A @ (resid / sigma))
        return TSeries(tf, y_mean + design(tf).T @ theta)


class BGLST(object):
    """Placeholder kept for API parity (empty in the reference too, ``spectral.py:207-208``)."""
    pass
