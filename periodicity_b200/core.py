"""numpy-only ``TSeries`` / ``FSeries`` containers for the GLS / PDM drop-ins.

The reference wraps ``xarray.DataArray`` (``src/periodicity/core.py:53-58``);
xarray is not part of this stack, and the hot path only touches a small part
of the container API.  This module restates exactly that
part on plain numpy arrays, keeping names and semantics:

``TSeries``  constructor defaults / length check / sort by time
             (``core.py:460-477``), ``time`` (``:479-481``), ``values``
             (``:60-66``), ``size`` / ``len`` (``:94-99``), ``baseline``
             (``:504-506``), ``median_dt`` / ``dt`` (``:508-519``), ``copy``
             (``:144-145``), scalar / elementwise arithmetic (``:158-187``),
             ``__getitem__`` (``:489-494``), ``fold`` / ``timeshift`` /
             ``timescale`` (``:537-544``).
``FSeries``  constructor (adds ``period = 1/frequency``, sorts ascending in
             frequency, ``core.py:859-881``), ``frequency`` / ``period``
             (``:883-889``), ``__getitem__`` (``:897-902``), NaN-aware
             ``argmax / argmin / amax / amin / max / min`` (``:202-240``),
             ``fmax / pmax`` (``:938-942``), ``find_peaks`` (``:283-317``),
             ``psort_by_peak`` / ``psort_by_prominence`` (``:944-950``),
             ``period_at_highest_peak`` / ``period_at_highest_prominence``
             (``:952-961``), ``median_df`` / ``df`` (``:911-922``).

Everything else in the reference's ``core.py`` (filters, envelopes, TFSeries,
...) is outside the hot path and intentionally absent.
"""
from numbers import Number

import numpy as np

__all__ = ["TSeries", "FSeries"]


class _Series(np.lib.mixins.NDArrayOperatorsMixin):
    """1-D values labelled by one monotonically increasing coordinate."""

    _coord_name = "index"
    __array_priority__ = 100

    def __init__(self, coord, values, assume_sorted):
        coord = np.asarray(coord)
        values = np.asarray(values)
        if coord.ndim != 1 or values.ndim != 1 or coord.size != values.size:
            raise ValueError("Input arrays have incompatible lengths.")
        if not assume_sorted and coord.size > 1 and np.any(coord[1:] < coord[:-1]):
            order = np.argsort(coord, kind="stable")
            coord = coord[order]
            values = values[order]
        self._coord = coord
        self._values = values
        self.attrs = {}

    # -- array protocol ------------------------------------------------------
    @property
    def values(self):
        return self._values

    @values.setter
    def values(self, new):
        new = np.asarray(new)
        if new.shape != self._values.shape:
            raise ValueError("replacement data must match the series shape")
        self._values = new

    @property
    def size(self):
        return self._values.size

    @property
    def shape(self):
        return self._values.shape

    @property
    def ndim(self):
        return 1

    @property
    def dtype(self):
        return self._values.dtype

    def __len__(self):
        return self._values.shape[0]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._values, dtype=dtype)

    def _like(self, values):
        new = type(self)(self._coord, values, assume_sorted=True)
        new.attrs.update(self.attrs)
        return new

    def copy(self):
        return self._like(self._values.copy())

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method not in ("__call__", "reduce"):
            return NotImplemented
        raw = []
        for x in inputs:
            if isinstance(x, _Series):
                if type(x) is not type(self) or (x is not self and not np.array_equal(x._coord, self._coord)):
                    raise ValueError("series are not aligned on the same coordinate")
                raw.append(x._values)
            elif isinstance(x, (Number, np.ndarray, np.generic)):
                raw.append(x)
            else:
                return NotImplemented
        if "out" in kwargs:
            kwargs["out"] = tuple(o._values if isinstance(o, _Series) else o for o in kwargs["out"])
        result = getattr(ufunc, method)(*raw, **kwargs)
        if method == "reduce":
            return result.item() if np.ndim(result) == 0 else result
        if isinstance(result, tuple):
            return tuple(self._like(r) for r in result)
        if np.ndim(result) == 0:
            return result
        return self._like(result)

    # -- NaN-aware reductions (core.py:202-262) -------------------------------
    def _reduce(self, func, axis=None, out=None, **kw):
        if axis not in (None, 0, -1) or out is not None:
            raise ValueError("series are one-dimensional; only full reductions are supported")
        kw.pop("keepdims", None)
        result = func(self._values, **{k: v for k, v in kw.items() if v is not None})
        return result.item() if hasattr(result, "item") else result

    def argmax(self, **kw):
        return int(self._reduce(np.nanargmax, **kw))

    def argmin(self, **kw):
        return int(self._reduce(np.nanargmin, **kw))

    def amax(self, **kw):
        return self._reduce(np.nanmax, **kw)

    def amin(self, **kw):
        return self._reduce(np.nanmin, **kw)

    def max(self, **kw):
        i = self.argmax()
        return self[i:i + 1]

    def min(self, **kw):
        i = self.argmin()
        return self[i:i + 1]

    def mean(self, **kw):
        return self._reduce(np.nanmean, **kw)

    def median(self, **kw):
        return self._reduce(np.nanmedian, **kw)

    def std(self, **kw):
        return self._reduce(np.nanstd, **kw)

    def var(self, **kw):
        return self._reduce(np.nanvar, **kw)

    def sum(self, **kw):
        return self._reduce(np.nansum, **kw)

    def __getitem__(self, key):
        coord = self._coord[key]
        values = self._values[key]
        if np.ndim(values) < 1:
            return values.item()
        new = type(self)(coord, values)
        return new

    def find_peaks(self, include_edges=False, prominence=0.0, **peak_kwargs):
        """Local maxima with prominences in ``attrs`` (``core.py:283-317``)."""
        from scipy import signal as _signal

        maxima, res = _signal.find_peaks(self._values, prominence=prominence, **peak_kwargs)
        if include_edges:
            maxima = np.hstack([0, maxima, -1])
            for key, val in res.items():
                fill = np.nan if val.dtype.kind == "f" else -1
                res[key] = np.hstack([fill, val, fill])
        res["indices"] = maxima
        peaks = self[maxima]
        peaks.attrs.update(res)
        return peaks

    def find_zero_crossings(self, height=None, delta=0.0):
        """Indices ``j`` with a sign change between samples ``j`` and ``j + 1`` (``core.py:341-367``); with
        ``height`` the near-zero minima of ``|values|`` instead, as the reference does."""
        if height is None:
            (ind,) = np.where(np.diff(np.signbit(self._values)))
            return ind
        from scipy import signal as _signal
        ind, _ = _signal.find_peaks(-np.abs(self._values), height=-height, prominence=delta)
        return ind

    def __repr__(self):
        return (f"<{type(self).__name__} ({self._coord_name}: {self.size})>\n"
                f"{self._coord_name}: {self._coord!r}\nvalues: {self._values!r}")


class TSeries(_Series):
    """Time series: ``TSeries(time=None, values=None, assume_sorted=False)``."""

    _coord_name = "time"

    def __init__(self, time=None, values=None, assume_sorted=False):
        if time is None:
            time = np.arange(len(values))
        if values is None:
            values = np.ones(len(time))
        super().__init__(time, values, assume_sorted)

    @property
    def time(self):
        return self._coord

    @property
    def baseline(self):
        return self._coord[-1] - self._coord[0]

    @property
    def median_dt(self):
        return np.median(np.diff(self._coord))

    @property
    def dt(self):
        if np.allclose(np.diff(self._coord), self.median_dt):
            return self.median_dt
        raise AttributeError(
            "The sampling period is only strictly defined for uniformly sampled signals. "
            "Use median_dt for a median value.")

    def tmax(self):
        return self.max().time.item()

    def timeshift(self, t0):
        return TSeries(self._coord + t0, self._values)

    def timescale(self, alpha):
        return TSeries(self._coord * alpha, self._values)

    def fold(self, period, t0=0):
        return TSeries(((self._coord - t0) / period) % 1, self._values)


class FSeries(_Series):
    """Frequency series: ``FSeries(frequency=None, values=None, assume_sorted=False)``."""

    _coord_name = "frequency"

    def __init__(self, frequency=None, values=None, assume_sorted=False):
        if values is None:
            values = np.ones(len(frequency))
        super().__init__(frequency, values, assume_sorted)

    @property
    def frequency(self):
        return self._coord

    @property
    def period(self):
        with np.errstate(divide="ignore", invalid="ignore"):
            return 1.0 / self._coord

    @property
    def median_df(self):
        return np.median(np.diff(self._coord))

    @property
    def df(self):
        if np.allclose(np.diff(self._coord), self.median_df):
            return self.median_df
        raise AttributeError(
            "The sampling period is only strictly defined for uniform frequency grids. "
            "Use median_df for a median value.")

    def fmax(self):
        return self.max().frequency.item()

    def pmax(self):
        return self.max().period.item()

    def psort_by_peak(self):
        peaks = self.find_peaks()
        return peaks.period[peaks.values.argsort()[::-1]]

    def psort_by_prominence(self):
        peaks = self.find_peaks()
        return peaks.period[peaks.attrs["prominences"].argsort()[::-1]]

    @property
    def period_at_highest_peak(self):
        return self.find_peaks().pmax()

    @property
    def period_at_highest_prominence(self):
        peaks = self.find_peaks()
        return peaks.period[np.nanargmax(peaks.attrs["prominences"])]

    def periods_at_half_max(self, peak_order=1, use_prominence=False):
        """``(lower, upper)`` periods where the periodogram falls to half the height (or half the prominence) of
        its ``peak_order``-th highest peak (``core.py:957-972``).  The level is ``values[peak] - height / 2`` as a
        scalar: the reference builds it as a one-sample series, which its own exact-join arithmetic cannot
        subtract from the slices (same defect as ``phase.py:65``); the index arithmetic is the reference's."""
        peaks = self.find_peaks()
        indices = peaks.attrs["indices"]
        heights = peaks.attrs["prominences"] if use_prominence else peaks.values
        jmax = heights.argsort()[-peak_order]
        idmax = indices[jmax]
        half = self._values[idmax] - heights[jmax] / 2
        hi = (self[:idmax] - half).find_zero_crossings()[-1]
        lo = (self[idmax:] - half).find_zero_crossings()[0]
        return self[idmax:].period[lo], self[:idmax].period[hi]
