"""Batched GLS over many light curves (the survey workload, BASELINE config C4).

The reference has no batch API: a survey is a Python loop over ``GLS()(signal)``
 and ``GLS.bootstrap`` is the same loop over resamples
(``spectral.py:145-150``).  ``gls_survey`` evaluates all curves in one
``pdc_gls_batch`` call; each curve keeps its own grid origin and spacing
(``df = 1/(baseline*n)``, ``fmin = 0.5*df`` as ``spectral.py:88-92``) and all
curves share the number of frequencies ``nf``.
"""
import numpy as np

from . import _ffi
from .core import TSeries

__all__ = ["gls_survey"]


def gls_survey(signals, errs=None, nf=10_000, n=5, fmin=None, psd=False, fit_mean=True, want_power=False,
               device=None, shard=False, top_k=None):
    """GLS peak search over a list of light curves.

    signals : sequence of ``TSeries`` (or array-likes, coerced like ``spectral.py:86-87``)
    errs    : optional sequence of per-sample uncertainties (all curves or none)
    nf      : number of trial frequencies per curve;  n : samples per peak (``GLS.n``)
    fmin    : optional common minimum frequency (default ``0.5*df`` per curve)

    top_k   : if set, also return the ``top_k`` highest periodogram *peaks* (local maxima, as
              ``FSeries.find_peaks`` defines them, ``core.py:283-317``) of every curve; the periodograms
              stay on the GPU and only ``[B, top_k]`` indices and powers come back (``pdc_peaks_topk_dev``)

    Returns a dict with ``fmin, df`` (per curve), ``argmax, max_power, best_frequency,
    best_period`` and, if ``want_power``, ``power`` of shape ``[len(signals), nf]``; with ``top_k``
    also ``peak_index, peak_power, peak_frequency`` of shape ``[len(signals), top_k]``.
    With ``shard=True`` (and ``torch.distributed`` initialised) the curves are split across
    ranks and every rank receives all results.
    """
    sigs = [s if isinstance(s, TSeries) else TSeries(values=s) for s in signals]
    B = len(sigs)
    if B == 0:
        raise ValueError("need at least one light curve")
    if errs is not None and len(errs) != B:
        raise ValueError("errs must have one entry per light curve")
    sizes = np.array([len(s) for s in sigs], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    t = np.concatenate([np.asarray(s.time, dtype=np.float64) for s in sigs])
    y = np.concatenate([np.asarray(s.values, dtype=np.float64) for s in sigs])
    w = None
    psd_scale = None
    if errs is not None:
        w = np.concatenate([np.asarray(e, dtype=np.float64) ** -2.0 for e in errs])
        if w.size != t.size:
            raise ValueError("Input arrays have incompatible lengths.")
    if psd:
        psd_scale = np.array([0.5 * (w[offsets[b]:offsets[b + 1]].sum() if w is not None else sizes[b])
                              for b in range(B)])
    df = np.array([1.0 / s.baseline / n for s in sigs])
    f0 = 0.5 * df if fmin is None else np.full(B, float(fmin))
    peaks = None
    if top_k is not None:
        if shard:
            raise ValueError("top_k and shard cannot be combined yet")
        import torch
        from . import dist
        ctx = _ffi.default_context(device)
        dev = torch.device("cuda", ctx.device)
        td, yd = torch.from_numpy(t).to(dev), torch.from_numpy(y).to(dev)
        wd = None if w is None else torch.from_numpy(w).to(dev)
        pw, argd, mxd = dist.gls_batch_torch(td, yd, wd, offsets, f0, df, nf, fit_mean, psd_scale, True, ctx=ctx)
        pidx = torch.empty((B, int(top_k)), dtype=torch.int64, device=dev)
        pval = torch.empty((B, int(top_k)), dtype=torch.float64, device=dev)
        ctx.peaks_topk_dev(pw.data_ptr(), B, nf, int(top_k), pidx.data_ptr(), pval.data_ptr(),
                           torch.cuda.current_stream(dev).cuda_stream)
        arg, mx = argd.cpu().numpy(), mxd.cpu().numpy()
        power = pw.cpu().numpy() if want_power else None
        peaks = (pidx.cpu().numpy(), pval.cpu().numpy())
    elif shard:
        from . import dist
        power, arg, mx = dist.gls_batch_sharded(t, y, w, offsets, f0, df, nf, fit_mean, psd_scale, want_power,
                                                device=device)
    else:
        power, arg, mx = _ffi.default_context(device).gls_batch(t, y, w, offsets, f0, df, nf, fit_mean=fit_mean,
                                                                psd_scale=psd_scale, want_power=want_power)
    best_f = f0 + df * arg
    out = {"fmin": f0, "df": df, "argmax": arg, "max_power": mx, "best_frequency": best_f,
           "best_period": 1.0 / best_f}
    if want_power:
        out["power"] = power
    if peaks is not None:
        out["peak_index"], out["peak_power"] = peaks
        out["peak_frequency"] = np.where(peaks[0] >= 0, f0[:, None] + df[:, None] * peaks[0], np.nan)
    return out
