"""ctypes binding of ``libperiodicity_b200.so`` (C ABI: ``include/periodicity_b200.h``).

The library is the only compute path: there is no CPU fallback.  If the shared
object is missing or no B200 is visible, the first call raises ``RuntimeError``.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PERIODICITY_B200_LIB: explicit path of the shared object (tuning builds, tools/tune_strip.py); default = in-tree build
LIB_PATH = os.environ.get("PERIODICITY_B200_LIB") or os.path.join(_HERE, "lib", "libperiodicity_b200.so")

PDC_OK, PDC_EINVAL, PDC_ECUDA, PDC_ENOMEM, PDC_ENODEVICE = range(5)
GLS_FIT_MEAN = 1
GLS_PSD = 2
STREAM_CTX = ctypes.c_void_p(-1).value  # PDC_STREAM_CTX: the ctx's own stream (0/None = CUDA default stream)

MAX_PEERS = 16


class Fanout(ctypes.Structure):
    """``pdc_fanout``: per-rank destination buffers of the fused epilogue + all-gather."""
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("power", ctypes.c_void_p * MAX_PEERS), ("best", ctypes.c_void_p * MAX_PEERS)]


_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)

#: every symbol declared in include/periodicity_b200.h -> (restype, argtypes)
SIGNATURES = {
    "pdc_version": (ctypes.c_int, []),
    "pdc_last_error": (ctypes.c_char_p, []),
    "pdc_ctx_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "pdc_ctx_create_multi": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "pdc_ctx_device_count": (ctypes.c_int, [ctypes.c_void_p]),
    "pdc_ctx_device_id": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "pdc_ctx_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "pdc_ctx_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "pdc_ctx_sm_count": (ctypes.c_int, [ctypes.c_void_p]),
    "pdc_ctx_launch_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "pdc_ctx_last_main_kernel_ms": (ctypes.c_double, [ctypes.c_void_p]),
    "pdc_ctx_main_kernel_ms_total": (ctypes.c_double, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]),
    "pdc_debug_umma_prof": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "pdc_ctx_last_gls_path": (ctypes.c_int, [ctypes.c_void_p]),
    "pdc_debug_umma_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "pdc_gls": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                               ctypes.c_int64, ctypes.c_uint, ctypes.c_double,
                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gls_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                   ctypes.c_int64, ctypes.c_uint, ctypes.c_double,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gls_freqs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint, ctypes.c_double,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gls_freqs_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gls_dev_fanout": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_uint, ctypes.c_double,
                                          ctypes.POINTER(Fanout), ctypes.c_void_p]),
    "pdc_gls_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_uint, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gls_batch_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_uint, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gls_multi": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int64, ctypes.c_int64, ctypes.c_uint, ctypes.c_double,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gls_multi_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                         ctypes.c_int64, ctypes.c_int64, ctypes.c_uint, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_pdm": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_aov": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                               ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_aov_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_ce": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                              ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_ce_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                              ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_gl_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p]),
    "pdc_stringlength": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "pdc_stringlength_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_peaks_topk": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_peaks_halfmax": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_peaks_halfmax_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_peaks_topk_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "pdc_pdm_dev_fanout": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                          ctypes.POINTER(Fanout), ctypes.c_void_p]),
    "pdc_pdm_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load the shared object and attach prototypes. Raises RuntimeError if absent."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C periodicity_b200/csrc`. periodicity_b200 has no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def _check(rc):
    if rc == PDC_OK:
        return
    msg = load_library().pdc_last_error().decode("utf-8", "replace")
    if rc == PDC_EINVAL:
        raise ValueError(msg)
    if rc == PDC_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def _f64(a):
    """C-contiguous float64 view/copy of ``a`` (what the C ABI consumes)."""
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data


class Context:
    """One CUDA device + stream + grow-only scratch (``pdc_ctx``), or -- ``device`` given as a list / tuple of
    ordinals -- a multi-device ctx (``pdc_ctx_create_multi``) whose host entry points shard the grid or the batch over
    the devices inside the library.  Not thread-safe."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = [int(d) for d in device]
            if not ids:
                raise ValueError("empty device list")
            arr = (ctypes.c_int * len(ids))(*ids)
            self.device = ids[0]
            self.devices = tuple(ids)
            _check(self._lib.pdc_ctx_create_multi(ctypes.byref(self._h), arr, len(ids)))
        else:
            self.device = int(device)
            self.devices = (self.device,)
            _check(self._lib.pdc_ctx_create(ctypes.byref(self._h), self.device))
        self._lock = threading.Lock()

    @property
    def device_count(self):
        return self._lib.pdc_ctx_device_count(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.pdc_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return self._lib.pdc_ctx_sm_count(self._h)

    @property
    def launch_count(self):
        return self._lib.pdc_ctx_launch_count(self._h)

    def last_main_kernel_ms(self):
        return self._lib.pdc_ctx_last_main_kernel_ms(self._h)

    def main_kernel_ms_total(self):
        """(cumulative ms, launch count) of the dominant kernel since the ctx was created."""
        cnt = ctypes.c_int64(0)
        ms = self._lib.pdc_ctx_main_kernel_ms_total(self._h, ctypes.byref(cnt))
        return ms, cnt.value

    def last_gls_path(self):
        """0 FP32 strip kernel, 1 tensor-core kernel, 2 the same with the precomputed fine operand, 3 the 2-CTA tensor-core kernel."""
        return self._lib.pdc_ctx_last_gls_path(self._h)

    def umma_prof(self, cap=1 << 16):
        """Clock stamps [jobs, 4] of the last tensor-core GLS launch (needs PDC_GLS_UMMA_PROF=1 at ctx creation)."""
        out = np.zeros((cap, 4), dtype=np.int64)
        n = self._lib.pdc_debug_umma_prof(self._h, _ptr(out), cap)
        if n < 0:
            raise RuntimeError("pdc_debug_umma_prof failed")
        return out[:min(n, cap)]

    def umma_trace(self):
        """Per-pair clock stamps [1024, 8] of block 0 (PDC_GLS_UMMA_PROF=1 and PDC_GLS_UMMA_DBG & 16)."""
        cap = 1 << 16
        out = np.zeros((cap, 4), dtype=np.int64)
        n = self._lib.pdc_debug_umma_prof(self._h, _ptr(out), cap)
        if n <= 2048:
            return np.zeros((0, 8), dtype=np.int64)
        return out[n - 2048:n].reshape(1024, 8)

    def synchronize(self):
        _check(self._lib.pdc_ctx_synchronize(self._h))

    # ---- host-pointer entry points (numpy in, numpy out) --------------------
    def gls(self, t, y, w, fmin, df, nf, fit_mean=True, psd_scale=None, j0=0, want_power=True):
        t = _f64(t)
        y = _f64(y)
        if t.ndim != 1 or t.shape != y.shape:
            raise ValueError("Input arrays have incompatible lengths.")
        if w is not None:
            w = _f64(w)
            if w.shape != t.shape:
                raise ValueError("Input arrays have incompatible lengths.")
        nf = int(nf)
        flags = (GLS_FIT_MEAN if fit_mean else 0) | (GLS_PSD if psd_scale is not None else 0)
        power = np.empty(nf, dtype=np.float64) if want_power else None
        arg = ctypes.c_int64(-1)
        mx = ctypes.c_double(float("nan"))
        with self._lock:
            _check(self._lib.pdc_gls(self._h, _ptr(t), _ptr(y), _ptr(w), t.size, float(fmin), float(df),
                                     int(j0), nf, flags, float(psd_scale if psd_scale is not None else 1.0),
                                     _ptr(power), ctypes.addressof(arg), ctypes.addressof(mx)))
        return power, arg.value, mx.value

    def gls_freqs(self, t, y, w, freqs, fit_mean=True, psd_scale=None):
        """GLS power at an arbitrary list of frequencies (``pdc_gls_freqs``): (power, argmax, max)."""
        t = _f64(t)
        y = _f64(y)
        freqs = _f64(freqs)
        if t.ndim != 1 or t.shape != y.shape:
            raise ValueError("Input arrays have incompatible lengths.")
        if w is not None:
            w = _f64(w)
            if w.shape != t.shape:
                raise ValueError("Input arrays have incompatible lengths.")
        flags = (GLS_FIT_MEAN if fit_mean else 0) | (GLS_PSD if psd_scale is not None else 0)
        power = np.empty(freqs.size, dtype=np.float64)
        arg = ctypes.c_int64(-1)
        mx = ctypes.c_double(float("nan"))
        with self._lock:
            _check(self._lib.pdc_gls_freqs(self._h, _ptr(t), _ptr(y), _ptr(w), t.size, _ptr(freqs), freqs.size, flags,
                                           float(psd_scale if psd_scale is not None else 1.0), _ptr(power),
                                           ctypes.addressof(arg), ctypes.addressof(mx)))
        return power, arg.value, mx.value

    def gls_batch(self, t, y, w, offsets, fmin, df, nf, fit_mean=True, psd_scale=None, want_power=True):
        t = _f64(t)
        y = _f64(y)
        if w is not None:
            w = _f64(w)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        B = offsets.size - 1
        fmin = _f64(np.broadcast_to(fmin, (B,)))
        df = _f64(np.broadcast_to(df, (B,)))
        if psd_scale is not None:
            psd_scale = _f64(np.broadcast_to(psd_scale, (B,)))
        nf = int(nf)
        if B < 1 or offsets[0] < 0 or offsets[-1] > t.size or np.any(np.diff(offsets) < 1):
            raise ValueError("offsets must be increasing and within the sample arrays")
        flags = (GLS_FIT_MEAN if fit_mean else 0) | (GLS_PSD if psd_scale is not None else 0)
        power = np.empty((B, nf), dtype=np.float64) if want_power else None
        arg = np.empty(B, dtype=np.int64)
        mx = np.empty(B, dtype=np.float64)
        with self._lock:
            _check(self._lib.pdc_gls_batch(self._h, _ptr(t), _ptr(y), _ptr(w), _ptr(offsets), B, _ptr(fmin),
                                           _ptr(df), nf, flags, _ptr(psd_scale), _ptr(power), _ptr(arg), _ptr(mx)))
        return power, arg, mx

    def gls_multi(self, t, Y, w, fmin, df, nf, fit_mean=True, psd_scale=None, j0=0, want_power=True):
        """GLS of the rows of ``Y`` [S, n], all sampled at the common times ``t`` (``pdc_gls_multi``)."""
        t = _f64(t)
        Y = np.atleast_2d(_f64(Y))
        if t.ndim != 1 or Y.shape[1] != t.size:
            raise ValueError("Input arrays have incompatible lengths.")
        if w is not None:
            w = _f64(w)
            if w.shape != t.shape:
                raise ValueError("Input arrays have incompatible lengths.")
        S, nf = Y.shape[0], int(nf)
        flags = (GLS_FIT_MEAN if fit_mean else 0) | (GLS_PSD if psd_scale is not None else 0)
        power = np.empty((S, nf), dtype=np.float64) if want_power else None
        arg = np.empty(S, dtype=np.int64)
        mx = np.empty(S, dtype=np.float64)
        with self._lock:
            _check(self._lib.pdc_gls_multi(self._h, _ptr(t), _ptr(Y), _ptr(w), t.size, S, float(fmin), float(df),
                                           int(j0), nf, flags, float(psd_scale if psd_scale is not None else 1.0),
                                           _ptr(power), _ptr(arg), _ptr(mx)))
        return power, arg, mx

    def pdm(self, t, x, periods, nb, nc):
        t = _f64(t)
        x = _f64(x)
        periods = _f64(periods)
        if t.ndim != 1 or t.shape != x.shape:
            raise ValueError("Input arrays have incompatible lengths.")
        theta = np.empty(periods.size, dtype=np.float64)
        arg = ctypes.c_int64(-1)
        mn = ctypes.c_double(float("nan"))
        with self._lock:
            _check(self._lib.pdc_pdm(self._h, _ptr(t), _ptr(x), t.size, _ptr(periods), periods.size,
                                     int(nb), int(nc), _ptr(theta), ctypes.addressof(arg), ctypes.addressof(mn)))
        return theta, arg.value, mn.value

    def aov(self, t, x, periods, nb):
        """Analysis-of-variance statistic for each trial period (``pdc_aov``): (theta, argmax, max)."""
        t = _f64(t)
        x = _f64(x)
        periods = _f64(periods)
        if t.ndim != 1 or t.shape != x.shape:
            raise ValueError("Input arrays have incompatible lengths.")
        theta = np.empty(periods.size, dtype=np.float64)
        arg = ctypes.c_int64(-1)
        mx = ctypes.c_double(float("nan"))
        with self._lock:
            _check(self._lib.pdc_aov(self._h, _ptr(t), _ptr(x), t.size, _ptr(periods), periods.size,
                                     int(nb), _ptr(theta), ctypes.addressof(arg), ctypes.addressof(mx)))
        return theta, arg.value, mx.value

    def ce(self, t, x, periods, nphi, nm):
        """Conditional entropy for each trial period (``pdc_ce``): (h, argmin, min)."""
        t = _f64(t)
        x = _f64(x)
        periods = _f64(periods)
        if t.ndim != 1 or t.shape != x.shape:
            raise ValueError("Input arrays have incompatible lengths.")
        h = np.empty(periods.size, dtype=np.float64)
        arg = ctypes.c_int64(-1)
        mn = ctypes.c_double(float("nan"))
        with self._lock:
            _check(self._lib.pdc_ce(self._h, _ptr(t), _ptr(x), t.size, _ptr(periods), periods.size,
                                    int(nphi), int(nm), _ptr(h), ctypes.addressof(arg), ctypes.addressof(mn)))
        return h, arg.value, mn.value

    def ce_dev(self, t_ptr, x_ptr, n, periods_ptr, np_, nphi, nm, h_ptr, argmin_ptr, min_ptr, stream=0):
        _check(self._lib.pdc_ce_dev(self._h, t_ptr, x_ptr, int(n), periods_ptr, int(np_), int(nphi), int(nm),
                                    h_ptr, argmin_ptr or None, min_ptr or None, stream or None))

    def gl(self, t, periods, m_max=12, nc=10):
        """Gregory-Loredo ln odds for each trial period from event arrival times (``pdc_gl``): (lnodds, argmax, max)."""
        t = _f64(t)
        periods = _f64(periods)
        if t.ndim != 1:
            raise ValueError("t must be one-dimensional")
        out = np.empty(periods.size, dtype=np.float64)
        arg = ctypes.c_int64(-1)
        mx = ctypes.c_double(float("nan"))
        with self._lock:
            _check(self._lib.pdc_gl(self._h, _ptr(t), t.size, _ptr(periods), periods.size, int(m_max), int(nc),
                                    _ptr(out), ctypes.addressof(arg), ctypes.addressof(mx)))
        return out, arg.value, mx.value

    def peaks_halfmax(self, values, peak_idx, height=None):
        """Indices (left, right) of the half-maximum crossings around each given peak of each row (host arrays);
        ``height`` defaults to the peak values (``use_prominence=False`` of ``periods_at_half_max``)."""
        v = np.atleast_2d(_f64(values))
        rows, n = v.shape
        idx = np.ascontiguousarray(np.atleast_2d(peak_idx), dtype=np.int64)
        if idx.shape[0] != rows:
            raise ValueError("peak_idx must have one row per row of values")
        k = idx.shape[1]
        h = None if height is None else np.ascontiguousarray(np.atleast_2d(_f64(height)))
        left = np.empty((rows, k), dtype=np.int64)
        right = np.empty((rows, k), dtype=np.int64)
        with self._lock:
            _check(self._lib.pdc_peaks_halfmax(self._h, _ptr(v), rows, n, k, _ptr(idx), _ptr(h), _ptr(left), _ptr(right)))
        return left, right

    def stringlength(self, t, m, periods):
        """String length of the scaled signal ``m`` for each trial period (host arrays)."""
        t = _f64(t)
        m = _f64(m)
        periods = _f64(periods)
        if t.ndim != 1 or t.shape != m.shape:
            raise ValueError("Input arrays have incompatible lengths.")
        ell = np.empty(periods.size, dtype=np.float64)
        arg = ctypes.c_int64(-1)
        mn = ctypes.c_double(float("nan"))
        with self._lock:
            _check(self._lib.pdc_stringlength(self._h, _ptr(t), _ptr(m), t.size, _ptr(periods), periods.size,
                                              _ptr(ell), ctypes.addressof(arg), ctypes.addressof(mn)))
        return ell, arg.value, mn.value

    def stringlength_dev(self, t_ptr, m_ptr, n, periods_ptr, np_, ell_ptr, argmin_ptr, min_ptr, stream=0):
        _check(self._lib.pdc_stringlength_dev(self._h, t_ptr, m_ptr, int(n), periods_ptr, int(np_), ell_ptr,
                                              argmin_ptr or None, min_ptr or None, stream or None))

    def peaks_topk(self, values, k):
        """(indices [rows, k], values [rows, k]) of the k highest local maxima of each row (host arrays)."""
        v = np.atleast_2d(_f64(values))
        rows, n = v.shape
        idx = np.empty((rows, int(k)), dtype=np.int64)
        val = np.empty((rows, int(k)), dtype=np.float64)
        with self._lock:
            _check(self._lib.pdc_peaks_topk(self._h, _ptr(v), rows, n, int(k), _ptr(idx), _ptr(val)))
        return idx, val

    def peaks_topk_dev(self, values_ptr, rows, n, k, idx_ptr, val_ptr, stream=0):
        _check(self._lib.pdc_peaks_topk_dev(self._h, values_ptr, int(rows), int(n), int(k), idx_ptr, val_ptr,
                                            stream or None))

    # ---- device-pointer entry points (raw addresses, stream ordered) ---------
    def gls_dev(self, t_ptr, y_ptr, w_ptr, n, fmin, df, j0, nf, flags, psd_scale, power_ptr, argmax_ptr,
                max_ptr, stream=0):
        _check(self._lib.pdc_gls_dev(self._h, t_ptr, y_ptr, w_ptr or None, int(n), float(fmin), float(df),
                                     int(j0), int(nf), int(flags), float(psd_scale), power_ptr or None,
                                     argmax_ptr or None, max_ptr or None, stream or None))

    def gls_dev_fanout(self, t_ptr, y_ptr, w_ptr, n, fmin, df, j0, nf, flags, psd_scale, fanout, stream=0):
        _check(self._lib.pdc_gls_dev_fanout(self._h, t_ptr, y_ptr, w_ptr or None, int(n), float(fmin), float(df),
                                            int(j0), int(nf), int(flags), float(psd_scale), ctypes.byref(fanout),
                                            stream or None))

    def pdm_dev_fanout(self, t_ptr, x_ptr, n, periods_ptr, np_, nb, nc, offset, fanout, stream=0):
        _check(self._lib.pdc_pdm_dev_fanout(self._h, t_ptr, x_ptr, int(n), periods_ptr, int(np_), int(nb), int(nc),
                                            int(offset), ctypes.byref(fanout), stream or None))

    def gls_batch_dev(self, t_ptr, y_ptr, w_ptr, offsets, fmin, df, nf, flags, psd_scale, power_ptr,
                      argmax_ptr, max_ptr, stream=0):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        B = offsets.size - 1
        fmin = _f64(np.broadcast_to(fmin, (B,)))
        df = _f64(np.broadcast_to(df, (B,)))
        if psd_scale is not None:
            psd_scale = _f64(np.broadcast_to(psd_scale, (B,)))
        _check(self._lib.pdc_gls_batch_dev(self._h, t_ptr, y_ptr, w_ptr or None, _ptr(offsets), B, _ptr(fmin),
                                           _ptr(df), int(nf), int(flags), _ptr(psd_scale), power_ptr or None,
                                           argmax_ptr or None, max_ptr or None, stream or None))

    def aov_dev(self, t_ptr, x_ptr, n, periods_ptr, np_, nb, theta_ptr, argmax_ptr, max_ptr, stream=0):
        _check(self._lib.pdc_aov_dev(self._h, t_ptr, x_ptr, int(n), periods_ptr, int(np_), int(nb),
                                     theta_ptr, argmax_ptr or None, max_ptr or None, stream or None))

    def pdm_dev(self, t_ptr, x_ptr, n, periods_ptr, np_, nb, nc, theta_ptr, argmin_ptr, min_ptr, stream=0):
        _check(self._lib.pdc_pdm_dev(self._h, t_ptr, x_ptr, int(n), periods_ptr, int(np_), int(nb), int(nc),
                                     theta_ptr, argmin_ptr or None, min_ptr or None, stream or None))


_default_ctx = {}
_default_lock = threading.Lock()


def umma_plan(B, nf, nmax, sm_count=148, fine=-1, cg2=-1, nsplit=0, chunk=0):
    """Work decomposition of the tensor-core GLS kernels for a call (pure host arithmetic: works without a GPU)."""
    out = np.zeros(11, dtype=np.int64)
    _check(load_library().pdc_debug_umma_plan(int(sm_count), int(B), int(nf), int(nmax), int(fine), int(cg2), int(nsplit),
                                              int(chunk), _ptr(out)))
    keys = ["path", "fine", "nC", "nt1", "cpt1", "nt2", "cpt2", "nsplit", "chunk_stages", "jobs", "fine_bytes"]
    return dict(zip(keys, (int(v) for v in out)))


def default_context(device=None):
    """Process-wide context for ``device`` (default: LOCAL_RANK or 0), created on first use.  ``device`` may be a
    list / tuple of ordinals: a multi-device context (one entry: the ordinary context of that device)."""
    if device is None:
        device = int(os.environ.get("PERIODICITY_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if isinstance(device, (list, tuple)):
        device = tuple(int(d) for d in device)
        if len(device) == 1:
            device = device[0]
    with _default_lock:
        ctx = _default_ctx.get(device)
        if ctx is None:
            ctx = Context(device)
            _default_ctx[device] = ctx
    return ctx
