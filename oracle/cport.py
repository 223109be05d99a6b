"""ctypes loader for ``oracle/liboracle.so`` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.orc_gls_exact.restype = ctypes.c_int
        lib.orc_gls_exact.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
        lib.orc_gls_exact_at.restype = ctypes.c_int
        lib.orc_gls_exact_at.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p]
        lib.orc_gls_exact_freqs.restype = ctypes.c_int
        lib.orc_gls_exact_freqs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_double, ctypes.c_void_p]
        lib.orc_pdm.restype = ctypes.c_int
        lib.orc_pdm.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.orc_num_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def gls_exact(t, y, err, fmin, df, nf, fit_mean=True, psd=False, j0=0):
    """Formula oracle O1: exact-sum GLS power on fmin + (j0 + j) df, j < nf."""
    lib = load()
    t, y = _f64(t), _f64(y)
    w = None if err is None else _f64(np.asarray(err, dtype=np.float64) ** -2.0)
    psd_scale = 0.5 * (w.sum() if w is not None else float(t.size))
    power = np.empty(int(nf))
    rc = lib.orc_gls_exact(t.ctypes.data, y.ctypes.data, None if w is None else w.ctypes.data, t.size,
                           float(fmin), float(df), int(j0), int(nf), int(bool(fit_mean)), int(bool(psd)),
                           float(psd_scale), power.ctypes.data)
    if rc:
        raise RuntimeError(f"orc_gls_exact failed ({rc})")
    return power


def gls_exact_at(t, y, err, fmin, df, jidx, fit_mean=True, psd=False):
    """Formula oracle O1 at the selected grid indices ``jidx``: f = fmin + jidx * df (parallel over indices)."""
    lib = load()
    t, y = _f64(t), _f64(y)
    jidx = np.ascontiguousarray(jidx, dtype=np.int64)
    w = None if err is None else _f64(np.asarray(err, dtype=np.float64) ** -2.0)
    psd_scale = 0.5 * (w.sum() if w is not None else float(t.size))
    power = np.empty(jidx.size)
    rc = lib.orc_gls_exact_at(t.ctypes.data, y.ctypes.data, None if w is None else w.ctypes.data, t.size,
                              float(fmin), float(df), jidx.ctypes.data, jidx.size, int(bool(fit_mean)),
                              int(bool(psd)), float(psd_scale), power.ctypes.data)
    if rc:
        raise RuntimeError(f"orc_gls_exact_at failed ({rc})")
    return power


def gls_exact_freqs(t, y, err, freqs, fit_mean=True, psd=False):
    """Formula oracle O1 at an arbitrary list of frequencies (non-uniform grids)."""
    lib = load()
    t, y, freqs = _f64(t), _f64(y), _f64(freqs)
    w = None if err is None else _f64(np.asarray(err, dtype=np.float64) ** -2.0)
    psd_scale = 0.5 * (w.sum() if w is not None else float(t.size))
    power = np.empty(freqs.size)
    rc = lib.orc_gls_exact_freqs(t.ctypes.data, y.ctypes.data, None if w is None else w.ctypes.data, t.size,
                                 freqs.ctypes.data, freqs.size, int(bool(fit_mean)), int(bool(psd)),
                                 float(psd_scale), power.ctypes.data)
    if rc:
        raise RuntimeError(f"orc_gls_exact_freqs failed ({rc})")
    return power


def pdm(t, x, periods, nb, nc):
    lib = load()
    t, x, periods = _f64(t), _f64(x), _f64(periods)
    theta = np.empty(periods.size)
    rc = lib.orc_pdm(t.ctypes.data, x.ctypes.data, t.size, periods.ctypes.data, periods.size, int(nb), int(nc),
                     theta.ctypes.data)
    if rc:
        raise RuntimeError(f"orc_pdm failed ({rc})")
    return theta


def num_threads():
    return load().orc_num_threads()
