"""Load the UNMODIFIED reference hot-path files -- authoring container only.

TEST / BENCH INFRASTRUCTURE (see ``oracle/__init__.py``).  ``/root/reference`` does not
exist on the GPU box; there the two files are found under ``oracle/_ref/periodicity/`` if
``oracle/make_ref.py`` staged them (git-ignored, byte-for-byte copies that travel with the
snapshot).  Used by ``tests/golden/make_golden.py`` (which writes the committed fixtures), by
CPU tests that skip themselves when neither location exists, and by ``bench.py``'s reference
arm / ``cpu_baseline`` leg (``kind: "reference"``).

The reference package cannot be imported as a whole (``core.py:6`` needs
xarray, which is not installed).  ``spectral.py`` and ``phase.py`` only need
numpy plus ``TSeries`` / ``FSeries`` from ``.core`` (``spectral.py:1-5``,
``phase.py:1-5``), so they are executed as-is with ``periodicity.core`` seeded
by the numpy-only stand-in below, which carries exactly the attributes the two
files touch: ``time, values, size, baseline, median_dt, __len__, copy`` (+ ``fold, max, min`` and scalar
arithmetic for ``StringLength``, see the note in the class)
(``core.py:60-66,94-99,144-145,460-511``) and ``frequency, values``
(``core.py:859-889``).
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PERIODICITY_REFERENCE_ROOT", "/root/reference")
# byte-for-byte staged copies of the two files (oracle/make_ref.py; git-ignored, travels to the GPU box)
STAGED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "periodicity")


def source_dir():
    """Directory holding the reference's spectral.py / phase.py: the reference tree when it exists (authoring
    container), else the staged copies under oracle/_ref (GPU box), else None."""
    for d in (os.path.join(REFERENCE_ROOT, "src", "periodicity"), STAGED_DIR):
        if os.path.isfile(os.path.join(d, "spectral.py")) and os.path.isfile(os.path.join(d, "phase.py")):
            return d
    return None


class _StubTSeries:
    def __init__(self, time=None, values=None, assume_sorted=False):
        if time is None:
            time = np.arange(len(values))
        if values is None:
            values = np.ones(len(time))
        time = np.asarray(time)
        values = np.asarray(values)
        if time.size != values.size:
            raise ValueError("Input arrays have incompatible lengths.")
        if not assume_sorted and np.any(np.diff(time) < 0):
            order = np.argsort(time, kind="stable")
            time, values = time[order], values[order]
        self.time = time
        self.values = values

    @property
    def size(self):
        return self.values.size

    def __len__(self):
        return self.values.size

    @property
    def baseline(self):
        return self.time[-1] - self.time[0]

    @property
    def median_dt(self):
        return np.median(np.diff(self.time))

    def copy(self):
        return _StubTSeries(self.time, self.values.copy(), assume_sorted=True)

    def __mul__(self, k):
        return _StubTSeries(self.time, self.values * k, assume_sorted=True)

    __rmul__ = __mul__

    def __add__(self, k):
        return _StubTSeries(self.time, self.values + k, assume_sorted=True)

    __radd__ = __add__

    # --- what StringLength touches (phase.py:45-51,64-66) -------------------------------------
    # core.py:543-544; the constructor re-sorts by phase (core.py:473-477)
    def fold(self, period, t0=0):
        return _StubTSeries(((self.time - t0) / period) % 1, self.values)

    # ASSUMPTION: scalar extrema.  The reference's Signal.max() returns a one-sample series
    # (core.py:217-220) that `signal - signal.max()` cannot broadcast (core.py:175-178); scalars are
    # what phase.py:64-65 needs to mean anything.
    def max(self):
        return np.nanmax(self.values)

    def min(self):
        return np.nanmin(self.values)

    def __sub__(self, k):
        return _StubTSeries(self.time, self.values - k, assume_sorted=True)

    def __truediv__(self, k):
        return _StubTSeries(self.time, self.values / k, assume_sorted=True)


class _StubFSeries:
    def __init__(self, frequency=None, values=None, assume_sorted=False):
        frequency = np.asarray(frequency)
        values = np.asarray(values)
        if not assume_sorted and np.any(np.diff(frequency) < 0):
            order = np.argsort(frequency, kind="stable")
            frequency, values = frequency[order], values[order]
        self.frequency = frequency
        self.values = values

    def amax(self):
        return np.nanmax(self.values)


def available():
    return source_dir() is not None


def load():
    """Return ``(spectral, phase)`` modules of the unmodified reference."""
    src = source_dir()
    if src is None:
        raise RuntimeError(f"reference files found neither under {REFERENCE_ROOT} nor under {STAGED_DIR}")
    pkg_name = "_periodicity_reference"
    if pkg_name + ".spectral" in sys.modules:
        return sys.modules[pkg_name + ".spectral"], sys.modules[pkg_name + ".phase"]
    pkg = types.ModuleType(pkg_name)
    pkg.__path__ = []
    core = types.ModuleType(pkg_name + ".core")
    core.TSeries = _StubTSeries
    core.FSeries = _StubFSeries
    sys.modules[pkg_name] = pkg
    sys.modules[pkg_name + ".core"] = core
    mods = []
    for name in ("spectral", "phase"):
        path = os.path.join(src, name + ".py")
        spec = importlib.util.spec_from_file_location(f"{pkg_name}.{name}", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkg_name}.{name}"] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


TSeries = _StubTSeries
FSeries = _StubFSeries
