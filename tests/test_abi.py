"""The C-ABI shared library loads and exports every symbol include/periodicity_b200.h declares.

No compute calls: this runs on the CPU-only box.  Also checks that the product
path fails loudly (no CPU fallback) when no CUDA device is visible.
"""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from periodicity_b200 import _ffi

HEADER = os.path.join(ROOT, "include", "periodicity_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"PDC_API\s+[\w\s\*]+?\b(pdc_\w+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.isfile(_ffi.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _ffi.SIGNATURES, f"{name} has no ctypes prototype in _ffi.SIGNATURES"
    for name in _ffi.SIGNATURES:
        assert name in declared, f"{name} bound in _ffi but not declared in the header"


def test_version_and_error_string():
    lib = _ffi.load_library()
    assert lib.pdc_version() == 200
    assert isinstance(lib.pdc_last_error(), bytes)


def test_ctx_create_rejects_null_and_bad_device():
    lib = _ffi.load_library()
    assert lib.pdc_ctx_create(None, 0) == _ffi.PDC_EINVAL
    h = ctypes.c_void_p()
    rc = lib.pdc_ctx_create(ctypes.byref(h), 10_000)
    assert rc in (_ffi.PDC_EINVAL, _ffi.PDC_ENODEVICE) and not h
    assert lib.pdc_ctx_destroy(None) == _ffi.PDC_OK
    assert lib.pdc_ctx_sm_count(None) == -1


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    with pytest.raises(RuntimeError, match="no usable CUDA device|no CPU fallback"):
        _ffi.Context(0)
    from periodicity_b200 import GLS
    import numpy as np
    with pytest.raises(RuntimeError):
        GLS()(np.sin(np.arange(50.0)))
