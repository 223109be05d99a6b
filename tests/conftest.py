import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes tens of seconds (full-size BASELINE configs); still part of -m gpu")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices in one process / box (skips itself otherwise)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        out = {k: z[k] for k in z.files}
    for k, v in list(out.items()):
        if v.ndim == 0:
            out[k] = v.item()
    return out


def opt(v):
    """Golden files store None as NaN (scalars) or an empty array."""
    if isinstance(v, np.ndarray):
        return None if v.size == 0 else v
    if isinstance(v, float) and np.isnan(v):
        return None
    return v


GLS_CASES = ["gls_sine100", "gls_c1_small", "gls_err", "gls_nofitmean", "gls_psd", "gls_jd_default",
             "gls_spotted_star"]
PDM_CASES = ["pdm_basic", "pdm_negative_t_nc3", "pdm_sparse", "pdm_integer_t_ties", "pdm_subharmonic",
             "pdm_nc1", "pdm_defaults"]
PDM_KW = ["nb", "nc", "p_min", "p_max", "n_periods", "do_subharmonic"]
SL_CASES = ["sl_basic", "sl_sparse_negative_t", "sl_integer_t_ties", "sl_3000"]


@pytest.fixture(scope="session")
def gpu_ctx():
    from periodicity_b200 import _ffi
    return _ffi.default_context(0)
