"""Host-to-host timing of the multi-device ctx (pdc_ctx_create_multi) in ONE ordinary Python process: the same C-ABI
call on 1 device and on all visible devices, for the BASELINE shapes that fit the run.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from periodicity_b200 import _ffi  # noqa: E402
import torch  # noqa: E402

ndev = torch.cuda.device_count()
one = _ffi.Context(0)
many = _ffi.Context(list(range(ndev))) if ndev > 1 else None


def timed(fn, reps):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps * 1e3, out


res = {"devices": ndev}
c2 = bench.make_gls_c2(100_000)
c3 = bench.make_pdm_c3(100_000)
c5 = bench.make_gls_c5(10_000_000 if "--full" in sys.argv else 1_000_000)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()   # noqa: E731
for name, wl, reps in (("C2_gls", c2, 20), ("C3_pdm", c3, 20), ("C5_gls" if "--full" in sys.argv else "C5_gls_tenth", c5, 2)):
    t, y = pin(wl["t"]), pin(wl["y"])
    rec = {}
    for tag, ctx in (("1", one), (str(ndev), many)):
        if ctx is None:
            continue
        if wl["kind"] == "gls":
            ms, out = timed(lambda: ctx.gls(t, y, None, wl["fmin"], wl["df"], wl["nf"]), reps)
        else:
            ms, out = timed(lambda: ctx.pdm(t, y, wl["periods"], wl["nb"], wl["nc"]), reps)
        rec[tag] = {"ms": ms, "evals_per_s": t.size * wl["nf"] / (ms * 1e-3), "arg": int(out[1])}
    if many is not None:
        rec["speedup"] = rec["1"]["ms"] / rec[str(ndev)]["ms"]
        rec["same_peak"] = rec["1"]["arg"] == rec[str(ndev)]["arg"]
    res[name] = rec
print(json.dumps(res))
