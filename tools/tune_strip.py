"""Dev aid (GPU box): time the GLS strip kernel of several builds of the library on the C2 shape.

usage: python tools/tune_strip.py lib1.so lib2.so ...      (each timed in its own process)
       python tools/tune_strip.py --one lib.so             (worker)
Env PDC_GLS_GEOM / PDC_GLS_THREE_TERM are honoured by the library as usual.
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(lib, weighted):
    import numpy as np
    from periodicity_b200 import _ffi
    _ffi.LIB_PATH = os.path.abspath(lib)
    import bench
    wl = bench.make_gls_c2(100_000)
    ctx = _ffi.Context(0)
    t, y = wl["t"], wl["y"]
    w = None
    if weighted:
        w = np.random.default_rng(0).uniform(0.5, 1.5, t.size) ** -2.0
    for _ in range(3):
        ctx.gls(t, y, w, wl["fmin"], wl["df"], wl["nf"])
    k0, c0 = ctx.main_kernel_ms_total()
    for _ in range(10):
        p, am, mx = ctx.gls(t, y, w, wl["fmin"], wl["df"], wl["nf"])
    k1, c1 = ctx.main_kernel_ms_total()
    ms = (k1 - k0) / (c1 - c0)
    print(json.dumps({"lib": os.path.basename(lib), "weighted": weighted, "kernel_ms": round(ms, 4),
                      "evals_per_s": t.size * wl["nf"] / ms * 1e3, "argmax": int(am), "max": float(mx)}))


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        one(sys.argv[2], len(sys.argv) > 3 and sys.argv[3] == "w")
    else:
        for lib in sys.argv[1:]:
            for wflag in ([[], ["w"]] if os.environ.get("TUNE_W", "1") == "1" else [[]]):
                r = subprocess.run([sys.executable, __file__, "--one", lib] + wflag, capture_output=True, text=True)
                sys.stdout.write(r.stdout if r.returncode == 0 else f"{lib}: FAILED {r.stderr[-300:]}\n")
                sys.stdout.flush()
