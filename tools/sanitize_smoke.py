"""Small calls through every entry point, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from periodicity_b200 import _ffi  # noqa: E402

ctx = _ffi.Context(0)
rng = np.random.default_rng(0)
n = 6000
t = np.sort(rng.uniform(0, 100, n))
y = np.sin(t) + rng.standard_normal(n)
w = rng.uniform(0.5, 2, n)
df = 1 / (t[-1] - t[0]) / 5
for ww in (None, w):
    ctx.gls(t, y, ww, 0.5 * df, df, 5000)                       # three-term strip
    ctx.gls(t, y, ww, 0.5 * df * 3, df * 3, 3000)               # rotation strip
    ctx.gls_multi(t, np.stack([y, y[::-1], 2 * y]), ww, 0.5 * df, df, 3000)
off = np.array([0, 1000, 2500, 6000])
ctx.gls_batch(t, y, None, off, np.full(3, 0.5 * df), np.full(3, df), 700)
P = np.linspace(1, 9, 700)
ctx.pdm(t, y, P, 10, 2)                                         # packed ATOMS path (n >= 4096)
ctx.pdm(t[:3000], y[:3000], P, 5, 2)                            # float2 path
tb = t.copy(); tb[5] = np.nan
ctx.pdm(tb, y, P, 10, 2)                                        # guarded path
ctx.pdm(t * 1e9, y, P, 10, 2)                                   # exact (large |t/P|) path
ctx.aov(t, y, P, 10)                                            # AoV epilogue on the same histograms
ctx.ce(t, y, P, 10, 5)                                          # conditional entropy: count histograms
ctx.ce(tb, y, P, 10, 5)                                         # ... exact / guarded path
m = (y - y.max()) / (2 * (y.max() - y.min())) + 0.25
ctx.stringlength(t, m, P)                                       # shared-memory sort
tt = np.sort(rng.uniform(0, 100, 20000)); mm = rng.uniform(-0.25, 0.25, 20000)
ctx.stringlength(tt, mm, P[:8])                                 # global-scratch sort
pw, _, _ = ctx.gls(t, y, None, 0.5 * df, df, 5000)
idx, val = ctx.peaks_topk(pw, 5)
ctx.peaks_halfmax(pw, idx)
# tensor-core formulations of the GLS sums (forced on: the problems here are below the automatic threshold)
for env in ({"PDC_GLS_UMMA": "1", "PDC_GLS_UMMA_CG2": "0", "PDC_GLS_UMMA_FINE": "0"},      # one CTA per tile, fine operand in the kernel
            {"PDC_GLS_UMMA": "1", "PDC_GLS_UMMA_CG2": "0", "PDC_GLS_UMMA_FINE": "1"},      # ... precomputed, bulk copies
            {"PDC_GLS_UMMA": "1", "PDC_GLS_UMMA_CG2": "1"}):                               # pairs of CTAs (cta_group::2)
    os.environ.update(env)
    cu = _ffi.Context(0)
    for ww in (None, w):
        cu.gls(t, y, ww, 0.5 * df, df, 5000)
    cu.gls_batch(t, y, None, off, np.full(3, 0.5 * df), np.full(3, df), 700)
    for k in env:
        os.environ.pop(k)
os.environ["PDC_BATCH_PIPE_BYTES"] = "1"                         # read at ctx creation: upload batches in pipelined runs
ctx2 = _ffi.Context(0)
ctx2.gls_batch(t, y, w, np.arange(17) * 375, np.full(16, 0.5 * df), np.full(16, df), 700)
os.environ["PDC_MULTI_MIN_EVALS"] = "1"                          # read at ctx creation: shard even these small problems
mctx = _ffi.Context([0, 0, 0])                                   # multi-device ctx: three workers on one GPU
mctx.gls(t, y, w, 0.5 * df, df, 5000)
mctx.pdm(t, y, P, 10, 2)
mctx.ce(t, y, P, 10, 5)
mctx.gls_batch(t, y, None, off, np.full(3, 0.5 * df), np.full(3, df), 700)
mctx.close()
print("sanitize smoke done")
