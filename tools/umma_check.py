"""GPU check of the tensor-core GLS kernel (gls_umma.cu) against the C oracle and against the strip kernel, with timing.
usage (under gpurun):  python tools/umma_check.py [small|c2|c4|c5 ...]    -> prints one line per case
The kernel choice is a ctx property read from the environment at ctx creation (PDC_GLS_UMMA=0|1), so the script
creates one ctx per setting."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cport  # noqa: E402
from periodicity_b200 import _ffi  # noqa: E402


CG2 = [None]


def make_ctx(umma, nsplit=None, chunk=None, prof=False, dbg=None):
    os.environ["PDC_GLS_UMMA"] = str(umma)
    for key, val in (("PDC_GLS_UMMA_NSPLIT", nsplit), ("PDC_GLS_UMMA_CHUNK", chunk), ("PDC_GLS_UMMA_PROF", 1 if prof else None),
                     ("PDC_GLS_UMMA_DBG", dbg), ("PDC_GLS_UMMA_CG2", CG2[0])):
        if val:
            os.environ[key] = str(val)
        else:
            os.environ.pop(key, None)
    return _ffi.Context(0)


def prof_summary(ctx):
    pr = ctx.umma_prof()
    if len(pr) == 0:
        return "no profile"
    pr = pr[:max(len(pr) - 2048, 0)]
    setup, loop, flush = pr[:, 1] - pr[:, 0], pr[:, 2] - pr[:, 1], pr[:, 3] - pr[:, 2]
    return (f"jobs {len(pr)}: clocks setup {np.median(setup):.0f} loop median {np.median(loop):.0f} max {loop.max():.0f} "
            f"flush median {np.median(flush):.0f} max {flush.max():.0f}")


def synth(N, T, nf, sigma, seed, weighted=False):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, T, N))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * nf * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + sigma * rng.standard_normal(N)
    err = rng.uniform(0.5, 2.0, N) if weighted else None
    return t, y, err, fmin, df


def errs(p, ref):
    peak = np.nanmax(np.abs(ref))
    e1 = np.nanmax(np.abs(p - ref)) / peak
    big = np.abs(ref) >= 1e-2 * peak
    e2 = np.nanmax(np.abs(p[big] - ref[big]) / np.abs(ref[big])) if big.any() else 0.0
    return e1, e2


def timed(ctx, fn, reps=5):
    fn()
    ctx.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ctx.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, ctx.last_main_kernel_ms()


def case_small(ctx_t, ctx_s):
    for (N, nf, wt, seed) in [(3000, 1600, False, 1), (3000, 1600, True, 2), (5000, 20000, False, 3), (777, 300, False, 4),
                              (20000, 10000, True, 5), (4097, 8321, False, 6)]:
        t, y, err, fmin, df = synth(N, 100.0, nf, 1.0, seed, wt)
        w = None if err is None else err ** -2.0
        ref = cport.gls_exact(t, y, err, fmin, df, nf)
        out = {}
        for name, ctx in (("umma", ctx_t), ("strip", ctx_s)):
            p, am, mx = ctx.gls(t, y, w, fmin, df, nf)
            e1, e2 = errs(p, ref)
            out[name] = (e1, e2, int(am) == int(np.nanargmax(ref)), int(np.isnan(p).sum()))
        print(f"small N={N} nf={nf} weighted={wt}: umma err {out['umma'][0]:.2e}/{out['umma'][1]:.2e} argmax_ok={out['umma'][2]} nan={out['umma'][3]}"
              f" | strip err {out['strip'][0]:.2e}/{out['strip'][1]:.2e}", flush=True)


def case_big(ctx_t, ctx_s, N, nf, T, tag, nsplits=(None,), chunks=(None,)):
    t, y, w, fmin, df = synth(N, T, nf, 1.0, 11)   # unweighted: w is None
    idx = np.unique(np.concatenate([np.arange(0, nf, max(1, nf // 4096)), np.arange(int(0.3137 * nf) - 200, int(0.3137 * nf) + 200)]))
    idx = idx[(idx >= 0) & (idx < nf)]
    ref = cport.gls_exact_at(t, y, w, fmin, df, idx)
    ps, ams, _ = ctx_s.gls(t, y, w, fmin, df, nf)
    ms_s, k_s = timed(ctx_s, lambda: ctx_s.gls(t, y, w, fmin, df, nf), 3)
    es = errs(ps[idx], ref)
    print(f"{tag} strip: e2e {ms_s:.3f} ms kernel {k_s:.3f} ms err {es[0]:.2e}/{es[1]:.2e}", flush=True)
    for ns, cs in [(a, b) for a in nsplits for b in chunks]:
        ctx = make_ctx(1, ns, cs, prof=True)
        pt, amt, _ = ctx.gls(t, y, w, fmin, df, nf)
        ms_t, k_t = timed(ctx, lambda: ctx.gls(t, y, w, fmin, df, nf), 3)
        et = errs(pt[idx], ref)
        d = errs(pt, ps)
        print(f"{tag} umma   {prof_summary(ctx)}")
        print(f"{tag} umma nsplit={ns} chunk={cs}: e2e {ms_t:.3f} ms kernel {k_t:.3f} ms ({N * nf / k_t * 1e-9:.2f}e12 evals/s) err {et[0]:.2e}/{et[1]:.2e} "
              f"vs strip {d[0]:.2e}/{d[1]:.2e} argmax {int(amt)} vs {int(ams)} nan={int(np.isnan(pt).sum())}", flush=True)


def case_c4(ctx_t, ctx_s, B=256):
    rng = np.random.default_rng(5)
    N, nf = 20000, 10000
    ts, ys, offs = [], [], [0]
    fmins, dfs = [], []
    for b in range(B):
        t = np.sort(rng.uniform(0, 27.4, N))
        df = 1 / (t[-1] - t[0]) / 5
        y = np.sin(2 * np.pi * (3.0 + 0.01 * b) * t) + rng.standard_normal(N)
        ts.append(t); ys.append(y); offs.append(offs[-1] + N); fmins.append(0.5 * df); dfs.append(df)
    t, y = np.concatenate(ts), np.concatenate(ys)
    offs = np.array(offs, dtype=np.int64)
    fmins, dfs = np.array(fmins), np.array(dfs)
    res = {}
    for name, ctx in (("umma", ctx_t), ("strip", ctx_s)):
        p, am, mx = ctx.gls_batch(t, y, None, offs, fmins, dfs, nf)
        ms, _ = timed(ctx, lambda: ctx.gls_batch(t, y, None, offs, fmins, dfs, nf), 2)
        k0, _ = ctx.main_kernel_ms_total()
        ctx.gls_batch(t, y, None, offs, fmins, dfs, nf)
        k = ctx.main_kernel_ms_total()[0] - k0          # all dominant-kernel launches of one call
        res[name] = (p, am, ms, k)
    ref0 = cport.gls_exact(ts[0], ys[0], None, fmins[0], dfs[0], nf)
    refl = cport.gls_exact(ts[-1], ys[-1], None, fmins[-1], dfs[-1], nf)
    pu = res["umma"][0].reshape(B, nf)
    e0, el = errs(pu[0], ref0), errs(pu[-1], refl)
    d = errs(res["umma"][0], res["strip"][0])
    print(f"c4 B={B}: umma e2e {res['umma'][2]:.2f} ms kernel {res['umma'][3]:.2f} ms ({B * N * nf / res['umma'][3] * 1e-9:.2f}e12/s) | strip kernel {res['strip'][3]:.2f} ms"
          f" | err curve0 {e0[0]:.2e}/{e0[1]:.2e} last {el[0]:.2e}/{el[1]:.2e} vs strip {d[0]:.2e} argmax_equal={bool((res['umma'][1] == res['strip'][1]).all())}", flush=True)


def case_dbg(N=65000, nf=100000, T=1470.0):
    """Timing decomposition of the tensor-core kernel (results are wrong by construction)."""
    t, y, w, fmin, df = synth(N, T, nf, 1.0, 11)
    for dbg in (None, 1, 2, 4, 32, 64, 33, 65, 5):
        ctx = make_ctx(1, prof=True, dbg=dbg)
        ms, k = timed(ctx, lambda: ctx.gls(t, y, w, fmin, df, nf), 3)
        print(f"dbg={dbg}: kernel {k:.3f} ms   {prof_summary(ctx)}", flush=True)


def case_trace(N=65000, nf=100000, T=1470.0):
    t, y, w, fmin, df = synth(N, T, nf, 1.0, 11)
    for dbg in (16,):
        ctx = make_ctx(1, prof=True, dbg=dbg)
        ctx.gls(t, y, w, fmin, df, nf)
        ctx.gls(t, y, w, fmin, df, nf)
        tr = ctx.umma_trace()
        base = tr[0, 0]
        print(f"--- trace dbg={dbg}: columns = pair, worker3 [after empty wait, after full arrive], issuer [start, after tempty, after full, after issue], "
              f"worker3 drain [begin, end]; clocks relative to pair 0")
        for p in list(range(40, 56)):
            r = tr[p]
            print(p, " ".join(f"{(v - base) if v else 0:8d}" for v in r))
        d = np.diff(tr[20:280, 0])
        print("period per pair (worker 3 after-empty stamps): median", np.median(d), "mean", d.mean())


if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "c2"]
    if "cg2" in which:
        CG2[0] = 1
    ctx_t, ctx_s = make_ctx(1), make_ctx(0)
    if "small" in which:
        case_small(ctx_t, ctx_s)
    if "c2" in which:
        case_big(ctx_t, ctx_s, 65000, 100000, 1470.0, "c2", nsplits=(None,))
    if "c2sweep" in which:
        case_big(ctx_t, ctx_s, 65000, 100000, 1470.0, "c2", nsplits=(7, 14, 22, 29, 37))
    if "c2chunk" in which:
        case_big(ctx_t, ctx_s, 65000, 100000, 1470.0, "c2", chunks=(None,))
    if "trace" in which:
        case_trace()
    if "dbg" in which:
        case_dbg()
    if "c4" in which:
        case_c4(ctx_t, ctx_s)
    if "c5" in which:
        case_big(ctx_t, ctx_s, 1000000, 1000000, 1470.0, "c5/10")
