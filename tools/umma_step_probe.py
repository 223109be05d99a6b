"""Where does a C2 step go?  Back-to-back device-resident calls (warm L2) against the library's own kernel timer.
usage (under gpurun): python tools/umma_step_probe.py"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from periodicity_b200 import _ffi, dist as pdist
import bench

wl = bench.make_gls_c2(100_000)
dev = torch.device("cuda:0")
t_d = torch.as_tensor(wl["t"], device=dev); y_d = torch.as_tensor(wl["y"], device=dev)
for umma, fine in ((1, None), (1, 0), (0, None)):
    os.environ["PDC_GLS_UMMA"] = str(umma)
    if fine is None: os.environ.pop("PDC_GLS_UMMA_FINE", None)
    else: os.environ["PDC_GLS_UMMA_FINE"] = str(fine)
    ctx = _ffi.Context(0)
    for _ in range(5):
        pdist.gls_torch(t_d, y_d, None, wl["fmin"], wl["df"], wl["nf"], ctx=ctx)
    torch.cuda.synchronize()
    k0, c0 = ctx.main_kernel_ms_total()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        pdist.gls_torch(t_d, y_d, None, wl["fmin"], wl["df"], wl["nf"], ctx=ctx)
    e1.record(); t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    k1, c1 = ctx.main_kernel_ms_total()
    print(f"umma={umma} fine={fine}: {e0.elapsed_time(e1) / n:.4f} ms per call back to back (host enqueue {t_enq / n * 1e3:.4f} ms), "
          f"main kernel {(k1 - k0) / (c1 - c0):.4f} ms, path {ctx.last_gls_path()}", flush=True)
