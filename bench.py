#!/usr/bin/env python
"""bench.py -- sample*frequency evaluations per second of the GLS / PDM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic input:
  gls_c2 (default)  GLS, Kepler-like 65,000 points x 1e5 frequencies per GPU
                    (BASELINE.json configs[1]).  With N GPUs the frequency grid is
                    N x 1e5 long and sharded across ranks (weak scaling), followed by
                    ONE all-gather of [power shard, local max, local argmax].
  pdm_c3            PDM 1e5 points x 1e5 trial periods, nb=10, nc=2 (configs[2]);
                    period grid sharded the same way.
  gls_c5            GLS 1e6 points x (1e7/8 per GPU) frequencies (configs[4] per-GPU share).
  gls_c4            batched GLS, 256 TESS-like curves x 20,000 points x 1e4 frequencies per GPU.
  gls_c4_full       the whole configs[3] survey (1e4 curves) on every GPU -- one-GPU record of the full size.

Printed line (rank 0): metric/value/unit/... as the driver contract asks, plus
  roofline      dominant kernel vs the FP32 issue roofline (the path is FP32-pipe bound, not
                HBM or tensor bound): achieved = evals x 20 FLOP / kernel time
                (CUDA events on the launching stream inside the library), peak = measured FFMA rate
                (profiles/pipes_r01.json; MEASURED_PEAKS.json has no FP32 entry).
  cpu_baseline  the oracle port of the reference algorithm timed on this box's host cores.
  e2e           same metric through the host-pointer C-ABI call (H2D + kernels + D2H in the timed region).

--impl reference times the reference's own CPU algorithm (numpy restatement in oracle/, the
reference is pure Python and cannot be installed without xarray) on the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_EVAL_GLS = 20.0      # 12 FP32 instructions = 20 FLOP per sample*frequency
OPS_PER_EVAL_PDM = 8.0        # 8 ops per sample*period
FP32_PEAK_TFLOPS_MEASURED = 72.3   # profiles/pipes_r01.json: 36,172 GFFMA/s x 2
FP32_PEAK_GINSTR_MEASURED = 36172.0  # same measurement as thread-instructions/s (125 per clk per SM)
# what gls_strip_kernel<16,128> executes in its three-term form (SASS hot loop: 282 instructions per
# 2 samples x 16 frequencies; 6 FFMA + 2 FADD per evaluation): the 20 FLOP of the accounting figure are NOT all executed
GLS_EXECUTED_INSTR_PER_EVAL = 282.0 / 32.0
GLS_EXECUTED_FLOP_PER_EVAL = 14.0
PDM_PEAK_GEVALS_MEASURED = 3841.3  # profiles/pipes_r01.json smem_private_u32_atoms: private-column ATOMS.ADD, updates/s


def measured_hbm_gbs():
    """HBM copy bandwidth the driver measured on this pool (MEASURED_PEAKS.json); 6650 = the profiling
    recipe's stated fallback if the file is absent."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_*.json); None if no capture exists."""
    path = os.path.join(ROOT, "profiles", name)
    try:
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
# synthetic workloads (SURVEY.md section 8d recipes, fixed seeds)
# ------------------------------------------------------------------------------------------
def make_gls_c2(nf_total):
    rng = np.random.default_rng(2)
    n = 65_000
    slots = np.sort(rng.choice(71_940, n, replace=False))
    t = slots * (29.4244 / 1440.0) + rng.uniform(0, 1 / 1440.0, n)
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * 100_000 * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + rng.standard_normal(n)
    return dict(kind="gls", t=t, y=y, fmin=fmin, df=df, nf=nf_total,
                name=f"GLS Kepler-like 65,000 points x {nf_total} frequencies (C2 per GPU)")


def make_gls_c1(nf_total):
    rng = np.random.default_rng(1)
    n = 1000
    t = np.sort(rng.uniform(0, 100.0, n))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    fsig = fmin + 0.3137 * 10_000 * df
    y = 1000 + np.sin(2 * np.pi * fsig * t + 0.3) + 0.5 * rng.standard_normal(n)
    return dict(kind="gls", t=t, y=y, fmin=fmin, df=df, nf=nf_total,
                name=f"GLS 1,000 points x {nf_total} frequencies (C1)")


def make_gls_multi(series):
    """C4-shaped survey sector whose light curves SHARE one time axis (pdc_gls_multi)."""
    n, nf = 20_000, 10_000
    rng = np.random.default_rng(4000)
    keep = rng.uniform(size=n + n // 50) > 0.01
    t = (np.arange(n + n // 50)[keep][:n]) * (2.0 / 1440.0) + rng.uniform(0, 0.2 / 1440.0, n)
    P = rng.uniform(0.5, 10.0, series)
    Y = 1000 + np.sin(2 * np.pi * t[None, :] / P[:, None]) + rng.standard_normal((series, n))
    df = 1 / (t[-1] - t[0]) / 5
    return dict(kind="gls_multi", t=t, y=Y, fmin=0.5 * df, df=df, nf=nf, S=series,
                name=f"GLS {series} series on shared timestamps x 20,000 points x 1e4 frequencies (C4 shape)")


def make_gls_c5(nf_total):
    rng = np.random.default_rng(5)
    n = 1_000_000
    t = np.sort(rng.uniform(0, 1000.0, n))
    df = 1 / (t[-1] - t[0]) / 5
    fmin = 0.5 * df
    y = 1000 + np.sin(2 * np.pi * 17.123 * t + 0.3) + rng.standard_normal(n)
    return dict(kind="gls", t=t, y=y, fmin=fmin, df=df, nf=nf_total,
                name=f"GLS 1e6 points x {nf_total} frequencies (C5 share)")


def make_pdm_c3(np_total):
    rng = np.random.default_rng(3)
    n = 100_000
    t = np.sort(rng.uniform(0, 1000.0, n))
    x = 1000 + np.sin(2 * np.pi * t / 3.7) + 0.8 * np.sin(4 * np.pi * t / 3.7) + rng.standard_normal(n)
    periods = np.linspace(1.0, 11.0, np_total)
    return dict(kind="pdm", t=t, y=x, periods=periods, nb=10, nc=2, nf=np_total,
                name=f"PDM 1e5 points x {np_total} trial periods, nb=10 nc=2 (C3)")


def make_sl(np_total):
    """String Length on a sparse light curve (the method's use case; SURVEY 8f row 2 has no BASELINE config)."""
    from oracle import stringlength_numpy
    rng = np.random.default_rng(7)
    n = 2000
    t = np.sort(rng.uniform(0, 1000.0, n))
    x = 12.0 + 0.4 * np.sin(2 * np.pi * t / 4.3) + 0.05 * rng.standard_normal(n)
    periods = stringlength_numpy.period_grid(t, dphi=0.1, n_periods=np_total)
    return dict(kind="sl", t=t, y=stringlength_numpy.scale(x), periods=periods, nf=np_total,
                name=f"String Length 2,000 points x {np_total} trial periods (phase.py:18-72)")


def make_gls_c4(curves):
    n, nf = 20_000, 10_000
    ts, ys, fm, dfs = [], [], [], []
    for b in range(curves):
        rng = np.random.default_rng(4000 + b)
        keep = rng.uniform(size=n + n // 50) > 0.01
        tt = (np.arange(n + n // 50)[keep][:n]) * (2.0 / 1440.0) + rng.uniform(0, 0.2 / 1440.0, n)
        P = rng.uniform(0.5, 10.0)
        ys.append(1000 + np.sin(2 * np.pi * tt / P) + rng.standard_normal(n))
        ts.append(tt)
        d = 1 / (tt[-1] - tt[0]) / 5
        dfs.append(d)
        fm.append(0.5 * d)
    return dict(kind="gls_batch", t=np.concatenate(ts), y=np.concatenate(ys),
                offsets=np.arange(curves + 1, dtype=np.int64) * n, fmin=np.array(fm), df=np.array(dfs), nf=nf,
                name=f"batched GLS {curves} TESS-like curves x 20,000 points x 1e4 frequencies (C4 share)")


# ------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def _run(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline (oracle port of the reference's CPU algorithm)
# ------------------------------------------------------------------------------------------
_POOL = None


def _pool(cores):
    """One worker pool for the whole run (its start-up is not part of any timed step)."""
    global _POOL
    if _POOL is None:
        import atexit
        from multiprocessing import Pool
        _POOL = Pool(cores)
        atexit.register(_POOL.terminate)
    return _POOL


_BATCH_MODE = {}


def _cpu_batch(jobs, key):
    """Run the per-curve reference jobs either in this process or over all cores, whichever a one-off
    calibration on 8 jobs found faster on this host (process pools do not always scale on shared vCPUs).
    Returns the number of cores used."""
    cores = os.cpu_count() or 1
    if key not in _BATCH_MODE:
        probe = jobs[: min(8, len(jobs))]
        t0 = time.perf_counter()
        for j in probe:
            _cpu_gls_one(j)
        t_serial = time.perf_counter() - t0
        _pool(cores).map(_cpu_gls_one, probe)          # warm the workers (imports)
        t0 = time.perf_counter()
        _pool(cores).map(_cpu_gls_one, probe)
        t_pool = time.perf_counter() - t0
        _BATCH_MODE[key] = cores if t_pool < t_serial else 1
    if _BATCH_MODE[key] == 1:
        for j in jobs:
            _cpu_gls_one(j)
        return 1
    _pool(cores).map(_cpu_gls_one, jobs)
    return cores


def _cpu_gls_one(job):
    from oracle import gls_numpy
    t, y, fmin, df, nf = job
    return float(np.nanmax(gls_numpy.gls_power(t, y, None, fmin, df, nf, True, False)))


def _cpu_sl_chunk(t, m, periods):
    from oracle import stringlength_numpy
    return stringlength_numpy.string_lengths(t, m, periods)


def cpu_reference_step(wl):
    """One step of the reference's CPU path on (a bounded sample of) the workload.
    Returns (evals processed, cores used, description of the sample)."""
    from oracle import gls_numpy, pdm_numpy
    if wl["kind"] == "gls":
        # the reference's own algorithm: FFT extirpolation, single-threaded numpy (spectral.py:11-40)
        gls_numpy.gls_power(wl["t"], wl["y"], None, wl["fmin"], wl["df"], wl["nf"], True, False)
        return wl["t"].size * wl["nf"], 1, "full workload, reference FFT-extirpolation algorithm, numpy, 1 thread"
    if wl["kind"] == "gls_multi":
        cores = os.cpu_count() or 1
        B = min(wl["S"], 8 * cores)
        jobs = [(wl["t"], wl["y"][b], wl["fmin"], wl["df"], wl["nf"]) for b in range(B)]
        used = _cpu_batch(jobs, "multi")
        return wl["t"].size * B * wl["nf"], used, (f"first {B} series, python loop over the reference algorithm on "
                                                   f"{used} core(s) (faster of in-process / Pool({cores}))")
    if wl["kind"] == "gls_batch":
        # the reference has no batch API: a survey is a loop over curves; mapped over all host cores here
        cores = os.cpu_count() or 1
        B = min(len(wl["offsets"]) - 1, 8 * cores)
        jobs = [(wl["t"][wl["offsets"][b]:wl["offsets"][b + 1]], wl["y"][wl["offsets"][b]:wl["offsets"][b + 1]],
                 wl["fmin"][b], wl["df"][b], wl["nf"]) for b in range(B)]
        used = _cpu_batch(jobs, "batch")
        return int(wl["offsets"][B]) * wl["nf"], used, (f"first {B} curves, python loop over the reference algorithm on "
                                                        f"{used} core(s) (faster of in-process / Pool({cores}); the "
                                                        "reference has no batch API)")
    cores = os.cpu_count() or 1
    if wl["kind"] == "sl":
        sample = wl["periods"][:: max(1, wl["periods"].size // (512 * cores))][: 512 * cores]
        _pool(cores).starmap(_cpu_sl_chunk, [(wl["t"], wl["y"], c) for c in np.array_split(sample, cores)])
        return wl["t"].size * sample.size, cores, (f"{sample.size} of {wl['periods'].size} trial periods (strided), "
                                                   f"multiprocessing.Pool({cores}) as phase.py:68-70")
    sample = wl["periods"][:: max(1, wl["periods"].size // (24 * cores))][: 24 * cores]
    pdm_numpy.pdm_pool(wl["t"], wl["y"], sample, wl["nb"], wl["nc"], cores, sort=True)
    return wl["t"].size * sample.size, cores, (f"{sample.size} of {wl['periods'].size} trial periods (strided), "
                                               f"multiprocessing.Pool({cores}) as phase.py:185-186")


def run_reference(args, wl, metric, unit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_reference_step(wl)
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        e, cores, desc = cpu_reference_step(wl)
        evals += e
    dt = time.perf_counter() - t0
    value = evals / dt
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "note": "reference CPU algorithm (oracle port; the reference package "
                   "needs xarray and cannot be installed here)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout. Libraries loaded below (NCCL prints its version banner to
    fd 1 on some boxes) must not add to it: fd 1 is pointed at stderr for the whole run and the result line
    is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gls_c2", choices=["gls_c2", "pdm_c3", "gls_c5", "gls_c5_full", "gls_c4", "gls_c4_full", "gls_c1", "gls_multi", "sl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1, GLS / PDM grids: 'p2p' = all-gather fused into the epilogue kernel over NVLink peer "
                         "memory (pdc_gls_dev_fanout), 'nccl' = one ncclAllGather after the kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    per_gpu = {"gls_c2": 100_000, "pdm_c3": 100_000, "gls_c5": 1_250_000, "gls_c5_full": 10_000_000,
               "gls_c4": 256, "gls_c4_full": 10_000, "gls_c1": 10_000, "gls_multi": 256, "sl": 100_000}[args.workload]
    total_units = per_gpu * max(world, 1)
    if args.workload == "gls_c2":
        wl = make_gls_c2(total_units)
    elif args.workload in ("gls_c5", "gls_c5_full"):
        wl = make_gls_c5(total_units)
    elif args.workload == "gls_c1":
        wl = make_gls_c1(total_units)
    elif args.workload == "gls_multi":
        wl = make_gls_multi(total_units)
    elif args.workload == "pdm_c3":
        wl = make_pdm_c3(total_units)
    elif args.workload == "sl":
        wl = make_sl(total_units)
    else:
        wl = make_gls_c4(total_units)
    metric = {"pdm": "PDM sample*period evaluations per second",
              "sl": "String Length sample*period evaluations per second"}.get(
                  wl["kind"], "GLS sample*frequency evaluations per second")
    unit = "evals/s"

    if args.impl == "reference":
        run_reference(args, wl, metric, unit)
        return

    import torch
    import torch.distributed as dist
    from periodicity_b200 import _ffi
    from periodicity_b200 import dist as pdist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = _ffi.default_context(local_rank)

    # ---- device-resident inputs ------------------------------------------------------------
    n = wl["t"].size
    t_pin = torch.from_numpy(wl["t"]).pin_memory()
    y_pin = torch.from_numpy(np.ascontiguousarray(wl["y"])).pin_memory()
    t_d = t_pin.to(dev)
    y_d = y_pin.to(dev)
    kind = wl["kind"]
    if kind == "gls_multi":
        b0, b1 = pdist.batch_shard_bounds(wl["S"], rank, world)
        units_local = n * wl["nf"] * (b1 - b0)
        evals_total = n * wl["nf"] * wl["S"]
        L = b1 - b0
        ym_d = y_d[b0:b1].contiguous()
        pm_arg = torch.empty(L, dtype=torch.int64, device=dev)
        pm_max = torch.empty(L, dtype=torch.float64, device=dev)
    elif kind == "gls_batch":
        B = len(wl["offsets"]) - 1
        b0, b1 = pdist.batch_shard_bounds(B, rank, world)
        off = wl["offsets"][b0:b1 + 1]
        units_local = int(off[-1] - off[0]) * wl["nf"]
        evals_total = int(wl["offsets"][-1]) * wl["nf"]
        L = b1 - b0
    else:
        start, stop, L = pdist.shard_bounds(wl["nf"], rank, world)
        units_local = n * (stop - start)
        evals_total = n * wl["nf"]
        if kind in ("pdm", "sl"):
            p_d = torch.from_numpy(wl["periods"][start:stop].copy()).to(dev)
            pfull_d = torch.from_numpy(wl["periods"]).to(dev)

    def step_device():
        """One pass of the hot path with inputs resident in HBM; returns (values, best idx, best val)."""
        if kind == "gls":
            if world > 1 and gather_mode[0] == "p2p":
                power, best = pdist.gls_sharded_p2p_torch(t_d, y_d, None, wl["fmin"], wl["df"], wl["nf"], ctx=ctx)
                return power, best[:, 0], best[:, 1]
            power, arg, mx = pdist.gls_torch(t_d, y_d, None, wl["fmin"], wl["df"], stop - start, j0=start, ctx=ctx)
            if world == 1:
                return power, mx, arg          # one GPU: nothing to exchange
            garg = (arg + start).to(torch.float64).reshape(())
            vals, bests, args_ = pdist.all_gather_packed(power, mx.reshape(()), garg, L)
        elif kind == "pdm":
            if world > 1 and gather_mode[0] == "p2p":
                theta, best = pdist.pdm_sharded_p2p_torch(t_d, y_d, pfull_d, wl["nb"], wl["nc"], ctx=ctx)
                return theta, best[:, 0], best[:, 1]
            theta, arg, mn = pdist.pdm_torch(t_d, y_d, p_d, wl["nb"], wl["nc"], ctx=ctx)
            if world == 1:
                return theta, mn, arg
            garg = (arg + start).to(torch.float64).reshape(())
            vals, bests, args_ = pdist.all_gather_packed(theta, mn.reshape(()), garg, L)
        elif kind == "sl":
            ell, arg, mn = pdist.stringlength_torch(t_d, y_d, p_d, ctx=ctx)
            if world == 1:
                return ell, mn, arg
            garg = (arg + start).to(torch.float64).reshape(())
            vals, bests, args_ = pdist.all_gather_packed(ell, mn.reshape(()), garg, L)
        elif kind == "gls_multi":
            ctx._lib.pdc_gls_multi_dev(ctx._h, t_d.data_ptr(), ym_d.data_ptr(), None, n, L, float(wl["fmin"]),
                                       float(wl["df"]), 0, wl["nf"], _ffi.GLS_FIT_MEAN, 1.0, None, pm_arg.data_ptr(),
                                       pm_max.data_ptr(), torch.cuda.current_stream(dev).cuda_stream or None)
            vals, bests, args_ = pdist.all_gather_packed(pm_max, pm_max.max(), pm_arg.to(torch.float64).max(), L)
        else:
            a, e = int(off[0]), int(off[-1])
            _, arg, mx = pdist.gls_batch_torch(t_d[a:e], y_d[a:e], None, off - off[0], wl["fmin"][b0:b1],
                                               wl["df"][b0:b1], wl["nf"], want_power=False, ctx=ctx)
            vals, bests, args_ = pdist.all_gather_packed(mx, mx.max(), arg.to(torch.float64).max(), L)
        return vals, bests, args_

    gather_mode = [args.gather if (world > 1 and kind in ("gls", "pdm")) else "nccl"]
    if gather_mode[0] == "p2p":
        try:                               # symmetric-memory rendezvous is collective: every rank tries, all agree
            step_device()
            torch.cuda.synchronize()
            okflag = torch.ones(1, device=dev)
        except Exception as exc:           # noqa: BLE001 -- transport set-up failed: fall back to the NCCL all-gather
            print(f"[bench] p2p gather unavailable on rank {rank}: {exc}", file=sys.stderr)
            okflag = torch.zeros(1, device=dev)
        dist.all_reduce(okflag, op=dist.ReduceOp.MIN)
        if okflag.item() < 1:
            gather_mode[0] = "nccl"

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ctx.launch_count
    kms0, kcnt0 = ctx.main_kernel_ms_total()
    barrier()
    for k in range(args.steps):
        flush.zero_()                      # evict L2 between timed iterations (not timed)
        ev[k][0].record()
        step_device()
        ev[k][1].record()
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0
    barrier()
    clocks = sampler.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kms1, kcnt1 = ctx.main_kernel_ms_total()
    main_kernel_ms = (kms1 - kms0) / max(1, kcnt1 - kcnt0)   # average launch of the dominant kernel, timed region
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = evals_total / (ms_per_step * 1e-3)

    # ---- e2e: host-pointer C-ABI call, pinned host inputs, H2D + D2H inside the timed region ---
    th, yh = t_pin.numpy(), y_pin.numpy()

    def step_e2e():
        if kind == "gls":
            p, a, m = ctx.gls(th, yh, None, wl["fmin"], wl["df"], stop - start, j0=start)
            return p
        if kind == "pdm":
            p, a, m = ctx.pdm(th, yh, wl["periods"][start:stop], wl["nb"], wl["nc"])
            return p
        if kind == "sl":
            p, a, m = ctx.stringlength(th, yh, wl["periods"][start:stop])
            return p
        if kind == "gls_multi":
            _, a, m = ctx.gls_multi(th, yh[b0:b1], None, wl["fmin"], wl["df"], wl["nf"], want_power=False)
            return m
        a_, e_ = int(off[0]), int(off[-1])
        _, a, m = ctx.gls_batch(th[a_:e_], yh[a_:e_], None, off - off[0], wl["fmin"][b0:b1], wl["df"][b0:b1],
                                wl["nf"], want_power=False)
        return m

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = evals_total * args.steps / e2e_s
    if kind == "gls_multi":
        h2d = 8 * n * (1 + (b1 - b0))
        d2h = out.nbytes * 2
    elif kind == "gls_batch":
        h2d = 2 * 8 * int(off[-1] - off[0])
        d2h = out.nbytes * 2
    else:
        h2d = 2 * 8 * n + (8 * (stop - start) if kind in ("pdm", "sl") else 0)
        d2h = out.nbytes + 16

    # ---- roofline of the dominant kernel ----------------------------------------------------
    if kind == "sl":
        # bitonic network: log2(Np)(log2(Np)+1)/2 stages, each reading and (worst case) writing every 12-byte
        # (key, index) record once = the algorithmic shared-memory traffic of the chosen algorithm
        npad = 1 << (n - 1).bit_length()
        lg = npad.bit_length() - 1
        smem_bytes = (stop - start) * (lg * (lg + 1) // 2) * npad * 12.0 * 2
        ach = smem_bytes / (main_kernel_ms * 1e-3) / 1e9
        peak = 148 * 128 * 1.965                    # 128 B/clk/SM shared-memory bandwidth at 1.965 GHz, GB/s
        roof = {"bound": "smem", "kernel": "sl_kernel", "achieved": ach, "peak": peak, "unit": "GB/s (shared memory)",
                "frac": ach / peak, "traffic": None, "kernel_ms": main_kernel_ms,
                "evals_per_s_kernel": units_local / (main_kernel_ms * 1e-3),
                "peak_source": "nominal 128 B/clk/SM x 148 SMs x 1.965 GHz (shared-memory bandwidth; the sort never "
                               "leaves the SM, HBM traffic is 16 B/sample per block)"}
    elif kind == "pdm":
        ach = units_local / (main_kernel_ms * 1e-3) / 1e9
        roof = {"bound": "smem", "kernel": "pdm_hist_kernel", "achieved": ach, "peak": PDM_PEAK_GEVALS_MEASURED,
                "unit": "Gevals/s", "frac": ach / PDM_PEAK_GEVALS_MEASURED,
                "traffic": ncu_traffic("ncu_pdm_hist_r01e.json") if args.workload == "pdm_c3" and world == 1 else None,
                "kernel_ms": main_kernel_ms,
                "peak_source": "profiles/pipes_r01.json smem_private_u32_atoms (one shared-memory ATOMS.ADD on a private "
                               "32-bit column word per sample update, 13.7 per clk per SM: the floor of the "
                               "kernel's histogram update); path is shared-memory/issue bound, not HBM or tensor bound"}
    else:
        ach = units_local * FLOP_PER_EVAL_GLS / (main_kernel_ms * 1e-3) / 1e12
        roof = {"bound": "fp32", "kernel": "gls_strip_kernel", "achieved": ach, "peak": FP32_PEAK_TFLOPS_MEASURED,
                "unit": "TFLOP/s", "frac": ach / FP32_PEAK_TFLOPS_MEASURED,
                "traffic": ncu_traffic("ncu_gls_strip_r01c.json") if args.workload == "gls_c2" and world == 1 else None,
                "kernel_ms": main_kernel_ms, "flop_per_eval": FLOP_PER_EVAL_GLS,
                "note": ("shared-timestamp kernel: rotation and window sums are shared by 8 series, so the 20 FLOP "
                         "per evaluation of the accounting figure are not all executed; frac > 1 is expected")
                if kind == "gls_multi" else None,
                "evals_per_s_kernel": units_local / (main_kernel_ms * 1e-3),
                "executed": None if kind == "gls_multi" else {
                    "instr_per_eval": GLS_EXECUTED_INSTR_PER_EVAL, "flop_per_eval": GLS_EXECUTED_FLOP_PER_EVAL,
                    "issue_frac": units_local * GLS_EXECUTED_INSTR_PER_EVAL / (main_kernel_ms * 1e-3) / 1e9
                    / FP32_PEAK_GINSTR_MEASURED,
                    "note": "three-term recurrence along the frequency axis: 8 FP32 instructions per evaluation "
                            "instead of the 12 of SURVEY 8d's accounting figure; `frac` stays on the 20-FLOP figure "
                            "as SURVEY 8d prescribes, issue_frac = executed instructions / measured FP32 issue peak"},
                "peak_source": "profiles/pipes_r01.json ffma_shared_operands x2 FLOP (measured on this pool's B200; "
                               "MEASURED_PEAKS.json has no FP32 entry; nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5)",
                "hbm": {"achieved_gbs": (32.0 * n + 6 * 8 * 2 * (units_local / n)) / (main_kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": measured_hbm_gbs()[0], "peak_source": measured_hbm_gbs()[1],
                        "note": "algorithmic bytes: 32 B/sample record + FP64 partial flush; the path is "
                                "compute bound, this line only shows how far from the HBM roofline it sits"}}

    # ---- cpu baseline (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        reps, evals = 0, 0
        cpu_reference_step(wl)             # untimed warm-up (imports, worker start-up, calibration)
        t0 = time.perf_counter()
        while True:
            e, cores, desc = cpu_reference_step(wl)
            evals += e
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 50:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": evals / dt, "unit": unit, "cores": cores, "kind": "port",
               "sample": f"{desc}; {reps} repetition(s) in {dt:.1f} s", "host_cpus": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"pdm": "f64 phase, integer (fixed-point) histograms", "sl": "f64"}.get(kind, "f32 sums, f64 phase/epilogue"),
            "data": "synthetic",
            "config": {"workload": wl["name"], "units_per_gpu": per_gpu, "sharding": "frequency grid" if kind == "gls"
                       else ("period grid" if kind in ("pdm", "sl") else "light-curve batch"),
                       "kernel": "glsm_strip_kernel" if kind == "gls_multi" else None,
                       "l2": "flushed between timed steps (256 MiB memset, not timed); per-step CUDA events summed",
                       "collective": ("none (1 GPU)" if world == 1 else
                                      ("all-gather fused into the epilogue kernel: stores to every rank's symmetric "
                                       "buffer over NVLink peer memory + 2 device barriers (no NCCL call)"
                                       if gather_mode[0] == "p2p" else
                                       "one NCCL all-gather of [values, best, index] per step"))},
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / args.steps * 1e3, "api": "pdc_gls / pdc_pdm host-pointer C-ABI call (ctypes)"},
            "gpu_launches": int(launches),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
