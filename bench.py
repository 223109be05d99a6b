#!/usr/bin/env python
"""bench.py -- sample*frequency evaluations per second of the GLS / PDM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-configs]

One "step" = one pass of the hot path over one batch of synthetic input.

PRIMARY line (the driver's `value`): `gls_c2`, BASELINE.json configs[1] -- GLS on a Kepler-like light curve,
65,000 points x 1e5 frequencies PER GPU.  With N GPUs the grid is N x 1e5 long and sharded across ranks (weak
scaling, as in round 1 so the records stay comparable) with the all-gather of power + arg-max fused into the
epilogue kernel over NVLink peer memory.

`configs` block (same JSON line, measured in the same run, default workload only): the other BASELINE.json
configs at their FULL, FIXED sizes, i.e. STRONG scaling when N > 1:
  C3_pdm        PDM 1e5 points x 1e5 trial periods, nb=10 nc=2 (configs[2]); period grid split over the N ranks
  C5_gls        GLS 1e6 points x 1e7 frequencies (configs[4]); frequency grid split over the N ranks
  C4_gls_batch  batched GLS 1e4 TESS-like curves x 20,000 points x 1e4 frequencies (configs[3]); curves split
  C1_gls        GLS 1,000 x 1e4 (configs[0], the reference's own CPU-sized case; N = 1 only: it must not be sharded)
Each record carries ms_per_step, value, roofline, e2e and a `parity` object.

Every record's `parity`: after the timed region rank 0 checks the GATHERED result (what the collective delivered)
against the C oracle (oracle/oracle.c: the reference's formula with exact sums / `PDM._pdm`) at ~64 strided grid
indices plus a window round the reported arg-extremum, that the reported arg-extremum is the extremum of the
gathered array, and -- where the reference's own algorithm fits in a second or so -- that its peak index is the
same.  A failed check makes the run exit non-zero AFTER printing the line.

Printed line (rank 0): metric/value/unit/... as the driver contract asks, plus
  roofline      dominant kernel: achieved = evals x 20 FLOP (SURVEY.md 8d's accounting figure) / kernel time (CUDA events
                on the launching stream inside the library).  GLS on the tensor-core kernel (gls_umma_kernel, the default for
                large calls): bound "tensor", peak = MEASURED_PEAKS.json bf16_tflops, plus `executed` (72 tensor FLOP per
                evaluation) and the same figure against the FP32 roofline.  GLS on gls_strip_kernel (small calls,
                PDC_GLS_UMMA=0): bound "fp32", peak = measured FFMA rate (profiles/r01/pipes_r01.json).  PDM: updates/s vs
                the measured shared-memory ATOMS.ADD rate.
  cpu_baseline  the reference's CPU path timed on this box's host cores: the UNMODIFIED reference files when
                oracle/_ref holds them (kind "reference"; staged by oracle/make_ref.py), else the numpy port (kind "port").
  e2e           same metric through the public host API with HOST buffers: N = 1 the host-pointer C-ABI call
                (pdc_gls / pdc_pdm / pdc_gls_batch); N > 1 the sharded API (`dist.gls_sharded_p2p`, ... = what
                `GLS(shard="p2p")` calls): every rank uploads from pinned host memory, runs its shard, the gather is
                fused into the epilogue, and rank 0 downloads the FULL result to host memory inside the timed region.

--impl reference times the reference's own CPU implementation on the same config/metric/unit (rank 0 only).
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

KERNEL_TAG = "r02"            # profiles/ncu_{gls_strip,pdm_hist}_<tag>.json: `ncu --set full` capture of the CURRENT kernels
FLOP_PER_EVAL_GLS = 20.0      # 12 FP32 instructions = 20 FLOP per sample*frequency
FP32_PEAK_TFLOPS_MEASURED = 72.3   # profiles/r01/pipes_r01.json: 36,172 GFFMA/s x 2
FP32_PEAK_GINSTR_MEASURED = 36172.0  # same measurement as thread-instructions/s (125 per clk per SM)
# what gls_strip_kernel<16,128> executes in its three-term form (SASS hot loop: 282 instructions per
# 2 samples x 16 frequencies; 6 FFMA + 2 FADD per evaluation): the 20 FLOP of the accounting figure are NOT all executed
GLS_EXECUTED_INSTR_PER_EVAL = 282.0 / 32.0
GLS_EXECUTED_FLOP_PER_EVAL = 14.0
# gls_umma_kernel (tcgen05): per evaluation 6 sums x 2 multiply-adds (angle addition: cos and sin slot), each taken as
# hi*hi + hi*lo + lo*hi in fp16 -> 36 tensor-core multiply-adds = 72 FLOP executed for the 20 FLOP of the accounting figure
GLS_UMMA_EXECUTED_FLOP_PER_EVAL = 72.0
PDM_PEAK_GEVALS_MEASURED = 3841.3  # profiles/r01/pipes_r01.json smem_private_u32_atoms: private-column ATOMS.ADD, updates/s

METRICS = {"pdm": "PDM sample*period evaluations per second",
           "sl": "String Length sample*period evaluations per second",
           "ce": "conditional-entropy sample*period evaluations per second"}
UNIT = "evals/s"


def measured_hbm_gbs():
    """HBM copy bandwidth the driver measured on this pool (MEASURED_PEAKS.json); 6650 = the profiling
    recipe's stated fallback if the file is absent."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_tflops(sustained=False):
    """Dense bf16/fp16 tensor peak measured by the driver on this pool's B200s (cuBLAS 8192^3): the burst figure for a
    kernel timed alone, the sustained one for seconds-long steps; fallback per B200_PROFILING.md."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        key = "bf16_tflops_sustained" if sustained else "bf16_tflops"
        return float(d[key]), f"MEASURED_PEAKS.json {key} (of measured)"
    except Exception:
        return (1400.0 if sustained else 1590.0), "B200_PROFILING.md fallback (of fallback)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of the CURRENT kernel (profiles/ncu_<kernel>_<KERNEL_TAG>.json); None (and the file
    name that was looked for) if this tree's kernel has not been captured."""
    name = f"ncu_{kernel}_{KERNEL_TAG}.json"
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f).get("dram_bytes_per_launch"), name
    except Exception:
        return None, name


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
                name=f"GLS 1e6 points x {nf_total} frequencies (C5)")


def make_pdm_c3(np_total, t_offset=0.0):
    rng = np.random.default_rng(3)
    n = 100_000
    t = np.sort(rng.uniform(0, 1000.0, n)) + t_offset
    x = 1000 + np.sin(2 * np.pi * t / 3.7) + 0.8 * np.sin(4 * np.pi * t / 3.7) + rng.standard_normal(n)
    periods = np.linspace(1.0, 11.0, np_total)
    return dict(kind="pdm", t=t, y=x, periods=periods, nb=10, nc=2, nf=np_total,
                name=f"PDM 1e5 points x {np_total} trial periods, nb=10 nc=2 (C3)" +
                     (f", time stamps offset by {t_offset:g} (Julian dates)" if t_offset else ""))


def make_ce_c3(np_total):
    """Conditional entropy (reference TODO, phase.py:13) on the C3 light curve: 10 phase x 5 magnitude bins."""
    wl = make_pdm_c3(np_total)
    wl.update(kind="ce", nb=10, nm=5,
              name=f"conditional entropy 1e5 points x {np_total} trial periods, 10 phase x 5 magnitude bins (C3 shape)")
    return wl


def make_sl(np_total, n=2000):
    """String Length on a sparse light curve (the method's use case; SURVEY 8f row 2 has no BASELINE config).
    n > 16,384 sorts in global scratch instead of shared memory (sl_kernel's fallback)."""
    from oracle import stringlength_numpy
    rng = np.random.default_rng(7)
    t = np.sort(rng.uniform(0, 1000.0, n))
    x = 12.0 + 0.4 * np.sin(2 * np.pi * t / 4.3) + 0.05 * rng.standard_normal(n)
    periods = stringlength_numpy.period_grid(t, dphi=0.1, n_periods=np_total)
    return dict(kind="sl", t=t, y=stringlength_numpy.scale(x), periods=periods, nf=np_total,
                name=f"String Length {n:,} points x {np_total} trial periods (phase.py:18-72)")


def make_gls_c4(curves, first=0, total=None):
    """TESS-like survey batch: curve b (global index) has its own seed 4000+b, so a rank can generate only its own
    curves [first, first + curves) of a `total`-curve survey."""
    n, nf = 20_000, 10_000
    ts, ys, fm, dfs = [], [], [], []
    for b in range(first, first + curves):
        rng = np.random.default_rng(4000 + b)
        keep = rng.uniform(size=n + n // 50) > 0.01
        tt = (np.arange(n + n // 50)[keep][:n]) * (2.0 / 1440.0) + rng.uniform(0, 0.2 / 1440.0, n)
        P = rng.uniform(0.5, 10.0)
        ys.append(1000 + np.sin(2 * np.pi * tt / P) + rng.standard_normal(n))
        ts.append(tt)
        d = 1 / (tt[-1] - tt[0]) / 5
        dfs.append(d)
        fm.append(0.5 * d)
    total = curves if total is None else total
    return dict(kind="gls_batch", t=np.concatenate(ts), y=np.concatenate(ys), first=first, total=total,
                offsets=np.arange(curves + 1, dtype=np.int64) * n, fmin=np.array(fm), df=np.array(dfs), nf=nf,
                name=f"batched GLS {total} TESS-like curves x 20,000 points x 1e4 frequencies (C4)")


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
# reference arm / cpu baseline: the reference's own CPU path
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


def reference_kind():
    """"reference" when the unmodified reference files are loadable (authoring container: /root/reference; GPU box:
    oracle/_ref staged by oracle/make_ref.py), else "port" (numpy restatement in oracle/)."""
    try:
        from oracle import refload
        if refload.available():
            refload.load()
            return "reference"
    except Exception as exc:  # noqa: BLE001
        print(f"[bench] unmodified reference not loadable, timing the port: {exc}", file=sys.stderr)
    return "port"


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


def _ref_gls(t, y, fmin, df, nf):
    """One GLS call of the UNMODIFIED reference class (spectral.py:53-135) on the bench grid: the recipes use the
    reference's default n=5 and fmin, and fmax = fmin + (nf - 1.5) df makes its np.arange yield exactly nf points."""
    from oracle import refload
    spectral, _ = refload.load()
    g = spectral.GLS(fmax=fmin + (nf - 1.5) * df)
    out = g(refload.TSeries(t, y, assume_sorted=True))
    if out.values.size != nf:
        raise RuntimeError(f"reference grid has {out.values.size} points, expected {nf}")
    return out.values


def _cpu_gls_one(job):
    t, y, fmin, df, nf, kind = job
    if kind == "reference":
        return float(np.nanmax(_ref_gls(t, y, fmin, df, nf)))
    from oracle import gls_numpy
    return float(np.nanmax(gls_numpy.gls_power(t, y, None, fmin, df, nf, True, False)))


def _cpu_sl_chunk(t, m, periods):
    from oracle import stringlength_numpy
    return stringlength_numpy.string_lengths(t, m, periods)


def _cpu_ce_chunk(t, x, periods, nb, nm):
    from oracle import ce_numpy
    return ce_numpy.ce(t, x, periods, nb, nm)


def cpu_reference_step(wl, kind):
    """One step of the reference's CPU path on (a bounded sample of) the workload.
    Returns (evals processed, cores used, description of the sample)."""
    from oracle import gls_numpy, pdm_numpy
    who = "unmodified reference files (oracle/_ref)" if kind == "reference" else "numpy port of the reference algorithm"
    if wl["kind"] == "gls":
        # the reference's own algorithm: FFT extirpolation, single-threaded numpy (spectral.py:11-40)
        if kind == "reference":
            _ref_gls(wl["t"], wl["y"], wl["fmin"], wl["df"], wl["nf"])
        else:
            gls_numpy.gls_power(wl["t"], wl["y"], None, wl["fmin"], wl["df"], wl["nf"], True, False)
        return wl["t"].size * wl["nf"], 1, f"full workload, GLS.__call__ (FFT extirpolation, spectral.py:74-135), {who}, 1 thread"
    if wl["kind"] == "gls_multi":
        cores = os.cpu_count() or 1
        B = min(wl["S"], 8 * cores)
        jobs = [(wl["t"], wl["y"][b], wl["fmin"], wl["df"], wl["nf"], kind) for b in range(B)]
        used = _cpu_batch(jobs, "multi")
        return wl["t"].size * B * wl["nf"], used, (f"first {B} series, python loop over GLS.__call__, {who}, on "
                                                   f"{used} core(s) (faster of in-process / Pool({cores}))")
    if wl["kind"] == "gls_batch":
        # the reference has no batch API: a survey is a loop over curves; mapped over all host cores here
        cores = os.cpu_count() or 1
        B = min(len(wl["offsets"]) - 1, 8 * cores)
        jobs = [(wl["t"][wl["offsets"][b]:wl["offsets"][b + 1]], wl["y"][wl["offsets"][b]:wl["offsets"][b + 1]],
                 wl["fmin"][b], wl["df"][b], wl["nf"], kind) for b in range(B)]
        used = _cpu_batch(jobs, "batch")
        return int(wl["offsets"][B]) * wl["nf"], used, (f"first {B} curves, python loop over GLS.__call__, {who}, on "
                                                        f"{used} core(s) (faster of in-process / Pool({cores}); the "
                                                        "reference has no batch API)")
    cores = os.cpu_count() or 1
    if wl["kind"] == "sl":
        sample = wl["periods"][:: max(1, wl["periods"].size // (512 * cores))][: 512 * cores]
        _pool(cores).starmap(_cpu_sl_chunk, [(wl["t"], wl["y"], c) for c in np.array_split(sample, cores)])
        return wl["t"].size * sample.size, cores, (f"{sample.size} of {wl['periods'].size} trial periods (strided), "
                                                   f"multiprocessing.Pool({cores}) as phase.py:68-70")
    if wl["kind"] == "ce":
        sample = wl["periods"][:: max(1, wl["periods"].size // (64 * cores))][: 64 * cores]
        _pool(cores).starmap(_cpu_ce_chunk, [(wl["t"], wl["y"], c, wl["nb"], wl["nm"]) for c in np.array_split(sample, cores)])
        return wl["t"].size * sample.size, cores, (f"{sample.size} of {wl['periods'].size} trial periods (strided), numpy "
                                                   f"histogram2d oracle over Pool({cores}); the reference has no implementation")
    nsample = 24 * cores
    if kind == "reference":
        # the reference's OWN fan-out: PDM.__call__ forks a multiprocessing.Pool(cores) per call and maps self._pdm
        # over linspace(p_min, p_max, n_periods) (phase.py:180-187); the sample is a coarser grid over the same range
        from oracle import refload
        _, phase = refload.load()
        p = wl["periods"]
        pdm = phase.PDM(nb=wl["nb"], nc=wl["nc"], p_min=float(p[0]), p_max=float(p[-1]), n_periods=nsample, cores=cores)
        pdm(refload.TSeries(wl["t"], wl["y"], assume_sorted=True))
        return wl["t"].size * nsample, cores, (f"{nsample} trial periods spanning the same range (of {p.size}), the "
                                               f"reference's own PDM.__call__ with its multiprocessing.Pool({cores}) "
                                               "(phase.py:151-195), unmodified files")
    sample = wl["periods"][:: max(1, wl["periods"].size // nsample)][:nsample]
    pdm_numpy.pdm_pool(wl["t"], wl["y"], sample, wl["nb"], wl["nc"], cores, sort=True)
    return wl["t"].size * sample.size, cores, (f"{sample.size} of {wl['periods'].size} trial periods (strided), "
                                               f"multiprocessing.Pool({cores}) as phase.py:185-186, numpy port")


def time_cpu_reference(wl, kind, budget_s=10.0, max_reps=50, warm=True):
    """cpu_baseline object: repeat the reference step for about `budget_s` seconds."""
    if warm:
        cpu_reference_step(wl, kind)       # untimed warm-up (imports, worker start-up, calibration)
    reps, evals = 0, 0
    t0 = time.perf_counter()
    while True:
        e, cores, desc = cpu_reference_step(wl, kind)
        evals += e
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= max_reps:
            break
    dt = time.perf_counter() - t0
    return {"value": evals / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{desc}; {reps} repetition(s) in {dt:.1f} s", "host_cpus": os.cpu_count()}


def workload_config(wl, per_gpu, world, scaling):
    """The `config` object: identical for our arm and the reference arm (the driver compares them)."""
    kind = wl["kind"]
    return {"workload": wl["name"], "units_per_gpu": per_gpu, "n_gpus": world, "scaling": scaling,
            "sharding": "frequency grid" if kind in ("gls", "gls_multi") else
                        ("period grid" if kind in ("pdm", "sl", "ce") else "light-curve batch"),
            "l2": "flushed between timed steps (256 MiB memset, not timed); per-step CUDA events summed",
            "collective": "none (1 GPU)" if world == 1 else
                          ("all-gather fused into the epilogue kernel: stores to every rank's symmetric buffer over "
                           "NVLink peer memory + 1 device barrier (no NCCL call); NCCL all-gather of "
                           "[values, best, index] for the batch / String Length workloads or with --gather nccl")}


def run_reference(args, wl, per_gpu, world, scaling, extra_configs):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = reference_kind()
    for _ in range(args.warmup):
        cpu_reference_step(wl, kind)
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        e, cores, desc = cpu_reference_step(wl, kind)
        evals += e
    dt = time.perf_counter() - t0
    value = evals / dt
    metric = METRICS.get(wl["kind"], "GLS sample*frequency evaluations per second")
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, per_gpu, world, scaling),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if extra_configs:
        cfgs = {}
        for name, mk in extra_configs:
            try:
                w2 = mk()
                budget = 1.0 if name == "C5_gls" else 8.0       # C5: one pass of the 2^26-point FFT path (~20 s, 4 GB)
                cb = time_cpu_reference(w2, kind, budget_s=budget, max_reps=20, warm=(name != "C5_gls"))
                cfgs[name] = {"workload": w2["name"], "value": cb["value"], "unit": UNIT, "cores": cb["cores"],
                              "kind": kind, "sample": cb["sample"]}
            except Exception as exc:  # noqa: BLE001
                cfgs[name] = {"error": repr(exc)}
        line["configs"] = cfgs
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


def strided_sel(n, center, count=64, halfwin=4):
    """~`count` strided indices of [0, n) + a window round `center` + both ends."""
    parts = [np.arange(0, n, max(1, n // count)), [n - 1]]
    if center is not None and center >= 0:
        parts.append(np.arange(max(0, center - halfwin), min(n, center + halfwin + 1)))
    return np.unique(np.concatenate(parts)).astype(np.int64)


class Env:
    """torch / torch.distributed / ctx state shared by all workloads of one run."""

    def __init__(self, gather):
        import torch
        import torch.distributed as dist
        from periodicity_b200 import _ffi
        from periodicity_b200 import dist as pdist
        self.torch, self.dist, self._ffi, self.pdist = torch, dist, _ffi, pdist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = _ffi.default_context(self.local_rank)
        self.gather = gather
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        tt = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def all_ok(self, ok):
        if self.world == 1:
            return bool(ok)
        tt = self.torch.tensor([1.0 if ok else 0.0], device=self.dev)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MIN)
        return tt.item() >= 1


def run_workload(env, wl, per_gpu, scaling, steps, warmup, cpu_budget_s=10.0, want_cpu=True, check_ref_peak=True):
    """Time one workload (device-resident `value`, `e2e`, roofline), check the gathered result against the oracle.
    Returns the record (rank 0) / None (other ranks) and the parity verdict (all ranks)."""
    torch, dist, pdist, _ffi, ctx = env.torch, env.dist, env.pdist, env._ffi, env.ctx
    rank, world, dev = env.rank, env.world, env.dev
    kind = wl["kind"]
    n = wl["t"].size if kind != "gls_batch" else None

    # ---- device-resident inputs ------------------------------------------------------------
    t_pin = torch.from_numpy(wl["t"]).pin_memory()
    y_pin = torch.from_numpy(np.ascontiguousarray(wl["y"])).pin_memory()
    t_d = t_pin.to(dev)
    y_d = y_pin.to(dev)
    if kind == "gls_multi":
        b0, b1 = pdist.batch_shard_bounds(wl["S"], rank, world)
        units_local = n * wl["nf"] * (b1 - b0)
        evals_total = n * wl["nf"] * wl["S"]
        L = b1 - b0
        ym_d = y_d[b0:b1].contiguous()
        pm_arg = torch.empty(L, dtype=torch.int64, device=dev)
        pm_max = torch.empty(L, dtype=torch.float64, device=dev)
    elif kind == "gls_batch":
        # wl holds only this rank's curves [first, first + B_local) of a `total`-curve survey
        off = wl["offsets"]
        B_local, B_total = len(off) - 1, wl["total"]
        units_local = int(off[-1] - off[0]) * wl["nf"]
        L = max(1, -(-B_total // world))
    else:
        start, stop, L = pdist.shard_bounds(wl["nf"], rank, world)
        units_local = n * (stop - start)
        evals_total = n * wl["nf"]
        if kind in ("pdm", "sl", "ce"):
            p_d = torch.from_numpy(wl["periods"][start:stop].copy()).to(dev)
            pfull_d = torch.from_numpy(wl["periods"]).to(dev)
    if kind == "gls_batch":
        tot = torch.tensor([float(units_local)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        evals_total = int(tot.item())

    gather_mode = [env.gather if (world > 1 and kind in ("gls", "pdm")) else "nccl"]

    def step_device():
        """One pass of the hot path with inputs resident in HBM; returns (values, best value(s), best index(es))."""
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
        elif kind == "ce":
            h, arg, mn = pdist.ce_torch(t_d, y_d, p_d, wl["nb"], wl["nm"], ctx=ctx)
            if world == 1:
                return h, mn, arg
            garg = (arg + start).to(torch.float64).reshape(())
            vals, bests, args_ = pdist.all_gather_packed(h, mn.reshape(()), garg, L)
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
            # survey batch: per-curve (max, argmax) only -- the periodograms stay on the GPU (power_out = NULL)
            _, arg, mx = pdist.gls_batch_torch(t_d, y_d, None, off, wl["fmin"], wl["df"], wl["nf"],
                                               want_power=False, ctx=ctx)
            if world == 1:
                return None, mx, arg
            vals, bests, args_ = pdist.all_gather_packed(mx, mx.max(), arg.to(torch.float64).max(), L)
            return None, vals, None
        return vals, bests, args_

    if gather_mode[0] == "p2p":
        try:                               # symmetric-memory rendezvous is collective: every rank tries, all agree
            step_device()
            torch.cuda.synchronize()
            ok = True
        except Exception as exc:           # noqa: BLE001 -- transport set-up failed: fall back to the NCCL all-gather
            print(f"[bench] p2p gather unavailable on rank {rank}: {exc}", file=sys.stderr)
            ok = False
        if not env.all_ok(ok):
            gather_mode[0] = "nccl"

    for _ in range(warmup):
        step_device()
    env.barrier()

    sampler = ClockSampler(env.local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    launches0 = ctx.launch_count
    kms0, kcnt0 = ctx.main_kernel_ms_total()
    env.barrier()
    for k in range(steps):
        env.flush.zero_()                  # evict L2 between timed iterations (not timed)
        ev[k][0].record()
        result = step_device()
        ev[k][1].record()
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0
    env.barrier()
    clocks = sampler.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    kms1, kcnt1 = ctx.main_kernel_ms_total()
    main_kernel_ms = (kms1 - kms0) / max(1, kcnt1 - kcnt0)   # average launch of the dominant kernel, timed region
    total_ms = env.max_over_ranks(total_ms)
    ms_per_step = total_ms / steps
    value = evals_total / (ms_per_step * 1e-3)

    # ---- parity of the gathered device-resident result (rank 0 checks, every rank learns the verdict) ----------
    parity = None
    if rank == 0:
        try:
            parity = check_parity(env, wl, result, gather_mode[0], check_ref_peak)
        except Exception as exc:  # noqa: BLE001
            parity = {"ok": False, "error": repr(exc)}
    parity_ok = env.all_ok(parity["ok"] if parity is not None else True)

    # ---- e2e: public host API, pinned host inputs, H2D + D2H inside the timed region ---------------------------
    th, yh = t_pin.numpy(), y_pin.numpy()
    e2e_api = [""]

    def step_e2e():
        if kind == "gls":
            if world > 1 and gather_mode[0] == "p2p":
                e2e_api[0] = "dist.gls_sharded_p2p (what GLS(shard='p2p') calls): pinned H2D on every rank, fused gather, full periodogram to host on rank 0"
                return pdist.gls_sharded_p2p(th, yh, None, wl["fmin"], wl["df"], wl["nf"], device=env.local_rank,
                                             root=0, copy=False)[0]
            if world > 1:
                e2e_api[0] = "dist.gls_sharded (GLS(shard=True)): pinned H2D, NCCL all-gather, full periodogram to host"
                return pdist.gls_sharded(th, yh, None, wl["fmin"], wl["df"], wl["nf"], device=env.local_rank)[0]
            e2e_api[0] = "pdc_gls host-pointer C-ABI call (ctypes)"
            return ctx.gls(th, yh, None, wl["fmin"], wl["df"], wl["nf"])[0]
        if kind == "pdm":
            if world > 1 and gather_mode[0] == "p2p":
                e2e_api[0] = "dist.pdm_sharded_p2p (what PDM(shard='p2p') calls): pinned H2D on every rank, fused gather, full theta array to host on rank 0"
                return pdist.pdm_sharded_p2p(th, yh, wl["periods"], wl["nb"], wl["nc"], device=env.local_rank,
                                             root=0, copy=False)[0]
            if world > 1:
                e2e_api[0] = "dist.pdm_sharded (PDM(shard=True))"
                return pdist.pdm_sharded(th, yh, wl["periods"], wl["nb"], wl["nc"], device=env.local_rank)[0]
            e2e_api[0] = "pdc_pdm host-pointer C-ABI call (ctypes)"
            return ctx.pdm(th, yh, wl["periods"], wl["nb"], wl["nc"])[0]
        if kind == "ce":
            e2e_api[0] = "pdc_ce host-pointer C-ABI call (ctypes), this rank's slice"
            return ctx.ce(th, yh, wl["periods"][start:stop], wl["nb"], wl["nm"])[0]
        if kind == "sl":
            e2e_api[0] = "pdc_stringlength host-pointer C-ABI call (ctypes), this rank's slice"
            return ctx.stringlength(th, yh, wl["periods"][start:stop])[0]
        if kind == "gls_multi":
            e2e_api[0] = "pdc_gls_multi host-pointer C-ABI call (ctypes), this rank's series"
            return ctx.gls_multi(th, yh[b0:b1], None, wl["fmin"], wl["df"], wl["nf"], want_power=False)[2]
        e2e_api[0] = ("pdc_gls_batch host-pointer C-ABI call (ctypes) on this rank's curves" +
                      ("" if world == 1 else " + one NCCL all-gather of (max, argmax), result on the host of rank 0"))
        _, a, m = ctx.gls_batch(th, yh, None, off, wl["fmin"], wl["df"], wl["nf"], want_power=False)
        if world == 1:
            return m
        packed = torch.full((L, 2), float("nan"), dtype=torch.float64, device=dev)
        packed[:B_local, 0] = torch.from_numpy(m).to(dev)
        packed[:B_local, 1] = torch.from_numpy(a.astype(np.float64)).to(dev)
        allp = torch.empty((world * L, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allp.view(-1), packed.view(-1))
        return allp.cpu().numpy() if rank == 0 else None

    e2e_warm = max(1, min(warmup, 3))
    for _ in range(e2e_warm):
        step_e2e()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = step_e2e()
    torch.cuda.synchronize()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    e2e_value = evals_total * steps / e2e_s
    if kind == "gls_multi":
        h2d = 8 * n * (1 + (b1 - b0))
        d2h = 16 * (b1 - b0)
    elif kind == "gls_batch":
        h2d = 2 * 8 * int(off[-1] - off[0])
        d2h = 16 * (B_local if world == 1 else world * L)
    elif world > 1 and kind in ("gls", "pdm"):
        h2d = 2 * 8 * n + (8 * wl["nf"] if kind == "pdm" else 0)       # per rank: every rank uploads the inputs
        d2h = 8 * wl["nf"] + 16 * world                                  # rank 0: full result + candidate table
    else:
        h2d = 2 * 8 * n + (8 * (stop - start) if kind in ("pdm", "sl", "ce") else 0)
        d2h = 8 * (stop - start) + 16

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
    elif kind in ("pdm", "ce"):
        ach = units_local / (main_kernel_ms * 1e-3) / 1e9
        traffic, tfile = ncu_traffic("pdm_hist" if kind == "pdm" else "ce_hist")
        roof = {"bound": "smem", "kernel": "pdm_hist_kernel" if kind == "pdm" else "ce_hist_kernel", "achieved": ach,
                "peak": PDM_PEAK_GEVALS_MEASURED, "unit": "Gevals/s", "frac": ach / PDM_PEAK_GEVALS_MEASURED,
                "traffic": traffic if world == 1 and wl["nf"] == 100_000 else None, "traffic_source": f"profiles/{tfile}",
                "kernel_ms": main_kernel_ms,
                "peak_source": "profiles/r01/pipes_r01.json smem_private_u32_atoms (one shared-memory ATOMS.ADD on a private "
                               "32-bit column word per sample update, 13.7 per clk per SM: the floor of the "
                               "kernel's histogram update); path is shared-memory/issue bound, not HBM or tensor bound"}
    elif kind != "gls_multi" and ctx.last_gls_path() in (1, 2, 3):
        # tensor-core formulation (gls_umma_kernel): bound = tcgen05 fp16 rate
        ach = units_local * FLOP_PER_EVAL_GLS / (main_kernel_ms * 1e-3) / 1e12
        executed = units_local * GLS_UMMA_EXECUTED_FLOP_PER_EVAL / (main_kernel_ms * 1e-3) / 1e12
        nsamp = n if n is not None else int(off[-1] - off[0])
        peak, peak_src = measured_tensor_tflops(sustained=main_kernel_ms > 100.0)
        traffic, tfile = ncu_traffic("gls_umma2" if ctx.last_gls_path() == 3 else "gls_umma")
        roof = {"bound": "tensor", "kernel": {1: "gls_umma_kernel", 2: "gls_umma_kernel<fine operand precomputed>",
                                              3: "gls_umma2_kernel (cta_group::2, fine operand precomputed)"}[ctx.last_gls_path()],
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": traffic if (kind == "gls" and world == 1 and nsamp == 65_000 and wl["nf"] == 100_000) else None,
                "traffic_source": f"profiles/{tfile}",
                "traffic_note": "almost all of it is the precomputed fine operand (16 KB per 16 samples, type and 128 fine "
                                "indices: 133 MB on C2 for the one-CTA kernel, 266 MB for the pair kernel), written once per "
                                "call and read once per launch from HBM -- about 4 % of the HBM bandwidth, traded for half of "
                                "the operand arithmetic; the algorithmic input is 1.9 MB",
                "kernel_ms": main_kernel_ms, "flop_per_eval": FLOP_PER_EVAL_GLS,
                "evals_per_s_kernel": units_local / (main_kernel_ms * 1e-3),
                "executed": {"flop_per_eval": GLS_UMMA_EXECUTED_FLOP_PER_EVAL, "tflops": executed, "frac_of_peak": executed / peak,
                             "note": "angle addition across blocks of the grid turns the six sums into fp16 GEMMs over the "
                                     "sample axis: 12 multiply-adds per evaluation, each as hi*hi + hi*lo + lo*hi (fp16 pairs "
                                     "carry 22 bits) = 72 executed tensor FLOP for the 20 FLOP of SURVEY 8d's accounting "
                                     "figure; `frac` stays on the 20-FLOP figure as the contract prescribes, "
                                     "frac_of_peak = executed tensor FLOP / measured cuBLAS bf16 peak"},
                "vs_fp32_roofline": {"frac": ach / FP32_PEAK_TFLOPS_MEASURED, "peak": FP32_PEAK_TFLOPS_MEASURED,
                                     "note": "the same 20-FLOP figure against the FP32 issue peak that bounds gls_strip_kernel "
                                             "(0.93 there): the tensor-core formulation is past the SIMT roofline"},
                "peak_source": peak_src,
                "hbm": {"achieved_gbs": (32.0 * nsamp + 6 * 8 * 2 * (units_local / max(1, nsamp))) / (main_kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": measured_hbm_gbs()[0], "peak_source": measured_hbm_gbs()[1],
                        "note": "algorithmic bytes: 32 B/sample record + fixed-point flush; compute bound"}}
    else:
        ach = units_local * FLOP_PER_EVAL_GLS / (main_kernel_ms * 1e-3) / 1e12
        traffic, tfile = ncu_traffic("gls_strip")
        nsamp = n if n is not None else int(off[-1] - off[0])
        roof = {"bound": "fp32", "kernel": "glsm_strip_kernel" if kind == "gls_multi" else "gls_strip_kernel",
                "achieved": ach, "peak": FP32_PEAK_TFLOPS_MEASURED,
                "unit": "TFLOP/s", "frac": ach / FP32_PEAK_TFLOPS_MEASURED,
                "traffic": traffic if (kind == "gls" and world == 1 and nsamp == 65_000 and wl["nf"] == 100_000) else None,
                "traffic_source": f"profiles/{tfile}",
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
                "peak_source": "profiles/r01/pipes_r01.json ffma_shared_operands x2 FLOP (measured on this pool's B200; "
                               "MEASURED_PEAKS.json has no FP32 entry; nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5)",
                "hbm": {"achieved_gbs": (32.0 * nsamp + 6 * 8 * 2 * (units_local / max(1, nsamp))) / (main_kernel_ms * 1e-3) / 1e9,
                        "peak_gbs": measured_hbm_gbs()[0], "peak_source": measured_hbm_gbs()[1],
                        "note": "algorithmic bytes: 32 B/sample record + FP64 partial flush; the path is "
                                "compute bound, this line only shows how far from the HBM roofline it sits"}}

    # ---- cpu baseline (rank 0, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        cpu = time_cpu_reference(wl, reference_kind(), budget_s=cpu_budget_s)

    rec = None
    if rank == 0:
        rec = {
            "metric": METRICS.get(kind, "GLS sample*frequency evaluations per second"),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None,
            "dtype": {"pdm": "f64 phase, integer (fixed-point) histograms", "ce": "f64 phase, integer count histograms",
                      "sl": "f64"}.get(kind, "fp16 hi/lo tensor-core products, f32 accumulation, f64 phase/epilogue"
                                       if roof.get("bound") == "tensor" else "f32 sums, f64 phase/epilogue"),
            "data": "synthetic",
            "config": workload_config(wl, per_gpu, world, scaling),
            "gather_used": "none" if world == 1 else gather_mode[0],
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / steps * 1e3, "api": e2e_api[0]},
            "gpu_launches": int(launches),
        }
    return rec, parity_ok


def check_parity(env, wl, result, gather_mode, check_ref_peak):
    """Rank 0: the gathered, device-resident result of the LAST timed step against the C oracle."""
    from oracle import cport, gls_numpy
    torch = env.torch
    kind, world = wl["kind"], env.world
    vals, bests, args_ = result
    out = {"oracle": "oracle/oracle.c (reference formula, exact float64 sums / PDM._pdm)", "tolerance": 1e-5}
    if kind in ("gls", "pdm", "sl", "ce"):
        nf = wl["nf"]
        full = vals.reshape(-1)[:nf].cpu().numpy()
        sign = +1 if kind == "gls" else -1
        if world == 1:
            g_idx, g_val = int(args_.reshape(-1)[0].item()), float(bests.reshape(-1)[0].item())
        else:
            g_idx, g_val = env.pdist.reduce_best(bests.cpu().numpy(), args_.cpu().numpy().astype(np.int64), sign)
        host_idx = int(np.nanargmax(full) if sign > 0 else np.nanargmin(full))
        arg_ok = (g_idx == host_idx) and (g_val == full[host_idx])
        sel = strided_sel(nf, g_idx)
        if kind == "gls":
            ref = cport.gls_exact_at(wl["t"], wl["y"], None, wl["fmin"], wl["df"], sel)
            peak = float(np.nanmax(full))
            max_rel = float(np.nanmax(np.abs(full[sel] - ref)) / peak)
            oracle_best = int(sel[np.nanargmax(ref)])
        elif kind == "pdm":
            ref = cport.pdm(wl["t"], wl["y"], wl["periods"][sel], wl["nb"], wl["nc"])
            max_rel = float(np.nanmax(np.abs(full[sel] - ref) / np.abs(ref)))
            oracle_best = int(sel[np.nanargmin(ref)])
        elif kind == "ce":
            from oracle import ce_numpy
            ref = ce_numpy.ce(wl["t"], wl["y"], wl["periods"][sel], wl["nb"], wl["nm"])
            max_rel = float(np.nanmax(np.abs(full[sel] - ref) / np.maximum(np.abs(ref), 1e-300)))
            oracle_best = int(sel[np.nanargmin(ref)])
            out["oracle"] = "oracle/ce_numpy.py (np.histogram2d statement; parity unpinned by the reference, phase.py:13 is a TODO)"
        else:
            from oracle import stringlength_numpy
            ref = stringlength_numpy.string_lengths(wl["t"], wl["y"], wl["periods"][sel])
            max_rel = float(np.nanmax(np.abs(full[sel] - ref) / np.abs(ref)))
            oracle_best = int(sel[np.nanargmin(ref)])
        out.update(max_rel=max_rel, checked=int(sel.size), argmax_index=g_idx,
                   argmax_ok=bool(arg_ok and oracle_best == g_idx))
        ok = max_rel <= 1e-5 and out["argmax_ok"]
        if kind == "gls" and check_ref_peak:
            fast = gls_numpy.gls_power(wl["t"], wl["y"], None, wl["fmin"], wl["df"], nf, True, False)
            out["reference_algorithm_argmax"] = int(np.nanargmax(fast))
            out["reference_peak_ok"] = bool(out["reference_algorithm_argmax"] == g_idx)
            ok = ok and out["reference_peak_ok"]
        out["ok"] = bool(ok)
        return out
    if kind == "gls_batch":
        # per-curve (max, argmax): device-resident run returned them for this rank's curves (N = 1) or gathered maxima
        off, nf = wl["offsets"], wl["nf"]
        B_local = len(off) - 1
        if world == 1:
            mx = bests.cpu().numpy()
            arg = args_.cpu().numpy()
        else:
            mx = bests.reshape(-1)[:B_local].cpu().numpy()      # rank 0's own curves come first in the gathered table
            arg = None
        curves = sorted(set([0, B_local // 2, B_local - 1]))
        worst, arg_ok, ref_ok = 0.0, True, True
        for b in curves:
            tb, yb = wl["t"][off[b]:off[b + 1]], wl["y"][off[b]:off[b + 1]]
            fast = gls_numpy.gls_power(tb, yb, None, wl["fmin"][b], wl["df"][b], nf, True, False)
            jb = int(np.nanargmax(fast)) if arg is None else int(arg[b])
            sel = strided_sel(nf, jb, count=16)
            ref = cport.gls_exact_at(tb, yb, None, wl["fmin"][b], wl["df"][b], sel)
            # the run keeps only the peaks: compare the peak value with the oracle's value at the same index
            worst = max(worst, abs(mx[b] - ref[np.searchsorted(sel, jb)]) / np.nanmax(ref))
            arg_ok = arg_ok and int(sel[np.nanargmax(ref)]) == jb
            ref_ok = ref_ok and int(np.nanargmax(fast)) == jb
        out.update(max_rel=float(worst), checked=len(curves), argmax_ok=bool(arg_ok), reference_peak_ok=bool(ref_ok),
                   note="survey run keeps per-curve (max, argmax) only; peak value vs oracle at the peak index, "
                        "peak index vs oracle window and vs the reference algorithm, on 3 of this rank's curves")
        out["ok"] = bool(worst <= 1e-5 and arg_ok and ref_ok)
        return out
    # gls_multi: per-series maxima of rank 0's series
    S0 = min(3, wl["S"])
    worst, arg_ok = 0.0, True
    mx = vals.reshape(-1).cpu().numpy()
    for s in range(S0):
        fast = gls_numpy.gls_power(wl["t"], wl["y"][s], None, wl["fmin"], wl["df"], wl["nf"], True, False)
        jb = int(np.nanargmax(fast))
        ref = cport.gls_exact_at(wl["t"], wl["y"][s], None, wl["fmin"], wl["df"], np.arange(max(0, jb - 2), jb + 3))
        worst = max(worst, abs(mx[s] - np.nanmax(ref)) / np.nanmax(ref))
    out.update(max_rel=float(worst), checked=S0, argmax_ok=bool(arg_ok), ok=bool(worst <= 1e-5))
    return out


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gls_c2",
                    choices=["gls_c2", "pdm_c3", "gls_c5", "gls_c5_full", "gls_c4", "gls_c4_full", "gls_c1", "gls_multi",
                             "sl", "sl_long", "ce_c3", "pdm_c3_jd"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="default workload only: skip the `configs` block (C3 / C5 / C4 / C1 at their BASELINE sizes)")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1 with --workload: keep the TOTAL size of the named config fixed and split it over the ranks")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1, GLS / PDM grids: 'p2p' = all-gather fused into the epilogue kernel over NVLink peer "
                         "memory (pdc_gls_dev_fanout), 'nccl' = one ncclAllGather after the kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    per_gpu = {"gls_c2": 100_000, "pdm_c3": 100_000, "gls_c5": 1_250_000, "gls_c5_full": 10_000_000,
               "gls_c4": 256, "gls_c4_full": 10_000, "gls_c1": 10_000, "gls_multi": 256, "sl": 100_000,
               "ce_c3": 100_000, "pdm_c3_jd": 100_000, "sl_long": 4_000}[args.workload]
    strong = args.strong or args.workload in ("gls_c5_full",)
    total_units = per_gpu if strong else per_gpu * max(world, 1)
    scaling = "strong" if (strong and world > 1) else "weak"

    def build(workload, units):
        if workload == "gls_c2":
            return make_gls_c2(units)
        if workload in ("gls_c5", "gls_c5_full"):
            return make_gls_c5(units)
        if workload == "gls_c1":
            return make_gls_c1(units)
        if workload == "gls_multi":
            return make_gls_multi(units)
        if workload == "pdm_c3":
            return make_pdm_c3(units)
        if workload == "pdm_c3_jd":
            return make_pdm_c3(units, t_offset=2_457_000.5)
        if workload == "ce_c3":
            return make_ce_c3(units)
        if workload == "sl":
            return make_sl(units)
        if workload == "sl_long":
            return make_sl(units, n=20_000)
        # survey batch: this rank's curves only
        from periodicity_b200 import dist as pdist
        b0, b1 = pdist.batch_shard_bounds(units, rank, world)
        return make_gls_c4(b1 - b0, first=b0, total=units)

    default_run = args.workload == "gls_c2" and not args.no_configs and not args.strong

    if args.impl == "reference":
        if rank != 0:
            return
        wl = build(args.workload, total_units) if args.workload not in ("gls_c4", "gls_c4_full") else \
            make_gls_c4(min(total_units, 8 * (os.cpu_count() or 1)), total=total_units)
        extra = None
        if default_run:
            extra = [("C3_pdm", lambda: make_pdm_c3(100_000)),
                     ("C4_gls_batch", lambda: make_gls_c4(8 * (os.cpu_count() or 1), total=10_000)),
                     ("C1_gls", lambda: make_gls_c1(10_000)),
                     ("C5_gls", lambda: make_gls_c5(10_000_000))]
        run_reference(args, wl, per_gpu, world, scaling, extra)
        return

    env = Env(args.gather)
    wl = build(args.workload, total_units)
    big = args.workload in ("gls_c5", "gls_c5_full", "gls_c4_full")
    steps = min(args.steps, 3) if big else args.steps
    rec, ok = run_workload(env, wl, per_gpu if not strong else -(-per_gpu // world), scaling, steps, args.warmup,
                           want_cpu=not args.no_cpu_baseline,
                           # the reference algorithm's peak index is asserted on the NAMED configs only.  The weak-scaled
                           # primary grid at N > 1 (N x 1e5 frequencies) is not one: it runs N x past the Kepler cadence's
                           # pseudo-Nyquist, where alias peaks of nearly equal height exist and the reference's own
                           # approximation error (up to 3 % of the peak there) decides which of them it reports --
                           # measured at N = 4, 8 in round 2: exact oracle and GPU agree on bin 31370, the FFT
                           # extirpolation picks an alias
                           check_ref_peak=args.workload not in ("gls_c5", "gls_c5_full") and (world == 1 or strong))
    all_ok = ok
    if default_run:
        cfgs = {}
        plan = [("C3_pdm", lambda: make_pdm_c3(100_000), 100_000, args.steps, True),
                ("C5_gls", lambda: make_gls_c5(10_000_000), 10_000_000, 2, False),
                ("C4_gls_batch", lambda: build("gls_c4_full", 10_000), 10_000, 2, False)]
        if world == 1:
            plan.append(("C1_gls", lambda: make_gls_c1(10_000), 10_000, args.steps, True))
        for name, mk, units, k, cpu_on in plan:
            try:
                w2 = mk()
                r2, ok2 = run_workload(env, w2, -(-units // world), "strong" if world > 1 else "weak", k, 3,
                                       cpu_budget_s=6.0, want_cpu=cpu_on and not args.no_cpu_baseline,
                                       check_ref_peak=(name != "C5_gls"))
                all_ok = all_ok and ok2
                if r2 is not None:
                    for drop in ("higher_is_better", "vs_baseline", "data", "unit"):
                        r2.pop(drop, None)
                    cfgs[name] = r2
                del w2
                env.torch.cuda.empty_cache()
            except Exception as exc:  # noqa: BLE001 -- a failed extra config must not cost the primary line
                import traceback
                traceback.print_exc(file=sys.stderr)
                if rank == 0:
                    cfgs[name] = {"error": repr(exc)}
                all_ok = False
        if rec is not None:
            rec["configs"] = cfgs
    if rank == 0:
        emit(rec)
    if world > 1:
        env.dist.destroy_process_group()
    if not all_ok:
        print("[bench] PARITY CHECK FAILED (see the `parity` objects of the printed line)", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
