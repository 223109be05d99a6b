"""Multi-GPU check of the fused epilogue + all-gather (run under torchrun on >= 2 GPUs)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from periodicity_b200 import _ffi, GLS, TSeries
from periodicity_b200 import dist as pdist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = _ffi.default_context(local)
rng = np.random.default_rng(2)
n, nf = 65_000, 100_000 * world + 7
t = np.sort(rng.uniform(0, 1470, n))
df = 1 / (t[-1] - t[0]) / 5
fmin = 0.5 * df
y = 1000 + np.sin(2 * np.pi * (fmin + 31370.4 * df) * t) + rng.standard_normal(n)
ref, ridx, rval = pdist.gls_sharded(t, y, None, fmin, df, nf, device=local)
p, idx, val = pdist.gls_sharded_p2p(t, y, None, fmin, df, nf, device=local)
ok = np.array_equal(p, ref) and idx == ridx and val == rval
# ... and both against the C oracle (formula with exact sums), strided over the WHOLE gathered grid + round the peak
from oracle import cport
sel = np.unique(np.concatenate([np.arange(0, nf, max(1, nf // 96)), np.arange(max(0, idx - 4), min(nf, idx + 5)), [nf - 1]]))
oref = cport.gls_exact_at(t, y, None, fmin, df, sel)
oracle_gls = float(np.max(np.abs(p[sel] - oref)) / np.max(oref))
ok = ok and oracle_gls <= 1e-5 and int(sel[np.argmax(oref)]) == idx
root_p, ridx2, _ = pdist.gls_sharded_p2p(t, y, None, fmin, df, nf, device=local, root=0, copy=False)   # bench.py's e2e form
ok = ok and ridx2 == idx and ((root_p is None) if rank != 0 else np.array_equal(root_p, ref))
p2, idx2, val2 = pdist.gls_sharded_p2p(t, 2 * y + 1, None, fmin, df, nf, device=local)   # reuse of the symmetric buffer
ok = ok and idx2 == ridx and np.nanmax(np.abs(p2 - ref)) < 2e-6 * rval
ls = GLS(fmin=fmin, fmax=fmin + (nf - 1.5) * df, shard="p2p", device=local)(TSeries(t, y))
ok = ok and np.array_equal(ls.values, ref)
# PDM: period grid sharded, fused gather vs NCCL gather
periods = np.linspace(1.0, 11.0, 20_001)
x = np.sin(2 * np.pi * t / 3.7) + rng.standard_normal(n)
th_ref, ai_ref, av_ref = pdist.pdm_sharded(t, x, periods, 10, 2, device=local)
th, ai, av = pdist.pdm_sharded_p2p(t, x, periods, 10, 2, device=local)
ok = ok and np.array_equal(th, th_ref) and ai == ai_ref and av == av_ref
psel = np.unique(np.concatenate([np.arange(0, periods.size, 211), np.arange(max(0, ai - 3), min(periods.size, ai + 4))]))
pref = cport.pdm(t, x, periods[psel], 10, 2)
oracle_pdm = float(np.max(np.abs(th[psel] - pref) / pref))
ok = ok and oracle_pdm <= 1e-5 and int(psel[np.argmin(pref)]) == ai
# timing: NCCL all-gather path vs fused path, device resident
td, yd = torch.from_numpy(t).to(dev), torch.from_numpy(y).to(dev)
start, stop, L = pdist.shard_bounds(nf, rank, world)
def step_nccl():
    power, arg, mx = pdist.gls_torch(td, yd, None, fmin, df, stop - start, j0=start, ctx=ctx)
    return pdist.all_gather_packed(power, mx.reshape(()), (arg + start).to(torch.float64).reshape(()), L)
def step_p2p():
    return pdist.gls_sharded_p2p_torch(td, yd, None, fmin, df, nf, ctx=ctx)
res = {}
for name, fn in (("nccl", step_nccl), ("p2p", step_p2p)):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[name] = float(ms.item())
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print({"world": world, "parity_all_ranks": bool(flag.item()), "oracle_max_rel": {"gls": oracle_gls, "pdm": oracle_pdm},
           "ms_per_step": res,
           "evals_per_s": {k: n * nf / (v * 1e-3) for k, v in res.items()}})
dist.destroy_process_group()
