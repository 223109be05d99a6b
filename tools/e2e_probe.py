import numpy as np, time, torch, sys
sys.path.insert(0,'.')
import bench
from periodicity_b200 import _ffi
wl=bench.make_gls_c2(100_000)
ctx=_ffi.default_context(0)
t_pin=torch.from_numpy(wl["t"]).pin_memory(); y_pin=torch.from_numpy(wl["y"]).pin_memory()
th,yh=t_pin.numpy(),y_pin.numpy()
def timeit(f,n=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("host call, power out      ms", timeit(lambda: ctx.gls(th,yh,None,wl["fmin"],wl["df"],wl["nf"])))
print("host call, no power       ms", timeit(lambda: ctx.gls(th,yh,None,wl["fmin"],wl["df"],wl["nf"],want_power=False)))
print("host call, pageable input ms", timeit(lambda: ctx.gls(wl["t"],wl["y"],None,wl["fmin"],wl["df"],wl["nf"])))
k0,c0=ctx.main_kernel_ms_total()
ctx.gls(th,yh,None,wl["fmin"],wl["df"],wl["nf"]); k1,c1=ctx.main_kernel_ms_total(); print("kernel ms",(k1-k0)/(c1-c0))
out=torch.empty(100_000,dtype=torch.float64).pin_memory()
d=torch.empty(100_000,dtype=torch.float64,device="cuda")
print("D2H 0.8MB pinned ms", timeit(lambda: out.copy_(d,non_blocking=False)))
o2=np.empty(100_000)
print("D2H 0.8MB pageable ms", timeit(lambda: torch.from_numpy(o2).copy_(d)))
print("H2D 0.52MB pinned ms", timeit(lambda: d[:65000].copy_(t_pin,non_blocking=False)))
