#!/bin/bash
# GPU-box check of the tensor-core GLS path: its own tests, smoke, the default bench line and the strip-kernel line.
#   usage (under gpurun): bash tools/umma_gpu_check.sh <tag>
TAG=${1:-r02u}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gls_umma_gpu.py tests/test_gls_gpu.py -m gpu -x -q --durations=5 > $OUT/pytest_gls_$TAG.log 2>&1; echo "pytest gls rc=$?"; tail -n 4 $OUT/pytest_gls_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/smoke_$TAG.log
python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/bench_default_$TAG.json 2> $OUT/bench_default_$TAG.err; echo "bench default rc=$?"
PDC_GLS_UMMA=0 python bench.py --gpus 1 --steps 20 --warmup 3 --no-configs --no-cpu-baseline > $OUT/bench_c2_strip_$TAG.json 2> $OUT/bench_c2_strip_$TAG.err; echo "bench strip rc=$?"
python - <<PY
import json
for name in ("bench_default_$TAG", "bench_c2_strip_$TAG"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % name).read().strip().splitlines()[-1])
    except Exception as e:
        print(name, "ERR", e); continue
    r = d.get("roofline") or {}
    print(name, "value %.4g ms %.4f e2e %.4g (%.4f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), "kernel", r.get("kernel"), "kernel_ms", r.get("kernel_ms"), "bound", r.get("bound"), "frac", r.get("frac"), "executed", (r.get("executed") or {}).get("frac_of_peak"), "parity", d.get("parity"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    for k, v in d.get("configs", {}).items():
        if "error" in v: print("   ", k, v); continue
        rr = v.get("roofline") or {}
        print("   ", k, "value %.4g" % v["value"], "ms", v.get("ms_per_step"), "e2e ms", v["e2e"].get("ms_per_step"), "kernel", rr.get("kernel"), rr.get("kernel_ms"), "parity", (v.get("parity") or {}).get("ok"), (v.get("parity") or {}).get("max_rel"))
PY
