#!/bin/bash
# GPU-box check of the tree as it is (1 GPU), with the driver's own commands: GPU tests (-x), smoke, both bench arms,
# the single-workload bench lines kept under profiles/, sanitizer.      usage (under gpurun): bash tools/final_check.sh <tag>
# (ncu evidence: tools/profile.sh <tag>; multi-GPU: tools/multi_gpu_check.sh <tag>)
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q --durations=8 > $OUT/final_pytest_gpu_$TAG.log 2>&1; echo "pytest -m gpu -x rc=$?"; tail -n 3 $OUT/final_pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/final_smoke_$TAG.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $OUT/final_bench_reference_$TAG.json 2> $OUT/final_bench_reference_$TAG.err; echo "bench reference rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/final_bench_default_$TAG.json 2> $OUT/final_bench_default_$TAG.err; echo "bench default rc=$?"
for wl in pdm_c3_jd ce_c3 gls_multi sl sl_long; do
  python bench.py --workload $wl --no-configs > $OUT/final_bench_${wl}_$TAG.json 2> $OUT/final_bench_${wl}_$TAG.err; echo "bench $wl rc=$?"
done
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $OUT/final_sanitize_memcheck_$TAG.log 2>&1; tail -n 2 $OUT/final_sanitize_memcheck_$TAG.log
compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $OUT/final_sanitize_racecheck_$TAG.log 2>&1; tail -n 2 $OUT/final_sanitize_racecheck_$TAG.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/final_bench_*_$TAG.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    r = d.get("roofline") or {}
    print(f.split("final_bench_")[1], "value %.4g ms %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"), "parity", (d.get("parity") or {}).get("ok"))
    for k, v in d.get("configs", {}).items():
        if "error" in v: print("   ", k, v); continue
        print("   ", k, "value %.4g" % v["value"], "ms", v.get("ms_per_step"), "parity", (v.get("parity") or {}).get("ok"))
PY
