#!/bin/bash
# GPU-box check of the tree as it is: GPU tests, smoke, the bench lines the driver takes, sanitizer.
# usage (under gpurun): bash tools/final_check.sh <tag>      (ncu evidence: tools/profile.sh <tag>)
TAG=${1:-r01f}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q --durations=8 > $OUT/final_pytest_gpu_$TAG.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 $OUT/final_pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/final_smoke_$TAG.log
python bench.py --workload gls_c4 --no-cpu-baseline > $OUT/bench_c4_$TAG.json 2> $OUT/bench_c4_$TAG.err; echo "bench c4 rc=$?"
python bench.py > $OUT/bench_c2_$TAG.json 2> $OUT/bench_c2_$TAG.err; echo "bench c2 rc=$?"
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $OUT/sanitize_memcheck_$TAG.log 2>&1; tail -n 2 $OUT/sanitize_memcheck_$TAG.log
compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $OUT/sanitize_racecheck_$TAG.log 2>&1; tail -n 2 $OUT/sanitize_racecheck_$TAG.log
cat $OUT/bench_c2_$TAG.json $OUT/bench_c4_$TAG.json | cut -c1-300
python - <<PY
import json
for w in ("c2", "c4"):
    d = json.loads(open("gpurun_out/bench_%s_$TAG.json" % w).read().strip().splitlines()[-1])
    print(w, "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.4g (%.3f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
