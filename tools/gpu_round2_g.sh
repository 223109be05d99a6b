#!/bin/bash
# GPU-box run G (1 GPU): tests + PDM timing after the plain / shifted split of the packed loop.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -q --durations=5 > $OUT/g_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 9 $OUT/g_pytest_gpu.log
for wl in pdm_c3 pdm_c3_jd ce_c3; do
  python bench.py --workload $wl --no-configs --no-cpu-baseline > $OUT/g_bench_$wl.json 2> $OUT/g_bench_$wl.err; echo "bench $wl rc=$?"
done
python - <<'PY'
import json
for wl in ("pdm_c3", "pdm_c3_jd", "ce_c3"):
    v = json.loads(open("gpurun_out/g_bench_%s.json" % wl).read().strip().splitlines()[-1])
    print(wl, "value %.4g ms %.4f kernel_ms %.4f frac %.3f e2e %.4g parity %s" % (v["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["parity"]["ok"]))
PY
