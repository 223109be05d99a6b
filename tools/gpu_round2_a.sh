#!/bin/bash
# GPU-box run A of round 2: full GPU test suite, the default bench line (with the configs block), the reference arm,
# and the PDM feed variants (gpurun_variants/*.so).
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $OUT/a_smi.txt 2>&1
nproc > $OUT/a_nproc.txt
python -m pytest tests -m gpu -x -q --durations=12 > $OUT/a_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 4 $OUT/a_pytest_gpu.log
python bench.py > $OUT/a_bench_default.json 2> $OUT/a_bench_default.err; echo "bench default rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/a_bench_reference.json 2> $OUT/a_bench_reference.err; echo "bench reference rc=$?"
bash tools/exp_pdm_variants.sh > $OUT/a_pdm_variants.txt 2>&1
cat $OUT/a_pdm_variants.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/a_bench_default.json").read().strip().splitlines()[-1])
print("C2 value %.4g ms %.3f kernel_ms %.4f e2e %.4g parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["parity"]))
for k, v in d.get("configs", {}).items():
    if "error" in v: print(k, v); continue
    print(k, "value %.4g ms %.3f kernel_ms %.4f frac %.3f e2e %.4g (%.3f ms) parity %s" % (v["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v["parity"]))
PY
