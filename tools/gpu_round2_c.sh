#!/bin/bash
# GPU-box run C of round 2: fused-tail kernels (3 launches per GLS call, 2 per PDM call), multi-device ctx.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q --durations=8 > $OUT/c_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 12 $OUT/c_pytest_gpu.log
python bench.py > $OUT/c_bench_default.json 2> $OUT/c_bench_default.err; echo "bench default rc=$?"; tail -n 3 $OUT/c_bench_default.err
python bench.py --workload ce_c3 --no-configs > $OUT/c_bench_ce.json 2> $OUT/c_bench_ce.err; echo "bench ce rc=$?"; cut -c1-400 $OUT/c_bench_ce.json
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $OUT/c_sanitize_memcheck.log 2>&1; tail -n 3 $OUT/c_sanitize_memcheck.log
compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $OUT/c_sanitize_racecheck.log 2>&1; tail -n 3 $OUT/c_sanitize_racecheck.log
for wl in gls_c2 pdm_c3 gls_c1; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${wl}_r02c.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/c_p_$wl.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:gls_strip -s 3 -c 1 -f -o $OUT/prof_gls_strip_r02c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/c_p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdm_hist -s 3 -c 1 -f -o $OUT/prof_pdm_hist_r02c python bench.py --workload pdm_c3 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/c_p5.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c_bench_default.json").read().strip().splitlines()[-1])
print("C2 value %.4g ms %.4f kernel_ms %.4f e2e %.4g (%.4f ms) launches %d parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["parity"]["ok"]))
for k, v in d.get("configs", {}).items():
    if "error" in v: print(k, v); continue
    print(k, "value %.4g ms %.4f kernel_ms %.4f frac %.3f e2e %.4g (%.4f ms) launches %d parity %s %.2e" % (v["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v["gpu_launches"], v["parity"]["ok"], v["parity"]["max_rel"]))
PY
