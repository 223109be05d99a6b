#!/bin/bash
# GPU-box run F of round 2 (1 GPU): tests of the final tree, bench lines, ncu evidence for profiles/ (tag r02f).
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q --durations=8 > $OUT/f_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 12 $OUT/f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/f_smoke.log
python bench.py > $OUT/f_bench_default.json 2> $OUT/f_bench_default.err; echo "bench default rc=$?"
for wl in pdm_c3_jd ce_c3 gls_multi sl; do
  python bench.py --workload $wl --no-configs > $OUT/f_bench_$wl.json 2> $OUT/f_bench_$wl.err; echo "bench $wl rc=$?"
done
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/f_bench_reference.json 2> $OUT/f_bench_reference.err; echo "bench reference rc=$?"
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $OUT/f_sanitize_memcheck.log 2>&1; tail -n 2 $OUT/f_sanitize_memcheck.log
compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $OUT/f_sanitize_racecheck.log 2>&1; tail -n 2 $OUT/f_sanitize_racecheck.log
bash tools/profile.sh r02f > $OUT/f_profile.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/f_bench_default.json").read().strip().splitlines()[-1])
print("C2 value %.4g ms %.4f kernel_ms %.4f e2e %.4g (%.4f ms) launches %d parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["parity"]["ok"]))
for k, v in d.get("configs", {}).items():
    if "error" in v: print(k, v); continue
    print(k, "value %.4g ms %.4f kernel_ms %.4f frac %.3f e2e %.4g (%.4f ms) launches %d parity %s %.2e" % (v["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v["gpu_launches"], v["parity"]["ok"], v["parity"]["max_rel"]))
for wl in ("pdm_c3_jd", "ce_c3", "gls_multi", "sl"):
    try:
        v = json.loads(open("gpurun_out/f_bench_%s.json" % wl).read().strip().splitlines()[-1])
        print(wl, "value %.4g ms %.4f kernel_ms %.4f frac %.3f e2e %.4g parity %s" % (v["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["parity"]))
    except Exception as e:
        print(wl, "ERR", e)
PY
