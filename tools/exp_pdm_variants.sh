#!/bin/bash
# GPU-box experiment: build-time variants of the packed loop of pdm_hist_kernel (gpurun_variants/*.so) against the in-tree build.
OUT=gpurun_out; mkdir -p $OUT
python bench.py --workload pdm_c3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/y_bench_c3_main.json 2> $OUT/y_bench_c3_main.err
for lib in gpurun_variants/*.so; do
  tag=$(basename $lib .so)
  [ -n "$VARIANT_TESTS" ] && { PERIODICITY_B200_LIB=$PWD/$lib python -m pytest tests/test_pdm_gpu.py -x -q > $OUT/y_pdm_tests_$tag.log 2>&1; echo "tests $tag rc=$?"; }
  PERIODICITY_B200_LIB=$PWD/$lib python bench.py --workload pdm_c3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/y_bench_c3_$tag.json 2> $OUT/y_bench_c3_$tag.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/y_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'e2e %.4g' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
