#!/bin/bash
# GPU-box experiment: pdm_hist_kernel with 1 or 2 trial periods per thread (PDC_PDM_PPT) and 4 or 8 samples per trip.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_pdm_gpu.py -x -q > $OUT/x_pdm_tests_default.log 2>&1; echo "tests default rc=$?"
PDC_PDM_PPT=2 python -m pytest tests/test_pdm_gpu.py -x -q > $OUT/x_pdm_tests_ppt2.log 2>&1; echo "tests ppt2 rc=$?"
PDC_PDM_PPT=1 python -m pytest tests/test_pdm_gpu.py -x -q > $OUT/x_pdm_tests_ppt1.log 2>&1; echo "tests ppt1 rc=$?"
for v in 1 2; do
  PDC_PDM_PPT=$v python bench.py --workload pdm_c3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/x_bench_c3_ppt$v.json 2> $OUT/x_bench_c3_ppt$v.err
done
for v in 1 2; do
  PERIODICITY_B200_LIB=$PWD/gpurun_variants/lib_u4.so PDC_PDM_PPT=$v python bench.py --workload pdm_c3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/x_bench_c3_ppt${v}_u4.json 2> $OUT/x_bench_c3_ppt${v}_u4.err
done
python bench.py --workload gls_c4_full --steps 3 --warmup 3 --no-cpu-baseline > $OUT/x_bench_c4_full.json 2> $OUT/x_bench_c4_full.err
tail -n 3 $OUT/x_pdm_tests_*.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/x_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'e2e %.4g' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
