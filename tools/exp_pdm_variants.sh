#!/bin/bash
# GPU-box experiment: build-time variants of the packed loop of pdm_hist_kernel
# (gpurun_variants/lib_c<CHAINS>_f<FIXUP>_p<PREFETCH>_i<L2INT>.so) x trial periods per thread (PDC_PDM_PPT).
OUT=gpurun_out; mkdir -p $OUT
rm -f $OUT/x_bench_*.json $OUT/x_bench_*.err $OUT/x_pdm_tests_*
for ppt in 2 4; do
  PDC_PDM_PPT=$ppt PERIODICITY_B200_LIB=$PWD/gpurun_variants/lib_c32_f1_p1_i1.so python -m pytest tests/test_pdm_gpu.py -x -q > $OUT/x_pdm_tests_i1_ppt$ppt.log 2>&1; echo "tests L2INT ppt$ppt rc=$?"
done
for lib in gpurun_variants/*.so; do
  tag=$(basename $lib .so)
  for ppt in 2 4; do
    PDC_PDM_PPT=$ppt PERIODICITY_B200_LIB=$PWD/$lib python bench.py --workload pdm_c3 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/x_bench_c3_${tag}_ppt$ppt.json 2> $OUT/x_bench_c3_${tag}_ppt$ppt.err
  done
done
tail -n 4 $OUT/x_pdm_tests_*.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/x_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.4f' % d['roofline']['kernel_ms'], 'e2e %.4g' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
