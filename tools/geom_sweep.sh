#!/bin/bash
# Dev aid: sweep strip-kernel geometries (PDC_GLS_GEOM) on the bench workloads.
for g in 9 10 11 12 13 14 15; do
  echo -n "geom $g c2: "
  PDC_GLS_GEOM=$g python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e kernel_ms %.4f step_ms %.4f frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
done
for g in 9 10 12 13 14; do
  echo -n "geom $g c5: "
  PDC_GLS_GEOM=$g python bench.py --no-cpu-baseline --workload gls_c5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.3e kernel_ms %.4f step_ms %.4f frac %.3f' % (d['value'], d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac']), d['clocks'])"
done
