for ns in 0 9 17 18 19 27 36 45 54 63; do
  PDC_GLS_NSPLIT=$ns python bench.py --workload gls_c2 --no-configs --no-cpu-baseline --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nsplit', $ns, 'ms', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), d['parity']['ok'])"
done
