#!/bin/bash
# GPU-box run I (2 GPUs): the driver's own commands (pytest -x, bench N=1 / N=2 with rc), CE after the ATOMS.ADD fix.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/i_pytest_gpu.log 2>&1; echo "pytest -m gpu -x rc=$?"; tail -n 3 $OUT/i_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/i_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/i_smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $OUT/i_bench_ref.json 2> $OUT/i_bench_ref.err; echo "bench reference rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/i_bench_1gpu.json 2> $OUT/i_bench_1gpu.err; echo "bench 1gpu rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $OUT/i_bench_ref2.json 2> $OUT/i_bench_ref2.err; echo "bench reference N=2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/i_bench_2gpu.json 2> $OUT/i_bench_2gpu.err; echo "bench 2gpu rc=$?"
python bench.py --workload ce_c3 --no-configs > $OUT/i_bench_ce.json 2> $OUT/i_bench_ce.err; echo "bench ce rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ce_hist -s 3 -c 1 -f -o $OUT/prof_ce_hist_r02i python bench.py --workload ce_c3 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/i_p1.log 2>&1
python - <<'PY'
import json
for f in ("i_bench_1gpu", "i_bench_2gpu", "i_bench_ce", "i_bench_ref", "i_bench_ref2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value %.4g ms %.4f e2e %.4g parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("parity") or {}).get("ok")), "kernel_ms", (d.get("roofline") or {}).get("kernel_ms"))
    for k, v in d.get("configs", {}).items():
        if "error" in v: print("  ", k, v); continue
        print("  ", k, "value %.4g" % v["value"], "ms %s" % v.get("ms_per_step"), "parity", (v.get("parity") or {}).get("ok"))
PY
