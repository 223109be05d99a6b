#!/bin/bash
# GPU-box run H of round 2 (8 GPUs): the whole GPU suite (torchrun p2p check on 8 ranks, multi-device ctx on 8 real
# devices), bench at N = 8 and N = 4 (default line: weak C2 + strong C3 / C5 / C4), multi-device ctx timing in one process.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/h_smi.txt
python -m pytest tests -m gpu -q --durations=5 > $OUT/h_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 8 $OUT/h_pytest_gpu.log
for N in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > $OUT/h_bench_${N}gpu.json 2> $OUT/h_bench_${N}gpu.err; echo "bench ${N}gpu rc=$?"; tail -n 3 $OUT/h_bench_${N}gpu.err
done
python tools/multi_ctx_bench.py --full > $OUT/h_multi_ctx.json 2> $OUT/h_multi_ctx.err; echo "multi ctx bench rc=$?"; cat $OUT/h_multi_ctx.json
python - <<'PY'
import json
for f in ("h_bench_8gpu", "h_bench_4gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "C2 value %.4g ms %.4f kernel_ms %.4f e2e %.4g (%.4f ms) gather %s parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gather_used"], d["parity"]["ok"]))
    for k, v in d.get("configs", {}).items():
        if "error" in v: print("  ", k, v); continue
        print("  ", k, "value %.4g ms %.4f kernel_ms %.4f frac %.3f e2e %.4g (%.4f ms) parity %s %.2e" % (v["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v["parity"]["ok"], v["parity"]["max_rel"]))
PY
