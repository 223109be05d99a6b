#!/bin/bash
# GPU-box run E of round 2 (2 GPUs): whole GPU suite incl. the torchrun p2p test and the multi-device ctx on two real
# devices, bench at N = 2 (default line with the strong-scaled configs block) and N = 1.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/e_smi.txt
python -m pytest tests -m gpu -x -q --durations=8 > $OUT/e_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 12 $OUT/e_pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/e_bench_2gpu.json 2> $OUT/e_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -n 5 $OUT/e_bench_2gpu.err
python bench.py > $OUT/e_bench_1gpu.json 2> $OUT/e_bench_1gpu.err; echo "bench 1gpu rc=$?"
python tools/multi_ctx_bench.py > $OUT/e_multi_ctx.json 2> $OUT/e_multi_ctx.err; echo "multi ctx bench rc=$?"; cat $OUT/e_multi_ctx.json
python - <<'PY'
import json
for f in ("e_bench_1gpu", "e_bench_2gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "C2 value %.4g ms %.4f kernel_ms %.4f e2e %.4g (%.4f ms) launches %d gather %s parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["gather_used"], d["parity"]))
    for k, v in d.get("configs", {}).items():
        if "error" in v: print("  ", k, v); continue
        print("  ", k, "value %.4g ms %.4f kernel_ms %.4f frac %.3f e2e %.4g (%.4f ms) parity %s %.2e" % (v["value"], v["ms_per_step"], v["roofline"]["kernel_ms"], v["roofline"]["frac"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v["parity"]["ok"], v["parity"]["max_rel"]))
PY
