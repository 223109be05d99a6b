#!/bin/bash
# GPU-box check on a multi-GPU box (gpurun --gpus N): the whole GPU suite (the torchrun p2p check and the multi-device
# ctx on real device lists run only here), bench.py at N and N/2 ranks (default line: weak C2 + strong C3 / C5 / C4),
# and the multi-device ctx timed in one process.    usage: bash tools/multi_gpu_check.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
nvidia-smi -L > $OUT/mg_smi_$TAG.txt
python -m pytest tests -m gpu -q --durations=5 > $OUT/mg_pytest_gpu_$TAG.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 8 $OUT/mg_pytest_gpu_$TAG.log
for N in $NG $((NG / 2)); do
  [ $N -lt 2 ] && continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > $OUT/mg_bench_${N}gpu_$TAG.json 2> $OUT/mg_bench_${N}gpu_$TAG.err; echo "bench ${N}gpu rc=$?"; tail -n 3 $OUT/mg_bench_${N}gpu_$TAG.err
done
python tools/multi_ctx_bench.py --full > $OUT/mg_multi_ctx_$TAG.json 2> $OUT/mg_multi_ctx_$TAG.err; echo "multi ctx bench rc=$?"; cat $OUT/mg_multi_ctx_$TAG.json
