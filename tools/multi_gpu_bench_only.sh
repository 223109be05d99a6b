#!/bin/bash
# Multi-GPU box, benches only (the GPU suite ran on the 2-GPU box): bench.py at all ranks (default line: weak C2 + strong
# C3 / C5 / C4 with the parity check of the gathered result) and the multi-device ctx in one process.
#   usage (gpurun --gpus N): bash tools/multi_gpu_bench_only.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $NG --steps 20 --warmup 3 > $OUT/mg_bench_${NG}gpu_$TAG.json 2> $OUT/mg_bench_${NG}gpu_$TAG.err; echo "bench ${NG}gpu rc=$?"; tail -n 2 $OUT/mg_bench_${NG}gpu_$TAG.err
python tools/multi_ctx_bench.py --full > $OUT/mg_multi_ctx_${NG}gpu_$TAG.json 2> $OUT/mg_multi_ctx_${NG}gpu_$TAG.err; echo "multi ctx bench rc=$?"; cat $OUT/mg_multi_ctx_${NG}gpu_$TAG.json
