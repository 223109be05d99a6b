#!/bin/bash
# Runs on the GPU box (under gpurun): ncu evidence of the tensor-core GLS path on the C2 shape (bench.py's default workload).
#   usage: bash tools/umma_profile.sh <tag>
TAG=${1:-r02u}
OUT=gpurun_out; mkdir -p $OUT
M="sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_tensor.sum,sm__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__cycles_elapsed.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum"
# 1. launch list (every launch with its device time; cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_gls_c2_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/up0.log 2>&1
# 2. pipe counters and one full capture of the hot kernel
ncu --metrics $M --clock-control none -k regex:gls_umma2?_kernel -s 3 -c 1 --csv --log-file $OUT/pipes_gls_umma_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/up1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gls_umma2?_kernel -s 3 -c 1 -f -o $OUT/prof_gls_umma_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/up2.log 2>&1
ls -la $OUT | tail -6
