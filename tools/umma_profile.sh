#!/bin/bash
# Runs on the GPU box (under gpurun): one full ncu capture + pipe counters of gls_umma_kernel on the C2 shape.
#   usage: bash tools/umma_profile.sh <tag>
TAG=${1:-r02u}
OUT=gpurun_out; mkdir -p $OUT
M="sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed_pipe_fp16.sum,sm__inst_executed_pipe_tensor.sum,sm__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum"
export PDC_GLS_UMMA=1
ncu --metrics $M --clock-control none -k regex:gls_umma -s 2 -c 1 --csv --log-file $OUT/pipes_gls_umma_$TAG.csv python tools/umma_check.py c2 > $OUT/up1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gls_umma -s 2 -c 1 -f -o $OUT/prof_gls_umma_$TAG python tools/umma_check.py c2 > $OUT/up2.log 2>&1
ls -la $OUT | tail -5
