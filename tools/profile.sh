#!/bin/bash
# Runs on the GPU box (under gpurun): ncu evidence for profiles/.  Usage: tools/profile.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
M="sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum"
# 1. launch lists (every launch with its device time; cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_gls_c2_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_pdm_c3_$TAG.csv python bench.py --workload pdm_c3 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_gls_c4_$TAG.csv python bench.py --workload gls_c4 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p3.log 2>&1
# 2. full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:gls_strip -s 3 -c 1 -f -o $OUT/prof_gls_strip_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdm_hist -s 3 -c 1 -f -o $OUT/prof_pdm_hist_$TAG python bench.py --workload pdm_c3 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p5.log 2>&1
# 3. the pipe counters SURVEY.md section 8d names
ncu --metrics $M --clock-control none -k regex:gls_strip -s 3 -c 1 --csv --log-file $OUT/pipes_gls_strip_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p6.log 2>&1
ncu --metrics $M --clock-control none -k regex:pdm_hist -s 3 -c 1 --csv --log-file $OUT/pipes_pdm_hist_$TAG.csv python bench.py --workload pdm_c3 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p7.log 2>&1
ncu --metrics $M --clock-control none -k regex:gls_strip -s 3 -c 1 --csv --log-file $OUT/pipes_gls_strip_c4_$TAG.csv python bench.py --workload gls_c4 --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $OUT/p8.log 2>&1
ls -la $OUT
