#!/bin/bash
# GPU-box check of the tree as it is: GPU tests, smoke, the bench lines the driver takes, launch lists and ncu evidence.
# usage (under gpurun): bash tools/final_check.sh <tag>
TAG=${1:-r01e}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q --durations=8 > $OUT/final_pytest_gpu_$TAG.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 $OUT/final_pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -n 1 $OUT/final_smoke_$TAG.log
python bench.py > $OUT/bench_c2_$TAG.json 2> $OUT/bench_c2_$TAG.err; echo "bench c2 rc=$?"
python bench.py --workload pdm_c3 > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; echo "bench c3 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_gls_c2_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/p1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_pdm_c3_$TAG.csv python bench.py --workload pdm_c3 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdm_hist -s 3 -c 1 -f -o $OUT/prof_pdm_hist_$TAG python bench.py --workload pdm_c3 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/p5.log 2>&1
M="sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum"
ncu --metrics $M --clock-control none -k regex:pdm_hist -s 3 -c 1 --csv --log-file $OUT/pipes_pdm_hist_$TAG.csv python bench.py --workload pdm_c3 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/p7.log 2>&1
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $OUT/sanitize_memcheck_$TAG.log 2>&1; tail -n 2 $OUT/sanitize_memcheck_$TAG.log
compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $OUT/sanitize_racecheck_$TAG.log 2>&1; tail -n 2 $OUT/sanitize_racecheck_$TAG.log
cat $OUT/bench_c2_$TAG.json $OUT/bench_c3_$TAG.json | cut -c1-400
