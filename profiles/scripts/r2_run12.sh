#!/bin/bash
# round 2, GPU call 12: re-run of the tests and the bench line after the k_moments accumulator change
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -8 > $O/gpu_tests_r2.log
cat $O/gpu_tests_r2.log
timeout 600 python bench.py > $O/bench_r2.json 2> $O/bench_r2.err
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_moments -c 1 --csv --log-file $O/kernel_metrics_r2_minimal.csv \
    python bench.py --steps 1 --warmup 0 --settings minimal --no-configs --no-e2e --no-cpu-baseline > $O/kernel_metrics_r2_minimal.log 2>&1
head -c 300 $O/bench_r2.json
