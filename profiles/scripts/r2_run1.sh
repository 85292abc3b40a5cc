#!/bin/bash
# round 2, GPU call 1: full GPU test suite, default bench line, ncu launch list + full capture of the new ENTROPY kernel
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_tests1.log
tail -5 gpurun_out/r2_tests1.log
timeout 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 1500 gpurun_out/r2_bench1.err
head -c 600 gpurun_out/r2_bench1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches1.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_launches1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_entropy_rank -c 1 -o gpurun_out/r2_entropy_rank -f \
    python bench.py --steps 1 --warmup 1 --series 200000 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_entropy.log 2>&1
ls -la gpurun_out | tail -8
