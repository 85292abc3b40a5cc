#!/bin/bash
# round 2, GPU call 2: GPU tests, bench line, ncu (full) of entropy / seq / moments at 200 k series
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_tests2.log
tail -5 gpurun_out/r2_tests2.log
timeout 900 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -c 800 gpurun_out/r2_bench2.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_entropy_rank|k_seq|k_basic|k_peaks" -c 4 -o gpurun_out/r2_kernels2 -f \
    python bench.py --steps 1 --warmup 0 --series 200000 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_moments" -c 1 -o gpurun_out/r2_moments2 -f \
    python bench.py --steps 1 --warmup 0 --settings minimal --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu2m.log 2>&1
ls -la gpurun_out | tail -6
