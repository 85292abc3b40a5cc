#!/bin/bash
# round 2, GPU call 3: tests, bench, A/B knobs (entropy row padding, FMA vs DMMA lag products), ncu of k_basic (tensor pipe) + entropy + moments
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_tests3.log
tail -4 gpurun_out/r2_tests3.log
timeout 900 python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
tail -c 600 gpurun_out/r2_bench3.err
Q="--no-configs --no-e2e --no-cpu-baseline --steps 3 --warmup 2"
TSFX_ENTROPY_PAD=1 timeout 300 python bench.py $Q > gpurun_out/r2_ab_pad1.json 2>/dev/null
TSFX_LAG=fma timeout 300 python bench.py $Q > gpurun_out/r2_ab_lagfma.json 2>/dev/null
timeout 300 python bench.py $Q > gpurun_out/r2_ab_default.json 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_entropy_rank|k_basic|k_seq" -c 3 -o gpurun_out/r2_kernels3 -f \
    python bench.py --steps 1 --warmup 0 --series 200000 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_moments" -c 1 -o gpurun_out/r2_moments3 -f \
    python bench.py --steps 1 --warmup 0 --settings minimal --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu3m.log 2>&1
ls -la gpurun_out | tail -4
