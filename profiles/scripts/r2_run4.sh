#!/bin/bash
# round 2, GPU call 4: tests, bench, A/B BASIC CTA width, ncu of peaks / entropy / moments
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests4.log
tail -6 gpurun_out/r2_tests4.log
timeout 900 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
tail -c 600 gpurun_out/r2_bench4.err
Q="--no-configs --no-e2e --no-cpu-baseline --steps 3 --warmup 2"
TSFX_BASIC_WPC=12 timeout 300 python bench.py $Q > gpurun_out/r2_ab_wpc12.json 2>/dev/null
TSFX_BASIC_WPC=24 timeout 300 python bench.py $Q > gpurun_out/r2_ab_wpc24.json 2>/dev/null
timeout 300 python bench.py $Q > gpurun_out/r2_ab_default4.json 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_entropy_rank|k_peaks" -c 2 -o gpurun_out/r2_kernels4 -f \
    python bench.py --steps 1 --warmup 0 --series 200000 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu4.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_moments" -c 1 -o gpurun_out/r2_moments4 -f \
    python bench.py --steps 1 --warmup 0 --settings minimal --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu4m.log 2>&1
ls -la gpurun_out | tail -4
