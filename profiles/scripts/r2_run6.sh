#!/bin/bash
# round 2, GPU call 6: GPU test suite, bench line, A/B of the SORTED CTA width and of the compact SEQ kernel
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 2>&1 | tail -40 > gpurun_out/r2_tests6.log
tail -6 gpurun_out/r2_tests6.log
Q="--no-configs --no-e2e --no-cpu-baseline --steps 3 --warmup 2"
timeout 120 python bench.py $Q > gpurun_out/r2_ab_default6.json 2>gpurun_out/r2_ab_default6.err
TSFX_SORTED_WPC=12 timeout 120 python bench.py $Q > gpurun_out/r2_ab_sorted12.json 2>/dev/null
TSFX_SEQ=general timeout 120 python bench.py $Q > gpurun_out/r2_ab_seqgeneral.json 2>/dev/null
timeout 400 python bench.py > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
tail -c 400 gpurun_out/r2_bench6.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_entropy_rank" -c 1 -o gpurun_out/r2_entropy1024 -f \
    python bench.py --steps 1 --warmup 0 --series 20000 --len 1024 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu6.log 2>&1
ls -la gpurun_out | tail -4
