#!/bin/bash
# round 2, GPU call 5: full GPU test suite (per-test timeout), A/B of the SORTED CTA width, ncu of the len-1024 ENTROPY path
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -40 > gpurun_out/r2_tests5.log
tail -6 gpurun_out/r2_tests5.log
Q="--no-configs --no-e2e --no-cpu-baseline --steps 3 --warmup 2"
TSFX_SORTED_WPC=12 timeout 300 python bench.py $Q > gpurun_out/r2_ab_sorted12.json 2>/dev/null
timeout 300 python bench.py $Q > gpurun_out/r2_ab_default5.json 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_entropy_rank" -c 1 -o gpurun_out/r2_entropy1024 -f \
    python bench.py --steps 1 --warmup 0 --series 20000 --len 1024 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu5.log 2>&1
ls -la gpurun_out | tail -4
