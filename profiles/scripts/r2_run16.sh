#!/bin/bash
# round 2, closing GPU call (about 40 s of run time left): the stage-(a) tests on the final tree
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_stage_a.py -q -x 2>&1 | tail -5 > gpurun_out/r2_tests16.log
cat gpurun_out/r2_tests16.log
