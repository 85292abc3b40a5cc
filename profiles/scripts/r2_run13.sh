#!/bin/bash
# round 2, GPU call 13: feature-selection tests (regression targets added)
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_selection.py -q --timeout 120 2>&1 | tail -30 > gpurun_out/r2_tests13.log
tail -8 gpurun_out/r2_tests13.log
