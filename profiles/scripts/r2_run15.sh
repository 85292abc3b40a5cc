#!/bin/bash
# round 2, last GPU call: the GPU test suite and the driver's smoke() on the final tree
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -6 > gpurun_out/gpu_tests_r2.log
cat gpurun_out/gpu_tests_r2.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1
cat gpurun_out/smoke_r2.log
