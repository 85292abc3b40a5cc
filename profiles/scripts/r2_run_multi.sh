#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): NCCL / symmetric-memory tests of the product, then the bench line at N GPUs
# for every placement mode (auto = multicast if the box has it, copy engines, P2P stores, NCCL all-gather)
set -x
N=${1:-2}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m | head -12
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -25 > gpurun_out/r2_multi_tests_$N.log
tail -5 gpurun_out/r2_multi_tests_$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
NCCL_DEBUG=WARN timeout 900 $TR bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 600 gpurun_out/r2_bench_${N}gpu.err
head -c 400 gpurun_out/r2_bench_${N}gpu.json
Q="--no-configs --no-e2e --no-cpu-baseline --steps 4 --warmup 2"
for mode in ${MODES:-copy store nccl multicast}; do
  timeout 300 $TR bench.py --gpus $N $Q --placement $mode > gpurun_out/r2_place_${mode}_${N}gpu.json 2> gpurun_out/r2_place_${mode}_${N}gpu.err
  head -c 200 gpurun_out/r2_place_${mode}_${N}gpu.json; tail -c 300 gpurun_out/r2_place_${mode}_${N}gpu.err
done
ls -la gpurun_out | tail -8
