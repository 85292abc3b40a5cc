#!/bin/bash
# One GPU-box call that produces every artefact committed under profiles/ for the round (run from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/scripts/final_run.sh'
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/final_tests.log
python bench.py > $O/bench_final.json 2> $O/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --series 250000 --len 1024 --no-cpu-baseline > $O/bench_len1024.json 2> $O/bench_len1024.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_basic -c 1 -f -o $O/r1_final_basic \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_basic.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_entropy -c 1 -f -o $O/r1_final_entropy \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $O/ncu_entropy.log 2>&1
cat $O/final_tests.log
head -c 600 $O/bench_final.json
