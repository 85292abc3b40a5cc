#!/bin/bash
# One GPU-box call that produces the single-GPU artefacts committed under profiles/ for round 2 (run from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/scripts/r2_final.sh'
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
Q="--no-configs --no-e2e --no-cpu-baseline"
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active
# 1. per-kernel DRAM bytes / warp instructions of one step at 1 M x 256 -> profiles/traffic_r2.json (bench.py quotes it)
timeout 600 ncu --metrics $M --clock-control none -c 12 --csv --log-file $O/kernel_metrics_r2.csv \
    python bench.py --steps 1 --warmup 0 $Q > $O/kernel_metrics_r2.log 2>&1
python profiles/scripts/kernel_metrics_to_json.py $O/kernel_metrics_r2.csv profiles/traffic_r2.json > $O/traffic_r2.log 2>&1
cp profiles/traffic_r2.json $O/traffic_r2.json
timeout 300 ncu --metrics $M --clock-control none -k regex:k_moments -c 1 --csv --log-file $O/kernel_metrics_r2_minimal.csv \
    python bench.py --steps 1 --warmup 0 --settings minimal $Q >> $O/kernel_metrics_r2.log 2>&1
# 2. tests and the bench lines
timeout 700 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -8 > $O/gpu_tests_r2.log
cat $O/gpu_tests_r2.log
timeout 600 python bench.py > $O/bench_r2.json 2> $O/bench_r2.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference.json 2> $O/bench_r2_reference.err
# 3. launch list (share of the step per kernel) and full captures of the changed kernels + the three untouched ones
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 $Q > $O/launches_r2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_basic|k_entropy_rank|k_seq_small|k_la|k_sorted|k_spectral|k_peaks" -c 7 -f -o $O/r2_final_kernels \
    python bench.py --steps 1 --warmup 0 --series 200000 $Q > $O/ncu_final.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_moments -c 1 -f -o $O/r2_final_moments \
    python bench.py --steps 1 --warmup 0 --settings minimal $Q >> $O/ncu_final.log 2>&1
# text summaries on the box (the reports themselves exceed the 64 MiB that travel back)
python profiles/scripts/ncu_summary.py $O/r2_final_kernels.ncu-rep "k_" 14 > $O/ncu_r2_summary.txt 2>&1
python profiles/scripts/ncu_summary.py $O/r2_final_moments.ncu-rep "k_moments" 12 >> $O/ncu_r2_summary.txt 2>&1
ncu -i $O/r2_final_kernels.ncu-rep --page raw --csv --kernel-name regex:k_basic 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for n,x in zip(h,v):
    if 'dmma' in n.lower() or 'pipe_tensor_cycles' in n: print(n, x)
" > $O/ncu_r2_basic_tensor_pipe.txt 2>&1
rm -f $O/*.ncu-rep
du -sh $O
head -c 500 $O/bench_r2.json
