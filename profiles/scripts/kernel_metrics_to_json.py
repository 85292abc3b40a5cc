#!/usr/bin/env python
"""profiles/kernel_metrics_rN.csv (ncu --metrics ... --csv of one bench.py step) -> profiles/traffic_rN.json,
the per-kernel figures bench.py quotes next to its live timings (dram bytes, warp instructions, pipe utilisation).

    python profiles/scripts/kernel_metrics_to_json.py [profiles/kernel_metrics_r2.csv [profiles/traffic_r2.json]]
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "kernel_metrics_r2.csv")
dst = sys.argv[2] if len(sys.argv) > 2 else src.replace("kernel_metrics", "traffic").replace(".csv", ".json")
rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
ix = {h: i for i, h in enumerate(rows[0])}
per = {}
ALIAS = {"k_entropy_rank": "k_entropy", "k_moments_dense": "k_moments", "k_seq_small": "k_seq"}      # one bench group per kernel family
for r in rows[1:]:
    if len(r) < len(rows[0]):
        continue
    m = re.search(r"(k_[a-z_]+)", r[ix["Kernel Name"]])
    if not m:
        continue
    key = (r[ix["ID"]], m.group(1))
    per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
out = {"_comment": "per launch, from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,"
                   "smsp__issue_active...,sm__inst_executed_pipe_fp64...,sm__warps_active...,gpu__time_duration.sum "
                   "--clock-control none python bench.py --steps 1 --warmup 0 --no-configs --no-e2e` (1 M series x 256, "
                   "ComprehensiveFCParameters); source: " + os.path.relpath(src, ROOT),
       "workload": {"series": 1000000, "len": 256, "settings": "comprehensive"}}
for (_id, name), v in per.items():
    kernel = name
    name = ALIAS.get(name, name)
    if name in out:
        continue                      # first launch of every kernel
    out[name] = {"kernel": kernel, "traffic_bytes": int(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]),
                 "dram_read_bytes": int(v["dram__bytes_read.sum"]), "dram_write_bytes": int(v["dram__bytes_write.sum"]),
                 "inst_executed": int(v["smsp__inst_executed.sum"]),
                 "issue_active_pct": v["smsp__issue_active.avg.pct_of_peak_sustained_active"],
                 "fp64_pipe_active_pct": v["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"],
                 "warps_active_pct": v["sm__warps_active.avg.pct_of_peak_sustained_active"],
                 "ncu_duration_ms": v["gpu__time_duration.sum"] / 1e6}
    for extra in ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",):
        if extra in v:
            out[name]["dmma_pipe_pct"] = v[extra]
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {a: v[a] for a in ("traffic_bytes", "inst_executed")}) for k, v in out.items() if k.startswith("k_")}, indent=1))
