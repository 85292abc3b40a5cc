#!/usr/bin/env python
"""bench.py -- series/sec of the feature-extraction hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--series S] [--len L]
                    [--settings comprehensive|efficient|minimal]

A "step" is one pass of the hot path (ComprehensiveFCParameters, 783 columns) over one batch of S synthetic
series of length L per GPU (default 1 000 000 x 256 = BASELINE.json configs[2], the configuration the
metric and the north-star target are quoted on).  Weak scaling: every rank owns S series (ids sharded
contiguously, no data-path collective inside the kernels); with N > 1 the step ends with the north star's
single all-gather of the [N*S x F] matrix, issued per row block on a side stream so it overlaps the kernels.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM, CUDA-event
timed, max over ranks); `e2e` = the same pass through the C-ABI host entry point (tsfx_extract_dense with
pinned HOST buffers: H2D of the values and D2H of the feature matrix inside the timed region).
`--impl reference` times the CPU path (the oracle port of the reference, all host cores) instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "series/sec extract_features ComprehensiveFCParameters"


def settings_by_name(name):
    from tsfresh_b200.settings import ComprehensiveFCParameters, EfficientFCParameters, MinimalFCParameters
    return {"comprehensive": ComprehensiveFCParameters, "efficient": EfficientFCParameters,
            "minimal": MinimalFCParameters}[name]()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    seed, count, length, name = args
    from oracle.extract import oracle_rows
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((count, length)).astype(np.float32).astype(np.float64)
    t0 = time.perf_counter()
    m = oracle_rows(list(x), settings_by_name(name))
    return m.shape, time.perf_counter() - t0


def usable_cores():
    """host cores this container may actually use: min(affinity mask, cgroup cpu.max quota)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(round(int(quota) / int(period)))))
    except Exception:
        pass
    return max(1, n)


def cpu_baseline(length, name, target_seconds, cores=None):
    """series/s of the CPU path (oracle port of the reference's per-series loop) on all host cores."""
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"          # the reference's own advice, docs/text/tsfresh_on_a_cluster.rst:225-231
    cores = cores or usable_cores()
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        # warm the workers (imports), calibrate on two series per core using the in-worker time, then size the
        # sample for ~target_seconds of wall clock
        pool.map(_cpu_worker, [(900 + i, 1, length, name) for i in range(cores)])
        res = pool.map(_cpu_worker, [(1000 + i, 2, length, name) for i in range(cores)])
        per_series = max(float(np.median([r[1] for r in res])) / 2.0, 1e-4)
        per_core = max(2, int(target_seconds / per_series))
        per_core = min(per_core, 2000)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(2000 + i, per_core, length, name) for i in range(cores)])
        wall = time.perf_counter() - t0
    n = per_core * cores
    return {"value": n / wall, "unit": "series/s", "cores": cores, "kind": "port",
            "note": "oracle port of the reference loop; measured in the build container at 0.90-0.98x the unmodified "
                    "reference's time per series (DESIGN.md section 5)",
            "sample": "%d series x len %d (%d per worker process, %d processes), wall %.2f s" % (n, length, per_core, cores, wall)}


def run_reference(args):
    """--impl reference: the CPU path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(args.len, args.settings, target_seconds=max(2.0, 60.0 / max(1, args.steps + args.warmup)))
        if i >= args.warmup:
            per_step.append(cb["value"])
    v = float(np.mean(per_step)) if per_step else cb["value"]
    cb = dict(cb, value=v)
    n_ref = cb["sample"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "series/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample_per_step": n_ref},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "series/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    return "%sFCParameters on %d series x len %d per GPU (BASELINE.json configs[2] shape), synthetic N(0,1) float32" % (
        args.settings.capitalize(), args.series, args.len)


# ----------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series", type=int, default=1_000_000)
    ap.add_argument("--len", type=int, default=256)
    ap.add_argument("--settings", default="comprehensive", choices=["comprehensive", "efficient", "minimal"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from tsfresh_b200 import _lib
    from tsfresh_b200.plan import Plan

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner with printf on fd 1 when the first
        # communicator is created, so fd 1 points at stderr until the result line is written
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    S, L = args.series, args.len
    plan = Plan(settings_by_name(args.settings))
    F = plan.n_cols
    # a non-default torch stream shared with the library, so torch's CUDA events bracket the library's launches
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = _lib.Context(local_rank, stream=stream.cuda_stream)
    dp = _lib.DevicePlan(ctx, plan)

    gen = torch.Generator(device=dev)
    gen.manual_seed(42 + 2 + rank)
    values = torch.randn((S, L), generator=gen, device=dev, dtype=torch.float32)
    out = torch.empty((S, F), device=dev, dtype=torch.float64)
    gathered = None
    comm_stream = None
    n_blocks = 8
    if world > 1:
        gathered = torch.empty((world, S, F), device=dev, dtype=torch.float64)
        comm_stream = torch.cuda.Stream(device=dev)

    def step(timing=False):
        if world == 1:
            dp.extract_dense_device(values.data_ptr(), S, L, out.data_ptr(), timing=timing)
            return
        # row blocks: kernels of block b+1 overlap the all-gather of block b (side stream)
        bs = (S + n_blocks - 1) // n_blocks
        for b in range(n_blocks):
            lo, hi = b * bs, min(S, (b + 1) * bs)
            if lo >= hi:
                break
            dp.extract_dense_device(values[lo:hi].data_ptr(), hi - lo, L, out[lo:hi].data_ptr(), timing=False)
            ev = torch.cuda.Event()
            ev.record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev)
                # gathered[r, lo:hi] <- rank r's block; contiguous per rank, so gather into a staging view
                dist.all_gather_into_tensor(gathered_stage[b][: world * (hi - lo) * F], out[lo:hi].reshape(-1))
        stream.wait_stream(comm_stream)

    gathered_stage = None
    if world > 1:
        bs = (S + n_blocks - 1) // n_blocks
        gathered_stage = [torch.empty(world * bs * F, device=dev, dtype=torch.float64) for _ in range(n_blocks)]
        del gathered

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    multi_stream = int(os.environ.get("TSFX_STREAMS", "1")) > 1      # groups overlap: per-group events need a separate pass
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    group_ms = {}
    e0.record(stream)
    for _ in range(args.steps):
        step(timing=(world == 1 and not multi_stream))
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    # per-group CUDA events: taken from the last timed step at N=1, from one extra (untimed) pass over this rank's
    # shard at N>1 (the timed steps there are split into row blocks for the all-gather overlap)
    if world > 1 or multi_stream:
        dp.extract_dense_device(values.data_ptr(), S, L, out.data_ptr(), timing=True)
        torch.cuda.synchronize()
    group_ms = ctx.timings()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * S / (ms_step / 1e3)
    launches_per_step = ctx.launch_count() * (1 if world == 1 else n_blocks)
    if rank != 0:
        group_ms = {}

    # ---------------- impute of the resident feature matrix: the HBM-bound pass of the framework
    impute_info = None
    if world == 1:
        ctx.impute_device(out.data_ptr(), S, F)                     # warm-up (allocations)
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record(stream)
        for _ in range(3):
            ctx.impute_device(out.data_ptr(), S, F)
        i1.record(stream)
        torch.cuda.synchronize()
        ims = i0.elapsed_time(i1) / 3
        # algorithmic bytes: the statistics sweep reads the matrix once; the replacement sweep only visits the
        # (row slice, column tile) blocks that hold a non-finite value
        impute_info = {"ms": ims, "algorithmic_GB": S * F * 8 / 1e9, "GBps": S * F * 8 / (ims * 1e-3) / 1e9,
                       "bound": "hbm", "kernels": "k_col_stats + k_col_reduce + k_impute_apply (+ one radix sort per NaN column)"}

    # ---------------- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        hv = torch.empty((S, L), dtype=torch.float32).pin_memory()
        hv.copy_(values.cpu())
        ho = torch.empty((S, F), dtype=torch.float64).pin_memory()
        hv_np, ho_np = hv.numpy(), ho.numpy()
        del out
        torch.cuda.empty_cache()
        for _ in range(2):
            dp.extract_dense(hv_np, out=ho_np)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 5))
        for _ in range(n_e2e):
            dp.extract_dense(hv_np, out=ho_np)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * S / dt, "unit": "series/s", "h2d_bytes_per_step": int(S * L * 4),
               "d2h_bytes_per_step": int(S * F * 8), "ms_per_step": dt * 1e3,
               "call": "tsfx_extract_dense (C ABI, pinned host buffers)"}

    if rank == 0:
        peaks, which = measured_peaks()
        roofline = None
        groups = {}
        if group_ms:
            # algorithmic bytes per series of one kernel group: the 4*L value bytes it reads plus the 8 bytes per
            # output column it writes (DESIGN.md "Roofline"); whole pass: 4L + 12 + 8F
            ncols = {}
            from tsfresh_b200 import plan as planmod
            for g, cnt in group_columns(plan).items():
                ncols[g] = cnt
            for g, ms in group_ms.items():
                by = S * 16 * F if g == "assemble" else S * (4 * L + 8 * ncols.get(g, 0))
                groups[g] = {"ms": ms, "columns": ncols.get(g, 0), "algorithmic_GBps": by / (ms * 1e-3) / 1e9}
            dom = max(group_ms, key=lambda g: group_ms[g])
            ach = groups[dom]["algorithmic_GBps"]
            traffic, limiter = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r1.json")))
                w = tj["workload"]
                if (w["series"], w["len"], w["settings"]) == (S, L, args.settings) and ("k_" + dom) in tj:
                    traffic = tj["k_" + dom]["traffic_bytes"] / 1e9
                    limiter = "ncu: fp64 pipe %.0f%% active, issue %.0f%% active" % (
                        tj["k_" + dom]["fp64_pipe_active_pct"], tj["k_" + dom]["issue_active_pct"])
            except Exception:
                pass
            # supplementary compute roofline: these kernels are bound by instruction issue, not by HBM.  Warp
            # instructions per launch come from the committed ncu capture of the same workload
            # (smsp__inst_executed.sum, profiles/traffic_r1.json); the rate is measured live; the peak is
            # 148 SMs x 4 schedulers x 1 warp instruction per clock at the sampled SM clock.
            issue = None
            try:
                inst = tj["k_" + dom].get("inst_executed") if traffic is not None else None
            except Exception:
                inst = None
            if inst:
                clk = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
                peak_i = 148 * 4 * clk
                a_i = inst / (group_ms[dom] * 1e-3)
                issue = {"bound": "instruction issue", "achieved": a_i / 1e12, "peak": peak_i / 1e12,
                         "unit": "T warp-inst/s", "frac": a_i / peak_i, "warp_inst_per_series": inst / S}
            try:                                           # every group's issue-slot utilisation, same recipe
                if (tj["workload"]["series"], tj["workload"]["len"], tj["workload"]["settings"]) == (S, L, args.settings):
                    clk = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
                    for g in groups:
                        ginst = tj.get("k_" + g, {}).get("inst_executed")
                        if ginst:
                            groups[g]["issue_frac"] = ginst / (groups[g]["ms"] * 1e-3) / (148 * 4 * clk)
            except Exception:
                pass
            roofline = {"kernel": "k_" + dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                        "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else "fallback 6650 GB/s",
                        "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": traffic, "traffic_unit": "GB per launch (ncu dram read+write)",
                        "algorithmic_GB_per_launch": S * (4 * L + 8 * ncols.get(dom, 0)) / 1e9,
                        "limiter": limiter or "the dominant kernel is instruction-issue bound, not HBM bound (DESIGN.md section 4)",
                        "compute_roofline": issue,
                        "whole_pass_GBps": S * (4 * L + 12 + 8 * F) / (ms_step * 1e-3) / 1e9,
                        "groups": groups}
            if impute_info:
                impute_info["frac_of_hbm_peak"] = impute_info["GBps"] / peaks["hbm_gbs"]
                roofline["impute"] = impute_info
        cb = None
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(L, args.settings, args.cpu_seconds)
        line = {
            "metric": METRIC, "value": value, "unit": "series/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "columns": F, "global_series": world * S,
                       "l2": "inputs (%.2f GB) + outputs (%.2f GB) per step exceed the 126 MB L2" % (S * L * 4 / 1e9, S * F * 8 / 1e9),
                       "parallelism": "ids sharded contiguously over %d rank(s)%s" % (world, "" if world == 1 else "; all-gather of the feature matrix per row block, overlapped")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline, "cpu_baseline": cb,
        }
        if saved_stdout is not None:
            sys.stdout.flush()
            try:
                import ctypes
                ctypes.CDLL(None).fflush(None)       # whatever C code buffered for "stdout" leaves through stderr too
            except Exception:
                pass
            os.dup2(saved_stdout, 1)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        if saved_stdout is not None and rank != 0:
            os.dup2(saved_stdout, 1)
        dist.destroy_process_group()


def group_columns(plan):
    """number of output columns each kernel group writes (mirrors group_of() in csrc/tsfx_api.cu)."""
    from tsfresh_b200.plan import CALC
    G = {"sorted": ["SYMMETRY_LOOKING", "HAS_DUPLICATE", "MEDIAN", "PERCENTAGE_OF_REOCCURRING_VALUES_TO_ALL_VALUES",
                    "PERCENTAGE_OF_REOCCURRING_DATAPOINTS_TO_ALL_DATAPOINTS", "SUM_OF_REOCCURRING_VALUES",
                    "SUM_OF_REOCCURRING_DATA_POINTS", "RATIO_VALUE_NUMBER_TO_TIME_SERIES_LENGTH", "QUANTILE",
                    "MEAN_N_ABSOLUTE_MAX", "CHANGE_QUANTILES", "FRIEDRICH_COEFFICIENTS", "MAX_LANGEVIN_FIXED_POINT"],
         "spectral": ["FFT_COEFFICIENT", "FFT_AGGREGATED", "SPKT_WELCH_DENSITY", "FOURIER_ENTROPY", "CWT_COEFFICIENTS"],
         "la": ["AR_COEFFICIENT", "AUGMENTED_DICKEY_FULLER"],
         "entropy": ["SAMPLE_ENTROPY", "APPROXIMATE_ENTROPY"],
         "seq": ["LEMPEL_ZIV_COMPLEXITY", "PERMUTATION_ENTROPY"], "peaks": ["NUMBER_CWT_PEAKS"]}
    rev = {}
    for g, names in G.items():
        for n in names:
            rev[CALC["TSFX_" + n]] = g
    out = {}
    for c in plan.descs["calc"]:
        g = rev.get(int(c), "basic")
        out[g] = out.get(g, 0) + 1
    return out


if __name__ == "__main__":
    main()
