#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: python profiles/ncu_top_lines.py REPORT.ncu-rep KERNEL_REGEX [N]
(wraps `ncu -i REPORT --page source --csv --print-source cuda,sass`)."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, agg = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and len(r) > 8 and r[0] not in ("", "Line No"):
            try:
                agg.append((fname, int(r[0]), r[1].strip(), int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0)))
            except ValueError:
                pass
    tot_i = sum(a[4] for a in agg) or 1
    tot_s = sum(a[3] for a in agg) or 1
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * a[3] / tot_s, 100.0 * a[4] / tot_i, a[0], a[1], a[2][:110]))


if __name__ == "__main__":
    main()
