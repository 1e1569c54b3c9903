#!/usr/bin/env python
"""Per-CUDA-source-line summary of one kernel of an ncu report captured with --import-source on.

usage: ncu_lines.py <report.ncu-rep> <kernel-regex> [launch-skip] [top-n]
Uses `ncu --page source --print-source cuda,sass --csv`: rows with a line number carry the metrics of all SASS
instructions attributed to that source line.
"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1:3]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k",
                      "regex:" + kre, "-s", skip, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, agg = "?", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        agg.append((fname, int(r[0]), r[1].strip()[:100], float(d["Instructions Executed"] or 0),
                    float(d["# Samples"] or 0), float(d["Thread Instructions Executed"] or 0),
                    float(d.get("stall_long_sb", 0) or 0), float(d.get("stall_short_sb", 0) or 0),
                    float(d.get("stall_mio", 0) or 0), float(d.get("L1 Wavefronts Shared", 0) or 0)))
tot = sum(a[3] for a in agg) or 1
tots = sum(a[4] for a in agg) or 1
print("total warp-instr %.0f samples %.0f" % (tot, tots))
for a in sorted(agg, key=lambda a: -a[4])[:topn]:
    print("%5.1f%% samp %5.1f%% inst (%4.1f lanes) lsb %4.0f ssb %4.0f mio %4.0f | %s:%d  %s"
          % (100 * a[4] / tots, 100 * a[3] / tot, a[5] / max(a[3], 1), a[6], a[7], a[8], a[0], a[1], a[2]))
