#!/usr/bin/env python
"""Aggregate an ncu SASS source page (csv) by CUDA source line using nvdisasm -g line markers.

usage: ncu_by_line.py <report.ncu-rep> <kernel-regex> <mangled-substring> [launch-skip] [cubin]
"""
import csv, io, re, subprocess, sys, collections
rep, kre, mangled = sys.argv[1:4]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
cubin = sys.argv[5] if len(sys.argv) > 5 else "/tmp/cub/siftb_api.sm_100a.cubin"
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# offset -> (file,line) for the function
line_of, cur, infun = {}, None, False
for l in dis:
    if l.startswith("//--------------------- .text."):
        infun = mangled in l
        continue
    if not infun:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "-s", skip, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
ci, sa, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[0].startswith("0x"):
        if len(r) > 1 and r[0] == "Kernel Name":
            break
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    key = line_of.get(a - base, ("?", 0))
    agg[key][0] += float(r[ci] or 0); agg[key][1] += float(r[sa] or 0); agg[key][2] += float(r[ti] or 0)
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total warp-instr %.0f samples %.0f" % (tot, tots))
src_cache = {}
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(__import__("os").environ.get("TOPN","45"))]:
    f, ln = key
    try:
        if f not in src_cache:
            src_cache[f] = open("/root/repo/sift_pyocl_b200/csrc/" + f).read().splitlines()
        text = src_cache[f][ln - 1].strip()[:90]
    except Exception:
        text = ""
    print("%5.1f%% samp %5.1f%% inst (%4.1f lanes) %s:%d  %s" % (100 * v[1] / tots, 100 * v[0] / tot, v[2] / max(v[0], 1), f, ln, text))
