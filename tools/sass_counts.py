#!/usr/bin/env python
"""Instruction-mnemonic counts per kernel of libsiftb200.so (cuobjdump -sass): the evidence that the blur kernel uses
TMA + mbarrier, the matcher VABSDIFF4.U8.ACC, etc.   usage: sass_counts.py [path/to/lib.so] > profiles/rNN_sass_counts.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "sift_pyocl_b200/libsiftb200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ["UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "VABSDIFF4", "IDP4A", "FFMA", "DFMA", "MUFU", "F2F", "LDS", "STS", "LDG", "STG",
        "ATOMG", "REDG", "RED", "SHFL", "MATCH", "VOTE", "BAR", "WARPSYNC", "HMMA", "IMMA", "UTCMMA"]
fn, counts, total = None, {}, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[fn], total[fn] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.\w+)*)", line)
    if m and fn:
        total[fn] += 1
        op, mods = m.group(1), m.group(2)
        for wkey in want:
            if op == wkey or (wkey == "RED" and op.startswith("RED")):
                key = op + (mods if op in ("UTMALDG", "SYNCS", "VABSDIFF4", "LDGSTS") else "")
                counts[fn][key] += 1
print("# %s: SASS instruction counts per kernel (static), selected mnemonics" % lib)
for fn in sorted(counts):
    c = counts[fn]
    print("%-60s total %5d  %s" % (fn[:60], total[fn], "  ".join("%s=%d" % kv for kv in sorted(c.items()))))
