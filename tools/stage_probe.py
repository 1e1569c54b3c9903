"""Median device time of every stage of the 4096 x 4096 / 3-octave workload (profiling events).
usage: stage_probe.py [reps] [images in flight]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
import sift_pyocl_b200 as sift
from sift_pyocl_b200.utils import multiscale_image

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 9
sift.par["OctaveMax"] = 3
plan = sift.SiftPlan(shape=(4096, 4096), dtype=np.float32, profile=True)
imgs = [torch.from_numpy(multiscale_image(4096, 1234 + i)).cuda() for i in range(2)]
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # images in flight
acc = {}
for j in range(depth):
    plan.submit(imgs[j % 2])
for r in range(reps + 2):
    plan.collect(records=False)
    events = plan.fetch_events()
    if r + depth < reps + 2:
        plan.submit(imgs[(r + depth) % 2])
    if r >= 2:
        per = {}
        for name, ms in events:
            key = name.split(" octave")[0]
            per[key] = per.get(key, 0.0) + ms
        for k, v in per.items():
            acc.setdefault(k, []).append(v)
print({k: round(float(np.median(v)), 4) for k, v in sorted(acc.items())})
