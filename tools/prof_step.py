"""Run a few steps of the 4096x4096 / 3-octave workload (device-resident input) for ncu captures."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
import sift_pyocl_b200 as sift
from sift_pyocl_b200.utils import multiscale_image

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
octaves = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sift.par["OctaveMax"] = octaves
plan = sift.SiftPlan(shape=(size, size), dtype=np.float32)
img = torch.from_numpy(multiscale_image(size, 1234)).cuda()
for _ in range(steps):
    plan.submit(img)
    n = plan.collect(records=False)
print("keypoints", n, plan.last_counts.tolist())
