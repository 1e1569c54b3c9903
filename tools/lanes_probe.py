"""Device-resident throughput of one SiftPlan (4096 x 4096 fp32, 3 octaves) against the number of compute lanes
(SIFTB_LANES, read at plan creation) and the number of images kept in flight.   usage: lanes_probe.py [K]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import sift_pyocl_b200 as sift  # noqa: E402
from sift_pyocl_b200.utils import multiscale_image  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 60
imgs = [torch.from_numpy(multiscale_image(4096, seed=1234 + i)).cuda() for i in range(2)]
sift.par["OctaveMax"] = 3


def run(plan, depth):
    pending, t0 = 0, None
    for it in range(K + 6):
        if it == 6:
            while pending:
                plan.collect(records=False)
                pending -= 1
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        if pending >= depth:
            plan.collect(records=False)
            pending -= 1
        plan.submit(imgs[it % 2])
        pending += 1
    while pending:
        plan.collect(records=False)
        pending -= 1
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / K * 1e3, 4)


out = {}
for lanes in (1, 2, 3):
    os.environ["SIFTB_LANES"] = str(lanes)
    plan = sift.SiftPlan(shape=(4096, 4096), dtype=np.float32)
    for depth in (1, 2, 3):
        out["lanes %d, %d in flight" % (lanes, depth)] = [run(plan, depth) for _ in range(3)]
    out["lanes %d MB" % lanes] = plan.memory / 1e6
    del plan
print(json.dumps(out))
