"""Throughput of 1 plan vs 2 plans (two compute streams) on device-resident 4096^2 images."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import sift_pyocl_b200 as sift
from sift_pyocl_b200.utils import multiscale_image

sift.par["OctaveMax"] = 3
imgs = [torch.from_numpy(multiscale_image(4096, 1234 + i)).cuda() for i in range(2)]
plans = [sift.SiftPlan(shape=(4096, 4096), dtype=np.float32) for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2)]
K = 40


def run(nplans):
    ps = plans[:nplans]
    for p in ps:
        p.submit(imgs[0]); p.collect(records=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for i in range(K):
        p = ps[i % nplans]
        if i >= nplans:
            n += p.collect(records=False)
        p.submit(imgs[i % 2])
    for j in range(nplans):
        n += ps[(K + j) % nplans].collect(records=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("plans=%d  %.3f ms/image  %.2f Mkp/s" % (nplans, 1e3 * dt / K, n / dt / 1e6))


for npl in range(1, len(plans) + 1):
    run(npl)
    run(npl)
