"""CPU baselines on the host this runs on: the oracle (oracle/libsiftref.so, C + OpenMP, same
arithmetic as the CUDA path) on all host threads and on one thread, for the whole keypoints() call, the convolution
stage alone (GB/s with the 8*W*H formula), matching and the affine warp.  One JSON line.  No GPU needed.
usage: cpu_baselines.py [size=4096]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import siftref  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = siftref.multiscale_image(size, 1234)
all_threads = siftref.num_threads()


def timed(fn, reps):
    fn()  # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2], out


out = {"host_threads": all_threads, "cpu_count": os.cpu_count(), "image": "%dx%d multiscale seed 1234" % (size, size)}
norm = siftref.normalize(img)
rng = np.random.default_rng(3)
n2, nq = 100000, 2000
d2 = np.minimum(rng.gamma(1.0, 28.0, (n2, 128)), 255).astype(np.uint8)
k2 = np.zeros(n2, siftref.dtype_kp) if hasattr(siftref, "dtype_kp") else None
if k2 is None:
    from sift_pyocl_b200._lib import dtype_kp
    k2 = np.zeros(n2, dtype_kp)
k2["desc"] = d2
k1 = k2[rng.choice(n2, nq, replace=False)].copy()
k1, k2 = k1.view(np.recarray), k2.view(np.recarray)
for label, nthr, reps in (("all_threads", all_threads, 5), ("one_thread", 1, 1)):
    siftref.set_num_threads(nthr)
    r = {"threads": nthr}
    dt, kp = timed(lambda: siftref.keypoints(img, octave_max=3), reps if nthr > 1 else 1)
    r["keypoints_3_octaves"] = {"s": dt, "keypoints": int(kp.size), "keypoints_per_s": kp.size / dt}
    for sigma in (1.2263, 3.09):  # first and last blur of an octave: 11 and 27 taps
        taps = siftref.gaussian_taps(sigma)
        dt, _ = timed(lambda: siftref.blur(norm, taps), reps)
        r["blur_%d_taps" % len(taps)] = {"s": dt, "GBps_8WH": 8.0 * img.size / dt / 1e9}
    dt, m = timed(lambda: siftref.match(k1, k2), reps)
    r["match_%dx%d" % (nq, n2)] = {"s": dt, "byte_sad_per_s": nq * n2 * 128 / dt,
                                   "extrapolated_100k_x_100k_s": dt * (100000 / nq)}
    dt, _ = timed(lambda: siftref.transform(img, [[1.01, -0.02], [0.015, 0.99]], [7.0, 5.0], 0.0), reps)
    r["transform"] = {"s": dt, "GBps_8WH": 8.0 * img.size / dt / 1e9}
    out[label] = r
siftref.set_num_threads(all_threads)
print(json.dumps(out))
