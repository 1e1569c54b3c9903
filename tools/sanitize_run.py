"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck): whole path in both
kernel-family variants, windows beyond the descriptor row table, buffer overflow, pipelined images on the plan's two
compute lanes, matcher (short list, segmented
long list, two queries per thread, L2 metric, device-side gathers), LinearAlign (warp of the resident frame)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import sift_pyocl_b200 as sift
from sift_pyocl_b200._lib import dtype_kp
from sift_pyocl_b200.utils import multiscale_image
from sift_pyocl_b200.alignment import transform

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for shape in ((300, 420), (513, 257)):
    img = multiscale_image(0, 5, shape)
    for dev in ("CPU", "GPU"):
        plan = sift.SiftPlan(template=img, devicetype=dev)
        kp = plan.keypoints(img)
        print(shape, dev, kp.size, plan.last_counts.tolist())
img = multiscale_image(0, 6, (256, 320))
print("init_sigma 3.0", sift.SiftPlan(template=img, init_sigma=3.0).keypoints(img).size)
print("overflow", sift.SiftPlan(template=img, PIX_PER_KP=300).keypoints(img).size)
plan = sift.SiftPlan(template=img)   # images in flight together: the plan's two compute lanes
imgs = [multiscale_image(0, 7 + i, (256, 320)) for i in range(5)]
print("two lanes", [k.size for k in plan.keypoints_many(imgs)], plan.memory)
mp = sift.MatchPlan()
print("matches", len(mp.match(kp, kp, raw_results=True)), mp.match(kp, kp).shape, mp.match_coords(kp, kp).shape)
rng = np.random.default_rng(0)
n1, n2 = (3000, 20000) if quick else (80000, 40000)
k1, k2 = np.zeros(n1, dtype_kp), np.zeros(n2, dtype_kp)
k1["desc"] = rng.integers(0, 80, (n1, 128))
k2["desc"] = rng.integers(0, 80, (n2, 128))
print("long lists", len(mp.match(k1.view(np.recarray), k2.view(np.recarray), raw_results=True)))
mp.metric = "l2"
print("l2", len(mp.match(k1[:2000].view(np.recarray), k2[:20000].view(np.recarray), raw_results=True)))
out = transform(img, [[1.01, 0.02], [-0.02, 0.99]], [1.5, -2.0], 0.0)
print("warp", out.shape)
la = sift.LinearAlign(img, extra=(3, 5))
res = la.align(np.roll(img, (2, 3), axis=(0, 1)), return_all=True)
print("align", None if res is None else (res["result"].shape, res["matching"].shape))
