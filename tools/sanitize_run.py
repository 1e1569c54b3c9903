"""Small whole-path + match + warp run for compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import sift_pyocl_b200 as sift
from sift_pyocl_b200.utils import multiscale_image
from sift_pyocl_b200.alignment import transform

for shape in ((300, 420), (513, 257)):
    img = multiscale_image(0, 5, shape)
    plan = sift.SiftPlan(template=img)
    kp = plan.keypoints(img)
    print(shape, kp.size, plan.last_counts.tolist())
mp = sift.MatchPlan()
print("matches", len(mp.match(kp, kp, raw_results=True)))
out = transform(img, [[1.01, 0.02], [-0.02, 0.99]], [1.5, -2.0], 0.0)
print("warp", out.shape)
