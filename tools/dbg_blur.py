import numpy as np, sys
sys.path.insert(0, '.')
from sift_pyocl_b200 import stages
from oracle import siftref as R
img = (255*np.random.default_rng(0).random((256, 512))).astype(np.float32)
for s in (1.2262734984654078, 3.0900155872895603):
    t = R.gaussian_taps(s)
    out = stages.blur(img, t)
    print(s, t.size, np.array_equal(out, R.blur(img, t)), abs(out-R.blur(img,t)).max())
