#!/usr/bin/env python
"""Generate tests/golden/reference_python.npz by RUNNING THE REFERENCE'S OWN python restatement.

The reference (pierrepaleo/sift_pyocl) cannot execute its OpenCL path in this image, but its test
suite ships a pure-numpy restatement of every keypoint stage, test/test_image_functions.py, which
its own unit tests use as the oracle for the kernels (test_image.py, test_keypoints.py).  This
script imports that file FROM /root/reference (tabs expanded -- it is python-2 era and mixes tabs
and spaces at :407), runs its functions on a small seeded image and stores inputs + outputs.
Nothing from the reference is copied into the repository: only the numeric vectors are committed.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The tests (tests/test_oracle_golden.py) read the .npz and never touch /root/reference.

Fixture construction follows test/test_image_setup.py:12-81 (numpy taps + scipy convolve1d
"reflect" pyramid) on a seeded synthetic image instead of scipy.misc.lena() (no longer shipped).
"""
import os
import sys

import numpy
import scipy.ndimage

REF = os.environ.get("SIFT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_functions():
    path = os.path.join(REF, "test", "test_image_functions.py")
    src = open(path).read().expandtabs(8)
    ns = {"__name__": "reference_test_image_functions"}
    exec(compile(src, path, "exec"), ns)
    return ns


def multiscale_image(h, w, seed):
    rng = numpy.random.default_rng(seed)
    img = numpy.zeros((h, w), numpy.float32)
    for k in (1, 2, 4, 8, 16, 32):
        hk, wk = -(-h // k), -(-w // k)
        r = rng.random((hk, wk), dtype=numpy.float32)
        z = r if k == 1 else scipy.ndimage.zoom(r, k, order=1)
        img += numpy.float32(numpy.sqrt(k)) * z[:h, :w].astype(numpy.float32)
    return img


def my_blur(img, sigma):  # test_image_setup.py:12-20
    ksize = int(numpy.ceil(8 * sigma + 1))
    if ksize % 2 == 0:
        ksize += 1
    x = numpy.arange(ksize) - (ksize - 1.0) / 2.0
    gaussian = numpy.exp(-(x / sigma) ** 2 / 2.0).astype(numpy.float32)
    gaussian /= gaussian.sum(dtype=numpy.float32)
    tmp1 = scipy.ndimage.convolve1d(img, gaussian, axis=-1, mode="reflect")
    return scipy.ndimage.convolve1d(tmp1, gaussian, axis=0, mode="reflect")


def main():
    ref = load_reference_functions()
    out = {}
    H, W = 160, 208
    raw = multiscale_image(H, W, seed=7)
    out["raw"] = raw
    l = ref["normalize_image"](raw)
    out["normalized"] = l
    out["shrunk"] = numpy.ascontiguousarray(ref["shrink"](l, 2, 2))

    # pyramid, test_image_setup.py:36-70
    initsigma, cursigma = 1.6, 0.5
    g = numpy.zeros((6, H, W), numpy.float32)
    g[0] = my_blur(l, numpy.sqrt(initsigma ** 2 - cursigma ** 2))
    sigmaratio = 2 ** (1 / 3.0)
    for i in range(1, 6):
        sigma = initsigma * sigmaratio ** (i - 1.0) * numpy.sqrt(sigmaratio ** 2 - 1.0)
        g[i] = my_blur(g[i - 1], sigma)
    DOGS = numpy.zeros((5, H, W), numpy.float32)
    for s in range(1, 6):
        DOGS[s - 1] = -(g[s] - g[s - 1])
    out["g"] = g  # DOGS[s] = g[s] - g[s+1] is recomputed (exactly) by the tests

    border_dist = numpy.int32(5)
    peakthresh = numpy.float32(255.0 * 0.04 / 3.0)
    EdgeThresh, EdgeThresh0 = numpy.float32(0.06), numpy.float32(0.08)
    nb_keypoints = 1000
    orisigma = numpy.float32(1.5)

    for octsize in (1, 2):
        for s in (1, 2, 3):
            tag = "o%d_s%d" % (octsize, s)
            kp_prev, n_ext = ref["my_local_maxmin"](DOGS, peakthresh, border_dist, octsize, EdgeThresh0, EdgeThresh,
                                                    nb_keypoints, s, W, H)
            out["maxmin_" + tag] = kp_prev[:n_ext].copy()
            if octsize != 1:
                continue
            # interpolation, test_image.py:205-252
            kp_int = kp_prev.copy()
            for i, k in enumerate(kp_int[:n_ext]):
                kp_int[i] = ref["my_interp_keypoint"](DOGS, s, int(k[1]), int(k[2]), 5, peakthresh, W, H)
            out["interp_" + tag] = kp_int[:n_ext].copy()
            kp_c, n_c = ref["my_compact"](kp_int.copy(), nb_keypoints)
            out["compact_" + tag] = kp_c[:n_c].copy()
            grad, ori = ref["my_gradient"](g[s])
            out["grad_" + tag] = grad.astype(numpy.float32)
            out["ori_" + tag] = ori.astype(numpy.float32)
            kp_o, n_o = ref["my_orientation"](kp_c.copy(), nb_keypoints, 0, n_c, grad, ori, octsize, orisigma)
            out["orient_" + tag] = kp_o[:n_o].copy()
            out["orient_nbase_" + tag] = numpy.int32(n_c)
            desc = ref["my_descriptor"](kp_o.copy(), grad, ori, octsize, 0, n_o)
            out["desc_" + tag] = desc
            print(tag, "extrema", n_ext, "interp", n_c, "orient", n_o, "desc", desc.shape)

    # matching, test_image_functions.py:384-426 (check_for_match gives ratio and argmin per query)
    d1 = numpy.concatenate([out["desc_o1_s%d" % s] for s in (1, 2, 3)]).astype(numpy.int32)
    rng = numpy.random.default_rng(11)
    perm = rng.permutation(d1.shape[0])
    d2 = numpy.clip(d1[perm] + rng.integers(-3, 4, d1.shape) * (rng.random(d1.shape) < 0.5), 0, 255).astype(numpy.int32)
    d2 = d2[: int(0.8 * d2.shape[0])]
    ratios, argmins = [], []
    for a in d1:
        r, m = ref["check_for_match"](a, d2)
        ratios.append(r)
        argmins.append(m)
    out["match_desc1"] = d1.astype(numpy.uint8)
    out["match_desc2"] = d2.astype(numpy.uint8)
    out["match_ratio"] = numpy.array(ratios, numpy.float64)
    out["match_argmin"] = numpy.array(argmins, numpy.int32)
    numpy.savez_compressed(os.path.join(HERE, "reference_python.npz"), **out)
    print("written", os.path.join(HERE, "reference_python.npz"))


if __name__ == "__main__":
    sys.exit(main())
