"""Shared helpers of the GPU parity tests."""
import numpy as np

RTOL = 1e-3  # north_star tolerance for x, y, sigma, theta


def ms(n, seed=1234, shape=None):
    from sift_pyocl_b200.utils import multiscale_image
    return multiscale_image(n, seed, shape)


def sort_rows(a):
    return a[np.lexsort(tuple(a[:, k] for k in range(a.shape[1] - 1, -1, -1)))]


def sort_kp(kp):
    return kp[np.lexsort((kp.angle, kp.scale, kp.y, kp.x))]


def same_records(a, b):
    """Two keypoint arrays hold the same set of 144-byte records (order is nondeterministic: atomic appends)."""
    a, b = sort_kp(a), sort_kp(b)
    return a.size == b.size and a.tobytes() == b.tobytes()


def compare_whole(sift, oracle, img, octave_max=0, **kw):
    """SiftPlan.keypoints(img) against the oracle: identical counts per octave and per (octave, scale, stage),
    bit-identical x, y, scale, angle and descriptors, records grouped by octave."""
    if octave_max:
        sift.par["OctaveMax"] = octave_max
    try:
        plan = sift.SiftPlan(template=img, **kw)
    finally:
        sift.par["OctaveMax"] = 100000
    kp = plan.keypoints(img)
    ref, info = oracle.keypoints(oracle.to_float(img) if img.dtype != np.float32 or img.ndim == 3 else img,
                                 init_sigma=kw.get("init_sigma") or 1.6, pix_per_kp=kw.get("PIX_PER_KP") or 10,
                                 octave_max=octave_max, return_all=True)
    assert np.array_equal(plan.last_counts, info["n_per_octave"][:plan.octave_max])
    assert np.array_equal(plan.stage_counts(), info["stage_counts"][:plan.octave_max])
    assert kp.size == ref.size
    a, b = sort_kp(kp), sort_kp(ref)
    for f in ("x", "y", "scale", "angle"):
        np.testing.assert_allclose(a[f], b[f], rtol=RTOL, atol=1e-6)
        assert np.array_equal(a[f], b[f]), "fp32 field %s is expected to be bit-identical to the oracle" % f
    assert np.array_equal(a.desc, b.desc)
    off = np.concatenate([[0], np.cumsum(plan.last_counts)])
    for o in range(plan.octave_max):
        assert same_records(kp[off[o]:off[o + 1]], ref[off[o]:off[o + 1]]), "octave %d" % o
    return plan, kp, ref


def desc_sets(n1=3000, n2=2500, seed=3):
    """Two synthetic keypoint lists with SIFT-like descriptor statistics; list 2 = permuted, perturbed rows of
    list 1, so a known fraction passes the ratio test."""
    from sift_pyocl_b200._lib import dtype_kp
    rng = np.random.default_rng(seed)
    d1 = np.minimum(rng.gamma(1.0, 28.0, (n1, 128)), 255).astype(np.uint8)
    perm = rng.permutation(n1)[:n2]
    noise = rng.integers(-2, 3, (n2, 128), dtype=np.int8) * (rng.random((n2, 128), dtype=np.float32) < 0.5)
    d2 = np.clip(d1[perm].astype(np.int16) + noise, 0, 255).astype(np.uint8)
    k1, k2 = np.zeros(n1, dtype_kp), np.zeros(n2, dtype_kp)
    k1["desc"], k2["desc"] = d1, d2
    k1["x"], k2["x"] = np.arange(n1), np.arange(n2)
    k1["y"], k2["y"] = rng.random(n1), rng.random(n2)
    return k1.view(np.recarray), k2.view(np.recarray), perm
