"""Host-side geometry of LinearAlign (sift_pyocl_b200/alignment.py) against the reference's expressions
(alignment.py:266-322, 353-356), restated here with the reference's own statements on an (m, 2) recarray of matched
keypoints.  No device needed."""
import numpy as np
import pytest

from sift_pyocl_b200._lib import dtype_kp


def _matching(n=300, seed=0, outliers=0):
    rng = np.random.default_rng(seed)
    m = np.zeros((n, 2), dtype_kp).view(np.recarray)
    m.x[:, 0], m.y[:, 0] = rng.random(n) * 900, rng.random(n) * 700
    m.scale[:, 0] = 1.6 + rng.random(n) * 3
    m.angle[:, 0] = rng.uniform(-3, 3, n)
    m.x[:, 1] = 1.02 * m.x[:, 0] - 0.03 * m.y[:, 0] + 4 + rng.normal(0, 0.2, n)
    m.y[:, 1] = 0.02 * m.x[:, 0] + 0.97 * m.y[:, 0] - 3 + rng.normal(0, 0.2, n)
    m.scale[:, 1] = m.scale[:, 0] * (1 + rng.normal(0, 0.01, n))
    m.angle[:, 1] = m.angle[:, 0] + rng.normal(0, 0.01, n)
    for i in range(outliers):  # gross mismatches: far away, rotated, rescaled
        m.x[i, 1] += 400.0 * (1 + i)
        m.angle[n - 1 - i, 1] += 2.5
        m.scale[n // 2 + i, 1] *= 30.0
    return m


def _reference_fit(matching):
    """alignment.py:278-282 with utils.matching_correction completed as pinv(X).y (test_transform.py:118-133)."""
    n = matching.shape[0]
    X = np.zeros((2 * n, 6))
    X[::2, 0], X[::2, 1], X[::2, 2] = matching.x[:, 0], matching.y[:, 0], 1
    X[1::2, 3], X[1::2, 4], X[1::2, 5] = matching.x[:, 0], matching.y[:, 0], 1
    y = np.zeros((2 * n, 1))
    y[::2, 0], y[1::2, 0] = matching.x[:, 1], matching.y[:, 1]
    t = np.dot(np.linalg.pinv(X), y).ravel()
    offset = np.array([t[5], t[2]], dtype=np.float32)
    matrix = np.empty((2, 2), dtype=np.float32)
    matrix[0, 0], matrix[0, 1] = t[4], t[3]
    matrix[1, 0], matrix[1, 1] = t[1], t[0]
    return matrix, offset


def test_pairs_layout_and_median_shift():
    from sift_pyocl_b200.alignment import median_shift, pairs_from_matching
    m = _matching(51, 1)
    p = pairs_from_matching(m)
    assert p.shape == (51, 8) and p.dtype == np.float32
    assert np.array_equal(p[:, 0], m.x[:, 0]) and np.array_equal(p[:, 5], m.y[:, 1]) and np.array_equal(p[:, 7], m.angle[:, 1])
    matrix, offset = median_shift(p)
    dx, dy = m[:, 1].x - m[:, 0].x, m[:, 1].y - m[:, 0].y          # alignment.py:271-274
    assert np.array_equal(matrix, np.identity(2, dtype=np.float32))
    assert np.array_equal(offset, np.array([+np.median(dy), +np.median(dx)], np.float32))


@pytest.mark.parametrize("n", [18, 300, 20000])
def test_affine_from_pairs_equals_reference_fit(n):
    from sift_pyocl_b200.alignment import affine_from_pairs, pairs_from_matching
    m = _matching(n, 2)
    matrix, offset = affine_from_pairs(pairs_from_matching(m))
    want_m, want_o = _reference_fit(m)
    assert matrix.dtype == np.float32 and offset.dtype == np.float32
    assert np.allclose(matrix, want_m, atol=2e-6) and np.allclose(offset, want_o, atol=2e-4)
    assert abs(matrix[1, 1] - 1.02) < 1e-3 and abs(matrix[0, 0] - 0.97) < 1e-3 and abs(offset[0] + 3) < 0.3


def test_inlier_mask_equals_reference_outlayer():
    from sift_pyocl_b200.alignment import affine_from_pairs, inlier_mask, pairs_from_matching
    m = _matching(400, 3, outliers=3)
    p = pairs_from_matching(m)
    # alignment.py:285-297
    dx, dy = m[:, 1].x - m[:, 0].x, m[:, 1].y - m[:, 0].y
    dangle = m[:, 1].angle - m[:, 0].angle
    dscale = np.log(m[:, 1].scale / m[:, 0].scale)
    distance = np.sqrt(dx * dx + dy * dy)
    outlayer = np.zeros(distance.shape, np.int8)
    outlayer += abs((distance - distance.mean()) / distance.std()) > 4
    outlayer += abs((dangle - dangle.mean()) / dangle.std()) > 4
    outlayer += abs((dscale - dscale.mean()) / dscale.std()) > 4
    keep = inlier_mask(p)
    assert np.array_equal(keep, outlayer == 0) and 0 < (~keep).sum() <= 9
    # re-fit on the inliers is closer to the true map than the contaminated fit
    bad, _ = affine_from_pairs(p)
    good, _ = affine_from_pairs(p[keep])
    truth = np.array([[0.97, 0.02], [-0.03, 1.02]], np.float32)
    assert abs(good - truth).max() < abs(bad - truth).max() and abs(good - truth).max() < 1e-3
    want_m, _ = _reference_fit(m[outlayer == 0])
    assert np.allclose(good, want_m, atol=2e-6)
    # identical pairs: zero spread -> 0/0 -> nobody is an outlier (the reference's NaN > 4 is False)
    same = np.tile(p[:1], (20, 1))
    assert inlier_mask(same).all()


def test_chain_transform_equals_reference_relative_mode():
    from sift_pyocl_b200.alignment import chain_transform
    rng = np.random.default_rng(5)
    rel = None
    want = None
    for _ in range(3):
        matrix = (np.identity(2) + rng.normal(0, 0.02, (2, 2))).astype(np.float32)
        offset = rng.normal(0, 5, 2).astype(np.float32)
        transfo = np.zeros((3, 3), dtype=np.float64)          # alignment.py:307-316
        transfo[:2, :2] = matrix
        transfo[0, 2] = offset[0]
        transfo[1, 2] = offset[1]
        transfo[2, 2] = 1
        want = transfo if want is None else np.dot(transfo, want)
        rel = chain_transform(rel, matrix, offset)
        assert np.array_equal(rel, want)


def test_residual_rms_equals_reference():
    from sift_pyocl_b200.alignment import affine_from_pairs, pairs_from_matching, residual_rms
    m = _matching(250, 6)
    p = pairs_from_matching(m)
    matrix, offset = affine_from_pairs(p)
    corr = np.dot(matrix, np.vstack((m[:, 0].y, m[:, 0].x))).T + offset.T - np.vstack((m[:, 1].y, m[:, 1].x)).T
    want = np.sqrt((corr * corr).sum(axis=-1).mean())           # alignment.py:353-356
    assert np.isclose(residual_rms(p, matrix, offset), want, rtol=1e-6) and want < 0.5
