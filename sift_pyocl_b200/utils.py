"""Host-side helpers (reference sift-src/utils.py).  No device code, no oracle."""
from math import ceil

import numpy


def calc_size(shape, blocksize):
    """Round ``shape`` up to a multiple of ``blocksize`` (reference utils.py:44-51; kept for API parity)."""
    if "__len__" in dir(blocksize):
        return tuple((int(i) + int(j) - 1) & ~(int(j) - 1) for i, j in zip(shape, blocksize))
    return tuple((int(i) + int(blocksize) - 1) & ~(int(blocksize) - 1) for i in shape)


def kernel_size(sigma, odd=False, cutoff=4):
    """Size of the Gaussian kernel for ``sigma`` (reference utils.py:54-64)."""
    size = int(ceil(2 * cutoff * sigma + 1))
    if odd and size % 2 == 0:
        size += 1
    return size


def sizeof(shape, dtype="uint8"):
    """Number of bytes of an array of ``shape`` and ``dtype`` (reference utils.py:76-83)."""
    itemsize = numpy.dtype(dtype).itemsize
    cnt = 1
    if "__len__" in dir(shape):
        for dim in shape:
            cnt *= dim
    else:
        cnt = int(shape)
    return cnt * itemsize


def affine_lstsq(x, y, xp, yp):
    """Least-squares (a, b, c, d, e, f) of ``x' = a x + b y + c ; y' = d x + e y + f``.

    The 2N x 6 system the reference builds (utils.py:156-189) is block structured: the even rows only involve
    (a, b, c), the odd rows only (d, e, f), both with the same N x 3 design matrix [x, y, 1].  Solving the two
    3-parameter problems gives the same solution as ``pinv(X) . y`` (test/test_transform.py:118-133) without the
    SVD of a 2N x 6 matrix.  Fewer than 3 independent points: minimum-norm solution, like pinv.
    """
    A = numpy.empty((len(x), 3), numpy.float64)
    A[:, 0], A[:, 1], A[:, 2] = x, y, 1.0
    rhs = numpy.empty((len(x), 2), numpy.float64)
    rhs[:, 0], rhs[:, 1] = xp, yp
    sol = numpy.linalg.lstsq(A, rhs, rcond=None)[0]  # columns: (a, b, c) and (d, e, f)
    return sol.T.ravel()


def matching_correction(matching):
    """Least-squares affine transform mapping keypoints[:, 0] onto keypoints[:, 1] of an (m, 2) recarray of
    matched keypoints.  Reference utils.py:156-189 builds the design matrix but the snapshot lost the solve and
    the return; completed as in the reference's own test (test/test_transform.py:131).  Returns (a, b, c, d, e, f)."""
    return affine_lstsq(matching.x[:, 0], matching.y[:, 0], matching.x[:, 1], matching.y[:, 1])


def multiscale_image(n, seed=1234, shape=None):
    """Seeded synthetic benchmark/test image:
    ``sum_k sqrt(k) * zoom(rng.random((n/k, n/k)), k, order=1)`` for k in 1,2,4,8,16,32, float32."""
    from scipy.ndimage import zoom
    h, w = (n, n) if shape is None else shape
    rng = numpy.random.default_rng(seed)
    img = numpy.zeros((h, w), numpy.float32)
    for k in (1, 2, 4, 8, 16, 32):
        hk, wk = -(-h // k), -(-w // k)
        r = rng.random((hk, wk), dtype=numpy.float32)
        z = r if k == 1 else zoom(r, k, order=1)
        img += numpy.float32(numpy.sqrt(k)) * z[:h, :w].astype(numpy.float32)
    return img
