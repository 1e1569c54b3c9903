"""LinearAlign: align images on a reference image with an affine transformation -- same public
surface as the reference's sift-src/alignment.py (keypoints + match + host least squares + warp).

Data flow per frame (everything between the two copies stays on the device, like the reference's
``buffers["input"]`` / ``ref_kp_gpu``, alignment.py:157,242-249):

    frame --H2D--> SiftPlan (records stay in HBM) --D2D--> MatchPlan (reference list resident)
          --> 32 bytes per matched pair --D2H--> host: shift / affine least squares / 4-sigma rejection
          --> 6 numbers --> warp of the still-resident frame --D2H--> aligned image
"""
import ctypes
import logging
from threading import Semaphore

import numpy

from . import _lib
from .match import MatchPlan
from .plan import SiftPlan
from .utils import affine_lstsq

logger = logging.getLogger("sift.alignment")

MIN_MATCHES_AFFINE = 3 * 6  # three points per degree of freedom (alignment.py:266)
OUTLIER_SIGMAS = 4.0         # alignment.py:294-296


def transform(image, matrix, offset, fill, out_shape=None, mode=1, device=0):
    """Inverse-mapped affine warp, bilinear (mode 1) or nearest (mode 0): transform.cl:22-108; an (H, W, 3)
    uint8 image goes through transform_RGB (transform.cl:116-203).  Host in, host out."""
    m = numpy.ascontiguousarray(numpy.asarray(matrix, numpy.float32).reshape(4))
    o = numpy.ascontiguousarray(numpy.asarray(offset, numpy.float32).reshape(2))
    colour = numpy.ndim(image) == 3
    image = numpy.ascontiguousarray(image, numpy.uint8 if colour else numpy.float32)
    h, w = image.shape[:2]
    oh, ow = (h, w) if out_shape is None else out_shape
    out = numpy.empty((oh, ow, 3) if colour else (oh, ow), image.dtype)
    fn = _lib.load().siftb_transform_rgb if colour else _lib.load().siftb_transform
    _lib.check(fn(_lib.ptr(image), h, w, _lib.ptr(out), oh, ow, m.ctypes.data_as(_lib.c_float_p),
                  o.ctypes.data_as(_lib.c_float_p), ctypes.c_float(fill), int(mode), int(device)))
    return out


# ---- geometry on the matched pairs: float32 [m, 8] = (x, y, scale, angle) of the reference keypoint, then of
# ---- its match in the frame (MatchPlan.match_coords).  Small pure functions, unit-tested on the CPU.
def pairs_from_matching(matching):
    """The [m, 8] pair array of a reference-style (m, 2) recarray of matched keypoints."""
    out = numpy.empty((matching.shape[0], 8), numpy.float32)
    for col, field in enumerate(("x", "y", "scale", "angle")):
        out[:, col] = matching[:, 0][field]
        out[:, 4 + col] = matching[:, 1][field]
    return out


def median_shift(pairs):
    """Pure translation (alignment.py:266-274): identity matrix, offset = median displacement as (dy, dx)."""
    dx = pairs[:, 4] - pairs[:, 0]
    dy = pairs[:, 5] - pairs[:, 1]
    return numpy.identity(2, dtype=numpy.float32), numpy.array([numpy.median(dy), numpy.median(dx)], numpy.float32)


def affine_from_pairs(pairs):
    """Least-squares affine map reference -> frame (alignment.py:278-282 + utils.matching_correction) in the
    warp's convention: frame (y, x) = matrix . reference (y, x) + offset."""
    a, b, c, d, e, f = affine_lstsq(pairs[:, 0], pairs[:, 1], pairs[:, 4], pairs[:, 5])
    return numpy.array([[e, d], [b, a]], numpy.float32), numpy.array([f, c], numpy.float32)


def inlier_mask(pairs, nsigma=OUTLIER_SIGMAS):
    """True for the pairs kept by the reference's validation (alignment.py:283-297): a pair is an outlier when its
    displacement length, its rotation or its log scale ratio lies more than ``nsigma`` standard deviations from
    the mean over all pairs."""
    dx = pairs[:, 4] - pairs[:, 0]
    dy = pairs[:, 5] - pairs[:, 1]
    quantities = (numpy.sqrt(dx * dx + dy * dy), pairs[:, 7] - pairs[:, 3], numpy.log(pairs[:, 6] / pairs[:, 2]))
    keep = numpy.ones(pairs.shape[0], bool)
    with numpy.errstate(divide="ignore", invalid="ignore"):
        for q in quantities:
            keep &= ~(numpy.abs((q - q.mean()) / q.std()) > nsigma)  # NaN (zero spread) counts as inlier, like the reference
    return keep


def chain_transform(previous, matrix, offset):
    """3x3 homogeneous form of (matrix, offset) composed onto ``previous`` (alignment.py:303-320, relative mode)."""
    step = numpy.identity(3, dtype=numpy.float64)
    step[:2, :2] = matrix
    step[:2, 2] = offset
    return step if previous is None else numpy.dot(step, previous)


def residual_rms(pairs, matrix, offset):
    """RMS distance between the mapped reference keypoints and their matches (alignment.py:353-356)."""
    ref_yx = numpy.stack((pairs[:, 1], pairs[:, 0]))
    img_yx = numpy.stack((pairs[:, 5], pairs[:, 4]))
    corr = (numpy.dot(matrix, ref_yx).T + numpy.asarray(offset).T) - img_yx.T
    return numpy.sqrt((corr * corr).sum(axis=-1).mean())


class LinearAlign(object):
    """Align images on a reference image based on an affine transformation (bi-linear + offset)
    (reference alignment.py:76-168)."""

    def __init__(self, image, devicetype="CPU", profile=False, device=None, max_workgroup_size=None,
                 ROI=None, extra=0, context=None, init_sigma=None):
        self.profile = bool(profile)
        self.events = []
        self.program = None
        self.buffers = {}
        shape = tuple(image.shape)
        if len(shape) not in (2, 3):
            raise RuntimeError("Unable to process image of shape %s" % (shape,))
        self.RGB = len(shape) == 3
        self.shape = shape[:2]
        self.ref = numpy.ascontiguousarray(image, numpy.uint8 if self.RGB else numpy.float32)
        pad = tuple(extra[:2]) if "__len__" in dir(extra) else (int(extra), int(extra))
        self.extra = pad
        self.outshape = tuple(int(n) + 2 * int(e) for n, e in zip(self.shape, pad))
        self.ROI = ROI
        self.ctx = context
        self.device = device
        self.sift = SiftPlan(template=self.ref, devicetype=devicetype, context=context, profile=self.profile,
                             device=device, max_workgroup_size=max_workgroup_size, init_sigma=init_sigma)
        self.match = MatchPlan(context=context, profile=self.profile, device=device,
                               max_workgroup_size=max_workgroup_size)
        self._set_reference(self.sift.keypoints(self.ref))
        self.fill_value = 0
        self.sem = Semaphore()
        self.relative_transfo = None

    def _set_reference(self, kp):
        """New reference keypoints: ROI filter on the host (alignment.py:149-154), then resident in the matcher
        (alignment.py:157 ``ref_kp_gpu``)."""
        if self.ROI is not None:
            inside = numpy.asarray(self.ROI)[(numpy.round(kp.y).astype(numpy.int32),
                                              numpy.round(kp.x).astype(numpy.int32))].astype(bool)
            logger.warning("Reducing keypoint list from %i to %i because of the ROI" % (kp.size, inside.sum()))
            kp = kp[inside]
        self.ref_kp = kp
        self.match.hold(0, self.ref_kp)

    def align(self, img, shift_only=False, return_all=False, double_check=False, relative=False, orsa=False):
        """Align image on reference image (reference alignment.py:227-360).

        :param img: numpy array containing the image to align to reference
        :param shift_only: fit a translation only (median displacement of the matched keypoints)
        :param return_all: return in addition to the image, keypoints, matching keypoints and transformations as a dict
        :param double_check: re-fit after dropping the matches more than 4 sigma away in displacement, rotation or scale
        :param relative: update reference keypoints with those from current image to perform relative alignment
        :return: aligned image, or all information, or None when no keypoint matches
        """
        logger.debug("ref_keypoints: %s" % self.ref_kp.size)
        frame = numpy.ascontiguousarray(img, numpy.uint8 if self.RGB else numpy.float32)  # alignment.py:237-240
        with self.sem:
            self.sift.submit(frame)               # one upload: keypoints now, the warp below
            n_kp = self.sift.collect(records=False)
            logger.debug("mod image keypoints: %s" % n_kp)
            pairs = self.match.match_coords(self.ref_kp, self.sift.device_keypoints(n_kp))
            n_match = pairs.shape[0]
            if n_match == 0:
                logger.warning("No matching keypoints")
                return None
            if orsa:
                logger.warning("feature is not available. No ORSA filtering")  # alignment.py:260-264
            enough = n_match >= MIN_MATCHES_AFFINE
            if shift_only or not enough:
                (logger.debug if shift_only else logger.warning)("Shift Only mode: Common keypoints: %s" % n_match)
                matrix, offset = median_shift(pairs)
            else:
                logger.debug("Common keypoints: %s" % n_match)
                matrix, offset = affine_from_pairs(pairs)
            if double_check and enough:
                logger.warning("Validating keypoints, %s,%s" % (matrix, offset))
                keep = inlier_mask(pairs)
                if not keep.all():
                    matrix, offset = affine_from_pairs(pairs[keep])
            kp = matching = None
            if return_all or relative:
                kp = self.sift.fetch_keypoints()
            if return_all:
                matching = self.match.last_pairs()
            if relative:  # the current frame becomes the reference of the next one; transforms accumulate
                self._set_reference(kp)
                self.relative_transfo = chain_transform(self.relative_transfo, matrix, offset)
                matrix = numpy.ascontiguousarray(self.relative_transfo[:2, :2], dtype=numpy.float32)
                offset = numpy.ascontiguousarray(self.relative_transfo[:2, 2], dtype=numpy.float32)
            fill = self.sift.buffers["min"].get()[0]  # alignment.py:345
            result = self.sift.warp_last(matrix, offset, float(fill), self.outshape, 1)
            if self.profile:
                self.events += [(name, ms) for name, ms in self.sift.fetch_events()
                                if name.startswith("transform") or name.startswith("copy D->H transformed")]
        if return_all:
            return {"result": result, "keypoint": kp, "matching": matching, "offset": offset, "matrix": matrix,
                    "rms": residual_rms(pairs, matrix, offset)}
        return result

    __call__ = align

    def log_profile(self):
        """Print the timing of every device operation of the underlying plans and of the warp
        (reference alignment.py:363-376)."""
        self.sift.log_profile()
        self.match.log_profile()
        t = 0.0
        for name, et in self.events:
            print("%50s:\t%.3fms" % (name, et))
            t += et
        print("%50s:\t%.3fms" % ("Total transform", t))
