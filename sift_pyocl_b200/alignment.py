"""LinearAlign: align images on a reference image with an affine transformation -- same public
surface as the reference's sift-src/alignment.py (keypoints + match + host least squares + warp)."""
import ctypes
import logging
from threading import Semaphore

import numpy

from . import _lib
from .match import MatchPlan
from .plan import SiftPlan
from .utils import matching_correction

logger = logging.getLogger("sift.alignment")


def transform(image, matrix, offset, fill, out_shape=None, mode=1, device=0):
    """Inverse-mapped affine warp, bilinear (mode 1) or nearest (mode 0): transform.cl:22-108; an (H, W, 3)
    uint8 image goes through transform_RGB (transform.cl:116-203)."""
    m = numpy.ascontiguousarray(numpy.asarray(matrix, numpy.float32).reshape(4))
    o = numpy.ascontiguousarray(numpy.asarray(offset, numpy.float32).reshape(2))
    if numpy.ndim(image) == 3:
        image = numpy.ascontiguousarray(image, numpy.uint8)
        h, w = image.shape[:2]
        oh, ow = (h, w) if out_shape is None else out_shape
        out = numpy.empty((oh, ow, 3), numpy.uint8)
        _lib.check(_lib.load().siftb_transform_rgb(_lib.ptr(image), h, w, _lib.ptr(out), oh, ow,
                                                   m.ctypes.data_as(_lib.c_float_p), o.ctypes.data_as(_lib.c_float_p),
                                                   ctypes.c_float(fill), int(mode), int(device)))
        return out
    image = numpy.ascontiguousarray(image, numpy.float32)
    h, w = image.shape
    oh, ow = (h, w) if out_shape is None else out_shape
    out = numpy.empty((oh, ow), numpy.float32)
    _lib.check(_lib.load().siftb_transform(_lib.ptr(image), h, w, _lib.ptr(out), oh, ow,
                                           m.ctypes.data_as(_lib.c_float_p), o.ctypes.data_as(_lib.c_float_p),
                                           ctypes.c_float(fill), int(mode), int(device)))
    return out


class LinearAlign(object):
    """Align images on a reference image based on an affine transformation (bi-linear + offset)
    (reference alignment.py:76-168)."""

    def __init__(self, image, devicetype="CPU", profile=False, device=None, max_workgroup_size=None,
                 ROI=None, extra=0, context=None, init_sigma=None):
        self.profile = bool(profile)
        self.events = []
        self.program = None
        self.ref = numpy.ascontiguousarray(image, numpy.float32)
        self.buffers = {}
        self.shape = image.shape
        if len(self.shape) == 3:
            self.RGB = True
            self.shape = self.shape[:2]
        elif len(self.shape) == 2:
            self.RGB = False
        else:
            raise RuntimeError("Unable to process image of shape %s" % (tuple(self.shape),))
        if "__len__" not in dir(extra):
            self.extra = (int(extra), int(extra))
        else:
            self.extra = extra[:2]
        self.outshape = tuple(i + 2 * j for i, j in zip(self.shape, self.extra))
        self.ROI = ROI
        self.ctx = context
        self.device = device
        self.sift = SiftPlan(template=image, context=context, profile=self.profile, device=device,
                             max_workgroup_size=max_workgroup_size, init_sigma=init_sigma)
        self.ref_kp = self.sift.keypoints(image)
        if self.ROI is not None:
            self.ref_kp = self._apply_roi(self.ref_kp)
        self.match = MatchPlan(context=context, profile=self.profile, device=device,
                               max_workgroup_size=max_workgroup_size)
        self.fill_value = 0
        self.sem = Semaphore()
        self.relative_transfo = None

    def _apply_roi(self, kp):  # alignment.py:149-154
        kpx = numpy.round(kp.x).astype(numpy.int32)
        kpy = numpy.round(kp.y).astype(numpy.int32)
        masked = self.ROI[(kpy, kpx)].astype(bool)
        logger.warning("Reducing keypoint list from %i to %i because of the ROI" % (kp.size, masked.sum()))
        return kp[masked]

    def align(self, img, shift_only=False, return_all=False, double_check=False, relative=False, orsa=False):
        """Align image on reference image (reference alignment.py:227-360).

        :param img: numpy array containing the image to align to reference
        :param return_all: return in addition to the image, keypoints, matching keypoints and transformations as a dict
        :param relative: update reference keypoints with those from current image to perform relative alignment
        :return: aligned image, or all information, or None when no keypoint matches
        """
        logger.debug("ref_keypoints: %s" % self.ref_kp.size)
        if self.RGB:
            data = numpy.ascontiguousarray(img, numpy.uint8)  # alignment.py:237-238
        else:
            data = numpy.ascontiguousarray(img, numpy.float32)
        with self.sem:
            kp = self.sift.keypoints(data)
            logger.debug("mod image keypoints: %s" % kp.size)
            raw_matching = self.match.match(self.ref_kp, kp, raw_results=True)
            len_match = raw_matching.shape[0]
            if len_match == 0:
                logger.warning("No matching keypoints")
                return
            matching = _lib.pair_records(self.ref_kp, raw_matching[:, 0], kp, raw_matching[:, 1])
            if orsa:
                logger.warning("feature is not available. No ORSA filtering")  # alignment.py:260-264
            if (len_match < 3 * 6) or (shift_only):  # 3 points per DOF
                if shift_only:
                    logger.debug("Shift Only mode: Common keypoints: %s" % len_match)
                else:
                    logger.warning("Shift Only mode: Common keypoints: %s" % len_match)
                dx = matching[:, 1].x - matching[:, 0].x
                dy = matching[:, 1].y - matching[:, 0].y
                matrix = numpy.identity(2, dtype=numpy.float32)
                offset = numpy.array([+numpy.median(dy), +numpy.median(dx)], numpy.float32)
            else:
                logger.debug("Common keypoints: %s" % len_match)
                matrix, offset = self._fit(matching)
            if double_check and (len_match >= 3 * 6):
                logger.warning("Validating keypoints, %s,%s" % (matrix, offset))
                dx = matching[:, 1].x - matching[:, 0].x
                dy = matching[:, 1].y - matching[:, 0].y
                dangle = matching[:, 1].angle - matching[:, 0].angle
                dscale = numpy.log(matching[:, 1].scale / matching[:, 0].scale)
                distance = numpy.sqrt(dx * dx + dy * dy)
                outlayer = numpy.zeros(distance.shape, numpy.int8)
                outlayer += abs((distance - distance.mean()) / distance.std()) > 4
                outlayer += abs((dangle - dangle.mean()) / dangle.std()) > 4
                outlayer += abs((dscale - dscale.mean()) / dscale.std()) > 4
                outlayersum = outlayer.sum()
                if outlayersum > 0 and not numpy.isinf(outlayersum):
                    matching2 = matching[outlayer == 0]
                    matrix, offset = self._fit(matching2)
            if relative:  # update stable part to perform a relative alignment
                self.ref_kp = kp
                if self.ROI is not None:
                    self.ref_kp = self._apply_roi(self.ref_kp)
                transfo = numpy.zeros((3, 3), dtype=numpy.float64)
                transfo[:2, :2] = matrix
                transfo[0, 2] = offset[0]
                transfo[1, 2] = offset[1]
                transfo[2, 2] = 1
                if self.relative_transfo is None:
                    self.relative_transfo = transfo
                else:
                    self.relative_transfo = numpy.dot(transfo, self.relative_transfo)
                matrix = numpy.ascontiguousarray(self.relative_transfo[:2, :2], dtype=numpy.float32)
                offset = numpy.ascontiguousarray(self.relative_transfo[:2, 2], dtype=numpy.float32)
            fill = self.sift.buffers["min"].get()[0]  # alignment.py:345
            result = transform(data, matrix, offset, float(fill), self.outshape, 1, self.sift.device)
        if return_all:
            corr = numpy.dot(matrix, numpy.vstack((matching[:, 0].y, matching[:, 0].x))).T + offset.T - \
                numpy.vstack((matching[:, 1].y, matching[:, 1].x)).T
            rms = numpy.sqrt((corr * corr).sum(axis=-1).mean())
            return {"result": result, "keypoint": kp, "matching": matching, "offset": offset, "matrix": matrix,
                    "rms": rms}
        return result

    __call__ = align

    @staticmethod
    def _fit(matching):  # alignment.py:278-282
        transform_matrix = matching_correction(matching)
        offset = numpy.array([transform_matrix[5], transform_matrix[2]], dtype=numpy.float32)
        matrix = numpy.empty((2, 2), dtype=numpy.float32)
        matrix[0, 0], matrix[0, 1] = transform_matrix[4], transform_matrix[3]
        matrix[1, 0], matrix[1, 1] = transform_matrix[1], transform_matrix[0]
        return matrix, offset

    def log_profile(self):
        """Print the timing of the underlying plans (reference alignment.py:363-376)."""
        self.sift.log_profile()
        self.match.log_profile()
