"""MatchPlan: brute-force keypoint matching -- same public surface as the reference's sift-src/match.py.

The metric is the reference's: L1 distance on the uint8 descriptors with a ratio test on the two best
distances, threshold ``par.MatchRatio ** 2`` (matching_gpu.cl:79-99, match.py:253).
"""
import ctypes
import logging
import threading

import numpy

from . import _lib
from .param import par

logger = logging.getLogger("sift.match")


class MatchPlan(object):
    """Plan to compare sets of SIFT keypoints and find common ones (reference match.py:52-127).

        mp = sift.MatchPlan()
        common = mp.match(kp1, kp2)      # recarray (m, 2) of dtype_kp, or int32 (m, 2) with raw_results=True
    """
    dtype_kp = _lib.dtype_kp

    def __init__(self, size=16384, devicetype="CPU", profile=False, device=None, max_workgroup_size=None,
                 roi=None, context=None):
        self.profile = bool(profile)
        self.events = []
        self.kpsize = int(size)
        self.buffers = {}
        self.programs = {}
        self.memory = None
        self.octave_max = None
        self.red_size = None
        self.ctx = context
        self.roi = None
        if device is None:
            self.device = 0
        elif "__len__" in dir(device):
            self.device = int(device[-1])
        else:
            self.device = int(device)
        self.devicetype = "GPU"
        self.max_workgroup_size = max_workgroup_size
        self._sem = threading.Semaphore()
        _lib.load()
        if roi is not None:
            self.set_roi(roi)

    def match(self, nkp1, nkp2, raw_results=False):
        """Calculate the matching of 2 keypoint lists (reference match.py:200-272).

        :param nkp1, nkp2: numpy 1D recarray of keypoints
        :param raw_results: if true return the 2D array of indexes of matching keypoints
        """
        assert len(nkp1.shape) == 1
        assert len(nkp2.shape) == 1
        valid_types = (numpy.ndarray, numpy.recarray)
        assert isinstance(nkp1, valid_types)
        assert isinstance(nkp2, valid_types)
        with self._sem:
            k1 = numpy.ascontiguousarray(nkp1, dtype=self.dtype_kp)
            k2 = numpy.ascontiguousarray(nkp2, dtype=self.dtype_kp)
            if min(k1.size, k2.size) > self.kpsize:  # match.py:241-243
                self.kpsize = min(k1.size, k2.size)
            pairs = numpy.empty((self.kpsize, 2), dtype=numpy.int32)
            n = ctypes.c_int()
            lib = _lib.load()
            _lib.check(lib.siftb_match_l1(_lib.ptr(k1), k1.size, _lib.ptr(k2), k2.size,
                                          numpy.float32(par.MatchRatio * par.MatchRatio), 0, self.device,
                                          _lib.ptr(pairs), self.kpsize, ctypes.byref(n)))
            size = min(n.value, self.kpsize)
            match = pairs[:size].copy()
            if raw_results:
                result = match
            else:
                result = _lib.pair_records(nkp1, match[:size, 0], nkp2, match[:size, 1])
        return result

    __call__ = match

    def set_roi(self, roi):
        """Define the region of interest (stored; like the reference it is not used by match(),
        match.py:312-321 -- LinearAlign filters keypoints by ROI on the host)."""
        with self._sem:
            self.roi = numpy.ascontiguousarray(roi, numpy.int8)

    def unset_roi(self):
        """Unset the region of interest (reference match.py:323-327)."""
        with self._sem:
            self.roi = None

    def reset_timer(self):
        with self._sem:
            self.events = []

    def log_profile(self):
        for name, et in self.events:
            print("%50s:\t%.3fms" % (name, et))
