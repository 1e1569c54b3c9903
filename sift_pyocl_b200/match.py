"""MatchPlan: brute-force keypoint matching -- same public surface as the reference's sift-src/match.py.

The metric is the reference's: L1 distance on the uint8 descriptors with a ratio test on the two best
distances, threshold ``par.MatchRatio ** 2`` (matching_gpu.cl:79-99, match.py:253).

Like the reference's plan (match.py:129-160) a MatchPlan owns persistent device buffers; keypoint lists may
be host recarrays or device-resident arrays (``DeviceRecords``, torch uint8 tensors, anything with
``__cuda_array_interface__``) in place of the reference's ``pyopencl.array.Array`` inputs (match.py:216-239).
"""
import ctypes
import logging
import threading

import numpy

from . import _lib
from .param import par

logger = logging.getLogger("sift.match")


class DeviceRecords(object):
    """``n`` keypoint records (dtype_kp, 144 bytes each) living in CUDA device memory: the stand-in for a
    ``pyopencl.array.Array`` of keypoints (reference match.py:211, alignment.py:157).  ``owner`` keeps the
    memory alive; ``get()`` copies the records to a host recarray."""

    def __init__(self, ptr, n, device=0, owner=None):
        self.ptr, self.size, self.device, self.owner = int(ptr), int(n), int(device), owner
        self.shape = (self.size,)
        self.dtype = _lib.dtype_kp
        self.__cuda_array_interface__ = {"shape": (max(self.size, 1) * 144,), "typestr": "|u1",
                                         "data": (self.ptr, False), "version": 3, "strides": None}

    def get(self):
        import torch
        t = torch.as_tensor(self, device="cuda:%d" % self.device)[:self.size * 144]
        return t.cpu().numpy().view(_lib.dtype_kp).view(numpy.recarray)


def _as_list(kp):
    """(pointer, n, on_device, keepalive) of a keypoint list given as host recarray or device array."""
    dptr = _lib.device_pointer(kp)
    if dptr is not None:
        if isinstance(kp, DeviceRecords):
            return ctypes.c_void_p(dptr), kp.size, 1, kp
        if hasattr(kp, "is_contiguous") and not kp.is_contiguous():
            kp = kp.contiguous()
            dptr = _lib.device_pointer(kp)
        nbytes = kp.numel() * kp.element_size() if hasattr(kp, "numel") else int(numpy.prod(kp.shape)) * kp.dtype.itemsize
        assert nbytes % 144 == 0, "device keypoint arrays must hold whole 144-byte dtype_kp records"
        return ctypes.c_void_p(dptr), nbytes // 144, 1, kp
    host = numpy.ascontiguousarray(kp, dtype=_lib.dtype_kp)
    return _lib.ptr(host), host.size, 0, host


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
        self._matcher = None
        self._held = [None, None]   # lists kept resident by hold()
        self._sizes = [0, 0]
        self._n_pairs = 0
        lib = _lib.load()
        handle = ctypes.c_void_p()
        _lib.check(lib.siftb_matcher_create(self.device, ctypes.byref(handle)), RuntimeError)
        self._matcher = handle
        self.queue = lib.siftb_matcher_stream(handle)
        if self.profile:
            lib.siftb_matcher_set_profile(handle, 1)
        self._metric = "l1"
        if roi is not None:
            self.set_roi(roi)

    @property
    def metric(self):
        """Distance between descriptors: "l1" (the reference's, matching_gpu.cl:79-99; default) or "l2" (squared
        Euclidean distance with the ratio test on the squared values -- an extra, not the reference's arithmetic)."""
        return self._metric

    @metric.setter
    def metric(self, name):
        name = str(name).lower()
        assert name in ("l1", "l2")
        with self._sem:
            _lib.check(_lib.load().siftb_matcher_set_metric(self._matcher, 1 if name == "l2" else 0))
            self._metric = name

    def __del__(self):
        m, self._matcher = getattr(self, "_matcher", None), None
        if m:
            try:
                _lib.load().siftb_matcher_destroy(m)
            except Exception:  # interpreter shutdown
                pass

    # ------------------------------------------------------------------------------------------
    def _load(self, which, kp):
        pointer, n, on_device, keep = _as_list(kp)
        _lib.check(_lib.load().siftb_matcher_set_list(self._matcher, which, pointer, n, on_device))
        self._sizes[which] = n
        del keep

    def hold(self, which, kp):
        """Copy the keypoint list ``kp`` into the plan's device buffer ``which`` (0: first, 1: second argument
        of match()) and keep it there: later ``match()`` calls that pass the SAME object for that argument skip
        the upload.  This is the reference's ``ref_kp_gpu`` (alignment.py:157): LinearAlign sends its reference
        keypoints once, not once per frame."""
        with self._sem:
            self._load(which, kp)
            self._held[which] = kp

    def release(self, which=None):
        """Forget the list(s) kept by hold()."""
        with self._sem:
            for w in ((0, 1) if which is None else (which,)):
                self._held[w] = None

    def _run(self, nkp1, nkp2):
        """Load the lists (unless held), run the matching kernel; returns the number of stored pairs."""
        for which, kp in ((0, nkp1), (1, nkp2)):
            if kp is not self._held[which] or kp is None:
                self._held[which] = None
                self._load(which, kp)
        n1, n2 = self._sizes
        if min(n1, n2) > self.kpsize:  # match.py:241-243
            self.kpsize = min(n1, n2)
        n = ctypes.c_int()
        _lib.check(_lib.load().siftb_matcher_run(self._matcher, numpy.float32(par.MatchRatio * par.MatchRatio),
                                                 self.kpsize, None, ctypes.byref(n)))
        self._n_pairs = min(n.value, self.kpsize)
        return self._n_pairs

    def match(self, nkp1, nkp2, raw_results=False):
        """Calculate the matching of 2 keypoint lists (reference match.py:200-272).

        :param nkp1, nkp2: numpy 1D recarray of keypoints, or an equivalent device-resident array
        :param raw_results: if true return the 2D array of indexes of matching keypoints
        """
        assert len(nkp1.shape) == 1 or _lib.device_pointer(nkp1) is not None
        assert len(nkp2.shape) == 1 or _lib.device_pointer(nkp2) is not None
        with self._sem:
            size = self._run(nkp1, nkp2)
            if raw_results:
                result = self._pairs(size)
            else:
                result = self._pair_records(size)
            self._collect_events()
        return result

    __call__ = match

    def match_coords(self, nkp1, nkp2):
        """Like match(), but returns only what a geometric fit needs: float32 [m, 8] =
        (x, y, scale, angle) of the keypoint of ``nkp1`` then of its match in ``nkp2``, gathered on the device
        (32 bytes per match cross the bus instead of the 288 of the two records)."""
        with self._sem:
            size = self._run(nkp1, nkp2)
            out = numpy.empty((size, 8), numpy.float32)
            if size:
                _lib.check(_lib.load().siftb_matcher_pair_coords(self._matcher, _lib.ptr(out)))
            self._collect_events()
        return out

    def last_pairs(self, raw_results=False):
        """Result of the most recent run again: index pairs or the (m, 2) recarray."""
        with self._sem:
            return self._pairs(self._n_pairs) if raw_results else self._pair_records(self._n_pairs)

    def _pairs(self, size):
        pairs = numpy.empty((size, 2), dtype=numpy.int32)
        if size:
            _lib.check(_lib.load().siftb_matcher_pairs(self._matcher, _lib.ptr(pairs)))
        return pairs

    def _pair_records(self, size):
        out = numpy.empty((size, 2), dtype=self.dtype_kp)
        if size:
            _lib.check(_lib.load().siftb_matcher_pair_records(self._matcher, _lib.ptr(out)))
        return out.view(numpy.recarray)

    def _collect_events(self):
        if not self.profile:
            return
        lib = _lib.load()
        names = ctypes.POINTER(ctypes.c_char_p)()
        ms = _lib.c_float_p()
        n = ctypes.c_int()
        _lib.check(lib.siftb_matcher_events(self._matcher, ctypes.byref(names), ctypes.byref(ms), ctypes.byref(n), 1))
        self.events += [(names[i].decode(), float(ms[i])) for i in range(n.value)]

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
        """If profiling is on, print the device time of every enqueued operation (match.py:329-345)."""
        t = 0.0
        for name, et in self.events:
            print("%50s:\t%.3fms" % (name, et))
            t += et
        print("_" * 80)
        print("%50s:\t%.3fms" % ("Total execution time", t))
