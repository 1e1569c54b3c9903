"""SiftPlan: keypoints of an image -- same public surface as the reference's sift-src/plan.py.

    import sift_pyocl_b200 as sift
    plan = sift.SiftPlan(template=img)           # or shape=..., dtype=...
    kp = plan.keypoints(img)                      # numpy.recarray of dtype_kp (x, y, scale, angle, desc)

Everything under the class boundary is new: the host stays Python and calls hand-written sm_100a
CUDA through the C ABI of include/siftb.h (ctypes).  There is no PyOpenCL and no CPU fallback.
"""
import ctypes
import logging
import threading
import time

import numpy

from . import _lib
from .param import par
from .utils import kernel_size

logger = logging.getLogger("sift.plan")


class _DeviceScalar(object):
    """Stand-in for the reference's ``plan.buffers["min"]`` pyopencl array (alignment.py:345 calls
    ``.get()[0]`` on it)."""

    def __init__(self):
        self.value = numpy.zeros(1, numpy.float32)

    def get(self):
        return self.value.copy()


class SiftPlan(object):
    """How to calculate a set of SIFT keypoints on an image (reference plan.py:70-119).

    Keyword arguments keep the reference's names and meaning.  ``devicetype``,
    ``max_workgroup_size`` and ``context`` (OpenCL notions) are accepted and ignored; ``device``
    selects the CUDA device: an integer ordinal, or the reference's (platform, device) 2-tuple of
    which the second entry is used.
    """
    converter = {numpy.dtype(numpy.uint8): "u8_to_float",
                 numpy.dtype(numpy.uint16): "u16_to_float",
                 numpy.dtype(numpy.uint32): "u32_to_float",
                 numpy.dtype(numpy.uint64): "u64_to_float",
                 numpy.dtype(numpy.int32): "s32_to_float",
                 numpy.dtype(numpy.int64): "s64_to_float"}
    sigmaRatio = 2.0 ** (1.0 / par.Scales)
    PIX_PER_KP = 10  # pre-allocate one keypoint slot per 10 pixels (reference plan.py:109)
    dtype_kp = _lib.dtype_kp

    def __init__(self, shape=None, dtype=None, devicetype="CPU", template=None,
                 profile=False, device=None, PIX_PER_KP=None,
                 max_workgroup_size=None, context=None, init_sigma=None):
        if init_sigma is None:
            init_sigma = par.InitSigma
        self._init_sigma = float(init_sigma)
        self.buffers = {"min": _DeviceScalar(), "max": _DeviceScalar()}
        self.programs = {}
        self._plan = None
        if template is not None:
            self.shape = tuple(template.shape)
            self.dtype = _np_dtype(template.dtype)
        else:
            self.shape = tuple(shape)
            self.dtype = numpy.dtype(dtype)
        if len(self.shape) == 3:
            self.RGB = True
            self.shape = self.shape[:2]
        elif len(self.shape) == 2:
            self.RGB = False
        else:
            raise RuntimeError("Unable to process image of shape %s" % (tuple(self.shape),))
        if PIX_PER_KP:
            self.PIX_PER_KP = int(PIX_PER_KP)
        self.profile = bool(profile)
        self.events = []
        self._sem = threading.Semaphore()
        # The reference runs orientation_gpu.cl + keypoints_gpu2.cl on devicetype "GPU" and the *_cpu.cl kernels on
        # "CPU" (plan.py:667-725); the two families give slightly different angles / descriptors.
        # Everything runs on the B200 here; ``devicetype`` selects whose NUMBERS are reproduced: "CPU" (the
        # reference's default, and the parity target of this package) or "GPU".
        self.devicetype = str(devicetype).upper() if devicetype is not None else "CPU"
        self.variant = "gpu" if self.devicetype == "GPU" else "cpu"
        self.max_workgroup_size = max_workgroup_size
        self.ctx = context
        self.queue = None
        if device is None:
            self.device = 0
        elif "__len__" in dir(device):
            self.device = int(device[-1])
        else:
            self.device = int(device)
        if self.RGB:
            if self.dtype != numpy.uint8:
                raise RuntimeError("invalid input format error (%s)" % (str(self.dtype)))
            code = _lib.RGB_CODE
        elif self.dtype in _lib.DTYPE_CODES:
            code = _lib.DTYPE_CODES[self.dtype]
        else:
            raise RuntimeError("invalid input format error (%s)" % (str(self.dtype)))
        self._code = code
        lib = _lib.load()
        handle = ctypes.c_void_p()
        octave_limit = int(par.OctaveMax) if par.OctaveMax < 64 else 0
        _lib.check(lib.siftb_plan_create(int(self.shape[0]), int(self.shape[1]), code, self.device,
                                         int(self.PIX_PER_KP), self._init_sigma, octave_limit, ctypes.byref(handle)),
                   RuntimeError)
        self._plan = handle
        if self.variant == "gpu":
            _lib.check(lib.siftb_plan_set_variant(handle, 1))
        self.kpsize = lib.siftb_plan_kpsize(handle)
        self._capacity = lib.siftb_plan_capacity(handle)
        self.octave_max = lib.siftb_plan_octaves(handle)
        self.scales = []  # in XY order, like the reference (plan.py:215)
        for o in range(self.octave_max):
            w, h = ctypes.c_int(), ctypes.c_int()
            lib.siftb_plan_octave_shape(handle, o, ctypes.byref(w), ctypes.byref(h))
            self.scales.append((numpy.int32(w.value), numpy.int32(h.value)))
        self.queue = lib.siftb_plan_stream(handle)
        if self.profile:
            lib.siftb_plan_set_profile(handle, 1)
        self.last_counts = numpy.zeros(self.octave_max, numpy.int32)
        self._out = None  # host record buffer, allocated on first use
        self._last_n = 0
        self._pending = 0
        self._keep = []
        logger.info("SiftPlan %s %s on CUDA device %d: %d octaves, kpsize %d, %.1f MB", self.shape, self.dtype,
                    self.device, self.octave_max, self.kpsize, self.memory / 1e6)

    @property
    def memory(self):
        """Device memory held by the plan, bytes (reference plan.py:226 _calc_memory).  Grows once, by the planes
        of a second image, the first time two images are in flight together (submit() / keypoints_many())."""
        return int(_lib.load().siftb_plan_device_bytes(self._plan)) if self._plan else 0

    def __del__(self):
        """Destructor: release all buffers (reference plan.py:203-211)."""
        plan, self._plan = getattr(self, "_plan", None), None
        if getattr(self, "_out_pinned", False):
            try:
                _lib.pinned_free(self._out)
            except Exception:
                pass
            self._out = None
        if plan:
            try:
                _lib.load().siftb_plan_destroy(plan)
            except Exception:  # interpreter shutdown
                pass

    # ------------------------------------------------------------------------------------------
    def _image_args(self, image):
        """Validate like the reference (plan.py:443-448) and return (pointer, flags, keepalive)."""
        assert tuple(image.shape[:2]) == self.shape
        dt = _np_dtype(image.dtype)
        assert dt in [self.dtype, numpy.dtype(numpy.float32)]
        if len(image.shape) == 3:  # interleaved colour: only uint8 RGB on an RGB plan (preprocess.cl:211 rgb_to_float)
            assert self.RGB and image.shape[2] == 3 and dt == numpy.uint8, "colour images must be (H, W, 3) uint8"
        is_f32 = dt == numpy.float32
        if self.RGB and not is_f32:
            assert len(image.shape) == 3
        flags = 0
        if is_f32 and self._code != 0:
            flags |= 2  # SIFTB_IS_F32
        self._last_colour = not is_f32 and self.RGB  # pixel format of the image the plan keeps on the device
        dptr = _lib.device_pointer(image)
        if dptr is not None:
            if hasattr(image, "is_contiguous") and not image.is_contiguous():
                image = image.contiguous()
                dptr = _lib.device_pointer(image)
            dev, stream = _lib.device_info(image)
            assert dev is None or dev == self.device, "image lives on CUDA device %s, the plan on %d" % (dev, self.device)
            if stream is not None:
                # the plan runs on its own non-blocking stream: order it after the stream that produces the image
                # (the reference shares one in-order queue with its pyopencl.array inputs, plan.py:451-456)
                _lib.check(_lib.load().siftb_plan_wait_stream(self._plan, ctypes.c_void_p(stream)))
            return ctypes.c_void_p(dptr), flags | 1, image
        if not image.flags["C_CONTIGUOUS"]:
            image = numpy.ascontiguousarray(image)
        return _lib.ptr(image), flags, image

    def _records(self, n):
        """Page-locked staging buffer for the D->H copy of ``n`` records (full PCIe speed).  It grows on demand: a
        buffer for the plan's whole capacity (2 * kpsize records = 483 MB of pinned memory at 4096 x 4096, 1.9 GB at
        8192 x 8192) would be two orders of magnitude more than images produce."""
        if self._out is None or self._out.shape[0] < n:
            if getattr(self, "_out_pinned", False):
                _lib.pinned_free(self._out)
            want = min(self._capacity, max(16384, int(1.5 * n)))
            try:
                self._out = _lib.pinned_empty((want,), self.dtype_kp)
                self._out_pinned = True
            except Exception:
                self._out = numpy.empty(want, dtype=self.dtype_kp)
                self._out_pinned = False
        return self._out

    def _collect_locked(self, records):
        """Wait for the oldest image in flight; returns its keypoints (or their number when ``records`` is false)."""
        lib = _lib.load()
        n = ctypes.c_int()
        mm = numpy.zeros(2, numpy.float32)
        rc = lib.siftb_plan_collect(self._plan, None, 0, ctypes.byref(n), self.last_counts.ctypes.data_as(_lib.c_int_p),
                                    mm.ctypes.data_as(_lib.c_float_p))
        if rc == _lib.SIFTB_EOVERFLOW:
            logger.warning("Keypoint counter overflow risk: counted %s / %s" % (n.value, self.kpsize))  # plan.py:771
        else:
            _lib.check(rc)
        self.buffers["min"].value[0], self.buffers["max"].value[0] = mm[0], mm[1]
        self._last_n = min(n.value, self._capacity)
        if not records:
            return self._last_n
        for octave, cnt in enumerate(self.last_counts):
            logger.info("in octave %i found %i kp" % (octave, cnt))  # plan.py:543
        if self.profile:
            self._fetch_events()
        return self._fetch_locked()

    def _fetch_locked(self):
        """Host recarray of the records of the most recently collected image."""
        n = self._last_n
        out = self._records(n)
        got = ctypes.c_int()
        _lib.check(_lib.load().siftb_plan_fetch_records(self._plan, _lib.ptr(out), n, ctypes.byref(got)))
        # fresh host array for the caller; copied as raw bytes (numpy copies a structured array field by field,
        # 4x slower than the memcpy this is)
        res = numpy.empty(got.value, dtype=self.dtype_kp)
        numpy.copyto(res.view(numpy.uint8), out[:got.value].view(numpy.uint8))
        return res.view(numpy.recarray)

    def keypoints(self, image):
        """Calculates the keypoints of the image (reference plan.py:432-567).

        :param image: 2-D array (3-D if RGB); numpy array, or a CUDA-resident array
                      (torch tensor / ``__cuda_array_interface__``) in place of a pyopencl Array
        :return: vector of keypoints, ``numpy.recarray`` of ``dtype_kp``
        """
        self.reset_timer()
        with self._sem:
            t0 = time.time()
            assert self._pending == 0, "images are in flight: collect() them first"
            pointer, flags, keep = self._image_args(image)
            _lib.check(_lib.load().siftb_plan_submit(self._plan, pointer, flags))
            output = self._collect_locked(True)
            del keep
            logger.info("Execution time: %.3fms" % (1000 * (time.time() - t0)))
        return output

    __call__ = keypoints

    # -- split form, for callers that overlap copies with compute (no reference equivalent) -----
    def submit(self, image):
        """Enqueue copy + all kernels for ``image`` and return immediately; pair with collect().

        Up to three images may be in flight: the H->D copies of images k+1, k+2 and the D->H copy of the records
        of image k then overlap the kernels of another image (results come back in submission order).
        Host images should live in page-locked memory (``pinned_empty``) for the copies to be asynchronous.
        """
        with self._sem:
            assert self._pending < 3, "three images are already in flight: call collect() first"
            pointer, flags, keep = self._image_args(image)
            _lib.check(_lib.load().siftb_plan_submit(self._plan, pointer, flags))
            self._keep.append(keep)
            self._pending += 1

    def collect(self, records=True):
        """Wait for the oldest submitted image and return its keypoints.  With ``records=False`` the records
        stay on the device (see device_records()) and only their number is returned."""
        with self._sem:
            assert self._pending > 0, "collect() without submit()"
            try:
                return self._collect_locked(records)
            finally:
                self._pending -= 1
                self._keep.pop(0)

    def keypoints_many(self, images):
        """Generator: keypoints of every image of ``images`` in order, with the copies of one image
        overlapping the kernels of the others (three images in flight)."""
        try:
            for image in images:
                self.submit(image)
                if self._pending == 3:
                    yield self.collect()
            while self._pending:
                yield self.collect()
        finally:
            # consumer stopped early / a submit raised: drain what is still in flight so that the plan stays usable
            while self._pending:
                self.collect(records=False)

    @staticmethod
    def pinned_empty(shape, dtype=numpy.float32, write_combined=False):
        """numpy array in page-locked host memory (asynchronous H<->D copies); ``write_combined`` for buffers
        that are only filled by the CPU and read by the device."""
        return _lib.pinned_empty(shape, dtype, write_combined)

    def device_records(self):
        """(device pointer of the record array, device pointer of the int32 record count) of the last
        run; valid until the next submit()/keypoints() on this plan."""
        recs, cnt = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(_lib.load().siftb_plan_result_dev(self._plan, ctypes.byref(recs), ctypes.byref(cnt)))
        return recs.value, cnt.value

    def wait_stream(self, stream):
        """Order the plan's queue after the work enqueued so far on ``stream`` (a cudaStream_t handle, e.g.
        ``torch.cuda.current_stream().cuda_stream``): needed when another stream produces a device-resident
        input image, or still reads the plan's device-resident records (device_records())."""
        _lib.check(_lib.load().siftb_plan_wait_stream(self._plan, ctypes.c_void_p(int(stream))))

    def hold_records(self, stream):
        """The device-resident records of the last collected image (device_records()) are not overwritten before
        the work enqueued so far on ``stream`` has finished.  Unlike wait_stream() this does not hold up the images
        already queued or the next ones: only the submit() that recycles the buffer (the third after the one that produced these records) waits."""
        _lib.check(_lib.load().siftb_plan_hold_records(self._plan, ctypes.c_void_p(int(stream))))

    def device_keypoints(self, n=None):
        """The keypoints of the last run as a device-resident array (``match.DeviceRecords``) -- what the reference
        gets by keeping results in a ``pyopencl.array`` (alignment.py:246-249); valid until the third submit()
        after the one that produced them (the plan cycles through three record buffers).  MatchPlan.match() accepts it
        directly, so the records never visit the host."""
        from .match import DeviceRecords
        recs, _ = self.device_records()
        return DeviceRecords(recs or 0, self._last_n if n is None else n, self.device, owner=self)

    def fetch_keypoints(self):
        """Host recarray of the keypoints of the last run (for callers that used ``collect(records=False)``)."""
        with self._sem:
            return self._fetch_locked()

    def warp_last(self, matrix, offset, fill, out_shape=None, mode=1, out=None):
        """Affine warp (transform.cl:22 / :116) of the image of the last run, which is still on the device: the
        reference uploads a frame once for both SIFT and the warp (alignment.py:242-246, 336-349).

        :param matrix, offset: output pixel (y, x) samples the image at ``matrix . (y, x) + offset``
        :param out: optional preallocated result (e.g. page-locked, see pinned_empty)
        """
        m = numpy.ascontiguousarray(numpy.asarray(matrix, numpy.float32).reshape(4))
        o = numpy.ascontiguousarray(numpy.asarray(offset, numpy.float32).reshape(2))
        oh, ow = self.shape if out_shape is None else (int(out_shape[0]), int(out_shape[1]))
        shape, dt = ((oh, ow, 3), numpy.uint8) if getattr(self, "_last_colour", False) else ((oh, ow), numpy.float32)
        if out is None:
            out = numpy.empty(shape, dt)
        assert out.shape == shape and out.dtype == dt and out.flags["C_CONTIGUOUS"]
        with self._sem:
            _lib.check(_lib.load().siftb_plan_warp_last(self._plan, m.ctypes.data_as(_lib.c_float_p),
                                                        o.ctypes.data_as(_lib.c_float_p), ctypes.c_float(fill),
                                                        int(mode), _lib.ptr(out), oh, ow, 0))
        return out

    @property
    def launches(self):
        """CUDA kernels launched by this plan so far."""
        return int(_lib.load().siftb_plan_launches(self._plan))

    def set_profile(self, enable):
        self.profile = bool(enable)
        _lib.load().siftb_plan_set_profile(self._plan, int(self.profile))

    def fetch_events(self):
        """[(stage name, device milliseconds)] of the last run when profiling is on."""
        self._fetch_events()
        return self.events

    # ------------------------------------------------------------------------------------------
    def stage_counts(self):
        """int[octave, scale-1, 3]: extrema found, kept after interpolation, after orientation
        assignment -- the counters the reference reads back at plan.py:642, 782 and 689."""
        c = numpy.zeros((self.octave_max, 3, 3), numpy.int32)
        _lib.check(_lib.load().siftb_plan_stage_counts(self._plan, c.ctypes.data_as(_lib.c_int_p)))
        return c

    def _fetch_events(self):
        lib = _lib.load()
        names = ctypes.POINTER(ctypes.c_char_p)()
        ms = _lib.c_float_p()
        n = ctypes.c_int()
        _lib.check(lib.siftb_plan_events(self._plan, ctypes.byref(names), ctypes.byref(ms), ctypes.byref(n)))
        self.events = [(names[i].decode(), float(ms[i])) for i in range(n.value)]

    def count_kp(self, output):
        """Print the number of keypoint per octave (reference plan.py:811-821)."""
        kpt = 0
        for octave, ksum in enumerate(self.last_counts):
            kpt += ksum
            print("octave %i kp count %i/%i size %s ratio:%s" % (octave, ksum, self.kpsize, self.scales[octave],
                                                                 1000.0 * ksum / self.scales[octave][1] / self.scales[octave][0]))
        print("Found total %i guess %s pixels per keypoint" % (kpt, self.shape[0] * self.shape[1] / max(kpt, 1)))

    def log_profile(self):
        """If profiling is on, print the device time of every stage (reference plan.py:826-847)."""
        t = orient = descr = 0.0
        if self.profile:
            for name, et in self.events:
                print("%50s:\t%.3fms" % (name, et))
                t += et
                if "orient" in name:
                    orient += et
                if "descriptors" in name:
                    descr += et
        print("_" * 80)
        print("%50s:\t%.3fms" % ("Total execution time", t))
        print("%50s:\t%.3fms" % ("Total Orientation assignment", orient))
        print("%50s:\t%.3fms" % ("Total Descriptors", descr))

    def reset_timer(self):
        """Resets the profiling timers (reference plan.py:849-854)."""
        with self._sem:
            self.events = []


def _np_dtype(dt):
    """numpy dtype of a numpy / torch / cupy dtype object."""
    try:
        return numpy.dtype(dt)
    except TypeError:
        name = str(dt).split(".")[-1]  # torch.float32 -> float32
        return numpy.dtype(name)


def gaussian_taps(sigma):
    """Normalised Gaussian taps the plan uses for ``sigma`` (reference plan.py:308-340)."""
    lib = _lib.load()
    taps = numpy.zeros(64, numpy.float32)
    n = ctypes.c_int()
    _lib.check(lib.siftb_gauss_taps(float(sigma), taps.ctypes.data_as(_lib.c_float_p), 64, ctypes.byref(n)))
    assert n.value == kernel_size(sigma, True)
    return taps[:n.value].copy()
