"""ctypes binding of libsiftb200.so (C ABI: include/siftb.h).

There is NO CPU fallback: if the CUDA library is missing or cannot be loaded the import of the
operator classes fails loudly.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C sift_pyocl_b200/csrc``.
"""
import ctypes
import os

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsiftb200.so")

dtype_kp = numpy.dtype([('x', numpy.float32),
                        ('y', numpy.float32),
                        ('scale', numpy.float32),
                        ('angle', numpy.float32),
                        ('desc', (numpy.uint8, 128))])  # reference plan.py:110-115

SIFTB_EOVERFLOW = -4

DTYPE_CODES = {numpy.dtype(numpy.float32): 0, numpy.dtype(numpy.uint8): 1, numpy.dtype(numpy.uint16): 2,
               numpy.dtype(numpy.uint32): 3, numpy.dtype(numpy.uint64): 4, numpy.dtype(numpy.int32): 5,
               numpy.dtype(numpy.int64): 6, numpy.dtype(numpy.float64): 7}
RGB_CODE = 8

c_void_p, c_int, c_float, c_double, c_u64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double,
                                             ctypes.c_uint64)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_float_p = ctypes.POINTER(ctypes.c_float)

# every symbol include/siftb.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "siftb_last_error": (ctypes.c_char_p, []),
    "siftb_version": (c_int, []),
    "siftb_device_count": (c_int, [c_int_p]),
    "siftb_host_alloc": (c_int, [ctypes.POINTER(c_void_p), c_u64]),
    "siftb_host_alloc_wc": (c_int, [ctypes.POINTER(c_void_p), c_u64]),
    "siftb_host_free": (c_int, [c_void_p]),
    "siftb_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_double, c_int, ctypes.POINTER(c_void_p)]),
    "siftb_plan_destroy": (c_int, [c_void_p]),
    "siftb_plan_octaves": (c_int, [c_void_p]),
    "siftb_plan_kpsize": (c_int, [c_void_p]),
    "siftb_plan_capacity": (c_int, [c_void_p]),
    "siftb_plan_octave_shape": (c_int, [c_void_p, c_int, c_int_p, c_int_p]),
    "siftb_plan_device_bytes": (c_u64, [c_void_p]),
    "siftb_plan_stream": (c_void_p, [c_void_p]),
    "siftb_plan_set_profile": (c_int, [c_void_p, c_int]),
    "siftb_plan_set_variant": (c_int, [c_void_p, c_int]),
    "siftb_plan_launches": (c_u64, [c_void_p]),
    "siftb_plan_device": (c_int, [c_void_p]),
    "siftb_plan_hold_records": (c_int, [c_void_p, c_void_p]),
    "siftb_plan_wait_stream": (c_int, [c_void_p, c_void_p]),
    "siftb_plan_keypoints": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int_p, c_int_p, c_float_p]),
    "siftb_plan_submit": (c_int, [c_void_p, c_void_p, c_int]),
    "siftb_plan_collect": (c_int, [c_void_p, c_void_p, c_int, c_int_p, c_int_p, c_float_p]),
    "siftb_plan_result_dev": (c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "siftb_plan_fetch_records": (c_int, [c_void_p, c_void_p, c_int, c_int_p]),
    "siftb_plan_events": (c_int, [c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_char_p)),
                                  ctypes.POINTER(c_float_p), c_int_p]),
    "siftb_plan_stage_counts": (c_int, [c_void_p, c_int_p]),
    "siftb_gauss_taps": (c_int, [c_double, c_float_p, c_int, c_int_p]),
    "siftb_minmax": (c_int, [c_void_p, c_int, c_int, c_float_p, c_float_p]),
    "siftb_normalize": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "siftb_to_float": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "siftb_blur": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "siftb_pyramid_octave": (c_int, [c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]),
    "siftb_gradient": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "siftb_local_maxmin": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int_p]),
    "siftb_interp": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_int_p]),
    "siftb_orientation": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                  c_int_p]),
    "siftb_descriptor": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "siftb_orientation_v": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                    c_int_p, c_int]),
    "siftb_descriptor_v": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int]),
    "siftb_matcher_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "siftb_matcher_destroy": (c_int, [c_void_p]),
    "siftb_matcher_set_profile": (c_int, [c_void_p, c_int]),
    "siftb_matcher_stream": (c_void_p, [c_void_p]),
    "siftb_matcher_set_metric": (c_int, [c_void_p, c_int]),
    "siftb_matcher_set_list": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int]),
    "siftb_matcher_run": (c_int, [c_void_p, c_float, c_int, c_void_p, c_int_p]),
    "siftb_matcher_pairs": (c_int, [c_void_p, c_void_p]),
    "siftb_matcher_pair_coords": (c_int, [c_void_p, c_void_p]),
    "siftb_matcher_pair_records": (c_int, [c_void_p, c_void_p]),
    "siftb_matcher_events": (c_int, [c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_char_p)),
                                     ctypes.POINTER(c_float_p), c_int_p, c_int]),
    "siftb_plan_warp_last": (c_int, [c_void_p, c_float_p, c_float_p, c_float, c_int, c_void_p, c_int, c_int, c_int]),
    "siftb_comm_unique_id": (c_int, [c_void_p]),
    "siftb_comm_init": (c_int, [c_int, c_int, c_void_p, c_int, ctypes.POINTER(c_void_p)]),
    "siftb_comm_destroy": (c_int, [c_void_p]),
    "siftb_allgather_kp": (c_int, [c_void_p, c_void_p, c_int, c_int_p, c_void_p, c_int, c_int_p]),
    "siftb_match_l1": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_int, c_int, c_void_p, c_int, c_int_p]),
    "siftb_transform": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float_p, c_float_p, c_float,
                                c_int, c_int]),
    "siftb_transform_rgb": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float_p, c_float_p, c_float,
                                    c_int, c_int]),
}

_lib = None


class SiftB200Error(RuntimeError):
    """Error reported by libsiftb200.so."""


def load():
    """Load libsiftb200.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("sift_pyocl_b200: CUDA library %s not built (run __graft_entry__.build() or "
                          "`make -C sift_pyocl_b200/csrc`); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, exc=SiftB200Error):
    """Map a C-ABI return code to the exception the reference would raise."""
    if rc == 0:
        return
    msg = load().siftb_last_error().decode("utf-8", "replace")
    if rc == -3:
        raise MemoryError(msg)  # reference plan.py:365-366
    raise exc("%s (code %d)" % (msg, rc))


def ptr(a):
    return a.ctypes.data_as(c_void_p)


def pair_records(kp1, idx1, kp2, idx2):
    """recarray (m, 2) of dtype_kp with [:, 0] = kp1[idx1] and [:, 1] = kp2[idx2] (the result layout of the
    reference's MatchPlan.match, match.py:267-270), gathered as raw 144-byte rows: numpy copies a structured
    array field by field, which is 3x slower than the memcpy this is."""
    m = len(idx1)
    out = numpy.empty((m, 2), dtype=dtype_kp)
    if m:
        ov = out.view(numpy.uint8).reshape(m, 2, dtype_kp.itemsize)
        for col, (kp, idx) in enumerate(((kp1, idx1), (kp2, idx2))):
            rows = numpy.ascontiguousarray(kp, dtype=dtype_kp).view(numpy.uint8).reshape(-1, dtype_kp.itemsize)
            ov[:, col, :] = numpy.take(rows, idx, axis=0)
    return out.view(numpy.recarray)


def device_pointer(obj):
    """Device pointer of a CUDA-resident array (torch tensor / __cuda_array_interface__), else None."""
    if hasattr(obj, "is_cuda") and hasattr(obj, "data_ptr"):
        return int(obj.data_ptr()) if obj.is_cuda else None
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        return int(cai["data"][0])
    return None


def device_info(obj):
    """(device ordinal or None if unknown, producer stream handle or None) of a CUDA-resident array.

    torch tensors: the tensor's device and torch's current stream on it (work enqueued there may still be
    writing the tensor).  ``__cuda_array_interface__`` v3: the optional ``stream`` entry (None = already
    synchronised, 1 = legacy default stream, 2 = per-thread default stream, else a cudaStream_t)."""
    if hasattr(obj, "is_cuda") and hasattr(obj, "data_ptr"):
        import torch
        return obj.device.index, int(torch.cuda.current_stream(obj.device).cuda_stream)
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        stream = cai.get("stream", None)
        dev = getattr(getattr(obj, "device", None), "id", None)  # cupy
        return dev, (int(stream) if stream is not None else None)
    return None, None


def pinned_empty(shape, dtype, write_combined=False):
    """numpy array backed by page-locked host memory (asynchronous H<->D copies).  ``write_combined``: for
    upload-only buffers (fill them once, never read them back on the CPU)."""
    lib = load()
    dtype = numpy.dtype(dtype)
    nbytes = int(numpy.prod(shape)) * dtype.itemsize
    p = c_void_p()
    alloc = lib.siftb_host_alloc_wc if write_combined else lib.siftb_host_alloc
    check(alloc(ctypes.byref(p), max(nbytes, 1)))
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
    arr = numpy.frombuffer(buf, dtype=dtype, count=int(numpy.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p.value
    return arr


_PINNED = {}


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p is not None:
        load().siftb_host_free(c_void_p(p))
