"""numpy front-ends of the stage-level C-ABI entry points (include/siftb.h "stage-level").

One function per reference kernel, used by the per-kernel parity tests exactly like the reference's
test/test_*.py exercise each .cl file.  All compute happens in libsiftb200.so on the GPU.
"""
import ctypes

import numpy

from . import _lib

_f32 = lambda a: numpy.ascontiguousarray(a, dtype=numpy.float32)  # noqa: E731


def minmax(img):
    img = _f32(img)
    mn, mx = ctypes.c_float(), ctypes.c_float()
    _lib.check(_lib.load().siftb_minmax(_lib.ptr(img), img.shape[0], img.shape[1], ctypes.byref(mn), ctypes.byref(mx)))
    return mn.value, mx.value


def normalize(img):
    img = _f32(img)
    out = numpy.empty_like(img)
    _lib.check(_lib.load().siftb_normalize(_lib.ptr(img), img.shape[0], img.shape[1], _lib.ptr(out)))
    return out


def to_float(img):
    img = numpy.ascontiguousarray(img)
    code = _lib.RGB_CODE if img.ndim == 3 else _lib.DTYPE_CODES[img.dtype]
    out = numpy.empty(img.shape[:2], numpy.float32)
    _lib.check(_lib.load().siftb_to_float(_lib.ptr(img), code, img.shape[0], img.shape[1], _lib.ptr(out)))
    return out


def blur(img, taps):
    img, taps = _f32(img), _f32(taps)
    out = numpy.empty_like(img)
    _lib.check(_lib.load().siftb_blur(_lib.ptr(img), img.shape[0], img.shape[1], _lib.ptr(taps), taps.size,
                                      _lib.ptr(out)))
    return out


def pyramid_octave(g0, init_sigma=1.6):
    """(G[0..5], DoG[0..4], G[3][::2, ::2]) of one octave."""
    g0 = _f32(g0)
    h, w = g0.shape
    G = numpy.empty((6, h, w), numpy.float32)
    G[0] = g0
    D = numpy.empty((5, h, w), numpy.float32)
    nxt = numpy.empty((h // 2, w // 2), numpy.float32)
    _lib.check(_lib.load().siftb_pyramid_octave(_lib.ptr(g0), h, w, init_sigma, _lib.ptr(G[1:]), _lib.ptr(D),
                                                _lib.ptr(nxt)))
    return G, D, nxt


def gradient(img):
    img = _f32(img)
    grad, ori = numpy.empty_like(img), numpy.empty_like(img)
    _lib.check(_lib.load().siftb_gradient(_lib.ptr(img), img.shape[0], img.shape[1], _lib.ptr(grad), _lib.ptr(ori)))
    return grad, ori


def local_maxmin(dogs, scale, octsize=1, cap=None):
    dogs = _f32(dogs)
    _, h, w = dogs.shape
    cap = h * w // 10 if cap is None else cap
    kp = numpy.zeros((cap, 4), numpy.float32)
    n = ctypes.c_int()
    _lib.check(_lib.load().siftb_local_maxmin(_lib.ptr(dogs), h, w, scale, octsize, _lib.ptr(kp), cap, ctypes.byref(n)))
    return kp[:min(n.value, cap)], n.value


def interp(dogs, kp, init_sigma=1.6):
    dogs, kp = _f32(dogs), _f32(kp)
    _, h, w = dogs.shape
    out = numpy.zeros_like(kp)
    n = ctypes.c_int()
    _lib.check(_lib.load().siftb_interp(_lib.ptr(dogs), h, w, _lib.ptr(kp), kp.shape[0], init_sigma, _lib.ptr(out),
                                        ctypes.byref(n)))
    return out[:n.value]


VARIANTS = {"cpu": 0, "gpu": 1}  # orientation_cpu.cl / keypoints_cpu.cl, or orientation_gpu.cl / keypoints_gpu2.cl


def orientation(kp, grad, ori, octsize=1, cap=None, variant="cpu"):
    kp, grad, ori = _f32(kp), _f32(grad), _f32(ori)
    n = kp.shape[0]
    cap = max(4 * n, 16) if cap is None else cap
    out = numpy.zeros((cap, 4), numpy.float32)
    m = ctypes.c_int()
    _lib.check(_lib.load().siftb_orientation_v(_lib.ptr(kp), n, _lib.ptr(grad), _lib.ptr(ori), grad.shape[0],
                                               grad.shape[1], octsize, _lib.ptr(out), cap, ctypes.byref(m),
                                               VARIANTS[variant]))
    return out[:min(m.value, cap)], m.value


def descriptor(kp, grad, ori, octsize=1, variant="cpu"):
    kp, grad, ori = _f32(kp), _f32(grad), _f32(ori)
    desc = numpy.zeros((kp.shape[0], 128), numpy.uint8)
    _lib.check(_lib.load().siftb_descriptor_v(_lib.ptr(kp), kp.shape[0], _lib.ptr(grad), _lib.ptr(ori), grad.shape[0],
                                              grad.shape[1], octsize, _lib.ptr(desc), VARIANTS[variant]))
    return desc
