"""ctypes front-end of the CPU oracle (oracle/siftref.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module; the product package ``sift_pyocl_b200`` never does.

Function names follow the reference kernels they restate (openCL/*.cl, sift-src/plan.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsiftref.so")

dtype_kp = np.dtype([("x", np.float32), ("y", np.float32), ("scale", np.float32), ("angle", np.float32),
                     ("desc", (np.uint8, 128))])  # plan.py:110-115

_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_u8_p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile oracle/libsiftref.so with the committed Makefile (gcc, seconds)."""
    src = os.path.join(_HERE, "siftref.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libsiftref.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.siftref_kernel_size.argtypes = [ctypes.c_double, ctypes.c_int]
        _lib.siftref_gaussian_taps.argtypes = [ctypes.c_double, ctypes.c_int, _c_float_p]
        _lib.siftref_gaussian_taps.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(_c_float_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return lib().siftref_num_threads()


def set_num_threads(n):
    lib().siftref_set_num_threads(int(n))


def kernel_size(sigma, odd=True):
    return lib().siftref_kernel_size(float(sigma), int(odd))


def gaussian_taps(sigma, size=None):
    size = kernel_size(sigma) if size is None else int(size)
    out = np.empty(size, np.float32)
    lib().siftref_gaussian_taps(float(sigma), size, _fp(out))
    return out


def num_octaves(h, w):
    return lib().siftref_num_octaves(int(h), int(w))


_DTYPE_CODE = {np.dtype(np.float32): 0, np.dtype(np.uint8): 1, np.dtype(np.uint16): 2, np.dtype(np.uint32): 3,
               np.dtype(np.uint64): 4, np.dtype(np.int32): 5, np.dtype(np.int64): 6, np.dtype(np.float64): 7}


def to_float(img):
    img = np.ascontiguousarray(img)
    if img.ndim == 3:
        assert img.dtype == np.uint8 and img.shape[2] == 3
        code, n = 8, img.shape[0] * img.shape[1]
    else:
        code, n = _DTYPE_CODE[img.dtype], img.size
    out = np.empty(img.shape[:2], np.float32)
    rc = lib().siftref_to_float(img.ctypes.data_as(ctypes.c_void_p), code, ctypes.c_long(n), _fp(out))
    assert rc == 0
    return out


def minmax(img):
    img = _f32(img)
    mn, mx = ctypes.c_float(), ctypes.c_float()
    lib().siftref_minmax(_fp(img), ctypes.c_long(img.size), ctypes.byref(mn), ctypes.byref(mx))
    return mn.value, mx.value


def normalize(img, mn=None, mx=None, max_out=255.0):
    out = _f32(img).copy()
    if mn is None:
        mn, mx = minmax(out)
    lib().siftref_normalize(_fp(out), ctypes.c_long(out.size), ctypes.c_float(mn), ctypes.c_float(mx),
                            ctypes.c_float(max_out))
    return out


def shrink(img, sw=2, sh=2):
    img = _f32(img)
    h, w = img.shape
    out = np.empty((h // sh, w // sw), np.float32)
    lib().siftref_shrink(_fp(img), _fp(out), sw, sh, w, h, out.shape[1], out.shape[0])
    return out


def convolve_h(img, taps):
    img, taps = _f32(img), _f32(taps)
    out = np.empty_like(img)
    lib().siftref_convolve_h(_fp(img), _fp(out), _fp(taps), taps.size, img.shape[1], img.shape[0])
    return out


def convolve_v(img, taps):
    img, taps = _f32(img), _f32(taps)
    out = np.empty_like(img)
    lib().siftref_convolve_v(_fp(img), _fp(out), _fp(taps), taps.size, img.shape[1], img.shape[0])
    return out


def blur(img, taps):
    return convolve_v(convolve_h(img, taps), taps)


def combine(u, a, v, b):
    u, v = _f32(u), _f32(v)
    out = np.empty_like(u)
    lib().siftref_combine(_fp(u), ctypes.c_float(a), _fp(v), ctypes.c_float(b), _fp(out), ctypes.c_long(u.size))
    return out


def gradient(img):
    img = _f32(img)
    grad, ori = np.empty_like(img), np.empty_like(img)
    lib().siftref_gradient(_fp(img), _fp(grad), _fp(ori), img.shape[1], img.shape[0])
    return grad, ori


def octave_sigmas(init_sigma=1.6, scales=3):
    """plan.py:297-306 / 602-618: the per-octave blur increments (python doubles)."""
    ratio = 2.0 ** (1.0 / scales)
    prev, out = float(init_sigma), []
    for _ in range(scales + 2):
        out.append(prev * (ratio ** 2 - 1.0) ** 0.5)
        prev *= ratio
    return out


def pyramid_octave(g0, init_sigma=1.6):
    """G[0..5], DoG[0..4] of one octave (plan.py:609-625)."""
    import math
    G = [_f32(g0)]
    ratio = 2.0 ** (1.0 / 3)
    prev = float(init_sigma)
    for _ in range(5):
        s = prev * math.sqrt(ratio ** 2 - 1.0)
        G.append(blur(G[-1], gaussian_taps(s)))
        prev *= ratio
    D = [combine(G[s + 1], -1.0, G[s], 1.0) for s in range(5)]
    return np.stack(G), np.stack(D)


PEAK_THRESH = np.float32(255.0 * 0.04 / 3.0)


def local_maxmin(dogs, scale, octsize=1, cap=None, border=5, peak=PEAK_THRESH, et0=0.08, et=0.06, kp=None, counter=0):
    dogs = _f32(dogs)
    _, h, w = dogs.shape
    cap = (h * w // 10) if cap is None else cap
    if kp is None:
        kp = -np.ones((cap, 4), np.float32)
    cnt = ctypes.c_int(counter)
    lib().siftref_local_maxmin(_fp(dogs), _fp(kp), border, ctypes.c_float(peak), octsize, ctypes.c_float(et0),
                               ctypes.c_float(et), ctypes.byref(cnt), kp.shape[0], scale, w, h)
    return kp, cnt.value


def interp_keypoint(dogs, kp, start, end, peak=PEAK_THRESH, init_sigma=1.6):
    dogs = _f32(dogs)
    _, h, w = dogs.shape
    kp = _f32(kp).copy()
    lib().siftref_interp_keypoint(_fp(dogs), _fp(kp), start, end, ctypes.c_float(peak), ctypes.c_float(init_sigma), w, h)
    return kp


def compact(kp, start, end):
    kp = _f32(kp).copy()
    n = lib().siftref_compact(_fp(kp), start, end, kp.shape[0])
    return kp, n


VARIANTS = {"cpu": 0, "gpu": 1, 0: 0, 1: 1}  # which reference kernels: *_cpu.cl, or orientation_gpu.cl + keypoints_gpu2.cl


def orientation(kp, grad, ori, start, end, octsize=1, orisigma=1.5, counter=None, variant="cpu"):
    kp = _f32(kp).copy()
    grad, ori = _f32(grad), _f32(ori)
    cnt = ctypes.c_int(end if counter is None else counter)
    lib().siftref_orientation_v(_fp(kp), _fp(grad), _fp(ori), ctypes.byref(cnt), octsize, ctypes.c_float(orisigma),
                                kp.shape[0], start, end, grad.shape[1], grad.shape[0], VARIANTS[variant])
    return kp, cnt.value


def descriptor(kp, grad, ori, start, end, octsize=1, variant="cpu"):
    kp = _f32(kp)
    grad, ori = _f32(grad), _f32(ori)
    desc = np.zeros((kp.shape[0], 128), np.uint8)
    lib().siftref_descriptor_v(_fp(kp), desc.ctypes.data_as(_c_u8_p), _fp(grad), _fp(ori), octsize, start, end,
                               grad.shape[1], grad.shape[0], VARIANTS[variant])
    return desc


def keypoints(img, init_sigma=1.6, octave_max=0, pix_per_kp=10, return_all=False, variant="cpu"):
    """Whole path (plan.py:432-567) on a 2-D float32 image.  Returns recarray[dtype_kp] (and details)."""
    img = _f32(img)
    h, w = img.shape
    cap = 2 * (h * w // pix_per_kp)  # kpsize is a per-octave limit in the reference (plan.py:243,748-752)
    out = np.zeros(cap, dtype_kp)
    noct = num_octaves(h, w)
    n_per_oct = np.zeros(noct, np.int32)
    stage = np.zeros((noct, 3, 3), np.int32)
    mm = np.zeros(2, np.float32)
    n = lib().siftref_keypoints_v(_fp(img), h, w, ctypes.c_double(init_sigma), int(octave_max), int(pix_per_kp),
                                  out.ctypes.data_as(ctypes.c_void_p), cap, n_per_oct.ctypes.data_as(_c_int_p), _fp(mm),
                                  stage.ctypes.data_as(_c_int_p), VARIANTS[variant])
    res = out[:min(n, cap)].view(np.recarray)
    if return_all:
        return res, {"n_per_octave": n_per_oct, "stage_counts": stage, "minmax": mm}
    return res


def match(kp1, kp2, ratio_th=np.float32(0.73 * 0.73), cap=None, metric="l1"):
    kp1 = np.ascontiguousarray(kp1, dtype=dtype_kp)
    kp2 = np.ascontiguousarray(kp2, dtype=dtype_kp)
    cap = max(16384, min(kp1.size, kp2.size)) if cap is None else cap  # match.py:241-243
    out = np.zeros((cap, 2), np.int32)
    n = lib().siftref_match_metric(kp1.ctypes.data_as(ctypes.c_void_p), kp2.ctypes.data_as(ctypes.c_void_p),
                                   out.ctypes.data_as(_c_int_p), cap, ctypes.c_float(ratio_th), kp1.size, kp2.size,
                                   {"l1": 0, "l2": 1}[metric])
    return out[:min(n, cap)]


def transform(img, matrix, offset, fill, out_shape=None, mode=1):
    img = _f32(img)
    h, w = img.shape
    oh, ow = (h, w) if out_shape is None else out_shape
    out = np.empty((oh, ow), np.float32)
    m = _f32(np.asarray(matrix).reshape(4))
    o = _f32(np.asarray(offset).reshape(2))
    lib().siftref_transform(_fp(img), _fp(out), _fp(m), _fp(o), w, h, ow, oh, ctypes.c_float(fill), mode)
    return out


def multiscale_image(n, seed=1234, shape=None):
    """Seeded synthetic test image: sum_k sqrt(k) * zoom(rand(n/k, n/k), k, order=1)."""
    from scipy.ndimage import zoom
    h, w = (n, n) if shape is None else shape
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    for k in (1, 2, 4, 8, 16, 32):
        hk, wk = -(-h // k), -(-w // k)
        r = rng.random((hk, wk), dtype=np.float32)
        z = r if k == 1 else zoom(r, k, order=1)
        img += np.float32(np.sqrt(k)) * z[:h, :w].astype(np.float32)
    return img


def transform_rgb(img, matrix, offset, fill, out_shape=None, mode=1):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    oh, ow = (h, w) if out_shape is None else out_shape
    out = np.empty((oh, ow, 3), np.uint8)
    m = _f32(np.asarray(matrix).reshape(4))
    o = _f32(np.asarray(offset).reshape(2))
    lib().siftref_transform_rgb(img.ctypes.data_as(_c_u8_p), out.ctypes.data_as(_c_u8_p), _fp(m), _fp(o), w, h, ow, oh,
                                ctypes.c_float(fill), mode)
    return out
