"""CPU-side checks: the C-ABI library exists, loads and exports every symbol include/siftb.h declares;
host logic of the operator classes that needs no device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "siftb.h")).read()
    return sorted(set(re.findall(r"\b(siftb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from sift_pyocl_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    lib.siftb_version.restype = ctypes.c_int
    assert lib.siftb_version() >= 100


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle or fall back to the CPU."""
    pkg = os.path.join(ROOT, "sift_pyocl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", "").replace("No device code, no oracle", ""), fn
            assert "siftref" not in src, fn


def test_errors_without_device_are_loud():
    from sift_pyocl_b200 import _lib
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sift_pyocl_b200 as sift
    with pytest.raises(RuntimeError):
        sift.SiftPlan(shape=(64, 64), dtype=np.float32)


def test_api_surface_matches_reference():
    import inspect
    import sift_pyocl_b200 as sift
    assert {"SiftPlan", "MatchPlan", "LinearAlign", "par", "version"} <= set(dir(sift))
    # reference plan.py:117-119, match.py:77, alignment.py:82-83,227
    assert list(inspect.signature(sift.SiftPlan.__init__).parameters)[1:] == [
        "shape", "dtype", "devicetype", "template", "profile", "device", "PIX_PER_KP", "max_workgroup_size",
        "context", "init_sigma"]
    assert list(inspect.signature(sift.MatchPlan.__init__).parameters)[1:] == [
        "size", "devicetype", "profile", "device", "max_workgroup_size", "roi", "context"]
    assert list(inspect.signature(sift.LinearAlign.__init__).parameters)[1:] == [
        "image", "devicetype", "profile", "device", "max_workgroup_size", "ROI", "extra", "context", "init_sigma"]
    assert list(inspect.signature(sift.LinearAlign.align).parameters)[1:] == [
        "img", "shift_only", "return_all", "double_check", "relative", "orsa"]
    assert list(inspect.signature(sift.MatchPlan.match).parameters)[1:] == ["nkp1", "nkp2", "raw_results"]
    assert sift.SiftPlan.dtype_kp.itemsize == 144
    assert sift.par.PeakThresh == 255.0 * 0.04 / 3.0 and sift.par.Scales == 3 and sift.par.MatchRatio == 0.73


def test_param_values_match_reference_table():
    from sift_pyocl_b200 import par  # reference param.py:52-79
    want = dict(OctaveMax=100000, DoubleImSize=0, order=3, InitSigma=1.6, BorderDist=5, Scales=3, EdgeThresh=0.06,
                EdgeThresh1=0.08, OriBins=36, OriSigma=1.5, OriHistThresh=0.8, MaxIndexVal=0.2, MagFactor=3,
                IndexSigma=1.0, IgnoreGradSign=0, MatchRatio=0.73, MatchXradius=1000000.0, MatchYradius=1000000.0,
                noncorrectlylocalized=0)
    for k, v in want.items():
        assert par[k] == v and getattr(par, k) == v
    with pytest.raises(AttributeError):
        par.nope


def test_utils_host_logic():
    from sift_pyocl_b200.utils import calc_size, kernel_size, matching_correction
    assert [kernel_size(s, True) for s in (1.2263, 1.5450, 1.9466, 2.4525, 3.0900, 1.5199)] == [11, 15, 17, 21, 27, 15]
    assert kernel_size(2.0) == 17 and kernel_size(1.0) == 9 and kernel_size(1.0, True) == 9
    assert calc_size((507, 209), (128, 1)) == (512, 209) and calc_size((100,), 64) == (128,)
    # matching_correction recovers a known affine map (completed per test_transform.py:118-133)
    rng = np.random.default_rng(0)
    from sift_pyocl_b200._lib import dtype_kp
    m = np.zeros((40, 2), dtype_kp).view(np.recarray)
    m[:, 0].x, m[:, 0].y = rng.random(40) * 100, rng.random(40) * 100
    a, b, c, d, e, f = 1.1, -0.1, 5.0, 0.05, 0.9, 7.0
    m[:, 1].x = a * m[:, 0].x + b * m[:, 0].y + c
    m[:, 1].y = d * m[:, 0].x + e * m[:, 0].y + f
    assert np.allclose(matching_correction(m), [a, b, c, d, e, f], atol=1e-4)


def test_pair_records_equals_structured_indexing():
    """_lib.pair_records (raw 144-byte row gather) == the reference's field-wise result[:, 0] = kp1[idx] form
    (match.py:267-270), including empty results and non-contiguous inputs."""
    import numpy as np
    from sift_pyocl_b200._lib import dtype_kp, pair_records
    rng = np.random.default_rng(2)
    k1, k2 = np.zeros(50, dtype_kp), np.zeros(80, dtype_kp)
    for k in (k1, k2):
        k["x"], k["y"] = rng.random(k.size), rng.random(k.size)
        k["scale"], k["angle"] = rng.random(k.size), rng.random(k.size)
        k["desc"] = rng.integers(0, 256, (k.size, 128))
    i1, i2 = rng.integers(0, 50, 33).astype(np.int32), rng.integers(0, 80, 33).astype(np.int32)
    want = np.recarray((33, 2), dtype_kp)
    want[:, 0], want[:, 1] = k1[i1], k2[i2]
    got = pair_records(k1.view(np.recarray), i1, k2[::1], i2)
    assert got.dtype == dtype_kp and got.shape == (33, 2) and got.tobytes() == want.tobytes()
    assert np.array_equal(got[:, 1].desc, k2["desc"][i2]) and np.array_equal(got[:, 0].x, k1["x"][i1])
    strided = np.concatenate([k2, k2])[::2]  # non-contiguous view
    assert pair_records(k1, i1, strided, i2 % strided.size).shape == (33, 2)
    assert pair_records(k1, i1[:0], k2, i2[:0]).shape == (0, 2)


def test_matching_correction_equals_pinv_form():
    """utils.matching_correction (two decoupled 3-parameter least-squares problems) == the pinv(X).y completion of
    the reference's 2N x 6 system (utils.py:156-189 + test/test_transform.py:118-133)."""
    import numpy as np
    from sift_pyocl_b200._lib import dtype_kp
    from sift_pyocl_b200.utils import matching_correction
    rng = np.random.default_rng(4)
    for n in (3, 18, 500):
        m = np.zeros((n, 2), dtype_kp).view(np.recarray)
        m.x[:, 0], m.y[:, 0] = rng.random(n) * 900, rng.random(n) * 700
        m.x[:, 1] = 1.02 * m.x[:, 0] - 0.03 * m.y[:, 0] + 4 + rng.normal(0, 0.2, n)
        m.y[:, 1] = 0.02 * m.x[:, 0] + 0.97 * m.y[:, 0] - 3 + rng.normal(0, 0.2, n)
        X = np.zeros((2 * n, 6))
        X[::2, 0], X[::2, 1], X[::2, 2] = m.x[:, 0], m.y[:, 0], 1
        X[1::2, 3], X[1::2, 4], X[1::2, 5] = m.x[:, 0], m.y[:, 0], 1
        y = np.zeros((2 * n, 1))
        y[::2, 0], y[1::2, 0] = m.x[:, 1], m.y[:, 1]
        want = np.dot(np.linalg.pinv(X), y).ravel()
        got = matching_correction(m)
        assert got.shape == (6,) and np.allclose(got, want, rtol=1e-9, atol=1e-9)
