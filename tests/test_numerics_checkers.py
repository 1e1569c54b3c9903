"""The device evaluates exp / atan2 / two kinds of division with its own cheaper double sequences.
Each sequence is restated operation by operation in C under tools/ and compared with the oracle's definition
(the host libm / IEEE division) -- exhaustively where the domain allows.  These tests build and run the checkers,
so the claim "same fp32 as the oracle" is re-established on every host the suite runs on."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")


def _build_and_run(tmp_path, name, args=(), timeout=600):
    exe = str(tmp_path / name)
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.check_call([GCC, "-O2", "-fopenmp", "-ffp-contract=off", "-mfma", os.path.join(ROOT, "tools", name + ".c"),
                           "-o", exe, "-lm"], env=env)
    env["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    out = subprocess.run([exe] + list(args), env=env, capture_output=True, text=True, timeout=timeout)
    return out.returncode, out.stdout


@pytest.mark.skipif(GCC is None, reason="no C compiler")
def test_exp_table_equals_libm_exhaustively(tmp_path):  # common.cuh cr_expf_neg, all fp32 x in [-16, 0]
    rc, out = _build_and_run(tmp_path, "exp_check")
    assert rc == 0 and "mismatches: 0" in out, out


@pytest.mark.skipif(GCC is None, reason="no C compiler")
def test_third_equals_division_exhaustively(tmp_path):  # k_orient smoothing, all 2^32 fp32 inputs
    rc, out = _build_and_run(tmp_path, "div3_check")
    assert rc == 0 and "float mismatches 0" in out, out


@pytest.mark.skipif(GCC is None, reason="no C compiler")
def test_div_by_reciprocal_equals_division(tmp_path):  # common.cuh div_by
    rc, out = _build_and_run(tmp_path, "divby_check", ["20000000"])
    assert rc == 0 and "mismatches: 0" in out, out


@pytest.mark.skipif(GCC is None, reason="no C compiler")
def test_atan2_equals_libm(tmp_path):  # common.cuh cr_atan2f_fast, also with the breakpoint choice perturbed
    for pert in ("1.0", "1.000001", "0.999999"):
        rc, out = _build_and_run(tmp_path, "atan2_check", ["20000000", pert])
        assert rc == 0 and "float mismatches=0" in out, out
    for seed in ("1.00000047", "0.99999953"):   # the division's reciprocal seed off by +-2^-21 (hardware dependent)
        rc, out = _build_and_run(tmp_path, "atan2_check", ["20000000", "1.0", seed])
        assert rc == 0 and "float mismatches=0" in out, out
