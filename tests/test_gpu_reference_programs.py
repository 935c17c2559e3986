"""Drop-in proof: the reference's OWN GPU test program (tests/test_gpu.cu: small_test,
options_test, inf_test, grad_check), compiled unmodified and linked against this repo's
libwarprnnt.so through this repo's include/rnnt.h (oracle/Makefile target ref_tests), must pass."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_test_gpu_on_b200lib")
TIME = os.path.join(ROOT, "oracle", "_ref", "ref_test_time_gpu_on_b200lib")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref programs not built")]


def test_reference_test_gpu_passes_on_this_library():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out
    assert "Tests pass" in out, out
    assert "mismatch" not in out, out


def test_reference_timing_harness_runs_on_this_library():
    """tests/test_time.cu <B> <T> <L> <A>: the program behind the README tables."""
    r = subprocess.run([TIME, "16", "150", "40", "28"], capture_output=True, text=True, timeout=600)
    assert "average 10 time cost" in r.stdout, r.stdout + r.stderr
