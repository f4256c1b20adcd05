"""Accuracy of csrc/fastmath.cuh (the lean FP64 log / exp / cbrt / division of the simulator kernel), checked
on the host: the header compiles for the CPU with the same sequence of IEEE operations as on the device (only the
hardware seeds differ, and they are refined far below the final rounding).  Reference: glibc long double."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ode-discovery-for-longitudinal-heterogeneous-treatment-effects-inference_b200", "csrc")

# maximum error in ulp tolerated per function (measured: see the comments in fastmath.cuh)
BOUNDS = {"div_fast": 0.501, "rcp_fast": 0.501, "log_ratio_path": 0.8, "log_ratio_generic": 1.0, "exp_fast": 1.1,
          "exp_fast_wide": 1.1, "cbrt_fast": 0.6, "div_small": 0.501}


@pytest.fixture(scope="module")
def report(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("fm") / "fastmath_check")
    src = os.path.join(ROOT, "tests", "native", "fastmath_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", CSRC, src, "-o", exe], check=True)
    out = subprocess.run([exe, "400000"], check=True, capture_output=True, text=True).stdout
    return {l.split()[0]: float(l.split()[1]) for l in out.strip().splitlines()}


@pytest.mark.parametrize("name", sorted(BOUNDS))
def test_ulp_bounds(report, name):
    assert report[name] <= BOUNDS[name], f"{name}: {report[name]} ulp"


def test_divisions_are_correctly_rounded_on_the_sample(report):
    assert report["div_fast_not_correctly_rounded"] == 0
    assert report["div_small_not_correctly_rounded"] == 0
