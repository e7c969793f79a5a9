"""csrc/f64_math.cuh -- the hand-rolled tanh / quotient / atanh of the fp64 check node -- against binary128 on the CPU.

The header is plain C++ under g++ (the CUDA build puts its coefficient tables into constant memory, the arithmetic is the
same), so its accuracy is pinned here without a GPU: tools/f64_math_check.cpp draws the arguments a check node can produce
(|m| <= 35, |r| <= 1 - 1.22e-15, log-uniform and saturated) and compares with libquadmath.
"""
import os
import re
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fp64_check_node_functions_against_binary128(tmp_path):
    exe = str(tmp_path / "f64_math_check")
    build = subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-o", exe,
                            os.path.join(REPO, "tools", "f64_math_check.cpp"), "-lquadmath"], capture_output=True, text=True)
    if build.returncode != 0 and "quadmath" in build.stderr:
        pytest.skip("libquadmath not available")
    assert build.returncode == 0, build.stderr
    out = subprocess.run([exe, "600000"], capture_output=True, text=True, check=True).stdout
    m = re.search(r"MAXULP tanh ([\d.]+) tanh_sat ([\d.]+) atanh ([\d.]+) div_diff (\d+)", out)
    assert m, out
    tanh_ulp, tanh_sat_ulp, atanh_ulp, div_diff = float(m.group(1)), float(m.group(2)), float(m.group(3)), int(m.group(4))
    assert tanh_ulp <= 3.0                       # measured 2.6
    assert tanh_sat_ulp <= 0.5001                # correctly rounded where the reference's results hinge on the last bit
    assert atanh_ulp <= 4.0                      # measured 3.6
    assert div_diff == 0                         # the quotient equals the IEEE one
    # the reference's clip constant IS tanh(17.5) rounded, and 2 atanh of it is what libm gives
    assert "tanh_half(35) = 0.99999999999999878" in out
    k = re.search(r"two_atanh\(clip\) = ([\d.]+) \(libm ([\d.]+)\)", out)
    assert k and k.group(1) == k.group(2)
