"""gpp_math.h is host+device: compile its scalar exp routines with g++ and bound their error against long double (CPU test)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_exp_accuracy(tmp_path):
  exe = str(tmp_path / "exp_accuracy")
  subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "host", "exp_accuracy.cpp")], check=True)
  out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
  worst_scaled, worst_mid, worst_poly, at0, underflow = map(float, out)
  assert worst_scaled < 4e-16          # table exp: relative error <= max(1, |x|) * 4e-16 (|x| eps: the rounding error x itself carries)
  assert worst_mid < 3e-15             # ... i.e. a few ulp for the exponents that matter (|x| <= 40)
  assert worst_poly < 5e-16            # polynomial exp: ~1 ulp
  assert at0 == 1.0 and underflow == 0.0
