"""`not gpu`: accuracy of the branch-free grouped sin/cos (csrc/rbm_trig.cuh) compiled for the HOST and compared with libm in
extended precision -- the device code is the same source.  Skipped when nvcc is unavailable."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

SRC = r"""
#include <cstdio>
#include <cmath>
#include <random>
#include "rbm_trig.cuh"
int main() {
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> u(-1, 1);
  double e64 = 0, e32 = 0;
  const double sc64[4] = {1.0, 20.0, 1000.0, rbm::kTrigFastMaxF64};
  const double sc32[4] = {1.0, 20.0, 1000.0, rbm::kTrigFastMaxF32};
  for (int it = 0; it < 4000000; ++it) {
    double x = u(g) * sc64[it % 4], s, c;
    rbm::sincos_core(x, s, c);
    e64 = fmax(e64, fmax(fabs((double)(s - sinl((long double)x))), fabs((double)(c - cosl((long double)x)))));
    float xf = (float)(u(g) * sc32[it % 4]), sf, cf;
    rbm::sincos_core(xf, sf, cf);
    e32 = fmax(e32, fmax(fabs(sf - sin((double)xf)), fabs(cf - cos((double)xf))));
  }
  double x3[3] = {0.0, 1e5, -99999.5}, s3[3], c3[3];
  rbm::sincos_group<3, double>(x3, s3, c3);
  double y3[3] = {0.5, 1e12, 3.0}, t3[3], d3[3];   // one out-of-range angle sends the whole group to libm
  rbm::sincos_group<3, double>(y3, t3, d3);
  double eg = fmax(fabs(s3[1] - sin(1e5)), fabs(t3[1] - sin(1e12)));
  printf("%.3e %.3e %.3e %g %g\n", e64, e32, eg, s3[0], c3[0]);
  return 0;
}
"""


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_grouped_sincos_accuracy(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = tmp_path / "t.cu"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "rigid_body_manipulation_b200", "csrc"),
                    "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    e64, e32, eg, s0, c0 = map(float, out)
    assert e64 < 2.5e-16, e64
    assert e32 < 2.5e-7, e32
    assert eg < 1e-15
    assert s0 == 0.0 and c0 == 1.0
