"""csrc/common.cuh FastDiv: the multiply-shift division the conv epilogue uses for its per-tile index arithmetic must
equal integer division for every 0 <= n < 2^31 it can meet.  The host half of the header (make_fastdiv) is compiled with
g++ and checked against `/` over dense small n, a stride over the whole range and the extremes, for the divisors the
kernels see (padded widths / heights, tile counts) and adversarial ones (powers of two +- 1, 2^31 - 1)."""
import os
import shutil
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "common.cuh"
#include <stdlib.h>
int main() {
  long bad = 0;
  const int ds[] = {1, 2, 3, 5, 7, 8, 21, 40, 78, 154, 306, 610, 612, 1023, 1024, 1025, 65535, 65536, 1000003,
                    (1 << 30) + 7, 2147483647};
  for (int d : ds) {
    const ryolo::FastDiv f = ryolo::make_fastdiv(d);
    for (long long n = 0; n < (1ll << 31); n += (n < 200000 ? 1 : 7919)) {
      const int q = (int)(((unsigned long long)(unsigned int)n * f.m) >> f.p);
      if (q != (int)(n / d)) bad++;
    }
    const long long n = (1ll << 31) - 1;
    if ((int)(((unsigned long long)(unsigned int)n * f.m) >> f.p) != (int)(n / d)) bad++;
  }
  printf("%ld\n", bad);
  return bad != 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fastdiv_equals_integer_division(tmp_path):
    cuda_inc = "/usr/local/cuda/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + cuda_inc, "-I" + os.path.join(REPO, "rotate-yolov3_b200", "csrc"),
                    "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout + out.stderr
