"""CPU test: pl-slam_b200/csrc/glibc_atan2f.cuh (the atan2f the reference's KeyLine angle resolves to; round-2 groundwork,
not yet used by the kernels) compiled for the host is bit-identical to this image's libm on random bit patterns and on
pixel-difference-like operands."""
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HARNESS = r'''
#include <stdio.h>
#include "glibc_atan2f.cuh"
int main(void) {
  unsigned long long bad = 0, n = 0; unsigned long long s = 88172645463325252ULL;
  for (long i = 0; i < 3000000; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    float x = plf_i2f_((int)(s & 0xffffffff)), y = plf_i2f_((int)(s >> 32));
    if (i % 3 == 0) { x = (float)((int)(s % 4001) - 2000) * 0.25f; y = (float)((int)((s >> 20) % 4001) - 2000) * 0.125f; }
    if (i % 3 == 1) { x = ((int)(s % 20000) - 10000) * 0.01f + 0.003f; y = ((int)((s >> 24) % 20000) - 10000) * 0.01f; }
    float a = atan2f(y, x), b = glibc_atan2f(y, x);
    if (!(a != a && b != b) && plf_f2i_(a) != plf_f2i_(b)) ++bad;
    a = atanf(y); b = glibc_atanf(y);
    if (!(a != a && b != b) && plf_f2i_(a) != plf_f2i_(b)) ++bad;
    ++n;
  }
  printf("%llu %llu\n", n, bad);
  return bad != 0;
}
'''


def test_port_is_bit_identical_to_libm(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text(HARNESS)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-x", "c", "-I", str(ROOT / "pl-slam_b200" / "csrc"), "-o", str(exe), str(src), "-lm"],
                   check=True, capture_output=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    n, bad = (int(v) for v in r.stdout.split())
    assert r.returncode == 0 and n == 3000000 and bad == 0
