"""Accuracy of the hardware-seeded reciprocal / reciprocal square root of the solve kernel (csrc/fast_math.h)
against the IEEE results, measured on the GPU by the standalone probe tools/fastmath_test.cu."""
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_reciprocals_within_3_ulp(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "fmt")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tools", "fastmath_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    m = re.search(r"rcp \S+ \(([\d.]+) ulp\)\s+rsqrt \S+ \(([\d.]+) ulp\)", out)
    assert m, out
    assert float(m.group(1)) <= 1.0 and float(m.group(2)) <= 3.0, out
