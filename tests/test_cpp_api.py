"""The C++ host mirror (include/sla.hpp) through its harness tests/cpp/test_api.cpp: host-only checks on CPU, the
reference's whole test-suite on the GPU."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "cpp"))


def binary(sla):
    import build as cpp_build
    return cpp_build.build()


def test_cpp_host_checks(sla):
    out = subprocess.run([binary(sla), "host"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all C++ API checks passed" in out.stdout


@pytest.mark.gpu
def test_cpp_reference_suite_on_gpu(sla, tmp_path):
    from helpers import fixtures
    fx = fixtures()
    path = os.path.join(tmp_path, "fixtures.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<I", 3))
        for name, (n, m) in (("small", (5, 5)), ("no_perfect", (9, 9)), ("large", (90, 900))):
            rp, c, v = fx[name + "_row_ptr"], fx[name + "_cols"], fx[name + "_vals"]
            f.write(struct.pack("<III", n, m, len(c)))
            f.write(np.ascontiguousarray(rp, dtype=np.uint32).tobytes())
            f.write(np.ascontiguousarray(c, dtype=np.uint32).tobytes())
            f.write(np.ascontiguousarray(v, dtype=np.float64).tobytes())
    out = subprocess.run([binary(sla), "gpu", path], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all C++ API checks passed" in out.stdout
