"""Builds tests/cpp/test_api (the C++ host mirror's test harness) against the in-tree libsla_b200.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "sparse_linear_assignment_b200")
OUT = os.path.join(HERE, "test_api")


def build():
    src = os.path.join(HERE, "test_api.cpp")
    deps = [src, os.path.join(ROOT, "include", "sla.hpp"), os.path.join(ROOT, "include", "sla.h"), os.path.join(PKG, "libsla_b200.so")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", OUT,
                    "-L", PKG, "-lsla_b200", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../sparse_linear_assignment_b200"],
                   check=True)
    return OUT


if __name__ == "__main__":
    build()
