"""The C++ host mirror (include/lz4net.hpp): compiles everywhere; runs on the GPU box against the oracle port."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_lz4net_hpp")


def _build():
    import oracle
    from lz4net_b200 import build
    build.build(); oracle.build()
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_lz4net_hpp.cpp"), "-o", EXE,
                           "-L" + os.path.join(ROOT, "lz4net_b200"), "-llz4b200", "-L" + os.path.join(ROOT, "oracle"), "-llz4_oracle",
                           "-Wl,-rpath," + os.path.join(ROOT, "lz4net_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")])


def test_cpp_mirror_compiles():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
