"""The C++ host mirror (include/fhe_b200.hpp) of the reference's Rust API: compiles against the C ABI on
the CPU box; on the GPU it runs the reference's own unit tests re-stated in C++ (tests/cpp/test_host_api.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_api")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "fhe_study_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"),
                           "-o", EXE, "-L" + libdir, "-lfhe_b200", "-Wl,-rpath," + libdir])
    return EXE


def test_cpp_host_mirror_compiles_and_links():
    import fhe_study_b200  # noqa: F401  (makes sure the library exists)

    assert os.path.exists(_build())


@pytest.mark.gpu
def test_cpp_host_mirror_runs_reference_unit_tests():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
