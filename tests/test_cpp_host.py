"""The C++ host mirror (better_search_rag_rust_b200/csrc/bsr_host.hpp) above the C ABI: compiled
with g++ and run as a separate program.  `cpu` mode checks host logic and the no-device error;
`gpu` mode runs the reference's search flow on cuda:0 and compares with the oracle bit for bit."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "better_search_rag_rust_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "build", "host_mirror_test")


def _build():
    from oracle import c_oracle as co
    co.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    newest = max(os.path.getmtime(p) for p in (src, os.path.join(PKG, "csrc", "bsr_host.hpp"),
                                               os.path.join(ROOT, "include", "bsr.h")))
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < newest:
        subprocess.check_call([
            "g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", src, "-o", EXE,
            "-L", PKG, "-lbsr_b200", "-L", os.path.join(ROOT, "oracle", "build"), "-lbsr_oracle",
            f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle', 'build')}"])
    return EXE


def test_cpp_host_mirror_cpu():
    out = subprocess.run([_build(), "cpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "cpu ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_gpu():
    out = subprocess.run([_build(), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "gpu ok" in out.stdout, out.stdout + out.stderr
