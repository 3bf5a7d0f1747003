"""The C ABI driven from plain C (tests/abi_smoke.c, compiled here with gcc against include/pop_arrowhead.h and linked to
libpop_arrowhead.so): no ctypes, no Python between the caller and the library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "piecewiseorthogonalpolynomials.jl_b200")
EXE = os.path.join(ROOT, "tests", "_build", "abi_smoke")


def build_smoke():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "abi_smoke.c")
    lib = os.path.join(PKG, "libpop_arrowhead.so")
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= max(os.path.getmtime(src), os.path.getmtime(lib)):
        return
    cmd = ["gcc", "-O1", "-Wall", "-Werror", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", EXE, "-L", PKG, "-lpop_arrowhead",
           "-Wl,-rpath," + PKG, "-lm"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_header_compiles_as_c99_and_links():
    """include/pop_arrowhead.h is valid C (not only C++), every entry point the program uses resolves at link time, and
    without a GPU the library reports POP_ERR_NOGPU instead of computing anything."""
    build_smoke()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_abi_smoke_on_gpu")
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 77, out.stdout + out.stderr
    assert "no CPU fallback" in out.stdout


@pytest.mark.gpu
def test_abi_smoke_on_gpu():
    build_smoke()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "errors ok" in out.stdout
