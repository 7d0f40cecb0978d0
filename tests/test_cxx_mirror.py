"""The C++ mirror of the reference's classes (include/cedr_b200.hpp) as a C++ caller sees
it: tests/cxx/test_cdr_mirror.cu drives cedr::qlt::QLT / cedr::caas::CAAS through device
kernels that copy the DeviceOp by value and through plain host loops (managed memory),
and compares with the CPU oracle bit for bit. Here: it must compile without a GPU
(nvcc cross-compiles), and on the GPU box it must print PASS."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CXX = os.path.join(HERE, "cxx")


def _build():
    import compose_b200.build as b
    from oracle import oracle_py
    b.build()
    oracle_py.build(ref=False)
    r = subprocess.run(["make", "-C", CXX, "-B", "test_cdr_mirror"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return os.path.join(CXX, "test_cdr_mirror")


def test_cxx_mirror_compiles():
    exe = _build()
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_cxx_mirror_runs_bitwise():
    exe = _build()
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=600)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:]
