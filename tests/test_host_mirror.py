"""C++ host mirror of the reference's interfaces (grape-vector-db_b200/host/gvdb_host.hpp):
compiles against the C ABI on a CPU box; its test program (tests/cpp/test_host_mirror.cpp, written
like the reference's own unit tests) runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "grape-vector-db_b200", "lib", "test_host_mirror")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "grape-vector-db_b200", "host"), "-s"])


def test_host_mirror_compiles(built):
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_host_mirror_runs(built):
    _build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all passed" in out.stdout
