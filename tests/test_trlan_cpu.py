"""Host logic of the multi-state solver (edipack_b200/csrc/trlan.hpp, the sp_eigh / ARPACK
replacement) on a dense CPU mock backend (tests/cpp/test_trlan.cpp): restart algebra,
convergence test, invariant-subspace handling, degenerate eigenvalues.  The device backend is
covered by tests/test_gpu_eigh.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def trlan_bin(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("trlan") / "test_trlan")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_trlan.cpp")])
    return exe


@pytest.mark.parametrize("args", [
    "300 1 10 1",      # ground state only, ARPACK-default basis of 10
    "300 4 20 2",      # a few states, several restarts
    "500 6 60 3",
    "40 5 40 4",       # basis = whole space: exact after one cycle
    "12 12 12 5",      # all eigenpairs requested
    "300 3 16 6 2",    # doubly degenerate ground state
    "1 1 1 3",         # one-state sector
    "2 2 20 3",        # Nblock larger than the sector
    "20 2 20 9",
])
def test_trlan_dense_mock(trlan_bin, args):
    out = subprocess.run([trlan_bin, *args.split()], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
