"""Symbolic phase of the multifrontal solver on the CPU (no GPU): tests/cpp/mf_tables_check.cpp executes the numeric phase the
CUDA kernels perform from the same tables (orderings, front structures, extend-add maps, pivot chunks, depth schedule) in plain
std::complex and checks the residual of A x = b."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("mf") / "mf_tables_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "mf_tables_check.cpp")], check=True)
    return exe


@pytest.mark.parametrize("args", [
    ("grid", "7", "5", "32", "144"),            # a single leaf front
    ("grid", "40", "25", "32", "80"),           # small fronts only
    ("grid", "64", "50", "16", "0"),            # every front through the large-front path, chunked pivots
    ("grid", "199", "99", "16", "144"),         # cfg2 system: mixed
    ("grid", "199", "99", "16", "144", "0", "2"),   # the same with tile-aligned leaves (excess unknowns pushed to the separator)
    ("grid", "64", "45", "9", "144", "0", "7"),     # every leaf excess pushed
    ("grid", "60", "47", "16", "144", "13", "2"),   # cross-shaped separators at the bottom of the tree (four children per front)
    ("grid", "300", "3", "8", "48"),            # degenerate strip
    ("graph3d", "12", "10", "8", "24", "100"),  # level-set bisection of a 3-D grid (Level-1 shim ordering)
    ("graph3d", "30", "1", "1", "4", "144"),    # a path graph
])
def test_tables_reproduce_the_solution(checker, args):
    res = subprocess.run([checker, *args], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "relative residual" in res.stdout
