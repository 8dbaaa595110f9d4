"""The facade's ConstraintBuilder / HorizontalLine against the golden vectors of the
reference's own unit tests (reference tests/test.cpp:45-448, all 13 TESTs; the three
IndicesTest cases are one entry here)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "osqp-solver_b200", "include", "gomp")


@pytest.fixture(scope="module")
def dump(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bin") / "builder_dump")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-DGOMP_FORCE_EIGEN_LITE", "-I", INC, "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "builder_dump.cpp")])
    return json.loads(subprocess.check_output([exe]))


with open(os.path.join(ROOT, "tests", "golden", "reference_tests.json")) as f:
    GOLD = json.load(f)


@pytest.mark.parametrize("case", sorted(GOLD))
def test_reference_golden(dump, case):
    got, want = dump[case], GOLD[case]
    for key, val in want.items():
        # the reference asserts isApprox / EXPECT_EQ; the facade reproduces the values exactly
        np.testing.assert_array_equal(np.asarray(got[key], dtype=float), np.asarray(val, dtype=float),
                                      err_msg=f"{case}.{key}")


def test_all_reference_tests_are_covered():
    assert len(GOLD) == 11  # 13 TESTs, IndicesTest.{position,velocity,acceleration} folded into one
