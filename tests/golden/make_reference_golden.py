"""Writes tests/golden/reference_tests.json: the expected values of the reference's own unit
tests (reference tests/test.cpp), transcribed by hand -- the reference's test binary cannot be
built here (gtest/Eigen/osqp-cpp are fetched from the network, reference tests/CMakeLists.txt).

Each entry cites the lines it was taken from.  `acceleration` and `all` carry the intended
matrices of tests/test.cpp:198-204 and :227-243, whose `expected_A <<` lists end in a stray
comma in the reference; the values are unambiguous.

Run:  python tests/golden/make_reference_golden.py
"""
import json
import math
import os


def eye_rows(cols, ones):
    rows = []
    for spec in ones:
        r = [0.0] * cols
        for c, v in spec:
            r[c] = float(v)
        rows.append(r)
    return rows


G = {}

# tests/test.cpp:45-80  ConstraintBuilderTest.linkingVelocityToPosition (N=2, W=3, rows 0..3)
G["linkingVelocityToPosition"] = dict(
    A=[[1, 0, -1, 0, 0, 0, 1, 0, 0, 0, 0, 0],
       [0, 1, 0, -1, 0, 0, 0, 1, 0, 0, 0, 0],
       [0, 0, 1, 0, -1, 0, 0, 0, 1, 0, 0, 0],
       [0, 0, 0, 1, 0, -1, 0, 0, 0, 1, 0, 0]],
    l=[0, 0, 0, 0], u=[0, 0, 0, 0])

# tests/test.cpp:82-100  LineUtilTest.XAxis  (line dir (2,0) through (1,1,1)); EXPECT_EQ = exact
G["LineUtil_XAxis"] = dict(distvec_norms=[0, 1, 1, math.sqrt(2)], dist_xy=[0, 1, 0], closest=[1.1, 1, 1])

# tests/test.cpp:102-122  IndicesTest.position / velocity / acceleration
G["Indices"] = dict(pos=[0, 2, 4], vel=[6, 8], acc=[14, 16])

# tests/test.cpp:124-153  ConstraintsTest.jointPosition (rows firstPositionConstraintIdx.., 6 rows)
G["jointPosition"] = dict(A=eye_rows(12, [[(i, 1)] for i in range(6)]), l=[1, 2, 1, 2, 1, 2], u=[3, 4, 3, 4, 3, 4])

# tests/test.cpp:155-182  ConstraintsTest.velocity
G["velocity"] = dict(A=eye_rows(12, [[(6 + i, 1)] for i in range(4)]), l=[1, 2, 1, 2], u=[3, 4, 3, 4])

# tests/test.cpp:184-209  ConstraintsTest.acceleration
G["acceleration"] = dict(A=eye_rows(12, [[(6, -1), (8, 1)], [(7, -1), (9, 1)]]), l=[1, 2], u=[3, 4])

# tests/test.cpp:211-248  ConstraintsTest.all
G["all"] = dict(
    A=eye_rows(12, [[(i, 1)] for i in range(6)] + [[(6 + i, 1)] for i in range(4)] + [[(6, -1), (8, 1)], [(7, -1), (9, 1)]]),
    l=[1, 2, 1, 2, 1, 2, 5, 6, 5, 6, 9, 10], u=[3, 4, 3, 4, 3, 4, 7, 8, 7, 8, 11, 12])

JAC_A = [[0, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0],
         [3, 4, 5, 0, 0, 0, 0, 0, 0, 0, 0, 0],
         [6, 7, 8, 0, 0, 0, 0, 0, 0, 0, 0, 0],
         [0, 0, 0, 0, 1, 2, 0, 0, 0, 0, 0, 0],
         [0, 0, 0, 3, 4, 5, 0, 0, 0, 0, 0, 0],
         [0, 0, 0, 6, 7, 8, 0, 0, 0, 0, 0, 0]]

# tests/test.cpp:250-303  ConstraintsTest.position3d_2 (stateful power-of-two FK)
lx, ly, lz = 11 + 0 + 1 + 2, 22 + 3 + 4 + 5, 33 + 6 + 7 + 8
ux, uy, uz = 44 + 0 + 1 + 2, 55 + 3 + 4 + 5, 66 + 6 + 7 + 8
G["position3d_2"] = dict(A=JAC_A, l=[lx - 1, ly - 2, lz - 4, lx - 8, ly - 16, lz - 32],
                         u=[ux - 1, uy - 2, uz - 4, ux - 8, uy - 16, uz - 32])

# tests/test.cpp:305-357  ConstraintsTest.position3d_1 (identity FK, trajectory of ones)
lx, ly, lz = 11 - 1 + 0 + 1 + 2, 22 - 1 + 3 + 4 + 5, 33 - 1 + 6 + 7 + 8
ux, uy, uz = 44 - 1 + 0 + 1 + 2, 55 - 1 + 3 + 4 + 5, 66 - 1 + 6 + 7 + 8
G["position3d_1"] = dict(A=JAC_A, l=[lx, ly, lz] * 2, u=[ux, uy, uz] * 2)

# tests/test.cpp:359-401  ConstraintsTest.position3d_jac_pow2 (bounds only are asserted)
lx, ly, lz = 11 - 2 + 0 + 2 + 4, 22 - 2 + 8 + 16 + 32, 33 - 2 + 64 + 128 + 256
ux, uy, uz = 44 - 2 + 0 + 2 + 4, 55 - 2 + 8 + 16 + 32, 66 - 2 + 64 + 128 + 256
G["position3d_jac_pow2"] = dict(l=[lx, ly, lz] * 2, u=[ux, uy, uz] * 2)

# tests/test.cpp:403-448  ConstraintsTest.ignore_velocity_trajectory (same expectations:
# the velocity half of the trajectory must not influence the linearisation)
G["ignore_velocity_trajectory"] = dict(l=[lx, ly, lz] * 2, u=[ux, uy, uz] * 2)

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tests.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1, sort_keys=True)
    print("wrote", out)
