// Dumps what the facade's ConstraintBuilder / HorizontalLine produce for the inputs of the
// reference's own unit tests (reference tests/test.cpp:45-448) as JSON on stdout.
// tests/test_builder_golden.py compares the dump with tests/golden/reference_tests.json.
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "constraints/constraint-builder.h"

static void dumpVec(const char *name, const QPVector &v, size_t start, size_t count, bool last = false) {
    std::printf("\"%s\": [", name);
    for (size_t i = 0; i < count; ++i) std::printf("%s%.17g", i ? ", " : "", v[start + i]);
    std::printf("]%s", last ? "" : ", ");
}
static void dumpRows(const QPConstraints &c, size_t start, size_t count) {
    const auto &[l, A, u] = c;
    std::printf("\"A\": [");
    for (size_t r = 0; r < count; ++r) {
        std::printf("%s[", r ? ", " : "");
        for (long long j = 0; j < A.cols(); ++j) std::printf("%s%.17g", j ? ", " : "", A.coeff((long long)(start + r), j));
        std::printf("]");
    }
    std::printf("], ");
    dumpVec("l", l, start, count);
    dumpVec("u", u, start, count, true);
}
static void beginCase(const char *name, bool first = false) { std::printf("%s\"%s\": {", first ? "" : ",\n", name); }
static void endCase() { std::printf("}"); }

static size_t firstPos(size_t w, size_t d) { return (w - 1) * d; }
static size_t firstVel(size_t w, size_t d) { return firstPos(w, d) + w * d; }
static size_t firstAcc(size_t w, size_t d) { return firstVel(w, d) + (w - 1) * d; }
static size_t first3d(size_t w, size_t d) { return firstAcc(w, d) + (w - 2) * d; }

static void constJac(double *out, const double (&vals)[9]) {
    for (int i = 0; i < 9; ++i) out[i] = vals[i];
}

int main() {
    std::printf("{\n");
    {
        beginCase("linkingVelocityToPosition", true);
        auto c = ConstraintBuilder<2>{3, {}, {}}.build();
        dumpRows(c, 0, 4);
        endCase();
    }
    {
        beginCase("LineUtil_XAxis");
        HorizontalLine line{{2, 0}, {1, 1, 1}};
        std::printf("\"distvec_norms\": [%.17g, %.17g, %.17g, %.17g], ", line.getDistanceVec({2, 1, 1}).norm(),
                    line.getDistanceVec({1, 2, 1}).norm(), line.getDistanceVec({1, 1, 2}).norm(),
                    line.getDistanceVec({1, 2, 2}).norm());
        std::printf("\"dist_xy\": [%.17g, %.17g, %.17g], ", line.getDistanceXY({2, 1, 1}), line.getDistanceXY({1, 2, 1}),
                    line.getDistanceXY({1, 1, 2}));
        QPVector3d p = line[QPVector3d{1.1, 1.2, 1.3}];
        std::printf("\"closest\": [%.17g, %.17g, %.17g]", p[0], p[1], p[2]);
        endCase();
    }
    {
        beginCase("Indices");
        auto b3 = ConstraintBuilder<2>{3, {}, {}};
        auto b4 = ConstraintBuilder<2>{4, {}, {}};
        std::printf("\"pos\": [%zu, %zu, %zu], \"vel\": [%zu, %zu], \"acc\": [%zu, %zu]", b3.nthPos(0), b3.nthPos(1),
                    b3.nthPos(2), b3.nthVelocity(0), b3.nthVelocity(1), b4.nthAcceleration(0), b4.nthAcceleration(1));
        endCase();
    }
    const size_t W = 3;
    {
        beginCase("jointPosition");
        auto c = ConstraintBuilder<2>{W, {}, {}}.positions(0, W - 1, constraints::inRange<2>({{1, 2}}, {{3, 4}})).build();
        dumpRows(c, firstPos(W, 2), W * 2);
        endCase();
    }
    {
        beginCase("velocity");
        auto c = ConstraintBuilder<2>{W, {}, {}}.velocities(0, W - 2, constraints::inRange<2>({{1, 2}}, {{3, 4}})).build();
        dumpRows(c, firstVel(W, 2), (W - 1) * 2);
        endCase();
    }
    {
        beginCase("acceleration");
        auto c = ConstraintBuilder<2>{W, {}, {}}.accelerations(0, W - 3, constraints::inRange<2>({{1, 2}}, {{3, 4}})).build();
        dumpRows(c, firstAcc(W, 2), (W - 2) * 2);
        endCase();
    }
    {
        beginCase("all");
        auto c = ConstraintBuilder<2>{W, {}, {}}
                     .positions(0, W - 1, constraints::inRange<2>({{1, 2}}, {{3, 4}}))
                     .velocities(0, W - 2, constraints::inRange<2>({{5, 6}}, {{7, 8}}))
                     .accelerations(0, W - 3, constraints::inRange<2>({{9, 10}}, {{11, 12}}))
                     .build();
        dumpRows(c, firstPos(W, 2), (W + W - 1 + W - 2) * 2);
        endCase();
    }
    ForwardKinematicsFun identity_fk = [](double *q) { return std::tuple<double, double, double>{q[0], q[1], q[2]}; };
    JacobianFun jac_0_8 = [](double *out, double *) { constJac(out, {0, 1, 2, 3, 4, 5, 6, 7, 8}); };
    JacobianFun jac_pow2 = [](double *out, double *) { constJac(out, {0, 1, 2, 4, 8, 16, 32, 64, 128}); };
    const auto box = constraints::inRange<3>({{11, 22, 33}}, {{44, 55, 66}});
    {
        // stateful FK: successive calls return (1,2,4), (8,16,32), ... -> pins "one FK call per
        // waypoint, in waypoint order" (reference tests/test.cpp:253-256)
        beginCase("position3d_2");
        int cnt = 0;
        ForwardKinematicsFun pow2_fk = [&cnt](double *) {
            double a = (double)(1 << cnt), b = (double)(1 << (cnt + 1)), c = (double)(1 << (cnt + 2));
            cnt += 3;
            return std::tuple<double, double, double>{a, b, c};
        };
        QPVector traj = QPVector::Constant(2 * 3 * 2, 1);
        auto c = ConstraintBuilder<3>{2, {RobotBall{pow2_fk, jac_0_8, 0, true}}, {}}.withObstacles(box, traj).build();
        dumpRows(c, first3d(2, 3), 6);
        endCase();
    }
    {
        beginCase("position3d_1");
        QPVector traj = QPVector::Constant(2 * 3 * 2, 1);
        auto c = ConstraintBuilder<3>{2, {RobotBall{identity_fk, jac_0_8, 0, true}}, {}}.withObstacles(box, traj).build();
        dumpRows(c, first3d(2, 3), 6);
        endCase();
    }
    {
        beginCase("position3d_jac_pow2");
        QPVector traj = QPVector::Constant(2 * 3 * 2, 2);
        auto c = ConstraintBuilder<3>{2, {RobotBall{identity_fk, jac_pow2, 0, true}}, {}}.withObstacles(box, traj).build();
        dumpRows(c, first3d(2, 3), 6);
        endCase();
    }
    {
        beginCase("ignore_velocity_trajectory");
        QPVector traj(12);
        traj << 2, 2, 2, 2, 2, 2, 1024, 1024, 1024, 1024, 1024, 1024;
        auto c = ConstraintBuilder<3>{2, {RobotBall{identity_fk, jac_pow2, 0, true}}, {}}.withObstacles(box, traj).build();
        dumpRows(c, first3d(2, 3), 6);
        endCase();
    }
    std::printf("\n}\n");
    return 0;
}
