// synthetic-arm.h -- a fixed serial chain of up to 7 revolute joints with analytic forward
// kinematics and geometric Jacobian.
//
// The reference's example needs the un-vendored Kinematics-UR5e-arm library
// (reference examples/solver-example.cpp:2,37-41) and ships no 7-DoF model, so the repo
// defines its own arm for the BASELINE.json workloads.  Geometry: modified
// Denavit-Hartenberg parameters of a Panda-like 7R arm; chains with fewer joints use the
// first N rows.  Plain C++ without library calls beyond sin/cos so that the same code
// compiles for the host generator and for device kernels.
#ifndef GOMP_SYNTHETIC_ARM_H
#define GOMP_SYNTHETIC_ARM_H

#include <cmath>

#if defined(__CUDACC__)
#define GOMP_HD __host__ __device__
#else
#define GOMP_HD
#endif

namespace gomp {

struct ArmGeometry {
    // modified DH (Craig): T_i = RotX(alpha_{i-1}) TransX(a_{i-1}) RotZ(theta_i) TransZ(d_i)
    double a[7], d[7], alpha[7];
    double flange; // tool offset along the last joint axis
};

GOMP_HD inline ArmGeometry pandaLikeGeometry() {
    const double h = 1.5707963267948966;
    ArmGeometry g = {{0.0, 0.0, 0.0, 0.0825, -0.0825, 0.0, 0.088},
                     {0.333, 0.0, 0.316, 0.0, 0.384, 0.0, 0.0},
                     {0.0, -h, h, h, -h, h, h},
                     0.107};
    return g;
}

// Pose bookkeeping for one pass down the chain: origins o[i] and joint axes z[i] (world).
template <int N>
struct ChainPose {
    double o[N + 1][3]; // o[i] = origin of frame i (i = 0..N-1), o[N] = tool point
    double z[N][3];     // rotation axis of joint i in the world frame
};

// Walks the first `upto` joints (1..N) and puts the point of interest `tool_offset` along the
// last visited joint axis.  upto = N, tool_offset = flange gives the gripper; a smaller `upto`
// gives a point on an inner link (used for the second collision ball).
template <int N>
GOMP_HD inline void chainPose(const ArmGeometry &g, const double *q, int upto, double tool_offset, ChainPose<N> &out) {
    // R = accumulated rotation (row-major 3x3), p = accumulated origin
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
    for (int i = 0; i < upto; ++i) {
        const double ca = cos(g.alpha[i]), sa = sin(g.alpha[i]);
        const double ct = cos(q[i]), st = sin(q[i]);
        // local transform of joint i
        const double L[9] = {ct, -st, 0.0, st * ca, ct * ca, -sa, st * sa, ct * sa, ca};
        const double t[3] = {g.a[i], -sa * g.d[i], ca * g.d[i]};
        double Rn[9], pn[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) Rn[3 * r + c] = R[3 * r] * L[c] + R[3 * r + 1] * L[3 + c] + R[3 * r + 2] * L[6 + c];
            pn[r] = p[r] + R[3 * r] * t[0] + R[3 * r + 1] * t[1] + R[3 * r + 2] * t[2];
        }
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
        for (int k = 0; k < 3; ++k) {
            p[k] = pn[k];
            out.o[i][k] = p[k];
            out.z[i][k] = R[3 * k + 2];
        }
    }
    for (int k = 0; k < 3; ++k) out.o[N][k] = p[k] + tool_offset * R[3 * k + 2];
    for (int i = upto; i < N; ++i)
        for (int k = 0; k < 3; ++k) {
            out.o[i][k] = out.o[N][k];
            out.z[i][k] = 0.0;
        }
}

template <int N>
GOMP_HD inline void armPoint(const ArmGeometry &g, const double *q, int upto, double tool_offset, double xyz[3]) {
    ChainPose<N> pose;
    chainPose<N>(g, q, upto, tool_offset, pose);
    for (int k = 0; k < 3; ++k) xyz[k] = pose.o[N][k];
}

// J (3 x N, row-major): column i = z_i x (p_tool - o_i) for the joints that move the point.
template <int N>
GOMP_HD inline void armJacobian(const ArmGeometry &g, const double *q, int upto, double tool_offset, double *J) {
    ChainPose<N> pose;
    chainPose<N>(g, q, upto, tool_offset, pose);
    for (int i = 0; i < N; ++i) {
        const double rx = pose.o[N][0] - pose.o[i][0], ry = pose.o[N][1] - pose.o[i][1], rz = pose.o[N][2] - pose.o[i][2];
        const double zx = pose.z[i][0], zy = pose.z[i][1], zz = pose.z[i][2];
        J[0 * N + i] = zy * rz - zz * ry;
        J[1 * N + i] = zz * rx - zx * rz;
        J[2 * N + i] = zx * ry - zy * rx;
    }
}

} // namespace gomp
#endif // GOMP_SYNTHETIC_ARM_H
