// horizontal-line.h -- the obstacle primitive of GOMP: an infinite horizontal line that a
// collision sphere has to pass over (or under).
//
// Drop-in for reference src/horizontal-line.h:6-104.  The line is A + t*D with D a unit
// vector in the XY plane.  All geometric predicates below are evaluated on the XY
// projection except isAbove, which looks at Z only.
#ifndef GOMP_HORIZONTAL_LINE_H
#define GOMP_HORIZONTAL_LINE_H

#include "utils.h"

class HorizontalLine {
public:
    HorizontalLine(const QPVector2d &direction, const Point &point, bool bypass_from_below = false)
        : D(unitXY(direction)), A(point), bypass_from_below(bypass_from_below) {}

    // Vector from P to its orthogonal projection X = A + ((P-A).D) D on the line.
    QPVector3d getDistanceVec(const Point &P) const {
        const double along = (P - A).dot(D);
        QPVector3d foot = A + along * D;
        return foot - P;
    }

    QPVector2d getDistanceVecXY(const Point &P) const {
        QPVector3d d = getDistanceVec(P);
        return QPVector2d{d[0], d[1]};
    }

    // Horizontal distance from P to the line.
    double getDistanceXY(const Point &P) const { return getDistanceVecXY(P).norm(); }

    // Closest point of the line to P.
    Point operator[](const Point &P) const { return P + getDistanceVec(P); }

    // True when the XY projections of P and Q lie strictly on different sides of the line.
    bool areOnOppositeSides(const Point &P, const Point &Q) const {
        return getDistanceVecXY(P).dot(getDistanceVecXY(Q)) < 0;
    }

    bool isClose(const Point &P, const RobotBall &b) const { return getDistanceXY(P) < b.radius; }

    // A waypoint collides with the line if the ball overlaps it horizontally or if the
    // trajectory crosses the line between this waypoint and either neighbour.
    bool hasCollision(int waypoint, const QPVector &trajectory_xyz, const RobotBall &b) const {
        const int count = (int)(trajectory_xyz.size() / 3);
        const Point here = at(trajectory_xyz, waypoint);
        if (isClose(here, b)) return true;
        if (waypoint > 0 && areOnOppositeSides(at(trajectory_xyz, waypoint - 1), here)) return true;
        if (waypoint + 1 < count && areOnOppositeSides(here, at(trajectory_xyz, waypoint + 1))) return true;
        return false;
    }

    // Clearance test on the side the line is to be passed on, with ERROR slack.
    bool isAbove(const Point &P, const RobotBall &b) const {
        const double dz = (P - A)[Axis::Z];
        return bypass_from_below ? dz <= -b.radius + ERROR : dz >= b.radius - ERROR;
    }

    bool bypassFromBelow() const { return bypass_from_below; }

    const QPVector3d &direction() const { return D; }
    const Point &point() const { return A; }

private:
    static QPVector3d unitXY(const QPVector2d &v) {
        const double len = v.norm();
        return QPVector3d{v[0] / len, v[1] / len, 0.0 / len};
    }
    static Point at(const QPVector &xyz, int waypoint) {
        return Point{xyz[3 * waypoint], xyz[3 * waypoint + 1], xyz[3 * waypoint + 2]};
    }

    const QPVector3d D; // unit direction, horizontal
    const Point A;      // a point on the line
    const bool bypass_from_below = false;
};

#endif // GOMP_HORIZONTAL_LINE_H
