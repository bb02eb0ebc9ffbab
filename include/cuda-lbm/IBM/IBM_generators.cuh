// IBM/IBM_generators.cuh — marker generators for the 2-D scenarios (reference src/IBM/IBM_generators.cuh:8,
// src/IBM/IBM_generators.cu:5-25).  Header-only here, so scenario files that only include IBM/IBMBody.cuh through
// scenarios/scenario.cuh still find create_cylinder.
#ifndef IBM_GENERATORS_H
#define IBM_GENERATORS_H

#include <cmath>
#include <vector>
#include "IBM/IBMBody.cuh"

// num_pts markers on the circle of radius r about (cx, cy), lattice units, first marker at angle 0, counter-clockwise.
// Evaluated in the same precision as the reference (float angle step, float cos/sin) so marker positions agree bit for bit.
inline IBMBody create_cylinder(float cx, float cy, float r, int num_pts = 16) {
    IBMBody body{num_pts, new float[2 * num_pts], new float[2 * num_pts]};
    const float step = 2 * M_PI / num_pts;
    for (int k = 0; k < num_pts; k++) {
        body.points[2 * k] = cx + r * cos(k * step);
        body.points[2 * k + 1] = cy + r * sin(k * step);
        body.velocities[2 * k] = body.velocities[2 * k + 1] = 0.0f;
    }
    return body;
}

// markers from an explicit list of (x, y) pairs
inline IBMBody body_from_points(const std::vector<float>& xy) {
    const int n = static_cast<int>(xy.size() / 2);
    IBMBody body{n, new float[2 * n], new float[2 * n]};
    for (int i = 0; i < 2 * n; i++) { body.points[i] = xy[i]; body.velocities[i] = 0.0f; }
    return body;
}

#endif  // IBM_GENERATORS_H
