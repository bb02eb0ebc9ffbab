// IBM/IBMBody.cuh — the marker container scenarios fill in add_bodies() (reference src/IBM/IBMBody.cuh:7-45).
// points / velocities are host arrays of `dimensions` floats per marker, AoS, allocated with new[]; after
// LBM<2>::allocate<Scenario>() they belong to the solver, which releases them in free() (the reference's IBMManager does
// the same, src/IBM/IBMManager.cuh:144-147).
#ifndef IBM_BODY_H
#define IBM_BODY_H

#include <math.h>
#include "util/utility.cuh"

struct IBMBody {
    int num_points;
    float* points;
    float* velocities;      // carried for interface parity; the reference's direct forcing targets u = 0 (IBM_impl.cuh:15)
};

inline void h_ibm_free(IBMBody body) {
    delete[] body.points;
    delete[] body.velocities;
}

#endif  // IBM_BODY_H
