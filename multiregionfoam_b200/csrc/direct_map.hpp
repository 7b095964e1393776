// direct_map.hpp -- arithmetic of the conformal-interface face / point maps (directMap variant).
//
// directMapInterfaceToInterfaceMapping::calcZone{A,B}ToZone{B,A}{Face,Point}Map
// (src/numerics/interfaceToInterfaceMappings/directMapInterfaceToInterfaceMapping/directMapInterfaceToInterfaceMapping.C:
// 147-168, 268-289, 389-410, 510-531) all run the same N^2 search: for every location of the receiving zone the FIRST
// location of the other zone (ascending) with  mag(to - from) < tol,  tol = relTol_ (0.001, :52) * min edge length (:153,
// :560-597); -1 where there is none, which the caller turns into the reference's FatalError (:170-181).
// Compiled by nvcc into k_direct_map (direct_map.cuh) and by g++ into the CPU emulator (tests/cpp/direct_map_emulate.cpp).
#pragma once
#include <stdint.h>

#include <cmath>

#if defined(__CUDACC__)
#define B200_DHD __host__ __device__ __forceinline__
#else
#define B200_DHD inline
#endif

namespace dmap
{

// mag(a - b) < tol with Foam's Vector arithmetic: componentwise difference, magSqr = x*x + y*y + z*z, mag = sqrt(magSqr).
// The squared pre-test only skips pairs that cannot pass (sqrt(d2) >= 2 tol (1 - ulp) > tol); the decision itself is the
// reference's expression.
B200_DHD bool matches(double ax, double ay, double az, double bx, double by, double bz, double tol, double fourTol2)
{
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    const double d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > fourTol2) return false;
    return sqrt(d2) < tol;
}

} // namespace dmap
