// fv_assemble.hpp -- arithmetic of the device-side coefficient refresh of the T equations.
//
// The two temperature regions of the reference assemble, every time step,
//   conductTemperature   (src/regions/conductTemperature/conductTemperature.C:135-142)
//       fvm::ddt(rho*cv, T) == fvm::laplacian(kappa, T)                              form B200_TEQN_CONDUCT
//   transportTemperature (src/regions/transportTemperature/transportTemperature.C:129-140)
//       rho*cp*(fvm::ddt(T) + fvm::div(phi, T)) == fvm::laplacian(kappa, T)          form B200_TEQN_TRANSPORT
// and hand diag / upper / lower / source to the coupled solver.  With a host-side solver library that is a
// host -> device copy of the whole matrix per step (as much traffic as 1-2 Krylov iterations); here the
// matrix is produced on the device, in the layout b200_sys_set_coeffs fills, from static geometry tables
// and the resident old-time field.
//
// Operator arithmetic restated from foam-extend 4.1 (source not in /root/reference):
//   EulerDdtScheme::fvmDdt(vf)        diag = rDeltaT*V                 source = rDeltaT*vf.oldTime()*V
//   EulerDdtScheme::fvmDdt(rho, vf)   diag = rDeltaT*rho*V             source = rDeltaT*rho*vf.oldTime()*V
//   gaussConvectionScheme::fvmDiv     lower = -w*phi; upper = lower + phi; negSumDiag()     (upwind: w = pos(phi))
//   gaussLaplacianScheme::fvmLaplacianUncorrected   upper = deltaCoeffs*gammaMagSf; negSumDiag()
//   lduMatrix::negSumDiag             for faces in order: Diag[l[f]] -= Lower[f]; Diag[u[f]] -= Upper[f]
//   dimensioned<scalar> * fvMatrix    every coefficient and the source times the scalar
//   A == B                            A - B, coefficient by coefficient
//   fvMatrix::solve                   addBoundaryDiag / addBoundarySource: patch by patch, face by face
// Products are rounded one by one (-fmad=false / -ffp-contract=off), left to right as written there.
//
// This header is compiled by nvcc into the kernels of fv_assemble.cuh and by g++ into the CPU emulator of
// tests/cpp/assemble_emulate.cpp, so the index logic and the order of the operations are checked without a GPU.
#pragma once
#include <stdint.h>

#include <vector>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

#define B200_TEQN_CONDUCT 0
#define B200_TEQN_TRANSPORT 1

namespace fvasm
{

struct FaceTerms
{
    double loDiv, upDiv, upLap; // convection lower / upper, laplacian upper (= lower)
};

// the per-face operator coefficients (kappaFace == nullptr: uniform kappa; phi == nullptr: no convection)
B200_HD FaceTerms face_terms(int form, int32_t f, double kappa, const double* kappaFace, const double* magSf,
                             const double* deltaCoeffs, const double* phi)
{
    FaceTerms t;
    const double gammaMagSf = (kappaFace ? kappaFace[f] : kappa) * magSf[f];
    t.upLap = deltaCoeffs[f] * gammaMagSf;
    t.loDiv = 0.0;
    t.upDiv = 0.0;
    if (form == B200_TEQN_TRANSPORT && phi)
    {
        const double ph = phi[f];
        const double w = ph >= 0.0 ? 1.0 : 0.0; // upwind weights: pos(faceFlux)
        t.loDiv = -w * ph;
        t.upDiv = t.loDiv + ph;
    }
    return t;
}

// off-diagonal coefficients of face f of the assembled equation
B200_HD void face_coeffs(int form, double rhoC, const FaceTerms& t, double& upper, double& lower)
{
    if (form == B200_TEQN_TRANSPORT)
    {
        upper = rhoC * t.upDiv - t.upLap;
        lower = rhoC * t.loDiv - t.upLap;
    }
    else
    {
        upper = 0.0 - t.upLap; // a diagonal ddt matrix minus the laplacian
        lower = 0.0 - t.upLap;
    }
}

// Row c of the assembled equation: diagonal and source, boundary contributions included.
//   l, u                 lduAddressing of the region
//   ownerStart[c..c+1]   faces with l[f] == c (contiguous in upper-triangular order)
//   losort, losortStart  faces with u[f] == c, ascending
//   bStart, bInt, bSrc   boundary faces of cell c in patch order: internalCoeffs / boundary source
B200_HD void cell_row(int form, int32_t c, double rhoC, double rDeltaT, double kappa, const double* kappaFace,
                      const double* V, const double* magSf, const double* deltaCoeffs, const double* phi,
                      const int32_t* ownerStart, const int32_t* losort, const int32_t* losortStart,
                      const int32_t* bStart, const double* bInt, const double* bSrc, double Told, double& diag,
                      double& source)
{
    // negSumDiag of the convection and laplacian operators: faces visited in ascending order, whichever
    // side of the face this cell is on
    double dDiv = 0.0, dLap = 0.0;
    int32_t io = ownerStart[c], eo = ownerStart[c + 1];
    int32_t in = losortStart[c], en = losortStart[c + 1];
    while (io < eo || in < en)
    {
        const int32_t fo = io < eo ? io : INT32_MAX;
        const int32_t fn = in < en ? losort[in] : INT32_MAX;
        if (fn < fo)
        { // this cell is the face's upper cell: Diag[u[f]] -= Upper[f]
            const FaceTerms t = face_terms(form, fn, kappa, kappaFace, magSf, deltaCoeffs, phi);
            dDiv -= t.upDiv;
            dLap -= t.upLap;
            in++;
        }
        else
        { // lower cell: Diag[l[f]] -= Lower[f]
            const FaceTerms t = face_terms(form, fo, kappa, kappaFace, magSf, deltaCoeffs, phi);
            dDiv -= t.loDiv;
            dLap -= t.upLap;
            io++;
        }
    }
    if (form == B200_TEQN_TRANSPORT)
    {
        diag = rhoC * (rDeltaT * V[c] + dDiv) - dLap;
        source = rhoC * (rDeltaT * Told * V[c]);
    }
    else
    {
        diag = rDeltaT * rhoC * V[c] - dLap;
        source = rDeltaT * rhoC * Told * V[c];
    }
    for (int32_t k = bStart[c]; k < bStart[c + 1]; k++)
    {
        diag += bInt[k];
        source += bSrc[k];
    }
}

// Host side: the row tables cell_row walks, from the region's lduAddressing and its boundary-face list.
struct RowTables
{
    std::vector<int32_t> ownerStart, losort, losortStart, bStart;
    std::vector<double> bInt, bSrc; // bucketed by cell, patch order kept within a cell
};

// returns false when the faces are not owner-sorted (the owner faces of a cell must be contiguous)
inline bool build_row_tables(int32_t N, int32_t F, const int32_t* l, const int32_t* u, int32_t nB, const int32_t* bCells,
                             const double* bIntCoeffs, const double* bSrcCoeffs, RowTables& T)
{
    for (int32_t f = 1; f < F; f++)
        if (l[f] < l[f - 1]) return false;
    T.ownerStart.assign(N + 1, 0);
    T.losortStart.assign(N + 1, 0);
    T.bStart.assign(N + 1, 0);
    T.losort.assign(F, 0);
    T.bInt.assign(nB, 0.0);
    T.bSrc.assign(nB, 0.0);
    for (int32_t f = 0; f < F; f++)
    {
        T.ownerStart[l[f] + 1]++;
        T.losortStart[u[f] + 1]++;
    }
    for (int32_t k = 0; k < nB; k++) T.bStart[bCells[k] + 1]++;
    for (int32_t c = 0; c < N; c++)
    {
        T.ownerStart[c + 1] += T.ownerStart[c];
        T.losortStart[c + 1] += T.losortStart[c];
        T.bStart[c + 1] += T.bStart[c];
    }
    std::vector<int32_t> fill(T.losortStart.begin(), T.losortStart.end() - 1);
    for (int32_t f = 0; f < F; f++) T.losort[fill[u[f]]++] = f; // stable: ascending faces per upper cell
    fill.assign(T.bStart.begin(), T.bStart.end() - 1);
    for (int32_t k = 0; k < nB; k++)
    {
        const int32_t p = fill[bCells[k]]++;
        T.bInt[p] = bIntCoeffs[k];
        T.bSrc[p] = bSrcCoeffs[k];
    }
    return true;
}

} // namespace fvasm
