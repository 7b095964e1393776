// CPU emulation of the device-side T-equation assembly (TEST INFRASTRUCTURE): the same fv_assemble.hpp the
// kernels k_asm_faces / k_asm_cells of multiregionfoam_b200/csrc/fv_assemble.cuh are built from, with the
// grid-stride thread loops replaced by plain loops and the slot permutation of the resident vectors emulated
// through an explicit slotOfCell table.  Build: g++ -O2 -ffp-contract=off (multiregionfoam_b200/build.py).
#include "../../multiregionfoam_b200/csrc/fv_assemble.hpp"

extern "C" int emu_assemble_T(int form, int nCells, int nFaces, const int32_t* l, const int32_t* u, double rhoC, double rDeltaT,
                              double kappa, const double* kappaFace, const double* V, const double* magSf,
                              const double* deltaCoeffs, const double* phi, int nB, const int32_t* bCells, const double* bInt,
                              const double* bSrc, const int32_t* slotOfCell, const double* xSlots, double* diagCell,
                              double* upper, double* lower, double* bSlots)
{
    fvasm::RowTables T;
    if (!fvasm::build_row_tables(nCells, nFaces, l, u, nB, bCells, bInt, bSrc, T)) return -6;
    const double* dPhi = form == B200_TEQN_TRANSPORT ? phi : nullptr;
    for (int32_t f = 0; f < nFaces; f++) // k_asm_faces
    {
        const fvasm::FaceTerms t = fvasm::face_terms(form, f, kappa, kappaFace, magSf, deltaCoeffs, dPhi);
        fvasm::face_coeffs(form, rhoC, t, upper[f], lower[f]);
    }
    for (int32_t c = 0; c < nCells; c++) // k_asm_cells
    {
        const int slot = slotOfCell[c];
        double d, src;
        fvasm::cell_row(form, c, rhoC, rDeltaT, kappa, kappaFace, V, magSf, deltaCoeffs, dPhi, T.ownerStart.data(),
                        T.losort.data(), T.losortStart.data(), T.bStart.data(), T.bInt.data(), T.bSrc.data(), xSlots[slot], d,
                        src);
        diagCell[c] = d;
        bSlots[slot] = src;
    }
    return 0;
}
