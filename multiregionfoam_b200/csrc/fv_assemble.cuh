// fv_assemble.cuh -- device-side coefficient refresh of the T equations: kernels and the two C-ABI entry
// points (b200_sys_set_fv_geometry, b200_sys_assemble_T).  Included by b200_ldu.cu after b200_sys (which owns the per-region FvRegionDev tables).
// Arithmetic: fv_assemble.hpp (shared with the CPU emulator of the tests).
//
// Both kernels are HBM-bound streams.  Per face: magSf, deltaCoeffs (+ phi, kappaFace) read, upper and lower
// written: 32-48 B.  Per cell: V, the row pointers, x (gathered through slotOfCell) read, diag and b written, and
// the face tables of the row's ~6 faces re-read (served by L2: the faces of a row were touched by the row's
// neighbours a few hundred threads earlier): ~72 B + the re-reads.
#pragma once

#include "fv_assemble.hpp"

namespace b200
{

__global__ void k_asm_faces(int form, int32_t nFaces, double rhoC, double kappa, const double* __restrict__ kappaFace,
                            const double* __restrict__ magSf, const double* __restrict__ delta, const double* __restrict__ phi,
                            double* __restrict__ upper, double* __restrict__ lower)
{
    for (int32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nFaces; f += gridDim.x * blockDim.x)
    {
        const fvasm::FaceTerms t = fvasm::face_terms(form, f, kappa, kappaFace, magSf, delta, phi);
        double up, lo;
        fvasm::face_coeffs(form, rhoC, t, up, lo);
        upper[f] = up;
        lower[f] = lo;
    }
}

__global__ void k_asm_cells(int form, int32_t nCells, double rhoC, double rDeltaT, double kappa, const double* __restrict__ kappaFace,
                            const double* __restrict__ V, const double* __restrict__ magSf, const double* __restrict__ delta,
                            const double* __restrict__ phi, const int* __restrict__ ownerStart, const int* __restrict__ losort,
                            const int* __restrict__ losortStart, const int* __restrict__ bStart, const double* __restrict__ bInt,
                            const double* __restrict__ bSrc, const int* __restrict__ slotOfCell /* of this region */,
                            const double* __restrict__ xSlots, double* __restrict__ diagCell /* of this region */,
                            double* __restrict__ bSlots)
{
    for (int32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nCells; c += gridDim.x * blockDim.x)
    {
        const int slot = slotOfCell[c];
        double d, src;
        fvasm::cell_row(form, c, rhoC, rDeltaT, kappa, kappaFace, V, magSf, delta, phi, ownerStart, losort, losortStart,
                        bStart, bInt, bSrc, xSlots[slot], d, src);
        diagCell[c] = d;
        bSlots[slot] = src;
    }
}

} // namespace b200

static int fv_launch_blocks(const b200_sys* s, int64_t n)
{
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)s->ctx->smCount * 8));
}

extern "C" int b200_sys_set_fv_geometry(b200_sys* s, int r, const double* V, const double* magSf, const double* deltaCoeffs,
                                        int32_t nBoundary, const int32_t* bCells, const double* bIntCoeffs,
                                        const double* bSrcCoeffs)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->finalized) return set_err(ctx, B200_ESTATE, "set_fv_geometry before finalize");
    if (r < 0 || r >= (int)s->regs.size()) return set_err(ctx, B200_EINVAL, "b200_sys_set_fv_geometry: bad region");
    const RegionHost& R = s->regs[r];
    if ((R.nCells && !V) || (R.nFaces && (!magSf || !deltaCoeffs)) || nBoundary < 0 ||
        (nBoundary && (!bCells || !bIntCoeffs || !bSrcCoeffs)))
        return set_err(ctx, B200_EINVAL, "b200_sys_set_fv_geometry: bad arguments");
    for (int32_t k = 0; k < nBoundary; k++)
        if (bCells[k] < 0 || bCells[k] >= R.nCells)
            return set_err(ctx, B200_EINVAL, "region %d boundary face %d: cell %d out of range", r, k, bCells[k]);
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    try
    {
        const int32_t N = R.nCells, F = R.nFaces;
        fvasm::RowTables T;
        // the row gather relies on the owner faces of a cell being contiguous (upper-triangular face order)
        if (!fvasm::build_row_tables(N, F, R.l.data(), R.u.data(), nBoundary, bCells, bIntCoeffs, bSrcCoeffs, T))
            return set_err(ctx, B200_EUNSUPPORTED, "region %d: faces are not in upper-triangular (owner-sorted) order", r);
        if (s->fv.size() != s->regs.size()) s->fv.resize(s->regs.size());
        std::unique_ptr<FvRegionDev> fresh(new FvRegionDev); // installed only when every table is on the device
        FvRegionDev& G = *fresh;
        const std::vector<double> hV(V, V + N), hMagSf(magSf, magSf + F), hDelta(deltaCoeffs, deltaCoeffs + F);
        CK(ctx, G.V.upload(hV, st));
        CK(ctx, G.magSf.upload(hMagSf, st));
        CK(ctx, G.delta.upload(hDelta, st));
        CK(ctx, G.ownerStart.upload(T.ownerStart, st));
        CK(ctx, G.losort.upload(T.losort, st));
        CK(ctx, G.losortStart.upload(T.losortStart, st));
        CK(ctx, G.bStart.upload(T.bStart, st));
        CK(ctx, G.bInt.upload(T.bInt, st));
        CK(ctx, G.bSrc.upload(T.bSrc, st));
        CK(ctx, G.phi.alloc(F));
        CK(ctx, G.kappaFace.alloc(F));
        CK(ctx, cudaStreamSynchronize(st)); // the host vectors above die here
        s->fv[r] = std::move(fresh);
    }
    catch (const std::exception& e)
    {
        return set_err(ctx, B200_ENOMEM, "b200_sys_set_fv_geometry: %s", e.what());
    }
    return B200_OK;
}

extern "C" int b200_sys_assemble_T(b200_sys* s, int r, int form, double rhoC, double rDeltaT, double kappa,
                                   const double* kappaFace, const double* phi)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->finalized) return set_err(ctx, B200_ESTATE, "assemble_T before finalize");
    if (r < 0 || r >= (int)s->regs.size() || (form != B200_TEQN_CONDUCT && form != B200_TEQN_TRANSPORT))
        return set_err(ctx, B200_EINVAL, "b200_sys_assemble_T: bad arguments");
    if (s->fv.size() != s->regs.size() || !s->fv[r])
        return set_err(ctx, B200_ESTATE, "assemble_T: region %d has no geometry (b200_sys_set_fv_geometry)", r);
    const RegionHost& R = s->regs[r];
    FvRegionDev& G = *s->fv[r];
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // per-step inputs: only what changed crosses the bus (NULL: keep the resident table)
    if (phi && R.nFaces)
    {
        CK(ctx, cudaMemcpyAsync(G.phi.p, phi, sizeof(double) * R.nFaces, cudaMemcpyHostToDevice, st));
        G.havePhi = true;
    }
    if (kappaFace && R.nFaces)
    {
        CK(ctx, cudaMemcpyAsync(G.kappaFace.p, kappaFace, sizeof(double) * R.nFaces, cudaMemcpyHostToDevice, st));
        G.haveKappaFace = true;
    }
    if (form == B200_TEQN_TRANSPORT && !G.havePhi && R.nFaces)
        return set_err(ctx, B200_ESTATE, "assemble_T: transport form needs the face flux phi at least once");
    const double* dPhi = (form == B200_TEQN_TRANSPORT && G.havePhi) ? G.phi.p : nullptr;
    const double* dKf = G.haveKappaFace ? G.kappaFace.p : nullptr;
    if (R.nFaces)
    {
        KScope k(s, B200_K_PACK);
        k_asm_faces<<<fv_launch_blocks(s, R.nFaces), 256, 0, st>>>(form, R.nFaces, rhoC, kappa, dKf, G.magSf.p, G.delta.p, dPhi,
                                                                  s->coef.p + R.faceOffset, s->coef.p + s->F + R.faceOffset);
        CK(ctx, cudaGetLastError());
    }
    if (R.nCells)
    {
        KScope k(s, B200_K_PACK);
        k_asm_cells<<<fv_launch_blocks(s, R.nCells), 256, 0, st>>>(
            form, R.nCells, rhoC, rDeltaT, kappa, dKf, G.V.p, G.magSf.p, G.delta.p, dPhi, G.ownerStart.p, G.losort.p,
            G.losortStart.p, G.bStart.p, G.bInt.p, G.bSrc.p, s->slotOfCell.p + R.cellOffset, s->vec[V_X].p,
            s->diagCell.p + R.cellOffset, s->vec[V_B].p);
        CK(ctx, cudaGetLastError());
    }
    // same invalidation as b200_sys_set_coeffs
    s->diagDirty = true;
    s->regionHasCoeffs[r] = 1;
    s->sellDirty = s->sellTDirty = true;
    s->precondValid = -1;
    s->fwd.packed[0] = s->fwd.packed[1] = s->bwd.packed[0] = s->bwd.packed[1] = false;
    return B200_OK;
}
