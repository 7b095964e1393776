/*---------------------------------------------------------------------------*\
  b200Binding.C -- see b200Binding.H.
\*---------------------------------------------------------------------------*/
#include "b200Binding.H"
#include "processorLduInterfaceField.H"
#include "regionCoupleFvPatch.H"
#include "regionCouplePolyPatch.H"
#include "fvMesh.H"
#include "Pstream.H"
#include "HashTable.H"

#include <cstdlib>
#include <cstring>
#include <fstream>

namespace Foam
{
    static b200_ctx* ctxPtr_ = NULL;
    // cache: first lduAddressing pointer of the system -> device system (topology change => new addressing)
    static HashTable<b200Binding::systemEntry*, const void*, Hash<const void*> > sysCache_;
}


b200_ctx* Foam::b200Binding::context()
{
    if (!ctxPtr_)
    {
        List<char> id(128, 0);
        if (Pstream::parRun())
        {
            if (Pstream::master()) check(b200_nccl_unique_id(id.begin()), "b200Binding::context()");
            Pstream::scatter(id);
        }
        int nDev = 1; // one GPU per rank of a node: device = rank % (GPUs of the node), B200_DEVICES overrides
        if (const char* e = getenv("B200_DEVICES")) nDev = max(1, atoi(e));
        const int rc = b200_ctx_create
        (
            Pstream::myProcNo() % nDev, Pstream::myProcNo(), Pstream::nProcs(),
            Pstream::parRun() ? id.begin() : NULL, &ctxPtr_
        );
        if (rc)
        {
            FatalErrorIn("b200Binding::context()")
                << "cannot create the B200 context: " << b200_last_error(NULL)
                << " (libb200ldu has no CPU fallback)" << abort(FatalError);
        }
    }
    return ctxPtr_;
}


void Foam::b200Binding::check(const int rc, const char* where)
{
    if (rc < 0)
    {
        FatalErrorIn(where) << "libb200ldu error " << rc << ": " << b200_last_error(ctxPtr_) << abort(FatalError);
    }
}


int Foam::b200Binding::solverId(const word& typeName)
{
    if (typeName == "cudaPCG" || typeName == "PCG" || typeName == "CG") return B200_SOLVER_PCG;
    if (typeName == "cudaPBiCGStab" || typeName == "BiCGStab" || typeName == "PBiCGStab") return B200_SOLVER_BICGSTAB;
    if (typeName == "cudaPBiCG" || typeName == "PBiCG" || typeName == "BiCG") return B200_SOLVER_PBICG;
    FatalErrorIn("b200Binding::solverId(const word&)") << "Unknown solver " << typeName << abort(FatalError);
    return -1;
}


Foam::word Foam::b200Binding::precondName(const dictionary& dict)
{
    word name("none");
    if (dict.found("preconditioner"))
    {
        if (dict.isDict("preconditioner")) dict.subDict("preconditioner").lookup("preconditioner") >> name;
        else dict.lookup("preconditioner") >> name;
    }
    return name;
}


int Foam::b200Binding::precondId(const word& name)
{
    if (name == "cudaDIC" || name == "DIC" || name == "FDIC") return B200_PRECOND_DIC;
    if (name == "cudaDILU" || name == "DILU") return B200_PRECOND_DILU;
    if (name == "Cholesky") return B200_PRECOND_CHOLESKY;
    if (name == "diagonal") return B200_PRECOND_DIAGONAL;
    if (name == "none") return B200_PRECOND_NONE;
    FatalErrorIn("b200Binding::precondId(const word&)")
        << "Unknown preconditioner " << name << " (GAMG etc. are not provided by libb200ldu)" << abort(FatalError);
    return -1;
}


// * * * * * * * * * * * * * * * interface extraction  * * * * * * * * * * * * * //

namespace Foam
{

// ragged -> CSR; rows: the rows to take (empty: all)
static void flatten
(
    const labelListList& addr, const scalarListList& w, const labelList& rows, const bool allRows,
    labelList& offsets, labelList& flatAddr, scalarField& flatW
)
{
    const label nRows = allRows ? addr.size() : rows.size();
    offsets.setSize(nRows + 1);
    label n = 0;
    for (label i = 0; i < nRows; i++)
    {
        offsets[i] = n;
        n += addr[allRows ? i : rows[i]].size();
    }
    offsets[nRows] = n;
    flatAddr.setSize(n);
    flatW.setSize(n);
    n = 0;
    for (label i = 0; i < nRows; i++)
    {
        const label row = allRows ? i : rows[i];
        forAll (addr[row], k)
        {
            flatAddr[n] = addr[row][k];
            flatW[n++] = w[row][k];
        }
    }
}

// The tables regionCoupleFvPatch::interpolate() applies to the shadow's patchInternalField
// (monolithicCouplingFvPatchField.C:214-217, 401-404): values of the shadow patch seen on the faces of rc.
// localParallel(): the interpolator lives on the two patches.  Otherwise (the pair is spread over processors, e.g. the
// shipped  simple; n (2 1 2)  of flowOverHeatedPlate) it lives on the global face zones: rows = zone faces, addresses =
// shadow zone faces; this rank needs the rows zoneAddressing() of its own faces, and the library is told which rank
// holds which shadow zone faces (pass 4 of describe, b200_sys_set_interface_pieces).
static void regionCoupleTables(const regionCoupleFvPatch& rc, b200Binding::ifaceInfo& I)
{
    const regionCouplePolyPatch& pp = refCast<const regionCouplePolyPatch>(rc.patch());
    I.zoneMode = Pstream::parRun() && !rc.localParallel();
    I.nPeerFaces = I.zoneMode ? pp.shadow().zone().size() : rc.shadow().size();
    if (I.zoneMode) I.zoneAddr = pp.zoneAddressing();
    if (pp.master())
    {
        // interpolate() = patchToPatch().slaveToMaster(): result[mf] = sum_k ff[masterAddr[mf][k]]*masterWeights[mf][k]
        flatten(pp.patchToPatch().masterAddr(), pp.patchToPatch().masterWeights(), I.zoneAddr, !I.zoneMode, I.ggiOffsets, I.ggiAddr, I.ggiWeights);
    }
    else
    {
        // interpolate() = shadow().patchToPatch().masterToSlave(): the master owns the interpolator
        flatten
        (
            pp.shadow().patchToPatch().slaveAddr(), pp.shadow().patchToPatch().slaveWeights(), I.zoneAddr, !I.zoneMode,
            I.ggiOffsets, I.ggiAddr, I.ggiWeights
        );
    }
}

} // End namespace Foam


void Foam::b200Binding::describe
(
    const UPtrList<const lduMatrix>& matrices,
    const List<lduInterfaceFieldPtrsList>& interfaces,
    List<List<ifaceInfo> >& ifaces
)
{
    const label nRows = matrices.size();
    ifaces.setSize(nRows);

    // ---- pass 1: which patches are interfaces, and of what kind
    forAll (matrices, r)
    {
        const lduInterfaceFieldPtrsList& ifs = interfaces[r];
        label n = 0;
        forAll (ifs, patchI) if (ifs.set(patchI)) n++;
        ifaces[r].setSize(n);
        n = 0;
        forAll (ifs, patchI)
        {
            if (!ifs.set(patchI)) continue;
            const lduInterfaceField& f = ifs[patchI];
            ifaceInfo& I = ifaces[r][n++];
            I.patch = patchI;
            if (isA<processorLduInterfaceField>(f))
            {
                I.kind = B200_IFACE_PROCESSOR;
                I.peerRank = refCast<const processorLduInterfaceField>(f).neighbProcNo();
                I.peerRegion = r;
                I.nPeerFaces = matrices[r].lduAddr().patchAddr(patchI).size();
            }
            else if (isA<regionCoupleFvPatch>(f.coupledInterface()))
            {
                I.kind = B200_IFACE_REGION_COUPLE;
                I.peerRank = Pstream::myProcNo();
            }
            else
            {
                FatalErrorIn("b200Binding::describe(...)")
                    << "lduInterfaceField of type " << f.type() << " on patch " << patchI
                    << " is not supported on the device (no CPU fallback)" << abort(FatalError);
            }
        }
    }

    // ---- pass 2: regionCouple: shadow (row, interface) and the interpolation tables
    forAll (matrices, r)
    {
        forAll (ifaces[r], i)
        {
            ifaceInfo& I = ifaces[r][i];
            if (I.kind != B200_IFACE_REGION_COUPLE) continue;
            const regionCoupleFvPatch& rc = refCast<const regionCoupleFvPatch>(interfaces[r][I.patch].coupledInterface());
            // the row of the shadow region: the matrix that lives on shadowRegion()'s addressing
            // (monolithicCouplingFvPatchField.C:183-187 looks the shadow FIELD up in shadowRegion(); the matrix of that
            // field is a row of the same coupledLduMatrix, multiRegionSystem.C:143-145)
            const lduAddressing* shadowAddr = &rc.shadowRegion().lduAddr();
            I.peerRegion = -1;
            forAll (matrices, q) if (&matrices[q].lduAddr() == shadowAddr) I.peerRegion = q;
            if (I.peerRegion < 0)
            {
                FatalErrorIn("b200Binding::describe(...)")
                    << "the shadow region of regionCouple patch " << rc.name() << " has no row in this coupled matrix"
                    << abort(FatalError);
            }
            I.peerIface = -1;
            forAll (ifaces[I.peerRegion], k) if (ifaces[I.peerRegion][k].patch == rc.shadowIndex()) I.peerIface = k;
            if (I.peerIface < 0)
            {
                FatalErrorIn("b200Binding::describe(...)")
                    << "shadow patch " << rc.shadowIndex() << " of regionCouple patch " << rc.name()
                    << " is not a coupled patch of its row" << abort(FatalError);
            }
            regionCoupleTables(rc, I);
        }
    }

    // ---- pass 3: processor patches: the peer's interface index.  Every rank publishes, per row, (neighbour rank,
    // interface index) of its processor interfaces in patch order; the k-th patch of rank A towards rank B in a row
    // pairs with the k-th patch of B towards A in the same row (decomposePar creates the two sides of a cut in the
    // same order on both processors).
    if (Pstream::parRun())
    {
        List<labelList> table(Pstream::nProcs());
        {
            DynamicList<label> mine;
            forAll (ifaces, r) forAll (ifaces[r], i)
            {
                if (ifaces[r][i].kind != B200_IFACE_PROCESSOR) continue;
                mine.append(r);
                mine.append(ifaces[r][i].peerRank);
                mine.append(i);
            }
            table[Pstream::myProcNo()] = labelList(mine);
        }
        Pstream::gatherList(table);
        Pstream::scatterList(table);
        forAll (ifaces, r) forAll (ifaces[r], i)
        {
            ifaceInfo& I = ifaces[r][i];
            if (I.kind != B200_IFACE_PROCESSOR) continue;
            label kMine = 0; // rank of this patch among my patches of row r towards the same neighbour
            for (label j = 0; j < i; j++)
                if (ifaces[r][j].kind == B200_IFACE_PROCESSOR && ifaces[r][j].peerRank == I.peerRank) kMine++;
            const labelList& theirs = table[I.peerRank];
            label k = 0;
            for (label e = 0; e + 2 < theirs.size() + 0 && I.peerIface < 0; e += 3)
            {
                if (theirs[e] == r && theirs[e + 1] == Pstream::myProcNo())
                {
                    if (k == kMine) I.peerIface = theirs[e + 2];
                    k++;
                }
            }
            if (I.peerIface < 0)
            {
                FatalErrorIn("b200Binding::describe(...)")
                    << "processor " << I.peerRank << " has no processor patch of row " << r << " towards processor "
                    << Pstream::myProcNo() << " that matches patch " << I.patch << abort(FatalError);
            }
        }

        // ---- pass 4: regionCouple pairs on the global zones: who holds which faces of the shadow zone.  Every rank
        // publishes (row, patch, interface index, nFaces, zoneAddressing...) of its zone-mode regionCouple patches.
        {
            List<labelList> zt(Pstream::nProcs());
            DynamicList<label> mine;
            forAll (ifaces, r) forAll (ifaces[r], i)
            {
                const ifaceInfo& I = ifaces[r][i];
                if (I.kind != B200_IFACE_REGION_COUPLE || !I.zoneMode) continue;
                mine.append(r);
                mine.append(I.patch);
                mine.append(i);
                mine.append(I.zoneAddr.size());
                forAll (I.zoneAddr, f) mine.append(I.zoneAddr[f]);
            }
            zt[Pstream::myProcNo()] = labelList(mine);
            Pstream::gatherList(zt);
            Pstream::scatterList(zt);
            forAll (ifaces, r) forAll (ifaces[r], i)
            {
                ifaceInfo& I = ifaces[r][i];
                if (I.kind != B200_IFACE_REGION_COUPLE || !I.zoneMode) continue;
                const regionCoupleFvPatch& rc = refCast<const regionCoupleFvPatch>(interfaces[r][I.patch].coupledInterface());
                DynamicList<label> pRank, pIface, pOff, pAddr;
                pOff.append(0);
                if (I.zoneAddr.size())   // an empty piece reads nothing
                {
                    forAll (zt, h)
                    {
                        const labelList& t = zt[h];
                        for (label e = 0; e + 3 < t.size(); e += 4 + t[e + 3])
                        {
                            if (t[e] != I.peerRegion || t[e + 1] != rc.shadowIndex() || t[e + 3] == 0) continue;
                            pRank.append(h);
                            pIface.append(t[e + 2]);
                            for (label f = 0; f < t[e + 3]; f++) pAddr.append(t[e + 4 + f]);
                            pOff.append(pAddr.size());
                        }
                    }
                }
                I.pieceRank = labelList(pRank);
                I.pieceIface = labelList(pIface);
                I.pieceOffsets = labelList(pOff);
                I.pieceZoneAddr = labelList(pAddr);
            }
        }
    }
}


Foam::b200Binding::systemEntry& Foam::b200Binding::system
(
    const UPtrList<const lduMatrix>& matrices,
    const List<lduInterfaceFieldPtrsList>& interfaces
)
{
    const char* where = "b200Binding::system(...)";
    const void* key = &matrices[0].lduAddr();
    if (sysCache_.found(key))
    {
        // cached: the patches cannot have changed (a topology change gives new lduAddressing objects), but attach() /
        // detach() flip the regionCouple patches and re-compute the interpolation (regionInterfaceType.C:543-627)
        systemEntry& E = *sysCache_[key];
        forAll (E.ifaces, r) forAll (E.ifaces[r], i)
        {
            ifaceInfo& I = E.ifaces[r][i];
            if (I.kind != B200_IFACE_REGION_COUPLE) continue;
            const regionCoupleFvPatch& rc = refCast<const regionCoupleFvPatch>(interfaces[r][I.patch].coupledInterface());
            check(b200_sys_set_interface_attached(E.sys, r, i, rc.coupled() ? 1 : 0), where);
            if (!rc.coupled()) continue;
            ifaceInfo now;
            regionCoupleTables(rc, now);
            if (now.ggiAddr != I.ggiAddr || now.ggiOffsets != I.ggiOffsets || now.ggiWeights != I.ggiWeights)
            {
                I.ggiOffsets = now.ggiOffsets;
                I.ggiAddr = now.ggiAddr;
                I.ggiWeights = now.ggiWeights;
                I.nPeerFaces = now.nPeerFaces;
                check
                (
                    b200_sys_set_interface_ggi(E.sys, r, i, I.nPeerFaces, I.ggiOffsets.begin(), I.ggiAddr.begin(), I.ggiWeights.begin()),
                    where
                );
            }
        }
        return E;
    }

    systemEntry* Eptr = new systemEntry;
    systemEntry& E = *Eptr;
    describe(matrices, interfaces, E.ifaces);
    check(b200_sys_create(context(), matrices.size(), &E.sys), where);
    forAll (matrices, r)
    {
        const lduAddressing& addr = matrices[r].lduAddr();
        check
        (
            b200_sys_set_region(E.sys, r, addr.size(), addr.lowerAddr().size(), addr.lowerAddr().begin(), addr.upperAddr().begin()),
            where
        );
    }
    forAll (matrices, r)
    {
        forAll (E.ifaces[r], i)
        {
            const ifaceInfo& I = E.ifaces[r][i];
            const unallocLabelList& fc = matrices[r].lduAddr().patchAddr(I.patch);
            const bool ggi = I.ggiOffsets.size() > 0;
            const int idx = b200_sys_add_interface
            (
                E.sys, r, I.kind, fc.size(), fc.begin(), I.peerRank, I.peerRegion, I.peerIface, I.nPeerFaces,
                ggi ? I.ggiOffsets.begin() : NULL, ggi ? I.ggiAddr.begin() : NULL, ggi ? I.ggiWeights.begin() : NULL
            );
            check(idx, where);
            if (idx != i)
            {
                FatalErrorIn(where) << "interface numbering out of step: " << idx << " != " << i << abort(FatalError);
            }
            if (I.kind == B200_IFACE_REGION_COUPLE)
            {
                const regionCoupleFvPatch& rc = refCast<const regionCoupleFvPatch>(interfaces[r][I.patch].coupledInterface());
                check(b200_sys_set_interface_attached(E.sys, r, i, rc.coupled() ? 1 : 0), where);
                if (I.zoneMode && I.pieceRank.size())
                {
                    labelList pieceRegion(I.pieceRank.size(), I.peerRegion);
                    check
                    (
                        b200_sys_set_interface_pieces
                        (
                            E.sys, r, i, I.pieceRank.size(), I.pieceRank.begin(), pieceRegion.begin(), I.pieceIface.begin(),
                            I.pieceOffsets.begin(), I.pieceZoneAddr.begin()
                        ),
                        where
                    );
                }
            }
        }
    }
    check(b200_sys_finalize(E.sys), where);
    sysCache_.insert(key, Eptr);
    return E;
}


void Foam::b200Binding::setCoeffs
(
    const systemEntry& E,
    const UPtrList<const lduMatrix>& matrices,
    const List<const FieldField<Field, scalar>*>& bouCoeffs,
    const List<const FieldField<Field, scalar>*>& intCoeffs
)
{
    forAll (matrices, r)
    {
        const lduMatrix& m = matrices[r];
        check
        (
            b200_sys_set_coeffs(E.sys, r, m.diag().begin(), m.upper().begin(), m.asymmetric() ? m.lower().begin() : NULL),
            "b200Binding::setCoeffs(...)"
        );
        forAll (E.ifaces[r], i)
        {
            const label patchI = E.ifaces[r][i].patch;
            check
            (
                b200_sys_set_interface_coeffs(E.sys, r, i, (*bouCoeffs[r])[patchI].begin(), (*intCoeffs[r])[patchI].begin()),
                "b200Binding::setCoeffs(...)"
            );
        }
    }
}


// * * * * * * * * * * * * * * * * * * dump  * * * * * * * * * * * * * * * * * * //

namespace Foam
{
    static void put(std::ofstream& os, const void* p, const size_t bytes) { os.write(reinterpret_cast<const char*>(p), bytes); }
    static void putI(std::ofstream& os, const int v) { put(os, &v, sizeof(int)); }
    static void putD(std::ofstream& os, const double v) { put(os, &v, sizeof(double)); }
    static void putLabels(std::ofstream& os, const UList<label>& l)
    {
        // WM_LABEL_SIZE=32 in the reference build (compile_commands.json): label == int32
        forAll (l, i) putI(os, l[i]);
    }
    static void putScalars(std::ofstream& os, const UList<scalar>& f)
    {
        forAll (f, i) putD(os, f[i]);
    }
    static void putName(std::ofstream& os, const word& w)
    {
        char buf[32];
        memset(buf, 0, sizeof(buf));
        strncpy(buf, w.c_str(), sizeof(buf) - 1);
        put(os, buf, sizeof(buf));
    }
}


void Foam::b200Binding::dump
(
    const fileName& file,
    const UPtrList<const lduMatrix>& matrices,
    const List<lduInterfaceFieldPtrsList>& interfaces,
    const List<const FieldField<Field, scalar>*>& bouCoeffs,
    const List<const FieldField<Field, scalar>*>& intCoeffs,
    const UPtrList<const scalarField>& x,
    const UPtrList<const scalarField>& b,
    const word& solverName,
    const word& preconName,
    const scalar tolerance,
    const scalar relTol,
    const label minIter,
    const label maxIter,
    const scalarField& history
)
{
    // layout: multiregionfoam_b200/dumpio.py (format "B200LDU1", little endian)
    List<List<ifaceInfo> > ifaces;
    describe(matrices, interfaces, ifaces);

    std::ofstream os(file.c_str(), std::ios::binary);
    if (!os.good())
    {
        FatalErrorIn("b200Binding::dump(...)") << "cannot open " << file << abort(FatalError);
    }
    put(os, "B200LDU1", 8);
    putI(os, Pstream::myProcNo());
    putI(os, Pstream::nProcs());
    putI(os, matrices.size());
    forAll (matrices, r)
    {
        const lduMatrix& m = matrices[r];
        const lduAddressing& addr = m.lduAddr();
        putI(os, addr.size());
        putI(os, addr.lowerAddr().size());
        putI(os, m.asymmetric() ? 0 : 1);
        putI(os, ifaces[r].size());
        putLabels(os, addr.lowerAddr());
        putLabels(os, addr.upperAddr());
        putScalars(os, m.diag());
        putScalars(os, m.upper());
        if (m.asymmetric()) putScalars(os, m.lower());
        putScalars(os, b[r]);
        putScalars(os, x[r]);
        forAll (ifaces[r], i)
        {
            const ifaceInfo& I = ifaces[r][i];
            const unallocLabelList& fc = addr.patchAddr(I.patch);
            const bool ggi = I.ggiOffsets.size() > 0;
            if (I.zoneMode)
            {
                FatalErrorIn("b200Binding::dump(...)")
                    << "patch " << I.patch << " of row " << r << ": regionCouple pairs spread over processors are not covered by the "
                    << "B200LDU1 format; record golden data with a decomposition that keeps the pair on one processor"
                    << abort(FatalError);
            }
            putI(os, I.kind);
            putI(os, fc.size());
            putI(os, I.peerRank);
            putI(os, I.peerRegion);
            putI(os, I.peerIface);
            putI(os, I.nPeerFaces);
            putI(os, ggi ? 1 : 0);
            putLabels(os, fc);
            putScalars(os, (*bouCoeffs[r])[I.patch]);
            putScalars(os, (*intCoeffs[r])[I.patch]);
            if (ggi)
            {
                putLabels(os, I.ggiOffsets);
                putLabels(os, I.ggiAddr);
                putScalars(os, I.ggiWeights);
            }
        }
    }
    putName(os, solverName);
    putName(os, preconName);
    putD(os, tolerance);
    putD(os, relTol);
    putI(os, minIter);
    putI(os, maxIter);
    putI(os, history.size());
    putScalars(os, history);
}
