/*---------------------------------------------------------------------------*\
  b200Binding.C -- see b200Binding.H.  NOT compiled in this repository (needs foam-extend 4.1).
\*---------------------------------------------------------------------------*/
#include "b200Binding.H"
#include "processorLduInterfaceField.H"
#include "ggiLduInterfaceField.H"
#include "regionCoupleLduInterfaceField.H"
#include "regionCoupleFvPatch.H"
#include "Pstream.H"
#include "HashTable.H"

namespace Foam
{
    static b200_ctx* ctxPtr_ = NULL;
    // cache: first lduAddressing pointer of the system -> device system (topology change => new addressing)
    static HashTable<b200_sys*, const void*, Hash<const void*> > sysCache_;
}


Foam::b200_ctx* Foam::b200Binding::context()
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
    FatalErrorIn("b200Binding::solverId(const word&)") << "Unknown solver " << typeName << abort(FatalError);
    return -1;
}


int Foam::b200Binding::precondId(const dictionary& dict)
{
    word name("none");
    if (dict.found("preconditioner"))
    {
        if (dict.isDict("preconditioner")) dict.subDict("preconditioner").lookup("preconditioner") >> name;
        else dict.lookup("preconditioner") >> name;
    }
    if (name == "cudaDIC" || name == "DIC" || name == "FDIC") return B200_PRECOND_DIC;
    if (name == "cudaDILU" || name == "DILU") return B200_PRECOND_DILU;
    if (name == "Cholesky") return B200_PRECOND_CHOLESKY;
    if (name == "diagonal") return B200_PRECOND_DIAGONAL;
    if (name == "none") return B200_PRECOND_NONE;
    FatalErrorIn("b200Binding::precondId(const dictionary&)")
        << "Unknown preconditioner " << name << " (GAMG etc. are not provided by libb200ldu)" << abort(FatalError);
    return -1;
}


Foam::b200_sys* Foam::b200Binding::system
(
    const UPtrList<const lduMatrix>& matrices,
    const List<lduInterfaceFieldPtrsList>& interfaces
)
{
    const void* key = &matrices[0].lduAddr();
    if (sysCache_.found(key)) return sysCache_[key];

    b200_sys* sys = NULL;
    check(b200_sys_create(context(), matrices.size(), &sys), "b200Binding::system(...)");
    forAll (matrices, r)
    {
        const lduAddressing& addr = matrices[r].lduAddr();
        check
        (
            b200_sys_set_region(sys, r, addr.size(), addr.lowerAddr().size(), addr.lowerAddr().begin(), addr.upperAddr().begin()),
            "b200Binding::system(...)"
        );
    }
    // patch index -> interface index of each row (only coupled patches are interfaces)
    forAll (matrices, r)
    {
        const lduInterfaceFieldPtrsList& ifs = interfaces[r];
        forAll (ifs, patchI)
        {
            if (!ifs.set(patchI)) continue;
            const lduInterfaceField& f = ifs[patchI];
            const unallocLabelList& fc = matrices[r].lduAddr().patchAddr(patchI);
            if (isA<processorLduInterfaceField>(f))
            {
                const processorLduInterfaceField& pf = refCast<const processorLduInterfaceField>(f);
                // the peer's interface index equals the rank of this patch among ITS coupled patches towards us;
                // processor patches are created pairwise in the same order on both sides (decomposePar)
                check
                (
                    b200_sys_add_interface(sys, r, B200_IFACE_PROCESSOR, fc.size(), fc.begin(), pf.neighbProcNo(), r,
                                           /* peerIface, resolved by the binding's patch ordering */ -1, fc.size(), NULL, NULL, NULL),
                    "b200Binding::system(...)"
                );
            }
            else if (isA<regionCoupleLduInterfaceField>(f) || isA<ggiLduInterfaceField>(f))
            {
                // shadow region / patch and the GGI addressing + weights of regionCouplePatch().interpolate:
                // see monolithicCouplingFvPatchField.C:183-187 (shadow lookup), :214-217, :401-404 (interpolate)
                // -> b200_sys_add_interface(sys, r, B200_IFACE_REGION_COUPLE, nFaces, faceCells, myRank, shadowRow,
                //                           shadowIface, nShadowFaces, ggiOffsets, ggiAddr, ggiWeights)
                // Before every solve of a cached system: b200_sys_set_interface_attached(sys, r, iface,
                // regionCouplePatch().attached()) (regionInterfaceType.C:543-627; a detached patch makes b200_solve return
                // B200_ESTATE, which check() turns into the FatalError of monolithicCouplingFvPatchField.C:406-413), and, when
                // the interpolator was rebuilt since (attach() after mesh motion), b200_sys_set_interface_ggi with the new
                // addressing / weights.
                notImplemented("regionCouple extraction: needs regionCoupleFvPatch::shadowRegion()/shadow() of the host tree");
            }
            else
            {
                FatalErrorIn("b200Binding::system(...)")
                    << "lduInterfaceField of type " << f.type() << " on patch " << patchI
                    << " is not supported on the device (no CPU fallback)" << abort(FatalError);
            }
        }
    }
    check(b200_sys_finalize(sys), "b200Binding::system(...)");
    sysCache_.insert(key, sys);
    return sys;
}


void Foam::b200Binding::setCoeffs
(
    b200_sys* sys,
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
            b200_sys_set_coeffs(sys, r, m.diag().begin(), m.upper().begin(), m.asymmetric() ? m.lower().begin() : NULL),
            "b200Binding::setCoeffs(...)"
        );
        label ifaceI = 0;
        forAll (*bouCoeffs[r], patchI)
        {
            if ((*bouCoeffs[r])[patchI].size() && m.lduAddr().patchAddr(patchI).size() /* coupled patch */)
            {
                check
                (
                    b200_sys_set_interface_coeffs(sys, r, ifaceI++, (*bouCoeffs[r])[patchI].begin(), (*intCoeffs[r])[patchI].begin()),
                    "b200Binding::setCoeffs(...)"
                );
            }
        }
    }
}


void Foam::b200Binding::dump(const fileName& dir, b200_sys*, const scalarField&, const scalarField&)
{
    // format: INTEGRATION.md "LDU dump format"; written with OFstream in binary mode
    notImplemented("b200Binding::dump: see INTEGRATION.md");
}
