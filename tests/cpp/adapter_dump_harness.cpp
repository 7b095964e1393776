// adapter_dump_harness.cpp -- drives adapters/b200LduSolvers/b200Binding.C (describe + dump) against MOCK objects built on
// the stand-in declarations of adapters/foamStub: a two-row coupled matrix (fluid 6 cells, solid 4 cells) whose rows are
// joined by a regionCouple pair with a non-conformal GGI (3 faces against 2) and, on rank 1 of 2, a processor patch.
// The stub declares the foam-extend interfaces without bodies; this file supplies bodies backed by plain tables, so
// the adapter's own logic - interface classification, shadow row / interface lookup through shadowRegion().lduAddr(),
// master / slave GGI table selection and flattening, processor peer resolution through Pstream::gatherList, and the
// B200LDU1 writer - runs for real.  tests/test_adapter_dump.py reads the file back with multiregionfoam_b200.dumpio.
//   usage: adapter_dump_harness <out file> <rank>
#include "b200Binding.H"
#include "processorLduInterfaceField.H"
#include "regionCoupleFvPatch.H"
#include "regionCouplePolyPatch.H"
#include "fvMesh.H"
#include "Pstream.H"

#include <map>

namespace Foam
{
error FatalError;
std::ostream& Info = std::cout;
bool mkDir(const fileName&) { return true; }

// ---- mock state ---------------------------------------------------------------------------------
static int g_rank = 0, g_nprocs = 1;
bool Pstream::parRun() { return g_nprocs > 1; }
bool Pstream::master() { return g_rank == 0; }
int Pstream::myProcNo() { return g_rank; }
int Pstream::nProcs() { return g_nprocs; }
// what the other rank would have published in describe()'s gatherList: (row, neighbour, interface index) triples
static std::map<int, labelList> g_otherTables;
// ... and in the second gatherList of describe() (zone pieces of regionCouple pairs spread over processors)
static std::map<int, labelList> g_otherZoneTables;
static int g_gatherCalls = 0;
void Pstream::gatherListHook(void* p)
{
    List<labelList>& l = *static_cast<List<labelList>*>(p);
    for (auto& kv : (g_gatherCalls == 0 ? g_otherTables : g_otherZoneTables)) l[kv.first] = kv.second;
    g_gatherCalls++;
}

struct AddrData
{
    labelList lower, upper;
    std::vector<labelList> patches;
    label nCells;
};
static std::map<const lduAddressing*, AddrData*> g_addr;
class mockAddressing : public lduAddressing
{
public:
    AddrData d;
    mockAddressing() { g_addr[this] = &d; }
    const unallocLabelList& lowerAddr() const { return d.lower; }
    const unallocLabelList& upperAddr() const { return d.upper; }
    const unallocLabelList& patchAddr(const label i) const { return d.patches[i]; }
};
label lduAddressing::size() const { return g_addr[this]->nCells; }

struct MatData
{
    const lduAddressing* addr;
    scalarField diag, upper, lower;
    bool asym;
};
static std::map<const lduMatrix*, MatData> g_mat;
const lduAddressing& lduMatrix::lduAddr() const { return *g_mat[this].addr; }
const scalarField& lduMatrix::diag() const { return g_mat[this].diag; }
const scalarField& lduMatrix::upper() const { return g_mat[this].upper; }
const scalarField& lduMatrix::lower() const { return g_mat[this].lower; }
bool lduMatrix::asymmetric() const { return g_mat[this].asym; }
bool lduMatrix::symmetric() const { return !g_mat[this].asym; }

static std::map<const lduInterfaceField*, const lduInterface*> g_ifaceOf;
const lduInterface& lduInterfaceField::coupledInterface() const { return *g_ifaceOf[this]; }

// regionCouple patch pair
struct RcData
{
    bool master, coupled;
    label size, shadowIndex;
    const regionCoupleFvPatch* shadow;
    const fvMesh* shadowMesh;
    const regionCouplePolyPatch* poly;
    word name;
};
static std::map<const regionCoupleFvPatch*, RcData> g_rc;
static std::map<const fvMesh*, const lduAddressing*> g_meshAddr;
static std::map<const regionCouplePolyPatch*, const regionCoupleFvPatch*> g_polyOwner;
static ggiZoneInterpolation g_interp; // owned by the master
static labelListList g_mAddr, g_sAddr;
static scalarListList g_mW, g_sW;
static word g_rcType("regionCouple");
const word& regionCoupleFvPatch::name() const { return g_rc[this].name; }
label regionCoupleFvPatch::size() const { return g_rc[this].size; }
const polyPatch& regionCoupleFvPatch::patch() const { return *g_rc[this].poly; }
bool regionCoupleFvPatch::coupled() const { return g_rc[this].coupled; }
bool regionCoupleFvPatch::master() const { return g_rc[this].master; }
static bool g_zoneMode = false; // the pair is spread over the two processors: interpolation on the global zones
bool regionCoupleFvPatch::localParallel() const { return !g_zoneMode; }
label regionCoupleFvPatch::shadowIndex() const { return g_rc[this].shadowIndex; }
const fvMesh& regionCoupleFvPatch::shadowRegion() const { return *g_rc[this].shadowMesh; }
const regionCoupleFvPatch& regionCoupleFvPatch::shadow() const { return *g_rc[this].shadow; }
const lduAddressing& fvMesh::lduAddr() const { return *g_meshAddr[this]; }
const word& polyPatch::name() const { return g_rcType; }
bool regionCouplePolyPatch::master() const { return g_rc[g_polyOwner[this]].master; }
bool regionCouplePolyPatch::attached() const { return true; }
const regionCouplePolyPatch& regionCouplePolyPatch::shadow() const { return *g_rc[g_rc[g_polyOwner[this]].shadow].poly; }
const ggiZoneInterpolation& regionCouplePolyPatch::patchToPatch() const { return g_interp; }
// zone members: reached only for pairs spread over processors (harness mode "zone")
static std::map<const regionCouplePolyPatch*, faceZone> g_zone;
static std::map<const faceZone*, label> g_zoneSize;
static std::map<const regionCouplePolyPatch*, labelList> g_zoneAddr;
label faceZone::size() const { return g_zoneSize[this]; }
const faceZone& regionCouplePolyPatch::zone() const { return g_zone[this]; }
const labelList& regionCouplePolyPatch::zoneAddressing() const { return g_zoneAddr[this]; }
const labelListList& ggiZoneInterpolation::masterAddr() const { return g_mAddr; }
const scalarListList& ggiZoneInterpolation::masterWeights() const { return g_mW; }
const labelListList& ggiZoneInterpolation::slaveAddr() const { return g_sAddr; }
const scalarListList& ggiZoneInterpolation::slaveWeights() const { return g_sW; }

class mockRcPatch : public regionCoupleFvPatch
{
public:
    const word& type() const { return g_rcType; }
};
class mockRcField : public lduInterfaceField
{
public:
    const word& type() const { return g_rcType; }
};
static word g_procType("processor");
class mockProcField : public lduInterfaceField, public processorLduInterfaceField
{
public:
    int nbr;
    const word& type() const { return g_procType; }
    int myProcNo() const { return g_rank; }
    int neighbProcNo() const { return nbr; }
};
class mockProcPatch : public lduInterface
{
public:
    const word& type() const { return g_procType; }
};

// members of the stub that the binding does not reach in this harness
dictionary const& dictionary::subDict(const word&) const { return *this; }
bool dictionary::found(const word&) const { return false; }
bool dictionary::isDict(const word&) const { return false; }
ITstream& dictionary::lookup(const word&) const { static ITstream s; return s; }
bool dictionary::remove(const word&) { return false; }
}

extern "C"
{ // the C ABI is not linked: describe() and dump() never call it
int b200_nccl_unique_id(void*) { return -1; }
int b200_ctx_create(int, int, int, const void*, b200_ctx**) { return -1; }
const char* b200_last_error(const b200_ctx*) { return "harness"; }
int b200_sys_create(b200_ctx*, int, b200_sys**) { return -1; }
int b200_sys_set_region(b200_sys*, int, int32_t, int32_t, const int32_t*, const int32_t*) { return -1; }
int b200_sys_add_interface(b200_sys*, int, int, int32_t, const int32_t*, int, int, int, int32_t, const int32_t*, const int32_t*, const double*) { return -1; }
int b200_sys_finalize(b200_sys*) { return -1; }
int b200_sys_set_coeffs(b200_sys*, int, const double*, const double*, const double*) { return -1; }
int b200_sys_set_interface_coeffs(b200_sys*, int, int, const double*, const double*) { return -1; }
int b200_sys_set_interface_attached(b200_sys*, int, int, int) { return -1; }
int b200_sys_set_interface_pieces(b200_sys*, int, int, int, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*) { return -1; }
int b200_sys_set_interface_ggi(b200_sys*, int, int, int32_t, const int32_t*, const int32_t*, const double*) { return -1; }
}

using namespace Foam;

static labelList L(std::initializer_list<label> v)
{
    labelList l(label(v.size()));
    label i = 0;
    for (label x : v) l[i++] = x;
    return l;
}
static scalarField F(std::initializer_list<scalar> v)
{
    scalarField f(label(v.size()));
    label i = 0;
    for (scalar x : v) f[i++] = x;
    return f;
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    g_rank = atoi(argv[2]);
    g_nprocs = 2;
    // fluid: 6 cells in a row (faces 0-1 .. 4-5), patches: 0 wall (uncoupled, no interface), 1 regionCouple (3 faces), 2 processor (1 face)
    // solid: 4 cells in a row, patches: 0 regionCouple (2 faces), 1 processor (2 faces)
    static mockAddressing aF, aS;
    aF.d.nCells = 6;
    aF.d.lower = L({0, 1, 2, 3, 4});
    aF.d.upper = L({1, 2, 3, 4, 5});
    aF.d.patches = {L({0}), L({1, 2, 3}), L({5})};
    aS.d.nCells = 4;
    aS.d.lower = L({0, 1, 2});
    aS.d.upper = L({1, 2, 3});
    aS.d.patches = {L({0, 1}), L({2, 3})};
    static lduMatrix mF, mS;
    g_mat[&mF] = MatData{&aF, F({10, 11, 12, 13, 14, 15}), F({-1, -1.5, -2, -2.5, -3}), F({-0.5, -0.75, -1, -1.25, -1.5}), true};
    g_mat[&mS] = MatData{&aS, F({20, 21, 22, 23}), F({-4, -5, -6}), scalarField(), false};
    static fvMesh meshF, meshS;
    g_meshAddr[&meshF] = &aF;
    g_meshAddr[&meshS] = &aS;
    static mockRcPatch rcF, rcS;
    static regionCouplePolyPatch ppF, ppS;
    g_polyOwner[&ppF] = &rcF;
    g_polyOwner[&ppS] = &rcS;
    g_rc[&rcF] = RcData{true, true, 3, 0, &rcS, &meshS, &ppF, word("interface")};   // master, shadow = solid patch 0
    g_rc[&rcS] = RcData{false, true, 2, 1, &rcF, &meshF, &ppS, word("top")};        // slave, shadow = fluid patch 1
    // GGI: master face i sees slave faces; slave face j sees master faces (rows rescaled to one)
    g_mAddr.setSize(3);
    g_mW.setSize(3);
    g_mAddr[0] = L({0});        g_mW[0] = F({1.0});
    g_mAddr[1] = L({0, 1});     g_mW[1] = F({0.5, 0.5});
    g_mAddr[2] = L({1});        g_mW[2] = F({1.0});
    g_sAddr.setSize(2);
    g_sW.setSize(2);
    g_sAddr[0] = L({0, 1});     g_sW[0] = F({2.0 / 3.0, 1.0 / 3.0});
    g_sAddr[1] = L({1, 2});     g_sW[1] = F({1.0 / 3.0, 2.0 / 3.0});
    static mockRcField fF, fS;
    g_ifaceOf[&fF] = &rcF;
    g_ifaceOf[&fS] = &rcS;
    static mockProcField pF, pS;
    static mockProcPatch ppatchF, ppatchS;
    pF.nbr = pS.nbr = 1 - g_rank;
    g_ifaceOf[&pF] = &ppatchF;
    g_ifaceOf[&pS] = &ppatchS;
    // the other rank's table: the same patch layout, so its interface indices are fluid: 1, solid: 1
    g_otherTables[1 - g_rank] = L({0, g_rank, 1, 1, g_rank, 1});

    if (argc > 3 && std::string(argv[3]) == "zone")
    {
        // The pair on its global zones: master zone 5 faces, slave zone 4; this rank (1) holds master zone faces 2, 3, 4 and
        // slave zone faces 2, 3, rank 0 the others.  patchToPatch() is the ZONE-level interpolator.
        g_zoneMode = true;
        g_zoneSize[&g_zone[&ppF]] = 5;
        g_zoneSize[&g_zone[&ppS]] = 4;
        g_zoneAddr[&ppF] = L({2, 3, 4});
        g_zoneAddr[&ppS] = L({2, 3});
        g_mAddr.setSize(5);
        g_mW.setSize(5);
        g_mAddr[0] = L({0});        g_mW[0] = F({1.0});
        g_mAddr[1] = L({0, 1});     g_mW[1] = F({0.5, 0.5});
        g_mAddr[2] = L({1, 2});     g_mW[2] = F({0.25, 0.75});
        g_mAddr[3] = L({2, 3});     g_mW[3] = F({0.5, 0.5});
        g_mAddr[4] = L({3});        g_mW[4] = F({1.0});
        g_sAddr.setSize(4);
        g_sW.setSize(4);
        g_sAddr[0] = L({0, 1});     g_sW[0] = F({0.5, 0.5});
        g_sAddr[1] = L({1, 2});     g_sW[1] = F({0.5, 0.5});
        g_sAddr[2] = L({2, 3});     g_sW[2] = F({0.375, 0.625});
        g_sAddr[3] = L({3, 4});     g_sW[3] = F({0.5, 0.5});
        // rank 0's zone table: (row, patch, interface index, nFaces, zone faces...) per zone-mode patch
        g_otherZoneTables[1 - g_rank] = L({0, 1, 0, 2, 0, 1, 1, 0, 0, 2, 0, 1});
    }

    UPtrList<const lduMatrix> matrices(2);
    matrices.set(0, &mF);
    matrices.set(1, &mS);
    List<lduInterfaceFieldPtrsList> ifaces(2);
    ifaces[0] = lduInterfaceFieldPtrsList(3);
    ifaces[0].set(1, &fF);
    ifaces[0].set(2, &pF);
    ifaces[1] = lduInterfaceFieldPtrsList(2);
    ifaces[1].set(0, &fS);
    ifaces[1].set(1, &pS);
    static FieldField<Field, scalar> bouF(3), intF(3), bouS(2), intS(2);
    static scalarField e0, bF1 = F({0.1, 0.2, 0.3}), iF1 = F({1.1, 1.2, 1.3}), bF2 = F({0.7}), iF2 = F({1.7});
    static scalarField bS0 = F({0.4, 0.5}), iS0 = F({1.4, 1.5}), bS1 = F({0.8, 0.9}), iS1 = F({1.8, 1.9});
    bouF.set(0, &e0); intF.set(0, &e0);
    bouF.set(1, &bF1); intF.set(1, &iF1);
    bouF.set(2, &bF2); intF.set(2, &iF2);
    bouS.set(0, &bS0); intS.set(0, &iS0);
    bouS.set(1, &bS1); intS.set(1, &iS1);
    List<const FieldField<Field, scalar>*> bou(2), inte(2);
    bou[0] = &bouF; bou[1] = &bouS;
    inte[0] = &intF; inte[1] = &intS;
    static scalarField xF = F({300, 301, 302, 303, 304, 305}), xS = F({310, 311, 312, 313});
    static scalarField sF = F({1, 2, 3, 4, 5, 6}), sS = F({7, 8, 9, 10});
    UPtrList<const scalarField> xs(2), bs(2);
    xs.set(0, &xF); xs.set(1, &xS);
    bs.set(0, &sF); bs.set(1, &sS);
    if (g_zoneMode)
    { // the dump format does not carry zone pieces: print what describe() found instead
        List<List<b200Binding::ifaceInfo> > info;
        b200Binding::describe(matrices, ifaces, info);
        forAll (info, r) forAll (info[r], i)
        {
            const b200Binding::ifaceInfo& I = info[r][i];
            std::cout << "iface " << r << " " << i << " kind " << I.kind << " zoneMode " << int(I.zoneMode) << " nPeerFaces " << I.nPeerFaces
                      << " peer " << I.peerRegion << " " << I.peerIface << " offsets";
            forAll (I.ggiOffsets, k) std::cout << " " << I.ggiOffsets[k];
            std::cout << " addr";
            forAll (I.ggiAddr, k) std::cout << " " << I.ggiAddr[k];
            std::cout << " w";
            forAll (I.ggiWeights, k) std::cout << " " << I.ggiWeights[k];
            std::cout << " pieces";
            forAll (I.pieceRank, k)
            {
                std::cout << " [" << I.pieceRank[k] << " " << I.pieceIface[k] << ":";
                for (label q = I.pieceOffsets[k]; q < I.pieceOffsets[k + 1]; q++) std::cout << " " << I.pieceZoneAddr[q];
                std::cout << "]";
            }
            std::cout << "\n";
        }
        return 0;
    }
    b200Binding::dump(fileName(std::string(argv[1])), matrices, ifaces, bou, inte, xs, bs, word("BiCGStab"), word("Cholesky"), 1e-15, 0.0, 0, 200,
                      F({0.5, 0.25, 0.125}));
    return 0;
}
