// b200_ldu.cu -- C ABI (include/b200_ldu.h) and host orchestration of the B200 LDU solver.
// Streams and launches only; all arithmetic is in kernels.cuh, all table construction in
// schedule.hpp.  There is no CPU fallback: every entry point needs a CUDA device.
#include "../../include/b200_ldu.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <thread>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <algorithm>
#include <memory>
#include <string>
#include <map>
#include <tuple>
#include <vector>

#include "kernels.cuh"
#include "schedule.hpp"

using namespace b200;

#define B200_VERSION 100

// ------------------------------------------------------------------------------------------ NCCL (dlopen)
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::string g_lastError;

static bool load_nccl(std::string& err)
{
    if (g_nccl.handle) return true;
    const char* cands[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so",
                           "/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl/lib/libnccl.so.2"};
    void* h = nullptr;
    for (const char* c : cands)
    {
        if (!c || !*c) continue;
        h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h)
    {
        err = std::string("cannot dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
        return false;
    }
#define LOADSYM(field, name)                                        \
    *(void**)(&g_nccl.field) = dlsym(h, name);                      \
    if (!g_nccl.field)                                              \
    {                                                               \
        err = std::string("libnccl lacks symbol ") + name;          \
        return false;                                               \
    }
    LOADSYM(GetUniqueId, "ncclGetUniqueId");
    LOADSYM(CommInitRank, "ncclCommInitRank");
    LOADSYM(CommDestroy, "ncclCommDestroy");
    LOADSYM(AllReduce, "ncclAllReduce");
    LOADSYM(Send, "ncclSend");
    LOADSYM(Recv, "ncclRecv");
    LOADSYM(GroupStart, "ncclGroupStart");
    LOADSYM(GroupEnd, "ncclGroupEnd");
    LOADSYM(GetErrorString, "ncclGetErrorString");
#undef LOADSYM
    g_nccl.handle = h;
    return true;
}

// ------------------------------------------------------------------------------------------ context
// Peer-to-peer transport of a context (nranks > 1): mailboxes for the fused all-reduce, an exchange area for the
// zone all-reduce, and the rendezvous directory the ranks use to hand each other cudaIpc handles.
struct PeerLink
{
    bool enabled = false;
    std::string dir;                   // /dev/shm/b200_<first bytes of the unique id>
    double* mbox = nullptr;            // [2][nranks][kMboxSlot]
    std::vector<double*> peerMbox;     // per rank: mapped mailbox (own: mbox)
    double** dPeerMbox = nullptr;      // the same on the device
    unsigned long long arSeq = 0;
    double* xbuf = nullptr;            // [2][nranks][kXchgCap + 8] exchange area (zone all-reduce without NCCL)
    std::vector<double*> peerX;
    double** dPeerX = nullptr;
    unsigned long long xSeq = 0;
    unsigned* counter = nullptr;       // last-block counter of the exchange kernels
    int* err = nullptr;                // device word: 2 = a peer did not answer
    int sysCount = 0;                  // systems finalized on this context (names the rendezvous files)
    std::vector<void*> opened;         // cudaIpcOpenMemHandle results to close
    std::vector<std::string> files;    // rendezvous files this rank wrote
};
constexpr int kXchgCap = 65536; // doubles per rank and parity in the exchange area

struct b200_ctx
{
    int device = 0, rank = 0, nranks = 1;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    PeerLink peer;
    std::string err;
    int64_t launches = 0;
    int smCount = 148;
    // result of the last b200_ggi_build (ggi_build.cuh), handed out by b200_ggi_fetch
    std::vector<int32_t> ggiOff, ggiAddr;
    std::vector<double> ggiW;
};

static int set_err(b200_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_lastError = buf;
    return code;
}

#define CK(ctx, call)                                                                                        \
    do                                                                                                       \
    {                                                                                                        \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return set_err(ctx, B200_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define NK(ctx, call)                                                                                           \
    do                                                                                                          \
    {                                                                                                           \
        ncclResult_t r_ = (call);                                                                               \
        if (r_ != ncclSuccess)                                                                                  \
            return set_err(ctx, B200_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
    } while (0)

template <class T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) count = 1;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    cudaError_t upload(const std::vector<T>& h, cudaStream_t st)
    {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess) return e;
        if (h.empty()) return cudaSuccess;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
};

// ------------------------------------------------------------------------------------------ peer-to-peer rendezvous
// The C ABI hands every rank the same 128-byte unique id (b200_nccl_unique_id broadcast by rank 0; any 128 random bytes
// do when NCCL is not used).  The ranks of one node meet in /dev/shm/b200_<id>: a rank publishes a record by writing
// <name>.tmp and renaming it, and reads a peer's record by polling for the file.
static std::string peer_dir(const void* uid)
{
    static const char* hex = "0123456789abcdef";
    const unsigned char* b = static_cast<const unsigned char*>(uid);
    std::string d = "/dev/shm/b200_";
    for (int i = 0; i < 16; i++)
    {
        d.push_back(hex[b[i] >> 4]);
        d.push_back(hex[b[i] & 15]);
    }
    return d;
}
static bool peer_publish(PeerLink& L, const std::string& name, const void* data, size_t n)
{
    const std::string path = L.dir + "/" + name, tmp = path + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(data, 1, n, f) == n;
    fclose(f);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) return false;
    L.files.push_back(path);
    return true;
}
static bool peer_fetch(const PeerLink& L, const std::string& name, std::vector<unsigned char>& data, double timeoutSec = 120.0)
{
    const std::string path = L.dir + "/" + name;
    const auto t0 = std::chrono::steady_clock::now();
    for (;;)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (f)
        {
            fseek(f, 0, SEEK_END);
            const long n = ftell(f);
            fseek(f, 0, SEEK_SET);
            data.resize((size_t)std::max(0l, n));
            const bool ok = n >= 0 && fread(data.data(), 1, data.size(), f) == data.size();
            fclose(f);
            return ok;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeoutSec) return false;
        std::this_thread::sleep_for(std::chrono::microseconds(500));
    }
}
// "p2p": the library's own transport only (no NCCL communicator: ranks may share a device); "nccl": NCCL only;
// unset / "auto": peer-to-peer when the rendezvous and the mappings succeed, NCCL otherwise
static int transport_mode()
{
    const char* e = getenv("B200_TRANSPORT");
    if (e && !strcmp(e, "p2p")) return 1;
    if (e && !strcmp(e, "nccl")) return 2;
    return 0;
}

// Peer-to-peer halo of one system (PeerLink of the context): one allocation [2 parities][recvTotal] doubles | [2][nPeers] flag
// words, mapped by the neighbours, which store their patch values straight into it (k_halo_push / k_blk_halo_push)
struct HaloLink
{
    bool enabled = false;
    unsigned char* buf = nullptr;
    size_t recvTotal = 0;
    std::vector<double*> peerData;             // per peer: its buffer, at the offset of this rank's segment (parity 0)
    std::vector<size_t> peerTotal;             // per peer: its recvTotal (parity stride)
    std::vector<unsigned long long*> peerFlag; // per peer: this rank's flag word in its buffer (parity 0)
    std::vector<int> peerNPeers;               // per peer: its number of segments (parity stride of the flags)
    unsigned long long seq = 0;
    unsigned* counters = nullptr;              // [nPeers] last-block counters of the push kernels
    std::vector<void*> opened;
    double* data(int par) const { return reinterpret_cast<double*>(buf) + (size_t)par * recvTotal; }
    unsigned long long* flags(int par, int nPeers) const
    {
        return reinterpret_cast<unsigned long long*>(buf + 2 * recvTotal * sizeof(double)) + (size_t)par * nPeers;
    }
    void release()
    {
        for (void* q : opened) cudaIpcCloseMemHandle(q);
        opened.clear();
        if (buf) cudaFree(buf);
        if (counters) cudaFree(counters);
        buf = nullptr;
        counters = nullptr;
        enabled = false;
    }
};

struct PipeDirMem
{
    DevBuf<int> gW, gCH, gShflMask, face, order, gLg, gRg, gKg, pFace, cFace;
    DevBuf<long long> gTermOff, gPOff, gCOff, gPFaceOff, gCFaceOff, stats;
    DevBuf<unsigned char> stream, gFast, pStream, cStream;
    PipeDev dev{};
    int smemBytes = 0;
    // second set of stream buffers holding the TRANSPOSED coefficients (preconditionT of PBiCG), created on first use
    DevBuf<unsigned char> streamT, pStreamT, cStreamT;
    PipeDev devT{};
    bool haveT = false;
    bool packed[2] = {false, false}; // [0]: dev holds the plain coefficients of the current matrix, [1]: devT the transposed ones
};

// static FV geometry of one region for the device-side coefficient refresh (fv_assemble.cuh)
struct FvRegionDev
{
    DevBuf<double> V, magSf, delta, phi, kappaFace, bInt, bSrc;
    DevBuf<int> ownerStart, losort, losortStart, bStart;
    bool havePhi = false, haveKappaFace = false;
};

// ------------------------------------------------------------------------------------------ system
enum VecId
{
    V_X = 0,
    V_B,
    V_R,
    V_RW,
    V_P,
    V_PH,
    V_V,
    V_S,
    V_SH,
    V_T,
    V_TMP,
    V_TMP2,
    V_COUNT
};

struct b200_sys
{
    b200_ctx* ctx = nullptr;
    std::vector<RegionHost> regs;
    bool finalized = false;
    int64_t N = 0, F = 0;
    double nGlobalCells = 0;

    // matrix
    DevBuf<double> diag, coef /* [upper(F) | lower(F)] */, ifCoefBou, ifCoefInt;
    std::vector<uint8_t> regionHasCoeffs;
    bool sellDirty = true, sellTDirty = true, diagDirty = false;
    DevBuf<double> diagCell; // cell-ordered diag as set by the caller (regions arrive one by one)
    // slot order (sweep schedule order) of all device vectors
    int64_t nSlots = 0;
    DevBuf<int> slotOfCell, cellOfSlot, gBase, gNT;
    DevBuf<double> stageCell; // cell-ordered staging for host <-> device copies
    int nGroups = 0;
    // SELL
    int64_t nSlices = 0, nEntries = 0;
    DevBuf<int> sliceOff, sellCol, sellSrc;
    DevBuf<double> sellVal, sellValT;
    DevBuf<unsigned> ifaceMask;
    // interfaces
    int nTouched = 0;
    DevBuf<int> ifRows, ifRowStart, ifEntCoef, ifEntSrc, ifEntCnt, ifGSrc, sendCells;
    DevBuf<double> ifGW, sendBuf, recvBuf;
    std::vector<int32_t> slotOfCellHost; // kept for rebuilding the interface plan (b200_sys_set_interface_ggi)
    bool ifacePlanDirty = false;
    int nDetached = 0; // regionCouple interfaces currently detached (regionInterfaceType::detach)
    std::vector<int> peers;
    std::vector<int32_t> sendOff, recvOff;
    HaloLink halo; // peer-to-peer halo of the processor patches
    int64_t nIfCoefs = 0;
    // sweeps
    PipeDirMem fwd, bwd;
    DevBuf<double> rD, rDraw;
    int precondValid = -1; // preconditioner id rD belongs to (-1: none)
    DevBuf<unsigned> ticket;
    unsigned ticketBase = 0;
    DevBuf<int> devErr;
    // vectors
    DevBuf<double> vec[V_COUNT];
    DevBuf<double> x0; // b200_x_save / b200_x_restore
    bool x0Valid = false;
    // scalars / reductions
    DevBuf<DevScalars> sc;
    DevBuf<double> partials;
    int pstride = 0;
    DevBuf<double> history;
    int* hostFlags = nullptr;  // mapped pinned: [0] done, [1] nIter
    int* devHostFlags = nullptr;
    // launch geometry
    int vecBlocks = 1, amulBlocks = 1, ifaceBlocks = 0;
    // profiling
    bool profiling = false;
    struct EvRec
    {
        int cls;
        cudaEvent_t a, b;
    };
    std::vector<EvRec> evRecs;
    std::vector<cudaEvent_t> evPool;
    double clsMs[B200_K_NCLASSES] = {0};
    int64_t clsLaunches[B200_K_NCLASSES] = {0};
    cudaEvent_t evSolveA = nullptr, evSolveB = nullptr;
    std::vector<cudaEvent_t> throttle;
    // device-side coefficient refresh (fv_assemble.cuh): static FV geometry per region, set on demand
    std::vector<std::unique_ptr<FvRegionDev>> fv;
};

struct KScope
{
    b200_sys* s;
    int cls;
    cudaEvent_t a = nullptr, b = nullptr;
    KScope(b200_sys* s_, int cls_) : s(s_), cls(cls_)
    {
        s->ctx->launches++;
        s->clsLaunches[cls]++;
        if (s->profiling)
        {
            auto get = [&]() {
                cudaEvent_t e;
                if (!s->evPool.empty())
                {
                    e = s->evPool.back();
                    s->evPool.pop_back();
                }
                else
                    cudaEventCreate(&e);
                return e;
            };
            a = get();
            b = get();
            cudaEventRecord(a, s->ctx->stream);
        }
    }
    ~KScope()
    {
        if (s->profiling)
        {
            cudaEventRecord(b, s->ctx->stream);
            s->evRecs.push_back({cls, a, b});
        }
    }
};

static void harvest_events(b200_sys* s)
{
    for (auto& r : s->evRecs)
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) s->clsMs[r.cls] += ms;
        s->evPool.push_back(r.a);
        s->evPool.push_back(r.b);
    }
    s->evRecs.clear();
}

// ------------------------------------------------------------------------------------------ C ABI: context
extern "C" int b200_version(void) { return B200_VERSION; }

extern "C" const char* b200_last_error(const b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_lastError.c_str(); }

extern "C" int b200_nccl_unique_id(void* out128)
{
    std::string err;
    if (!out128) return set_err(nullptr, B200_EINVAL, "b200_nccl_unique_id: null output");
    if (!load_nccl(err)) return set_err(nullptr, B200_ENCCL, "%s", err.c_str());
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return set_err(nullptr, B200_ENCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return B200_OK;
}

// Map one buffer of every rank into this process: publish this rank's cudaIpc handle as <tag>, fetch the peers'.
static bool peer_map_all(b200_ctx* c, const std::string& tag, void* mine, std::vector<double*>& mapped, std::string& why)
{
    PeerLink& L = c->peer;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, mine) != cudaSuccess)
    {
        why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(cudaGetLastError());
        return false;
    }
    if (!peer_publish(L, "r" + std::to_string(c->rank) + "." + tag, &h, sizeof(h)))
    {
        why = "cannot write the rendezvous record in " + L.dir;
        return false;
    }
    mapped.assign(c->nranks, nullptr);
    mapped[c->rank] = static_cast<double*>(mine);
    for (int r = 0; r < c->nranks; r++)
    {
        if (r == c->rank) continue;
        std::vector<unsigned char> rec;
        if (!peer_fetch(L, "r" + std::to_string(r) + "." + tag, rec) || rec.size() != sizeof(h))
        {
            why = "rank " + std::to_string(r) + " did not publish " + tag;
            return false;
        }
        memcpy(&h, rec.data(), sizeof(h));
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
        {
            why = std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(cudaGetLastError());
            return false;
        }
        L.opened.push_back(q);
        mapped[r] = static_cast<double*>(q);
    }
    return true;
}

static bool peer_setup(b200_ctx* c, const void* uid, std::string& why)
{
    PeerLink& L = c->peer;
    L.dir = peer_dir(uid);
    mkdir(L.dir.c_str(), 0700); // EEXIST: another rank was first
    const size_t mboxDoubles = 2 * (size_t)c->nranks * kMboxSlot;
    const size_t xDoubles = 2 * (size_t)c->nranks * (kXchgCap + 8);
    if (cudaMalloc((void**)&L.mbox, mboxDoubles * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&L.xbuf, xDoubles * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&L.counter, 64) != cudaSuccess || cudaMalloc((void**)&L.err, 64) != cudaSuccess ||
        cudaMalloc((void**)&L.dPeerMbox, sizeof(double*) * c->nranks) != cudaSuccess ||
        cudaMalloc((void**)&L.dPeerX, sizeof(double*) * c->nranks) != cudaSuccess)
    {
        why = std::string("cudaMalloc: ") + cudaGetErrorString(cudaGetLastError());
        return false;
    }
    cudaMemset(L.mbox, 0, mboxDoubles * sizeof(double));
    cudaMemset(L.xbuf, 0, xDoubles * sizeof(double));
    cudaMemset(L.counter, 0, 64);
    cudaMemset(L.err, 0, 64);
    cudaDeviceSynchronize(); // the zeros are in place before any peer can write
    if (!peer_map_all(c, "mbox", L.mbox, L.peerMbox, why)) return false;
    if (!peer_map_all(c, "xbuf", L.xbuf, L.peerX, why)) return false;
    cudaMemcpy(L.dPeerMbox, L.peerMbox.data(), sizeof(double*) * c->nranks, cudaMemcpyHostToDevice);
    cudaMemcpy(L.dPeerX, L.peerX.data(), sizeof(double*) * c->nranks, cudaMemcpyHostToDevice);
    L.enabled = true;
    return true;
}

static void peer_teardown(b200_ctx* c)
{
    PeerLink& L = c->peer;
    for (void* q : L.opened) cudaIpcCloseMemHandle(q);
    L.opened.clear();
    for (const std::string& f : L.files) unlink(f.c_str());
    L.files.clear();
    if (!L.dir.empty()) rmdir(L.dir.c_str()); // succeeds for the last rank to leave
    cudaFree(L.mbox);
    cudaFree(L.xbuf);
    cudaFree(L.counter);
    cudaFree(L.err);
    cudaFree(L.dPeerMbox);
    cudaFree(L.dPeerX);
    L = PeerLink();
}

extern "C" int b200_ctx_create(int device, int rank, int nranks, const void* ncclUniqueIdBytes, b200_ctx** out)
{
    if (!out || nranks < 1 || rank < 0 || rank >= nranks) return set_err(nullptr, B200_EINVAL, "b200_ctx_create: bad arguments");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(nullptr, B200_ECUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return set_err(nullptr, B200_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    std::unique_ptr<b200_ctx> c(new b200_ctx);
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    CK(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(nullptr, cudaGetDeviceProperties(&prop, device));
    c->smCount = prop.multiProcessorCount;
    CK(nullptr, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (nranks > 1)
    {
        std::string err;
        if (!ncclUniqueIdBytes) return set_err(nullptr, B200_EINVAL, "nranks > 1 needs the 128-byte unique id all ranks share");
        const int mode = transport_mode();
        if (mode != 1)
        { // NCCL communicator: the transport in "nccl" mode, the fall-back and the zone all-reduce otherwise
            if (!load_nccl(err)) return set_err(nullptr, B200_ENCCL, "%s", err.c_str());
            ncclUniqueId id;
            memcpy(&id, ncclUniqueIdBytes, 128);
            ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
            if (r != ncclSuccess) return set_err(nullptr, B200_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
        }
        if (mode != 2)
        {
            std::string why;
            if (!peer_setup(c.get(), ncclUniqueIdBytes, why))
            {
                peer_teardown(c.get());
                if (mode == 1) return set_err(nullptr, B200_ECUDA, "B200_TRANSPORT=p2p: %s", why.c_str());
                if (getenv("B200_TRANSPORT_INFO")) fprintf(stderr, "[b200] rank %d: peer-to-peer transport unavailable (%s), using NCCL\n", rank, why.c_str());
            }
            else if (getenv("B200_TRANSPORT_INFO"))
                fprintf(stderr, "[b200] rank %d: peer-to-peer transport up (%s)\n", rank, c->peer.dir.c_str());
        }
    }
    *out = c.release();
    return B200_OK;
}

extern "C" int b200_ctx_destroy(b200_ctx* ctx)
{
    if (!ctx) return B200_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    peer_teardown(ctx);
    if (ctx->comm) g_nccl.CommDestroy(ctx->comm);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B200_OK;
}

extern "C" int64_t b200_launch_count(const b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------ C ABI: system build
extern "C" int b200_sys_create(b200_ctx* ctx, int nRegions, b200_sys** out)
{
    if (!ctx || !out || nRegions < 1) return set_err(ctx, B200_EINVAL, "b200_sys_create: bad arguments");
    b200_sys* s = new b200_sys;
    s->ctx = ctx;
    s->regs.resize(nRegions);
    s->regionHasCoeffs.assign(nRegions, 0);
    *out = s;
    return B200_OK;
}

extern "C" int b200_sys_destroy(b200_sys* s)
{
    if (!s) return B200_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    harvest_events(s);
    for (auto e : s->evPool) cudaEventDestroy(e);
    for (auto e : s->throttle) cudaEventDestroy(e);
    if (s->evSolveA) cudaEventDestroy(s->evSolveA);
    if (s->evSolveB) cudaEventDestroy(s->evSolveB);
    if (s->hostFlags) cudaFreeHost(s->hostFlags);
    s->halo.release();
    delete s;
    return B200_OK;
}

extern "C" int b200_sys_set_region(b200_sys* s, int r, int32_t nCells, int32_t nFaces, const int32_t* lowerAddr,
                                   const int32_t* upperAddr)
{
    if (!s) return B200_EINVAL;
    if (s->finalized) return set_err(s->ctx, B200_ESTATE, "system already finalized");
    if (r < 0 || r >= (int)s->regs.size() || nCells < 0 || nFaces < 0 || (nFaces > 0 && (!lowerAddr || !upperAddr)))
        return set_err(s->ctx, B200_EINVAL, "b200_sys_set_region: bad arguments");
    RegionHost& R = s->regs[r];
    R.nCells = nCells;
    R.nFaces = nFaces;
    R.l.assign(lowerAddr, lowerAddr + nFaces);
    R.u.assign(upperAddr, upperAddr + nFaces);
    for (int32_t f = 0; f < nFaces; f++)
        if (R.l[f] < 0 || R.u[f] >= nCells || R.l[f] >= R.u[f])
            return set_err(s->ctx, B200_EINVAL, "region %d face %d: need 0 <= lower < upper < nCells", r, f);
    R.ifaces.clear();
    R.set = true;
    return B200_OK;
}

extern "C" int b200_sys_add_interface(b200_sys* s, int r, int kind, int32_t nFaces, const int32_t* faceCells, int peerRank,
                                      int peerRegion, int peerIface, int32_t nPeerFaces, const int32_t* ggiOffsets,
                                      const int32_t* ggiAddr, const double* ggiWeights)
{
    if (!s) return B200_EINVAL;
    if (s->finalized) return set_err(s->ctx, B200_ESTATE, "system already finalized");
    if (r < 0 || r >= (int)s->regs.size() || !s->regs[r].set) return set_err(s->ctx, B200_EINVAL, "add_interface: region %d not set", r);
    if (kind != B200_IFACE_REGION_COUPLE && kind != B200_IFACE_PROCESSOR)
        return set_err(s->ctx, B200_EUNSUPPORTED, "unknown interface kind %d (no CPU fallback for foreign lduInterfaceFields)", kind);
    if (nFaces < 0 || (nFaces > 0 && !faceCells) || peerRank < 0 || peerRank >= s->ctx->nranks || nPeerFaces < 0)
        return set_err(s->ctx, B200_EINVAL, "add_interface: bad arguments");
    RegionHost& R = s->regs[r];
    IfaceHost I;
    I.kind = kind;
    I.nFaces = nFaces;
    I.faceCells.assign(faceCells, faceCells + nFaces);
    for (int i = 0; i < nFaces; i++)
        if (faceCells[i] < 0 || faceCells[i] >= R.nCells) return set_err(s->ctx, B200_EINVAL, "add_interface: faceCells[%d] out of range", i);
    I.peerRank = peerRank;
    I.peerRegion = peerRegion;
    I.peerIface = peerIface;
    I.nPeerFaces = nPeerFaces;
    I.identity = (ggiOffsets == nullptr);
    if (I.identity)
    {
        if (nPeerFaces != nFaces) return set_err(s->ctx, B200_EINVAL, "identity interface needs nPeerFaces == nFaces");
    }
    else
    {
        if (!ggiAddr || !ggiWeights) return set_err(s->ctx, B200_EINVAL, "GGI interface needs addr and weights");
        I.ggiOffsets.assign(ggiOffsets, ggiOffsets + nFaces + 1);
        int nnz = ggiOffsets[nFaces];
        I.ggiAddr.assign(ggiAddr, ggiAddr + nnz);
        I.ggiWeights.assign(ggiWeights, ggiWeights + nnz);
    }
    R.ifaces.push_back(std::move(I));
    return (int)R.ifaces.size() - 1;
}

// regionCouple patch whose shadow patch decomposePar spread over several ranks (include/b200_ldu.h)
extern "C" int b200_sys_set_interface_pieces(b200_sys* s, int r, int iface, int nPieces, const int32_t* pieceRank, const int32_t* pieceRegion,
                                             const int32_t* pieceIface, const int32_t* pieceOffsets, const int32_t* pieceZoneAddr)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (s->finalized) return set_err(ctx, B200_ESTATE, "b200_sys_set_interface_pieces: system already finalized");
    if (r < 0 || r >= (int)s->regs.size() || iface < 0 || iface >= (int)s->regs[r].ifaces.size())
        return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_pieces: no interface %d of region %d", iface, r);
    IfaceHost& I = s->regs[r].ifaces[iface];
    if (I.kind != B200_IFACE_REGION_COUPLE || I.identity)
        return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_pieces: a regionCouple interface with GGI tables (zone face labels) is needed");
    if (nPieces < 0 || (nPieces && (!pieceRank || !pieceRegion || !pieceIface || !pieceOffsets)))
        return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_pieces: bad arguments");
    std::vector<IfaceHost::Piece> pieces((size_t)nPieces);
    std::vector<char> seen((size_t)I.nPeerFaces, 0);
    for (int k = 0; k < nPieces; k++)
    {
        IfaceHost::Piece& P = pieces[k];
        P.rank = pieceRank[k];
        P.region = pieceRegion[k];
        P.iface = pieceIface[k];
        const int32_t a = pieceOffsets[k], b = pieceOffsets[k + 1];
        if (P.rank < 0 || P.rank >= ctx->nranks || P.region < 0 || P.iface < 0 || a < 0 || b < a || (b > a && !pieceZoneAddr))
            return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_pieces: piece %d: bad rank / region / interface / offsets", k);
        P.zoneAddr.assign(pieceZoneAddr + a, pieceZoneAddr + b);
        for (int32_t z : P.zoneAddr)
        {
            if (z < 0 || z >= I.nPeerFaces) return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_pieces: piece %d: zone face %d outside the shadow zone of %d faces", k, z, I.nPeerFaces);
            if (seen[z]) return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_pieces: zone face %d is held by two pieces", z);
            seen[z] = 1;
        }
    }
    for (int32_t z : I.ggiAddr)
        if (nPieces && !seen[z]) return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_pieces: the GGI tables address zone face %d, which no piece holds", z);
    I.pieces = std::move(pieces);
    return B200_OK;
}

static int upload_pipe_dir(b200_sys* s, const PipeSchedule& S, const PipeSchedule::Dir& D, int dir, PipeDirMem& M)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    auto ll = [](const std::vector<int64_t>& v) { return std::vector<long long>(v.begin(), v.end()); };
    CK(ctx, M.gW.upload(D.gW, st));
    CK(ctx, M.gCH.upload(D.gCH, st));
    CK(ctx, M.gShflMask.upload(D.gShflMask, st));
    // unified stream: generic groups only (the fast groups use the split streams)
    std::vector<long long> genOff = ll(D.gGenOff);
    CK(ctx, M.gTermOff.upload(genOff, st));
    std::vector<unsigned char> stream((size_t)D.nGenTerms * 12, 0);
    std::vector<int> genFace((size_t)D.nGenTerms, -1);
    for (int g = 0; g < S.nGroups; g++)
    {
        if (D.gFast[g]) continue;
        const int W = D.gW[g], nT = S.gNT[g];
        unsigned char* gs = stream.data() + (size_t)D.gGenOff[g] * 12;
        for (int step = 0; step < nT; step++)
            memcpy(gs + (size_t)step * W * 384 + (size_t)W * 256, &D.code[D.gTermOff[g] + (int64_t)step * W * 32], (size_t)W * 128);
        if (W) memcpy(&genFace[D.gGenOff[g]], &D.face[D.gTermOff[g]], sizeof(int) * (size_t)nT * W * 32);
    }
    CK(ctx, M.stream.upload(stream, st));
    CK(ctx, M.face.upload(genFace, st));
    // split streams
    CK(ctx, M.gFast.upload(D.gFast, st));
    CK(ctx, M.gLg.upload(D.gLg, st));
    CK(ctx, M.gRg.upload(D.gRg, st));
    CK(ctx, M.gKg.upload(D.gKg, st));
    std::vector<long long> pOff = ll(D.gPOff), cOff = ll(D.gCOff), pfOff = ll(D.gPFaceOff), cfOff = ll(D.gCFaceOff);
    CK(ctx, M.gPOff.upload(pOff, st));
    CK(ctx, M.gCOff.upload(cOff, st));
    CK(ctx, M.gPFaceOff.upload(pfOff, st));
    CK(ctx, M.gCFaceOff.upload(cfOff, st));
    CK(ctx, M.pStream.upload(D.pStream, st));
    CK(ctx, M.cStream.upload(D.cStream, st));
    CK(ctx, M.pFace.upload(D.pFace, st));
    CK(ctx, M.cFace.upload(D.cFace, st));
    if (dir < 0) CK(ctx, M.order.upload(S.orderB, st));
    CK(ctx, cudaStreamSynchronize(st)); // host vectors go out of scope after return
    // shared memory: 256 B of barriers / counters | nStages stages, each one block of kNH steps (records, input
    // vectors, hdr); the generic single-warp path uses its own layout inside the same allocation.  At most
    // ~110 KB per CTA so that two groups are co-resident per SM.
    // direct mode (kernels.cuh, B200_PROD_DIRECT): a stage holds the consumer's operands only (C-block + hdr)
    const int fastStage = kProdDirect ? kSweepBlock * D.maxCStep : D.maxFastStage;
    const int stageBytes = (std::max(fastStage, D.maxGenStage) + 127) / 128 * 128;
    // two CTAs per SM, so that every group of up to 2 x SMs groups runs from the start: (228 KB - 2 x 1 KB reserved) / 2 =
    // 113 KB each, static shared memory included.  (One CTA per SM with a ring twice as deep - B200_SWEEP_SMEM_KB=220
    // B200_SWEEP_STAGES=8 - is 3-4 % faster for an isolated precondition() of C2, 176 groups on 148 SMs, but slower inside
    // the Krylov loop: 173 / 168 us per sweep against 166 / 159.)
    int nStages = getenv("B200_SWEEP_STAGES") ? atoi(getenv("B200_SWEEP_STAGES")) : (kProdDirect ? 8 : 4);
    nStages = std::max(2, std::min(nStages, kSweepMaxStages));
    const int smemCapKB = getenv("B200_SWEEP_SMEM_KB") ? atoi(getenv("B200_SWEEP_SMEM_KB")) : 113;
    while (nStages > 2 && kSweepSmemHeader + 256 + nStages * stageBytes > smemCapKB * 1024) nStages--;
    // the producers of the fast path work in kProdSets sets on consecutive blocks; a set must see every fill of "its"
    // stages (kernels.cuh, split_producer_sets)
    if (kProdSets > 1 && nStages >= kProdSets) nStages -= nStages % kProdSets;
    if (kProdSets > 1 && nStages < kProdSets) return set_err(ctx, B200_ESTATE, "sweep ring: %d stages of %d B do not fit %d producer sets", nStages, stageBytes, kProdSets);
    M.smemBytes = kSweepSmemHeader + nStages * stageBytes;
    M.dev.nGroups = S.nGroups;
    M.dev.dir = dir;
    M.dev.nStages = nStages;
    M.dev.stageBytes = stageBytes;
    M.dev.gBase = s->gBase.p;
    M.dev.gNT = s->gNT.p;
    M.dev.gW = M.gW.p;
    M.dev.gCH = M.gCH.p;
    M.dev.gShflMask = M.gShflMask.p;
    M.dev.gTermOff = M.gTermOff.p;
    M.dev.order = dir < 0 ? M.order.p : nullptr;
    M.dev.stream = M.stream.p;
    M.dev.face = M.face.p;
    M.dev.gFast = M.gFast.p;
    M.dev.gLg = M.gLg.p;
    M.dev.gRg = M.gRg.p;
    M.dev.gKg = M.gKg.p;
    M.dev.gPOff = M.gPOff.p;
    M.dev.gCOff = M.gCOff.p;
    M.dev.gPFaceOff = M.gPFaceOff.p;
    M.dev.gCFaceOff = M.gCFaceOff.p;
    M.dev.pStream = M.pStream.p;
    M.dev.cStream = M.cStream.p;
    M.dev.pFace = M.pFace.p;
    M.dev.cFace = M.cFace.p;
    M.dev.stats = nullptr;
    M.dev.l2Ahead = getenv("B200_SWEEP_L2AHEAD") ? std::max(0, atoi(getenv("B200_SWEEP_L2AHEAD"))) : kL2Ahead;
    M.dev.debugFlags = getenv("B200_SWEEP_DEBUG") ? atoi(getenv("B200_SWEEP_DEBUG")) : 0;
    M.dev.spinLimit = getenv("B200_SWEEP_SPIN_LIMIT") ? std::max(1, atoi(getenv("B200_SWEEP_SPIN_LIMIT"))) : kSweepSpinLimit;
    if (getenv("B200_SWEEP_INFO"))
    { // developer print: shape of the split streams
        std::map<std::tuple<int, int, int, int>, int> hist;
        for (int g = 0; g < S.nGroups; g++) hist[std::make_tuple((int)D.gFast[g], D.gLg[g], D.gRg[g], D.gKg[g])]++;
        fprintf(stderr, "[b200] sweep dir %+d: %d groups, %lld non-canonical steps (%lld linked, %lld dual), stage %d B x %d;", dir, S.nGroups,
                (long long)D.nGeneralSteps, (long long)D.nLinkedGeneralSteps, (long long)D.nDualSteps, stageBytes, nStages);
        for (auto& kv : hist)
            fprintf(stderr, " fast=%d Lg=%d Rg=%d Kg=%d: %d groups;", std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first),
                    std::get<3>(kv.first), kv.second);
        fprintf(stderr, "\n");
    }
    M.packed[0] = M.packed[1] = false;
    M.haveT = false;
    return B200_OK;
}

// Peer-to-peer halo of a system: every rank publishes the cudaIpc handle of its receive buffer and the table
// (peer rank, offset, count) of its segments; a rank finds in each neighbour's record where its own data has to go:
// in the segment the neighbour keeps for this rank (matchIndex[p] < 0: one segment per neighbour rank), or in the
// segment with the given index (block systems: one segment per interface, matchIndex[p] = the peer's interface index).
// Counts and offsets in doubles.  Collective over the ranks that have segments; `k` names the rendezvous records.
static int setup_halo_link_generic(b200_ctx* ctx, int k, const std::vector<int>& peers, const std::vector<int64_t>& sendCount,
                                   const std::vector<int64_t>& recvOff, const std::vector<int>& matchIndex, HaloLink& H)
{
    PeerLink& L = ctx->peer;
    const int nPeers = (int)peers.size();
    H.recvTotal = (size_t)recvOff.back();
    const size_t bytes = 2 * H.recvTotal * sizeof(double) + 2 * (size_t)nPeers * sizeof(unsigned long long);
    CK(ctx, cudaMalloc((void**)&H.buf, bytes));
    CK(ctx, cudaMemset(H.buf, 0, bytes));
    CK(ctx, cudaMalloc((void**)&H.counters, sizeof(unsigned) * nPeers));
    CK(ctx, cudaMemset(H.counters, 0, sizeof(unsigned) * nPeers));
    CK(ctx, cudaDeviceSynchronize());
    struct Seg
    {
        int32_t peer, pad;
        int64_t off, count;
    };
    struct Head
    {
        cudaIpcMemHandle_t h;
        int64_t recvTotal;
        int32_t nPeers, pad;
    };
    std::vector<unsigned char> rec(sizeof(Head) + sizeof(Seg) * nPeers);
    Head* hd = reinterpret_cast<Head*>(rec.data());
    CK(ctx, cudaIpcGetMemHandle(&hd->h, H.buf));
    hd->recvTotal = (int64_t)H.recvTotal;
    hd->nPeers = nPeers;
    hd->pad = 0;
    Seg* sg = reinterpret_cast<Seg*>(rec.data() + sizeof(Head));
    for (int p = 0; p < nPeers; p++) sg[p] = Seg{peers[p], 0, recvOff[p], recvOff[p + 1] - recvOff[p]};
    const std::string tag = "sys" + std::to_string(k);
    if (!peer_publish(L, "r" + std::to_string(ctx->rank) + "." + tag, rec.data(), rec.size()))
        return set_err(ctx, B200_ECUDA, "peer-to-peer halo: cannot write the rendezvous record in %s", L.dir.c_str());
    H.peerData.assign(nPeers, nullptr);
    H.peerTotal.assign(nPeers, 0);
    H.peerFlag.assign(nPeers, nullptr);
    H.peerNPeers.assign(nPeers, 0);
    std::map<int, void*> mapped; // one mapping per neighbour rank
    for (int p = 0; p < nPeers; p++)
    {
        if (peers[p] == ctx->rank) continue; // a pair of patches on this rank: nothing to map (block systems)
        std::vector<unsigned char> theirs;
        if (!peer_fetch(L, "r" + std::to_string(peers[p]) + "." + tag, theirs) || theirs.size() < sizeof(Head))
            return set_err(ctx, B200_ECUDA, "peer-to-peer halo: rank %d did not publish %s", peers[p], tag.c_str());
        const Head* th = reinterpret_cast<const Head*>(theirs.data());
        const Seg* ts = reinterpret_cast<const Seg*>(theirs.data() + sizeof(Head));
        int j = -1;
        if (matchIndex[p] >= 0)
            j = (matchIndex[p] < th->nPeers && ts[matchIndex[p]].peer == ctx->rank) ? matchIndex[p] : -1;
        else
            for (int q = 0; q < th->nPeers; q++)
                if (ts[q].peer == ctx->rank) j = q;
        if (j < 0 || ts[j].count != sendCount[p])
            return set_err(ctx, B200_EINVAL, "peer-to-peer halo: rank %d expects %lld values from rank %d, which sends %lld", peers[p],
                           (long long)(j < 0 ? -1 : ts[j].count), ctx->rank, (long long)sendCount[p]);
        void* q = nullptr;
        if (mapped.count(peers[p]))
            q = mapped[peers[p]];
        else
        {
            CK(ctx, cudaIpcOpenMemHandle(&q, th->h, cudaIpcMemLazyEnablePeerAccess));
            H.opened.push_back(q);
            mapped[peers[p]] = q;
        }
        H.peerData[p] = static_cast<double*>(q) + ts[j].off;
        H.peerTotal[p] = (size_t)th->recvTotal;
        H.peerFlag[p] = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(q) + 2 * (size_t)th->recvTotal * sizeof(double)) + j;
        H.peerNPeers[p] = th->nPeers;
    }
    H.enabled = true;
    return B200_OK;
}

static int setup_halo_link(b200_sys* s)
{
    b200_ctx* ctx = s->ctx;
    const int k = ctx->peer.sysCount++;
    if (s->peers.empty()) return B200_OK;
    const int nPeers = (int)s->peers.size();
    std::vector<int64_t> sendCount(nPeers), recvOff(s->recvOff.begin(), s->recvOff.end());
    for (int p = 0; p < nPeers; p++) sendCount[p] = s->sendOff[p + 1] - s->sendOff[p];
    return setup_halo_link_generic(ctx, k, s->peers, sendCount, recvOff, std::vector<int>(nPeers, -1), s->halo);
}

// Interface / halo plan of the system (schedule.hpp, IfacePlan): built at finalize, and rebuilt (first == false: the
// coefficient arrays keep their contents) when the GGI interpolation of an interface was replaced.
static int build_iface_plan(b200_sys* s, bool first)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    IfacePlan P;
    try
    {
        P.build(s->regs, ctx->rank, s->slotOfCellHost, s->nSlots);
    }
    catch (const std::exception& e)
    {
        return set_err(ctx, B200_EINVAL, "interface plan: %s", e.what());
    }
    s->nTouched = (int)P.rows.size();
    s->nIfCoefs = P.nCoefs;
    s->peers = P.peers;
    s->sendOff = P.sendOff;
    s->recvOff = P.recvOff;
    CK(ctx, s->ifRows.upload(P.rows, st));
    CK(ctx, s->ifRowStart.upload(P.rowStart, st));
    CK(ctx, s->ifEntCoef.upload(P.entCoef, st));
    CK(ctx, s->ifEntSrc.upload(P.entSrc, st));
    CK(ctx, s->ifEntCnt.upload(P.entCnt, st));
    CK(ctx, s->ifGSrc.upload(P.gSrc, st));
    CK(ctx, s->ifGW.upload(P.gW, st));
    CK(ctx, s->ifaceMask.upload(P.sliceMask, st));
    CK(ctx, s->sendCells.upload(P.sendCells, st));
    CK(ctx, s->sendBuf.alloc(P.sendOff.back()));
    CK(ctx, s->recvBuf.alloc(P.recvOff.back()));
    if (first)
    {
        CK(ctx, s->ifCoefBou.alloc(P.nCoefs));
        CK(ctx, s->ifCoefInt.alloc(P.nCoefs));
        CK(ctx, cudaMemsetAsync(s->ifCoefBou.p, 0, (P.nCoefs ? P.nCoefs : 1) * sizeof(double), st));
        CK(ctx, cudaMemsetAsync(s->ifCoefInt.p, 0, (P.nCoefs ? P.nCoefs : 1) * sizeof(double), st));
    }
    CK(ctx, cudaStreamSynchronize(st)); // host vectors go out of scope after return
    s->ifacePlanDirty = false;
    if (first && ctx->nranks > 1 && ctx->peer.enabled)
    {
        const int rc = setup_halo_link(s);
        if (rc) return rc;
    }
    else if (!first && s->halo.enabled && s->halo.recvTotal != (size_t)s->recvOff.back())
        return set_err(ctx, B200_ESTATE, "the processor patches of a finalized system cannot change");
    return B200_OK;
}

extern "C" int b200_sys_finalize(b200_sys* s)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (s->finalized) return B200_OK;
    for (size_t r = 0; r < s->regs.size(); r++)
        if (!s->regs[r].set) return set_err(ctx, B200_ESTATE, "region %zu not set", r);
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    try
    {
        int64_t co = 0, fo = 0;
        for (auto& R : s->regs)
        {
            R.cellOffset = co;
            R.faceOffset = fo;
            co += R.nCells;
            fo += R.nFaces;
        }
        GlobalLdu g;
        g.build(s->regs);
        s->N = g.N;
        s->F = g.F;
        PipeSchedule S;
        S.forceGeneric = getenv("B200_FORCE_GENERIC") != nullptr;
        if (getenv("B200_SWEEP_NO_SEAM_MERGE")) S.mergeSeams = false;
        S.build(g, s->regs);
        s->nSlots = S.nSlots;
        s->nGroups = S.nGroups;
        CK(ctx, s->slotOfCell.upload(S.slotOfCell, st));
        CK(ctx, s->cellOfSlot.upload(S.cellOfSlot, st));
        CK(ctx, s->gBase.upload(S.gBase, st));
        CK(ctx, s->gNT.upload(S.gNT, st));
        {
            SellLayout sell;
            sell.build(g, S);
            s->nSlices = sell.nSlices;
            s->nEntries = sell.nEntries;
            CK(ctx, s->sliceOff.upload(sell.sliceOff, st));
            CK(ctx, s->sellCol.upload(sell.col, st));
            CK(ctx, s->sellSrc.upload(sell.src, st));
            CK(ctx, s->sellVal.alloc(sell.nEntries));
            CK(ctx, cudaStreamSynchronize(st));
        }
        s->slotOfCellHost = S.slotOfCell;
        {
            const int rcI = build_iface_plan(s, true);
            if (rcI) return rcI;
        }
        int rc = upload_pipe_dir(s, S, S.fwd, +1, s->fwd);
        if (rc) return rc;
        rc = upload_pipe_dir(s, S, S.bwd, -1, s->bwd);
        if (rc) return rc;
    }
    catch (const std::exception& e)
    {
        return set_err(ctx, B200_EINVAL, "b200_sys_finalize: %s", e.what());
    }
    // matrix + vectors + scalars
    CK(ctx, s->diag.alloc(s->nSlots));
    CK(ctx, cudaMemsetAsync(s->diag.p, 0, std::max<size_t>(1, s->nSlots) * sizeof(double), st));
    CK(ctx, s->coef.alloc(2 * s->F));
    CK(ctx, s->rD.alloc(s->nSlots));
    CK(ctx, s->rDraw.alloc(s->nSlots));
    CK(ctx, s->stageCell.alloc(s->N));
    CK(ctx, s->diagCell.alloc(s->N));
    for (int v = 0; v < V_COUNT; v++)
    {
        CK(ctx, s->vec[v].alloc(s->nSlots));
        CK(ctx, cudaMemsetAsync(s->vec[v].p, 0, std::max<size_t>(1, s->nSlots) * sizeof(double), st));
    }
    {
        // the attribute is per kernel function, not per system: several systems may be alive (one per region in the
        // partitioned loop), so it only ever grows (per device)
        static int maxSmemSet[64] = {0};
        int& seen = maxSmemSet[ctx->device & 63];
        const int maxSmem = std::max(seen, std::max(s->fwd.smemBytes, s->bwd.smemBytes));
        if (maxSmem > 227 * 1024) return set_err(ctx, B200_EUNSUPPORTED, "sweep stage needs %d bytes of shared memory", maxSmem);
        seen = maxSmem;
        CK(ctx, cudaFuncSetAttribute(k_sweep<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
        CK(ctx, cudaFuncSetAttribute(k_sweep<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxSmem));
    }
    CK(ctx, s->sc.alloc(1));
    CK(ctx, cudaMemsetAsync(s->sc.p, 0, sizeof(DevScalars), st));
    CK(ctx, s->ticket.alloc(1));
    CK(ctx, cudaMemsetAsync(s->ticket.p, 0, sizeof(unsigned), st));
    s->ticketBase = 0;
    CK(ctx, s->devErr.alloc(1));
    CK(ctx, cudaMemsetAsync(s->devErr.p, 0, sizeof(int), st));
    CK(ctx, s->history.alloc(kHistOnDevice));
    // launch geometry: persistent-style grids sized in multiples of the SM count
    const int sm = ctx->smCount;
    auto cdiv = [](int64_t a, int64_t b) { return (a + b - 1) / b; };
    s->vecBlocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(s->nSlots, 256 * 2), (int64_t)sm * 8));
    s->amulBlocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(s->nSlices, 8), (int64_t)sm * 16));
    s->ifaceBlocks = (int)cdiv(s->nTouched, 128);
    s->pstride = std::max(s->vecBlocks, s->amulBlocks + s->ifaceBlocks) + 32;
    CK(ctx, s->partials.alloc((size_t)s->pstride * kMaxDots));
    CK(ctx, cudaMemsetAsync(s->partials.p, 0, (size_t)s->pstride * kMaxDots * sizeof(double), st));
    CK(ctx, cudaHostAlloc((void**)&s->hostFlags, 4 * sizeof(int), cudaHostAllocMapped));
    memset(s->hostFlags, 0, 4 * sizeof(int));
    CK(ctx, cudaHostGetDevicePointer((void**)&s->devHostFlags, s->hostFlags, 0));
    CK(ctx, cudaEventCreate(&s->evSolveA));
    CK(ctx, cudaEventCreate(&s->evSolveB));
    // global cell count (gAverage denominator)
    s->nGlobalCells = (double)s->N;
    if (ctx->nranks > 1)
    {
        double* d = s->vec[V_TMP].p;
        double hN = (double)s->N;
        CK(ctx, cudaMemcpyAsync(d, &hN, sizeof(double), cudaMemcpyHostToDevice, st));
        if (ctx->peer.enabled)
        {
            PeerLink& L = ctx->peer;
            ctx->launches++;
            k_peer_allreduce<<<1, 32, 0, st>>>(d, 1, PeerAR{L.dPeerMbox, L.mbox, ctx->rank, ctx->nranks, ++L.arSeq}, L.err);
            CK(ctx, cudaGetLastError());
        }
        else
            NK(ctx, g_nccl.AllReduce(d, d, 1, ncclDouble, ncclSum, ctx->comm, st));
        CK(ctx, cudaMemcpyAsync(&hN, d, sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(ctx, cudaStreamSynchronize(st));
        s->nGlobalCells = hN;
    }
    CK(ctx, cudaStreamSynchronize(st));
    s->finalized = true;
    return B200_OK;
}

extern "C" int64_t b200_sys_num_cells(const b200_sys* s) { return s ? s->N : 0; }
extern "C" int64_t b200_sys_num_faces(const b200_sys* s) { return s ? s->F : 0; }

extern "C" int b200_sys_set_coeffs(b200_sys* s, int r, const double* diag, const double* upper, const double* lower)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->finalized) return set_err(ctx, B200_ESTATE, "set_coeffs before finalize");
    if (r < 0 || r >= (int)s->regs.size() || !diag || (!upper && s->regs[r].nFaces > 0))
        return set_err(ctx, B200_EINVAL, "b200_sys_set_coeffs: bad arguments");
    CK(ctx, cudaSetDevice(ctx->device));
    const RegionHost& R = s->regs[r];
    cudaStream_t st = ctx->stream;
    if (R.nCells) CK(ctx, cudaMemcpyAsync(s->diagCell.p + R.cellOffset, diag, sizeof(double) * R.nCells, cudaMemcpyHostToDevice, st));
    s->diagDirty = true; // permuted to slot order once all regions are in (ensure_sell)
    if (R.nFaces)
    {
        CK(ctx, cudaMemcpyAsync(s->coef.p + R.faceOffset, upper, sizeof(double) * R.nFaces, cudaMemcpyHostToDevice, st));
        // symmetric matrix: lower aliases upper - filled on the device, the coefficients cross the bus once
        if (lower)
            CK(ctx, cudaMemcpyAsync(s->coef.p + s->F + R.faceOffset, lower, sizeof(double) * R.nFaces, cudaMemcpyHostToDevice, st));
        else
            CK(ctx, cudaMemcpyAsync(s->coef.p + s->F + R.faceOffset, s->coef.p + R.faceOffset, sizeof(double) * R.nFaces,
                                    cudaMemcpyDeviceToDevice, st));
    }
    s->regionHasCoeffs[r] = 1;
    s->sellDirty = s->sellTDirty = true;
    s->precondValid = -1;
    s->fwd.packed[0] = s->fwd.packed[1] = s->bwd.packed[0] = s->bwd.packed[1] = false;
    return B200_OK;
}

#include "fv_assemble.cuh"

extern "C" int b200_sys_set_interface_coeffs(b200_sys* s, int r, int iface, const double* bouCoeffs, const double* intCoeffs)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->finalized) return set_err(ctx, B200_ESTATE, "set_interface_coeffs before finalize");
    if (r < 0 || r >= (int)s->regs.size() || iface < 0 || iface >= (int)s->regs[r].ifaces.size() || !bouCoeffs)
        return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_coeffs: bad arguments");
    CK(ctx, cudaSetDevice(ctx->device));
    const IfaceHost& I = s->regs[r].ifaces[iface];
    if (I.nFaces)
    {
        CK(ctx, cudaMemcpyAsync(s->ifCoefBou.p + I.coefOffset, bouCoeffs, sizeof(double) * I.nFaces, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaMemcpyAsync(s->ifCoefInt.p + I.coefOffset, intCoeffs ? intCoeffs : bouCoeffs, sizeof(double) * I.nFaces,
                                cudaMemcpyHostToDevice, ctx->stream));
    }
    return B200_OK;
}

extern "C" int b200_sys_set_interface_attached(b200_sys* s, int r, int iface, int attached)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (r < 0 || r >= (int)s->regs.size() || iface < 0 || iface >= (int)s->regs[r].ifaces.size())
        return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_attached: bad arguments");
    IfaceHost& I = s->regs[r].ifaces[iface];
    if (I.kind != B200_IFACE_REGION_COUPLE) return set_err(ctx, B200_EINVAL, "only regionCouple interfaces attach / detach");
    if (I.attached != (attached != 0))
    {
        I.attached = attached != 0;
        s->nDetached += I.attached ? -1 : 1;
    }
    return B200_OK;
}

extern "C" int b200_sys_set_interface_ggi(b200_sys* s, int r, int iface, int32_t nPeerFaces, const int32_t* ggiOffsets,
                                          const int32_t* ggiAddr, const double* ggiWeights)
{
    if (!s) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (r < 0 || r >= (int)s->regs.size() || iface < 0 || iface >= (int)s->regs[r].ifaces.size())
        return set_err(ctx, B200_EINVAL, "b200_sys_set_interface_ggi: bad arguments");
    IfaceHost& I = s->regs[r].ifaces[iface];
    if (I.kind != B200_IFACE_REGION_COUPLE) return set_err(ctx, B200_EINVAL, "only regionCouple interfaces carry a GGI interpolation");
    if (!ggiOffsets)
    {
        if (nPeerFaces != I.nFaces) return set_err(ctx, B200_EINVAL, "identity interface needs nPeerFaces == nFaces");
        I.identity = true;
        I.ggiOffsets.clear();
        I.ggiAddr.clear();
        I.ggiWeights.clear();
    }
    else
    {
        if (!ggiAddr || !ggiWeights) return set_err(ctx, B200_EINVAL, "GGI interface needs addr and weights");
        const int nnz = ggiOffsets[I.nFaces];
        if (ggiOffsets[0] != 0 || nnz < 0) return set_err(ctx, B200_EINVAL, "bad GGI offsets");
        for (int i = 0; i < I.nFaces; i++)
            if (ggiOffsets[i + 1] < ggiOffsets[i]) return set_err(ctx, B200_EINVAL, "GGI offsets must not decrease");
        for (int k = 0; k < nnz; k++)
            if (ggiAddr[k] < 0 || ggiAddr[k] >= nPeerFaces) return set_err(ctx, B200_EINVAL, "GGI address %d out of range of the shadow patch", ggiAddr[k]);
        I.identity = false;
        I.ggiOffsets.assign(ggiOffsets, ggiOffsets + I.nFaces + 1);
        I.ggiAddr.assign(ggiAddr, ggiAddr + nnz);
        I.ggiWeights.assign(ggiWeights, ggiWeights + nnz);
    }
    I.nPeerFaces = nPeerFaces;
    if (s->finalized) s->ifacePlanDirty = true; // rebuilt at the next use
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ launch helpers
static int ensure_sell(b200_sys* s, bool transpose)
{
    b200_ctx* ctx = s->ctx;
    for (size_t r = 0; r < s->regs.size(); r++)
        if (!s->regionHasCoeffs[r]) return set_err(ctx, B200_ESTATE, "region %zu has no coefficients", r);
    if (s->nDetached)
        for (size_t r = 0; r < s->regs.size(); r++)
            for (size_t i = 0; i < s->regs[r].ifaces.size(); i++)
                if (!s->regs[r].ifaces[i].attached)
                    return set_err(ctx, B200_ESTATE,
                                   "regionCouple interface %zu of region %zu is detached: the coupled matrix-vector product needs attached "
                                   "patches (monolithicCouplingFvPatchField::initInterfaceMatrixUpdate is fatal for a detached patch)",
                                   i, r);
    if (s->ifacePlanDirty)
    {
        const int rc = build_iface_plan(s, false);
        if (rc) return rc;
    }
    if (s->nSlots == 0) return B200_OK;
    if (s->diagDirty)
    {
        KScope k(s, B200_K_PACK);
        k_to_slots<<<s->vecBlocks, 256, 0, ctx->stream>>>((size_t)s->nSlots, s->cellOfSlot.p, s->diagCell.p, s->diag.p);
        s->diagDirty = false;
    }
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((s->nEntries + 255) / 256, (int64_t)ctx->smCount * 16));
    if (!transpose && s->sellDirty)
    {
        KScope k(s, B200_K_PACK);
        k_pack_sell<<<blocks, 256, 0, ctx->stream>>>((size_t)s->nEntries, s->sellSrc.p, s->coef.p, s->sellVal.p, 0);
        s->sellDirty = false;
    }
    if (transpose && s->sellTDirty)
    {
        if (s->sellValT.n != (size_t)s->nEntries) CK(ctx, s->sellValT.alloc(s->nEntries));
        KScope k(s, B200_K_PACK);
        k_pack_sell<<<blocks, 256, 0, ctx->stream>>>((size_t)s->nEntries, s->sellSrc.p, s->coef.p, s->sellValT.p, (int)s->F);
        s->sellTDirty = false;
    }
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

// all-reduce red[0..nd) over ranks, then apply the scalar op
static int reduce_finish(b200_sys* s, const PartCounts& cnt, int nd, int op, int force)
{
    b200_ctx* ctx = s->ctx;
    const bool multi = ctx->nranks > 1;
    if (multi && ctx->peer.enabled)
    { // sums, all-reduce over the peer mailboxes and the scalar recurrences in one launch
        KScope k(s, B200_K_REDUCE);
        PeerLink& L = ctx->peer;
        const PeerAR ar{L.dPeerMbox, L.mbox, ctx->rank, ctx->nranks, ++L.arSeq};
        k_finalize<<<1, 1024, 0, ctx->stream>>>(s->partials.p, s->pstride, cnt, nd, s->sc.p, op, 1, force, s->history.p, s->devHostFlags, ar,
                                                 L.err);
        CK(ctx, cudaGetLastError());
        return B200_OK;
    }
    {
        KScope k(s, B200_K_REDUCE);
        k_finalize<<<1, 1024, 0, ctx->stream>>>(s->partials.p, s->pstride, cnt, nd, s->sc.p, op, multi ? 0 : 1, force,
                                                 s->history.p, s->devHostFlags, PeerAR{nullptr, nullptr, 0, 1, 0}, nullptr);
    }
    if (multi)
    {
        {
            KScope k(s, B200_K_HALO);
            double* red = reinterpret_cast<double*>(s->sc.p); // DevScalars::red is the first member
            NK(ctx, g_nccl.AllReduce(red, red, nd, ncclDouble, ncclSum, ctx->comm, ctx->stream));
        }
        KScope k(s, B200_K_REDUCE);
        k_scalar_op<<<1, 1, 0, ctx->stream>>>(s->sc.p, op, force, s->history.p, s->devHostFlags);
    }
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

// y = A x (or A^T x) with nd fused dots of y: nd=1: (y,d0); nd=2: (y,d0),(y,y).  Leaves the
// partials of the dots in quantities 0..nd-1 with counts amulBlocks + ifaceBlocks.
// op 0: y = A x (nd fused dots against d0).  op 1: y = d0 - A x as lduMatrix::residual rounds it.  op 2: y = sumA.
static int launch_amul(b200_sys* s, const double* x, double* y, int nd, const double* d0, bool transpose, int force,
                       PartCounts* cntOut, int op = 0)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    int rc = ensure_sell(s, transpose);
    if (rc) return rc;
    const double* val = transpose ? s->sellValT.p : s->sellVal.p;
    const double* ifc = transpose ? s->ifCoefInt.p : s->ifCoefBou.p;
    // halo exchange of the shadow-side patchInternalField (processorFvPatchField init/update)
    const unsigned long long* haloFlags = nullptr;
    int nHaloPeers = 0;
    unsigned long long haloSeq = 0;
    const double* recvPtr = s->recvBuf.p;
    if (!s->peers.empty() && s->halo.enabled)
    { // peer-to-peer: one push per neighbour, no send / receive buffers, no collective call
        HaloLink& H = s->halo;
        haloSeq = ++H.seq;
        const int par = (int)(haloSeq & 1ull);
        nHaloPeers = (int)s->peers.size();
        haloFlags = H.flags(par, nHaloPeers);
        recvPtr = H.data(par);
        KScope k(s, B200_K_HALO);
        for (int p = 0; p < nHaloPeers; p++)
        {
            const int ns = s->sendOff[p + 1] - s->sendOff[p];
            k_halo_push<<<std::max(1, (ns + 255) / 256), 256, 0, st>>>(ns, s->sendCells.p + s->sendOff[p], x,
                                                                     H.peerData[p] + (size_t)par * H.peerTotal[p],
                                                                     H.peerFlag[p] + (size_t)par * H.peerNPeers[p], haloSeq, H.counters + p, s->sc.p,
                                                                     1);
        }
    }
    else if (!s->peers.empty())
    {
        const int nSend = s->sendOff.back();
        if (nSend)
        {
            KScope k(s, B200_K_PACK);
            k_halo_pack<<<(nSend + 255) / 256, 256, 0, st>>>(nSend, s->sendCells.p, x, s->sendBuf.p, s->sc.p, 1);
        }
        KScope k(s, B200_K_HALO);
        NK(ctx, g_nccl.GroupStart());
        for (size_t p = 0; p < s->peers.size(); p++)
        {
            const int ns = s->sendOff[p + 1] - s->sendOff[p], nr = s->recvOff[p + 1] - s->recvOff[p];
            if (ns) NK(ctx, g_nccl.Send(s->sendBuf.p + s->sendOff[p], ns, ncclDouble, s->peers[p], ctx->comm, st));
            if (nr) NK(ctx, g_nccl.Recv(s->recvBuf.p + s->recvOff[p], nr, ncclDouble, s->peers[p], ctx->comm, st));
        }
        NK(ctx, g_nccl.GroupEnd());
    }
    if (s->nSlots > 0)
    {
        KScope k(s, B200_K_AMUL);
        const unsigned* mask = s->nTouched ? s->ifaceMask.p : nullptr;
        if (op == 1)
            k_amul<0, 1><<<s->amulBlocks, 256, 0, st>>>((int)s->nSlots, (int)s->nSlices, s->diag.p, s->sliceOff.p, s->sellCol.p, val, x, y,
                                                        d0, mask, s->partials.p, s->pstride, s->sc.p, force);
        else if (op == 2)
            k_amul<0, 2><<<s->amulBlocks, 256, 0, st>>>((int)s->nSlots, (int)s->nSlices, s->diag.p, s->sliceOff.p, s->sellCol.p, val, x, y,
                                                        nullptr, mask, s->partials.p, s->pstride, s->sc.p, force);
        else if (nd == 0)
            k_amul<0><<<s->amulBlocks, 256, 0, st>>>((int)s->nSlots, (int)s->nSlices, s->diag.p, s->sliceOff.p, s->sellCol.p, val, x, y,
                                                     nullptr, mask, s->partials.p, s->pstride, s->sc.p, force);
        else if (nd == 1)
            k_amul<1><<<s->amulBlocks, 256, 0, st>>>((int)s->nSlots, (int)s->nSlices, s->diag.p, s->sliceOff.p, s->sellCol.p, val, x, y,
                                                     d0, mask, s->partials.p, s->pstride, s->sc.p, force);
        else
            k_amul<2><<<s->amulBlocks, 256, 0, st>>>((int)s->nSlots, (int)s->nSlices, s->diag.p, s->sliceOff.p, s->sellCol.p, val, x, y,
                                                     d0, mask, s->partials.p, s->pstride, s->sc.p, force);
    }
    if (s->nTouched > 0)
    {
        KScope k(s, B200_K_IFACE);
#define IFACE_ARGS                                                                                                           \
    s->nTouched, s->ifRows.p, s->ifRowStart.p, s->ifEntCoef.p, s->ifEntSrc.p, s->ifEntCnt.p, s->ifGSrc.p, s->ifGW.p, ifc, x, \
        recvPtr, y, d0, s->partials.p, s->pstride, s->amulBlocks, s->sc.p, force, haloFlags, nHaloPeers, haloSeq
        if (op == 1)
            k_iface<0, 1><<<s->ifaceBlocks, 128, 0, st>>>(IFACE_ARGS);
        else if (op == 2)
            k_iface<0, 2><<<s->ifaceBlocks, 128, 0, st>>>(IFACE_ARGS);
        else if (nd == 0)
            k_iface<0><<<s->ifaceBlocks, 128, 0, st>>>(IFACE_ARGS);
        else if (nd == 1)
            k_iface<1><<<s->ifaceBlocks, 128, 0, st>>>(IFACE_ARGS);
        else
            k_iface<2><<<s->ifaceBlocks, 128, 0, st>>>(IFACE_ARGS);
#undef IFACE_ARGS
    }
    CK(ctx, cudaGetLastError());
    if (cntOut)
        for (int k = 0; k < kMaxDots; k++) cntOut->n[k] = (s->nSlots > 0 ? s->amulBlocks : 0) + s->ifaceBlocks;
    return B200_OK;
}

template <int MODE>
static int launch_sweep(b200_sys* s, PipeDirMem& M, PipeDev dev, const double* a, const double* b, double* out, int force)
{
    b200_ctx* ctx = s->ctx;
    if (s->nGroups == 0) return B200_OK;
    dev.stats = M.dev.stats;
    // calcReciprocalD (MODE 2) is the once-per-solve preconditioner construction: accounted with the packing kernels
    KScope k(s, MODE == 2 ? B200_K_PACK : (dev.dir > 0 ? B200_K_SWEEP_FWD : B200_K_SWEEP_BWD));
    // debug counters and time stamps live in separately compiled instantiations, the product path (0) carries none:
    // 1 = per-block time stamps + counters, 2 (debug flag 2) = cheap per-group counters only
    if (dev.stats && !(dev.debugFlags & 2))
        k_sweep<MODE, 1><<<s->nGroups, kSweepThreads, M.smemBytes, ctx->stream>>>(dev, a, b, out, s->ticket.p, s->ticketBase, s->devErr.p, s->sc.p,
                                                                                 force);
    else if (dev.stats)
        k_sweep<MODE, 2><<<s->nGroups, kSweepThreads, M.smemBytes, ctx->stream>>>(dev, a, b, out, s->ticket.p, s->ticketBase, s->devErr.p, s->sc.p,
                                                                                 force);
    else
        k_sweep<MODE, 0><<<s->nGroups, kSweepThreads, M.smemBytes, ctx->stream>>>(dev, a, b, out, s->ticket.p, s->ticketBase, s->devErr.p, s->sc.p,
                                                                                 force);
    s->ticketBase += (unsigned)s->nGroups;
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

static int fill_sentinel(b200_sys* s, double* a, double* b, int force)
{
    if (s->nSlots == 0) return B200_OK;
    KScope k(s, B200_K_VECTOR);
    k_fill2_sentinel<<<s->vecBlocks, 256, 0, s->ctx->stream>>>((size_t)s->nSlots, a, b, s->sc.p, force);
    CK(s->ctx, cudaGetLastError());
    return B200_OK;
}

static int pack_stream(b200_sys* s, const PipeDev& dev, const double* c1, const double* c2, const double* rD, int prodMode)
{
    if (s->nGroups == 0) return B200_OK;
    KScope k(s, B200_K_PACK);
    k_pack_stream<<<dim3(s->nGroups, 8), 256, 0, s->ctx->stream>>>(dev, c1, c2, rD, prodMode);
    CK(s->ctx, cudaGetLastError());
    return B200_OK;
}

static int ensure_transposed_buffers(b200_sys* s, PipeDirMem& M)
{
    if (M.haveT) return B200_OK;
    b200_ctx* ctx = s->ctx;
    auto clone = [&](DevBuf<unsigned char>& dst, const DevBuf<unsigned char>& src) -> cudaError_t {
        cudaError_t e = dst.alloc(src.n);
        if (e != cudaSuccess || src.n == 0) return e;
        return cudaMemcpyAsync(dst.p, src.p, src.n, cudaMemcpyDeviceToDevice, ctx->stream); // static codes / meta
    };
    CK(ctx, clone(M.streamT, M.stream));
    CK(ctx, clone(M.pStreamT, M.pStream));
    CK(ctx, clone(M.cStreamT, M.cStream));
    M.devT = M.dev;
    M.devT.stream = M.streamT.p;
    M.devT.pStream = M.pStreamT.p;
    M.devT.cStream = M.cStreamT.p;
    M.haveT = true;
    M.packed[1] = false;
    return B200_OK;
}

// Preconditioner construction: calcReciprocalD as a forward sweep in division mode, then the
// pre-multiplied sweep coefficients rD[row]*lower[f] / rD[row]*upper[f] packed into the streams.
// transposed: preconditionT swaps the roles of upper and lower (DILUPreconditioner::preconditionT);
// its coefficients live in a second set of stream buffers so that PBiCG can alternate.
static int ensure_precond(b200_sys* s, int precond, bool transposed)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    if (precond == B200_PRECOND_NONE) return B200_OK;
    const int v = transposed ? 1 : 0;
    const bool sweeps = precond >= B200_PRECOND_DIC;
    if (s->precondValid == precond && (!sweeps || (s->fwd.packed[v] && s->bwd.packed[v]))) return B200_OK;
    if (s->nSlots == 0)
    {
        s->precondValid = precond;
        return B200_OK;
    }
    int rc;
    if ((rc = ensure_sell(s, false))) return rc; // diag in slot order
    const double* cU = s->coef.p;
    const double* cL = (precond == B200_PRECOND_DIC) ? s->coef.p : s->coef.p + s->F; // DIC uses upper both ways
    if (s->precondValid != precond)
    {
        if (precond == B200_PRECOND_DIAGONAL)
        {
            KScope k(s, B200_K_VECTOR);
            k_invert<<<s->vecBlocks, 256, 0, st>>>((size_t)s->nSlots, s->diag.p, s->rD.p, s->cellOfSlot.p);
        }
        else
        {
            if ((rc = pack_stream(s, s->fwd.dev, cU, cL, nullptr, 1))) return rc;
            if ((rc = fill_sentinel(s, s->rDraw.p, nullptr, 1))) return rc;
            if ((rc = launch_sweep<2>(s, s->fwd, s->fwd.dev, s->diag.p, nullptr, s->rDraw.p, 1))) return rc;
            {
                KScope k(s, B200_K_VECTOR);
                k_invert<<<s->vecBlocks, 256, 0, st>>>((size_t)s->nSlots, s->rDraw.p, s->rD.p, s->cellOfSlot.p);
            }
        }
        s->precondValid = precond;
        s->fwd.packed[0] = s->fwd.packed[1] = s->bwd.packed[0] = s->bwd.packed[1] = false;
    }
    if (sweeps && !(s->fwd.packed[v] && s->bwd.packed[v]))
    {
        if (transposed)
        {
            if ((rc = ensure_transposed_buffers(s, s->fwd))) return rc;
            if ((rc = ensure_transposed_buffers(s, s->bwd))) return rc;
        }
        const PipeDev& fdev = transposed ? s->fwd.devT : s->fwd.dev;
        const PipeDev& bdev = transposed ? s->bwd.devT : s->bwd.dev;
        if ((rc = pack_stream(s, fdev, transposed ? cU : cL, nullptr, s->rD.p, 0))) return rc;
        if ((rc = pack_stream(s, bdev, transposed ? cL : cU, nullptr, s->rD.p, 0))) return rc;
        s->fwd.packed[v] = s->bwd.packed[v] = true;
    }
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

// w = M^-1 r (transposed: M^-T r).  tmp: scratch for the forward result.  If prefilled, the caller
// already wrote the sentinel into tmp and w (fused into the preceding vector kernel).
static int launch_precondition(b200_sys* s, int precond, const double* r, double* w, double* tmp, bool prefilled, int force,
                               bool transposed = false)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    if (s->nSlots == 0) return B200_OK;
    if (precond == B200_PRECOND_NONE)
    {
        KScope k(s, B200_K_VECTOR);
        k_copy<<<s->vecBlocks, 256, 0, st>>>((size_t)s->nSlots, r, w, s->sc.p, force);
        CK(ctx, cudaGetLastError());
        return B200_OK;
    }
    if (precond == B200_PRECOND_DIAGONAL)
    {
        KScope k(s, B200_K_VECTOR);
        k_mul<<<s->vecBlocks, 256, 0, st>>>((size_t)s->nSlots, s->rD.p, r, w, s->sc.p, force);
        CK(ctx, cudaGetLastError());
        return B200_OK;
    }
    int rc;
    if (!prefilled && (rc = fill_sentinel(s, tmp, w, force))) return rc;
    if ((rc = launch_sweep<0>(s, s->fwd, transposed ? s->fwd.devT : s->fwd.dev, s->rD.p, r, tmp, force))) return rc;
    if ((rc = launch_sweep<1>(s, s->bwd, transposed ? s->bwd.devT : s->bwd.dev, tmp, nullptr, w, force))) return rc;
    return B200_OK;
}

static PartCounts vec_counts(const b200_sys* s)
{
    PartCounts c;
    for (int k = 0; k < kMaxDots; k++) c.n[k] = s->nSlots > 0 ? s->vecBlocks : 0;
    return c;
}

// ------------------------------------------------------------------------------------------ vectors in / out
static int upload_vec(b200_sys* s, int v, const double* const* h)
{
    b200_ctx* ctx = s->ctx;
    for (size_t r = 0; r < s->regs.size(); r++)
    {
        const RegionHost& R = s->regs[r];
        if (!h[r] && R.nCells) return set_err(ctx, B200_EINVAL, "null host vector for region %zu", r);
        if (R.nCells)
            CK(ctx, cudaMemcpyAsync(s->stageCell.p + R.cellOffset, h[r], sizeof(double) * R.nCells, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (s->nSlots)
    {
        KScope k(s, B200_K_PACK);
        k_to_slots<<<s->vecBlocks, 256, 0, ctx->stream>>>((size_t)s->nSlots, s->cellOfSlot.p, s->stageCell.p, s->vec[v].p);
        CK(ctx, cudaGetLastError());
    }
    return B200_OK;
}

static int download_vec(b200_sys* s, const double* dev, double* const* h)
{
    b200_ctx* ctx = s->ctx;
    if (s->N)
    {
        KScope k(s, B200_K_PACK);
        k_from_slots<<<s->vecBlocks, 256, 0, ctx->stream>>>((size_t)s->N, s->slotOfCell.p, dev, s->stageCell.p);
        CK(ctx, cudaGetLastError());
    }
    for (size_t r = 0; r < s->regs.size(); r++)
    {
        const RegionHost& R = s->regs[r];
        if (!h[r] && R.nCells) return set_err(ctx, B200_EINVAL, "null host vector for region %zu", r);
        if (R.nCells)
            CK(ctx, cudaMemcpyAsync(h[r], s->stageCell.p + R.cellOffset, sizeof(double) * R.nCells, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return B200_OK;
}

static int check_device_error(b200_sys* s)
{
    int e = 0;
    CK(s->ctx, cudaMemcpyAsync(&e, s->devErr.p, sizeof(int), cudaMemcpyDeviceToHost, s->ctx->stream));
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    if (e)
    {
        cudaMemsetAsync(s->devErr.p, 0, sizeof(int), s->ctx->stream);
        return set_err(s->ctx, B200_EDEVICE, "sweep kernel timed out waiting for a dependency (device error %d)", e);
    }
    if (s->ctx->peer.enabled)
    {
        CK(s->ctx, cudaMemcpy(&e, s->ctx->peer.err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e)
        {
            cudaMemset(s->ctx->peer.err, 0, sizeof(int));
            return set_err(s->ctx, B200_EDEVICE, "a peer rank did not answer a peer-to-peer all-reduce (device error %d)", e);
        }
    }
    return B200_OK;
}

extern "C" int b200_upload(b200_sys* s, const double* const* x, const double* const* b)
{
    if (!s || !s->finalized) return s ? set_err(s->ctx, B200_ESTATE, "upload before finalize") : B200_EINVAL;
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc = B200_OK;
    if (x) rc = upload_vec(s, V_X, x);
    if (!rc && b) rc = upload_vec(s, V_B, b);
    if (rc) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    return B200_OK;
}

extern "C" int b200_download(b200_sys* s, double* const* x)
{
    if (!s || !s->finalized || !x) return s ? set_err(s->ctx, B200_ESTATE, "download before finalize") : B200_EINVAL;
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc = download_vec(s, s->vec[V_X].p, x);
    if (rc) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    return B200_OK;
}

extern "C" int b200_x_save(b200_sys* s)
{
    if (!s || !s->finalized) return s ? set_err(s->ctx, B200_ESTATE, "x_save before finalize") : B200_EINVAL;
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    if (s->x0.n != (size_t)s->nSlots) CK(s->ctx, s->x0.alloc(s->nSlots));
    if (s->nSlots) CK(s->ctx, cudaMemcpyAsync(s->x0.p, s->vec[V_X].p, sizeof(double) * s->nSlots, cudaMemcpyDeviceToDevice, s->ctx->stream));
    s->x0Valid = true;
    return B200_OK;
}

extern "C" int b200_x_restore(b200_sys* s)
{
    if (!s || !s->finalized) return s ? set_err(s->ctx, B200_ESTATE, "x_restore before finalize") : B200_EINVAL;
    if (!s->x0Valid) return set_err(s->ctx, B200_ESTATE, "x_restore without x_save");
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    if (s->nSlots) CK(s->ctx, cudaMemcpyAsync(s->vec[V_X].p, s->x0.p, sizeof(double) * s->nSlots, cudaMemcpyDeviceToDevice, s->ctx->stream));
    return B200_OK;
}

extern "C" int b200_host_register(b200_ctx* ctx, void* p, uint64_t bytes)
{
    if (!ctx || !p) return set_err(ctx, B200_EINVAL, "b200_host_register: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    if (bytes) CK(ctx, cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
    return B200_OK;
}

extern "C" int b200_host_unregister(b200_ctx* ctx, void* p)
{
    if (!ctx || !p) return set_err(ctx, B200_EINVAL, "b200_host_unregister: null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaHostUnregister(p));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ Krylov drivers
static int upload_scalars(b200_sys* s, const b200_solver_opts* o, int histCap)
{
    DevScalars h;
    memset(&h, 0, sizeof(h));
    h.tolerance = o->tolerance;
    h.relTol = o->relTol;
    h.minIter = o->minIter;
    h.maxIter = o->maxIter;
    h.nGlobalCells = s->nGlobalCells;
    h.histCap = std::min(histCap, (int)kHistOnDevice);
    s->hostFlags[0] = 0;
    s->hostFlags[1] = 0;
    s->hostFlags[2] = 0;
    CK(s->ctx, cudaMemcpyAsync(s->sc.p, &h, sizeof(h), cudaMemcpyHostToDevice, s->ctx->stream));
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream)); // h is on the stack
    return B200_OK;
}

// Common head of every solver: wA = A x, normFactor, r = b - wA, initial residual, stop check.
static int solve_head(b200_sys* s, int op, double* Ax, double* r, double* rw, double* zero1, double* zero2)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    int rc;
    const size_t n = (size_t)s->nSlots;
    if ((rc = launch_amul(s, s->vec[V_X].p, Ax, 0, nullptr, false, 1, nullptr))) return rc;
    // xRef = gAverage(x); tmp = A * xRef (coupledIterativeSolver::normFactor)
    if (n)
    {
        KScope k(s, B200_K_VECTOR);
        k_sum<<<s->vecBlocks, 256, 0, st>>>(n, s->vec[V_X].p, s->partials.p, s->pstride);
    }
    if ((rc = reduce_finish(s, vec_counts(s), 1, OP_XREF, 1))) return rc;
    if (n)
    {
        KScope k(s, B200_K_VECTOR);
        k_fill_xref<<<s->vecBlocks, 256, 0, st>>>(n, s->vec[V_TMP2].p, s->sc.p);
    }
    if ((rc = launch_amul(s, s->vec[V_TMP2].p, s->vec[V_TMP].p, 0, nullptr, false, 1, nullptr))) return rc;
    if (n)
    {
        KScope k(s, B200_K_VECTOR);
        k_init_residual<<<s->vecBlocks, 256, 0, st>>>(n, s->vec[V_B].p, Ax, s->vec[V_TMP].p, r, rw, zero1, zero2, s->partials.p,
                                                      s->pstride);
    }
    if ((rc = reduce_finish(s, vec_counts(s), 3, op, 1))) return rc;
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

// Keep the host at most kAhead iterations ahead of the device so that the done flag is seen soon.
// Leaving the loop: a single rank breaks as soon as its host sees the mapped done flag.  With several ranks every
// iteration enqueues NCCL calls (the all-reduce of reduce_finish, the halo exchange of launch_amul), so all hosts
// must enqueue exactly the same number of iterations although each of them reads the flag at another moment:
// the device publishes the loop iteration d that set done (hostFlags[2] = d + 2, identical on all ranks because it
// follows from all-reduced scalars), the flag is certainly visible once the event of iteration d has been waited
// for - which throttle_wait does at iteration d + kAhead - and every rank stops exactly there (or at maxIter).
// Iterations enqueued after d are no-ops on the device (every kernel returns on sc->done) whose collectives still pair up.
static int throttle_ahead(const b200_sys* s) { return s->ctx->nranks > 1 ? 3 : 8; }
static int throttle_wait(b200_sys* s, int it)
{
    const int kAhead = throttle_ahead(s);
    if ((int)s->throttle.size() < 8)
    {
        s->throttle.resize(8, nullptr);
        for (auto& e : s->throttle)
            if (!e) CK(s->ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (it >= kAhead) CK(s->ctx, cudaEventSynchronize(s->throttle[(it - kAhead) % 8]));
    return B200_OK;
}
static int throttle_mark(b200_sys* s, int it)
{
    CK(s->ctx, cudaEventRecord(s->throttle[it % 8], s->ctx->stream));
    return B200_OK;
}
// After solve_head: with several ranks wait for the head, so that "converged before the first iteration" is seen by
// every host (the loop is then skipped everywhere).
static int loop_enter(b200_sys* s, bool* skip)
{
    *skip = false;
    if (s->ctx->nranks > 1)
    {
        CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
        *skip = *(volatile int*)&s->hostFlags[2] != 0;
    }
    return B200_OK;
}
// top of a solver-loop iteration (after throttle_wait): true when this host has to leave the loop
static bool loop_leave(const b200_sys* s, int it)
{
    if (s->ctx->nranks <= 1) return *(volatile int*)&s->hostFlags[0] != 0;
    const int f = *(volatile int*)&s->hostFlags[2];
    return f != 0 && it >= (f - 2) + throttle_ahead(s);
}

// PBiCG::solve (foam/matrices/lduMatrix/solvers/PBiCG/PBiCG.C): PCG's scalar recurrences (rho = (wA, rT),
// alpha = rho / (wA, pT)) over a primal and a shadow (transposed) system: Tmul and preconditionT next to Amul and
// precondition.
static int solve_pbicg(b200_sys* s, const b200_solver_opts* o)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)s->nSlots;
    double *x = s->vec[V_X].p, *rA = s->vec[V_R].p, *rT = s->vec[V_RW].p, *pA = s->vec[V_P].p, *pT = s->vec[V_PH].p;
    double *wA = s->vec[V_V].p, *wT = s->vec[V_T].p, *zA = s->vec[V_S].p, *zT = s->vec[V_SH].p, *tmp = s->vec[V_TMP2].p;
    int rc;
    if ((rc = solve_head(s, OP_NORM_INIT_PCG, wA, rA, nullptr, pA, pT))) return rc;
    if ((rc = launch_amul(s, x, wT, 0, nullptr, true, 1, nullptr))) return rc; // wT = A^T x
    if (n)
    {
        KScope k(s, B200_K_VECTOR);
        k_sub<<<s->vecBlocks, 256, 0, st>>>(n, s->vec[V_B].p, wT, rT);
    }
    if ((rc = ensure_precond(s, o->precond, false))) return rc;
    if ((rc = ensure_precond(s, o->precond, true))) return rc;
    PartCounts cnt;
    bool skipLoop;
    if ((rc = loop_enter(s, &skipLoop))) return rc;
    for (int it = 0; it < o->maxIter && !skipLoop; it++)
    {
        if ((rc = throttle_wait(s, it))) return rc;
        if (loop_leave(s, it)) break;
        if ((rc = launch_precondition(s, o->precond, rA, zA, tmp, false, 0, false))) return rc; // wA = M^-1 rA
        if ((rc = launch_precondition(s, o->precond, rT, zT, tmp, false, 0, true))) return rc;  // wT = M^-T rT
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_dot<<<s->vecBlocks, 256, 0, st>>>(n, zA, rT, s->partials.p, s->pstride, s->sc.p); // wArT
        }
        if ((rc = reduce_finish(s, vec_counts(s), 1, OP_PCG_RHO, 0))) return rc;
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_pbicg_p<<<s->vecBlocks, 256, 0, st>>>(n, zA, pA, zT, pT, s->sc.p);
        }
        if ((rc = launch_amul(s, pT, wT, 0, nullptr, true, 0, nullptr))) return rc; // wT = A^T pT
        if ((rc = launch_amul(s, pA, wA, 1, pT, false, 0, &cnt))) return rc;       // wA = A pA, wApT
        if ((rc = reduce_finish(s, cnt, 1, OP_PCG_ALPHA, 0))) return rc;
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_pbicg_xr<<<s->vecBlocks, 256, 0, st>>>(n, x, pA, rA, wA, rT, wT, s->partials.p, s->pstride, s->sc.p);
        }
        if ((rc = reduce_finish(s, vec_counts(s), 1, OP_PCG_RESIDUAL, 0))) return rc;
        if ((rc = throttle_mark(s, it))) return rc;
    }
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

static int solve_bicgstab(b200_sys* s, const b200_solver_opts* o)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)s->nSlots;
    double *x = s->vec[V_X].p, *r = s->vec[V_R].p, *rw = s->vec[V_RW].p, *p = s->vec[V_P].p, *ph = s->vec[V_PH].p;
    double *v = s->vec[V_V].p, *sv = s->vec[V_S].p, *sh = s->vec[V_SH].p, *t = s->vec[V_T].p, *tmp = s->vec[V_TMP2].p;
    int rc;
    // p doubles as the Amul helper of the head (bicgStabSolver: "Multiplication helper p")
    if ((rc = solve_head(s, OP_NORM_INIT_BICGSTAB, p, r, rw, nullptr, v))) return rc;
    if (n)
    {
        KScope k(s, B200_K_VECTOR);
        k_fill<<<s->vecBlocks, 256, 0, st>>>(n, p, 0.0); // p = 0
    }
    if ((rc = ensure_precond(s, o->precond, false))) return rc;
    const bool sweeps = o->precond >= B200_PRECOND_DIC;
    PartCounts cnt;
    bool skipLoop;
    if ((rc = loop_enter(s, &skipLoop))) return rc;
    for (int it = 0; it < o->maxIter && !skipLoop; it++)
    {
        if ((rc = throttle_wait(s, it))) return rc;
        if (loop_leave(s, it)) break;
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_bicg_p<<<s->vecBlocks, 256, 0, st>>>(n, r, p, v, rw, sweeps ? tmp : nullptr, sweeps ? ph : nullptr, s->partials.p,
                                                   s->pstride, s->sc.p);
        }
        if ((rc = launch_precondition(s, o->precond, p, ph, tmp, true, 0))) return rc;
        if ((rc = launch_amul(s, ph, v, 1, rw, false, 0, &cnt))) return rc;
        cnt.n[1] = n ? s->vecBlocks : 0; // quantity 1 = (r, r) from k_bicg_p
        if ((rc = reduce_finish(s, cnt, 2, OP_BICG_ALPHA, 0))) return rc;
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_bicg_s<<<s->vecBlocks, 256, 0, st>>>(n, r, v, sv, sweeps ? tmp : nullptr, sweeps ? sh : nullptr, s->sc.p);
        }
        if ((rc = launch_precondition(s, o->precond, sv, sh, tmp, true, 0))) return rc;
        if ((rc = launch_amul(s, sh, t, 2, sv, false, 0, &cnt))) return rc;
        if ((rc = reduce_finish(s, cnt, 2, OP_BICG_OMEGA, 0))) return rc;
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_bicg_xr<<<s->vecBlocks, 256, 0, st>>>(n, x, ph, sh, sv, t, r, rw, s->partials.p, s->pstride, s->sc.p);
        }
        if ((rc = reduce_finish(s, vec_counts(s), 2, OP_BICG_RESIDUAL, 0))) return rc;
        if ((rc = throttle_mark(s, it))) return rc;
    }
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

static int solve_pcg(b200_sys* s, const b200_solver_opts* o)
{
    b200_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)s->nSlots;
    double *x = s->vec[V_X].p, *rA = s->vec[V_R].p, *pA = s->vec[V_P].p, *wA = s->vec[V_V].p, *z = s->vec[V_PH].p;
    double* tmp = s->vec[V_TMP2].p;
    int rc;
    if ((rc = solve_head(s, OP_NORM_INIT_PCG, wA, rA, nullptr, pA, nullptr))) return rc;
    if ((rc = ensure_precond(s, o->precond, false))) return rc;
    const bool sweeps = o->precond >= B200_PRECOND_DIC;
    if (sweeps && (rc = fill_sentinel(s, tmp, z, 1))) return rc;
    PartCounts cnt;
    bool skipLoop;
    if ((rc = loop_enter(s, &skipLoop))) return rc;
    for (int it = 0; it < o->maxIter && !skipLoop; it++)
    {
        if ((rc = throttle_wait(s, it))) return rc;
        if (loop_leave(s, it)) break;
        if ((rc = launch_precondition(s, o->precond, rA, z, tmp, true, 0))) return rc; // wA = M^-1 rA
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_dot<<<s->vecBlocks, 256, 0, st>>>(n, z, rA, s->partials.p, s->pstride, s->sc.p); // wArA
        }
        if ((rc = reduce_finish(s, vec_counts(s), 1, OP_PCG_RHO, 0))) return rc;
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_pcg_p<<<s->vecBlocks, 256, 0, st>>>(n, z, pA, s->sc.p);
        }
        if ((rc = launch_amul(s, pA, wA, 1, pA, false, 0, &cnt))) return rc; // wA = A pA, wApA
        if ((rc = reduce_finish(s, cnt, 1, OP_PCG_ALPHA, 0))) return rc;
        if (n)
        {
            KScope k(s, B200_K_VECTOR);
            k_pcg_xr<<<s->vecBlocks, 256, 0, st>>>(n, x, pA, rA, wA, sweeps ? tmp : nullptr, sweeps ? z : nullptr, s->partials.p,
                                                   s->pstride, s->sc.p);
        }
        if ((rc = reduce_finish(s, vec_counts(s), 1, OP_PCG_RESIDUAL, 0))) return rc;
        if ((rc = throttle_mark(s, it))) return rc;
    }
    CK(ctx, cudaGetLastError());
    return B200_OK;
}

extern "C" int b200_solve_resident(b200_sys* s, const b200_solver_opts* o, b200_perf* perf, double* history, int historyCap)
{
    if (!s || !o || !perf) return s ? set_err(s->ctx, B200_EINVAL, "b200_solve: null argument") : B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    if (!s->finalized) return set_err(ctx, B200_ESTATE, "solve before finalize");
    if (o->solver != B200_SOLVER_PCG && o->solver != B200_SOLVER_BICGSTAB && o->solver != B200_SOLVER_PBICG)
        return set_err(ctx, B200_EINVAL, "unknown solver %d", o->solver);
    if (o->precond < B200_PRECOND_NONE || o->precond > B200_PRECOND_CHOLESKY)
        return set_err(ctx, B200_EINVAL, "unknown preconditioner %d", o->precond);
    if (o->maxIter < 0 || o->minIter < 0) return set_err(ctx, B200_EINVAL, "negative iteration bounds");
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_sell(s, false);
    if (rc) return rc;
    // like the reference, every solve constructs its preconditioner from the current coefficients
    s->precondValid = -1;
    if ((rc = upload_scalars(s, o, history ? historyCap : 0))) return rc;
    CK(ctx, cudaEventRecord(s->evSolveA, ctx->stream));
    rc = (o->solver == B200_SOLVER_PCG) ? solve_pcg(s, o) : (o->solver == B200_SOLVER_PBICG) ? solve_pbicg(s, o) : solve_bicgstab(s, o);
    if (rc) return rc;
    CK(ctx, cudaEventRecord(s->evSolveB, ctx->stream));
    DevScalars h;
    CK(ctx, cudaMemcpyAsync(&h, s->sc.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(ctx, cudaEventElapsedTime(&ms, s->evSolveA, s->evSolveB));
    perf->initialResidual = h.initialResidual;
    perf->finalResidual = h.finalResidual;
    perf->nIterations = h.nIter;
    perf->converged = h.converged;
    perf->singular = h.singular;
    perf->normFactor = h.normFactor;
    perf->deviceMs = ms;
    if (history && historyCap > 0)
    {
        int nh = std::min({historyCap, h.nIter + 1, (int)kHistOnDevice});
        CK(ctx, cudaMemcpy(history, s->history.p, sizeof(double) * nh, cudaMemcpyDeviceToHost));
    }
    if (s->profiling) harvest_events(s);
    return check_device_error(s);
}

extern "C" int b200_solve(b200_sys* s, const b200_solver_opts* o, double* const* x, const double* const* b, b200_perf* perf,
                          double* history, int historyCap)
{
    if (!s || !x || !b) return s ? set_err(s->ctx, B200_EINVAL, "b200_solve: null argument") : B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "solve before finalize");
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc;
    if ((rc = upload_vec(s, V_X, x))) return rc;
    if ((rc = upload_vec(s, V_B, b))) return rc;
    if ((rc = b200_solve_resident(s, o, perf, history, historyCap))) return rc;
    if ((rc = download_vec(s, s->vec[V_X].p, x))) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ test hooks
extern "C" int b200_amul(b200_sys* s, const double* const* x, double* const* y, int transpose)
{
    if (!s || !x || !y) return B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "amul before finalize");
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc;
    if ((rc = upload_vec(s, V_P, x))) return rc;
    if ((rc = launch_amul(s, s->vec[V_P].p, s->vec[V_V].p, 0, nullptr, transpose != 0, 1, nullptr))) return rc;
    if ((rc = download_vec(s, s->vec[V_V].p, y))) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    if (s->profiling) harvest_events(s);
    return B200_OK;
}

// lduMatrix::residual / lduMatrix::sumA: the operations a smoother or the GAMG agglomeration asks of the matrix
extern "C" int b200_residual(b200_sys* s, const double* const* x, const double* const* b, double* const* r)
{
    if (!s || !x || !b || !r) return B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "residual before finalize");
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc;
    if ((rc = upload_vec(s, V_P, x))) return rc;
    if ((rc = upload_vec(s, V_S, b))) return rc;
    if ((rc = launch_amul(s, s->vec[V_P].p, s->vec[V_V].p, 0, s->vec[V_S].p, false, 1, nullptr, 1))) return rc;
    if ((rc = download_vec(s, s->vec[V_V].p, r))) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    if (s->profiling) harvest_events(s);
    return B200_OK;
}

extern "C" int b200_sum_a(b200_sys* s, double* const* sumA)
{
    if (!s || !sumA) return B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "sumA before finalize");
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc;
    // x is not read by the kernels of this mode, but the halo exchange of launch_amul still runs (and pairs up across
    // ranks); V_P holds whatever the last operation left there
    if ((rc = launch_amul(s, s->vec[V_P].p, s->vec[V_V].p, 0, nullptr, false, 1, nullptr, 2))) return rc;
    if ((rc = download_vec(s, s->vec[V_V].p, sumA))) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    if (s->profiling) harvest_events(s);
    return B200_OK;
}

extern "C" int b200_precondition(b200_sys* s, int precond, const double* const* r, double* const* w, int transpose)
{
    if (!s || !r || !w) return B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "precondition before finalize");
    if (precond < B200_PRECOND_NONE || precond > B200_PRECOND_CHOLESKY) return set_err(s->ctx, B200_EINVAL, "unknown preconditioner");
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc;
    for (size_t q = 0; q < s->regs.size(); q++)
        if (!s->regionHasCoeffs[q]) return set_err(s->ctx, B200_ESTATE, "region %zu has no coefficients", q);
    if ((rc = upload_vec(s, V_S, r))) return rc;
    if ((rc = ensure_precond(s, precond, transpose != 0))) return rc;
    if ((rc = launch_precondition(s, precond, s->vec[V_S].p, s->vec[V_SH].p, s->vec[V_TMP2].p, false, 1, transpose != 0))) return rc;
    if ((rc = download_vec(s, s->vec[V_SH].p, w))) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    if (s->profiling) harvest_events(s);
    return check_device_error(s);
}

// DICSmoother / DILUSmoother::smooth (and foam-extend's ILU smoother with the Cholesky-type preconditioner): nSweeps times
//   rA = residual(psi, source);  rA = M^-1 rA;  psi += rA        (include/b200_ldu.h)
extern "C" int b200_smooth(b200_sys* s, int precond, int nSweeps, double* const* x, const double* const* b)
{
    if (!s || !x || !b || nSweeps < 0) return B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "smooth before finalize");
    if (precond < B200_PRECOND_DIAGONAL || precond > B200_PRECOND_CHOLESKY) return set_err(s->ctx, B200_EINVAL, "b200_smooth: no such smoother (preconditioner %d)", precond);
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    for (size_t q = 0; q < s->regs.size(); q++)
        if (!s->regionHasCoeffs[q]) return set_err(ctx, B200_ESTATE, "region %zu has no coefficients", q);
    int rc;
    if ((rc = upload_vec(s, V_P, x))) return rc;
    if ((rc = upload_vec(s, V_S, b))) return rc;
    if ((rc = ensure_precond(s, precond, false))) return rc;
    for (int sweep = 0; sweep < nSweeps; sweep++)
    {
        if ((rc = launch_amul(s, s->vec[V_P].p, s->vec[V_V].p, 0, s->vec[V_S].p, false, 1, nullptr, 1))) return rc; // lduMatrix::residual
        if ((rc = launch_precondition(s, precond, s->vec[V_V].p, s->vec[V_SH].p, s->vec[V_TMP2].p, false, 1, false))) return rc;
        if (s->nSlots)
        {
            KScope k(s, B200_K_VECTOR);
            k_add_inplace<<<s->vecBlocks, 256, 0, ctx->stream>>>((size_t)s->nSlots, s->vec[V_P].p, s->vec[V_SH].p);
            CK(ctx, cudaGetLastError());
        }
    }
    if ((rc = download_vec(s, s->vec[V_P].p, x))) return rc;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (s->profiling) harvest_events(s);
    return check_device_error(s);
}

extern "C" int b200_get_rD(b200_sys* s, int precond, double* const* rD)
{
    if (!s || !rD) return B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "get_rD before finalize");
    if (precond < B200_PRECOND_DIAGONAL || precond > B200_PRECOND_CHOLESKY) return set_err(s->ctx, B200_EINVAL, "preconditioner has no rD");
    CK(s->ctx, cudaSetDevice(s->ctx->device));
    int rc;
    if ((rc = ensure_precond(s, precond, false))) return rc;
    if ((rc = download_vec(s, s->rD.p, rD))) return rc;
    CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
    return check_device_error(s);
}

extern "C" int b200_reduce(b200_sys* s, const double* const* a, const double* const* b, double* out2)
{
    if (!s || !a || !b || !out2) return B200_EINVAL;
    if (!s->finalized) return set_err(s->ctx, B200_ESTATE, "reduce before finalize");
    b200_ctx* ctx = s->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_vec(s, V_S, a))) return rc;
    if ((rc = upload_vec(s, V_SH, b))) return rc;
    if (s->nSlots)
    {
        KScope k(s, B200_K_VECTOR);
        k_dot_mag<<<s->vecBlocks, 256, 0, ctx->stream>>>((size_t)s->nSlots, s->vec[V_S].p, s->vec[V_SH].p, s->partials.p, s->pstride);
    }
    if ((rc = reduce_finish(s, vec_counts(s), 2, OP_STORE_RED, 1))) return rc;
    DevScalars h;
    CK(ctx, cudaMemcpyAsync(&h, s->sc.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    out2[0] = h.red[0];
    out2[1] = h.red[1];
    return B200_OK;
}

// ------------------------------------------------------------------------------------------ profiling
extern "C" int b200_debug_sweep_stats(b200_sys* s, int dir, int enable, long long* out, int cap)
{
    if (!s || !s->finalized) return B200_EINVAL;
    b200_ctx* ctx = s->ctx;
    PipeDirMem& M = dir > 0 ? s->fwd : s->bwd;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t n = (size_t)kStatsStride * (size_t)s->nGroups;
    if (out && M.stats.p)
        CK(ctx, cudaMemcpy(out, M.stats.p, sizeof(long long) * std::min<size_t>(n, (size_t)cap), cudaMemcpyDeviceToHost));
    if (enable)
    {
        if (M.stats.n != n) CK(ctx, M.stats.alloc(n));
        CK(ctx, cudaMemset(M.stats.p, 0, sizeof(long long) * std::max<size_t>(n, 1)));
        M.dev.stats = M.stats.p;
    }
    else
        M.dev.stats = nullptr;
    return s->nGroups;
}

extern "C" int b200_set_profiling(b200_sys* s, int enable)
{
    if (!s) return B200_EINVAL;
    cudaStreamSynchronize(s->ctx->stream);
    harvest_events(s);
    s->profiling = enable != 0;
    return B200_OK;
}

extern "C" int b200_get_kernel_times(b200_sys* s, double* msPerClass, int64_t* launchesPerClass, int reset)
{
    if (!s) return B200_EINVAL;
    cudaStreamSynchronize(s->ctx->stream);
    harvest_events(s);
    for (int k = 0; k < B200_K_NCLASSES; k++)
    {
        if (msPerClass) msPerClass[k] = s->clsMs[k];
        if (launchesPerClass) launchesPerClass[k] = s->clsLaunches[k];
        if (reset)
        {
            s->clsMs[k] = 0;
            s->clsLaunches[k] = 0;
        }
    }
    return B200_OK;
}

#include "ggi_build.cuh"

// ------------------------------------------------------------------------------------------ partitioned face transfer
extern "C" int b200_ggi_interpolate(b200_ctx* ctx, int32_t nTo, int32_t nFrom, const int32_t* offsets, const int32_t* addr,
                                    const double* weights, const double* ff, int nComp, double* result)
{
    if (!ctx || nTo < 0 || nFrom < 0 || nComp < 1 || !offsets || !result) return set_err(ctx, B200_EINVAL, "b200_ggi_interpolate: bad arguments");
    CK(ctx, cudaSetDevice(ctx->device));
    if (nTo == 0) return B200_OK;
    const int nnz = offsets[nTo];
    for (int k = 0; k < nnz; k++)
        if (addr[k] < 0 || addr[k] >= nFrom) return set_err(ctx, B200_EINVAL, "GGI address %d out of range", addr[k]);
    DevBuf<int> dOff, dAddr;
    DevBuf<double> dW, dF, dR;
    cudaStream_t st = ctx->stream;
    CK(ctx, dOff.alloc(nTo + 1));
    CK(ctx, dAddr.alloc(nnz));
    CK(ctx, dW.alloc(nnz));
    CK(ctx, dF.alloc((size_t)nFrom * nComp));
    CK(ctx, dR.alloc((size_t)nTo * nComp));
    CK(ctx, cudaMemcpyAsync(dOff.p, offsets, sizeof(int) * (nTo + 1), cudaMemcpyHostToDevice, st));
    if (nnz)
    {
        CK(ctx, cudaMemcpyAsync(dAddr.p, addr, sizeof(int) * nnz, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(dW.p, weights, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
    }
    if (nFrom) CK(ctx, cudaMemcpyAsync(dF.p, ff, sizeof(double) * nFrom * nComp, cudaMemcpyHostToDevice, st));
    ctx->launches++;
    k_ggi_interpolate<<<(nTo * nComp + 127) / 128, 128, 0, st>>>(nTo, dOff.p, dAddr.p, dW.p, dF.p, nComp, dR.p);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(result, dR.p, sizeof(double) * nTo * nComp, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return B200_OK;
}

extern "C" int b200_patch_face_to_global(b200_ctx* ctx, int32_t nLocal, const int32_t* faceToGlobalAddr, const double* pField,
                                         int nComp, int32_t nZoneFaces, double* gField)
{
    if (!ctx || nLocal < 0 || nZoneFaces < 0 || nComp < 1 || !gField) return set_err(ctx, B200_EINVAL, "b200_patch_face_to_global: bad arguments");
    CK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < nLocal; i++)
        if (faceToGlobalAddr[i] < 0 || faceToGlobalAddr[i] >= nZoneFaces) return set_err(ctx, B200_EINVAL, "faceToGlobalAddr out of range");
    DevBuf<int> dAddr;
    DevBuf<double> dP, dG;
    cudaStream_t st = ctx->stream;
    CK(ctx, dAddr.alloc(nLocal));
    CK(ctx, dP.alloc((size_t)nLocal * nComp));
    CK(ctx, dG.alloc((size_t)nZoneFaces * nComp));
    CK(ctx, cudaMemsetAsync(dG.p, 0, sizeof(double) * std::max<size_t>(1, (size_t)nZoneFaces * nComp), st));
    if (nLocal)
    {
        CK(ctx, cudaMemcpyAsync(dAddr.p, faceToGlobalAddr, sizeof(int) * nLocal, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(dP.p, pField, sizeof(double) * nLocal * nComp, cudaMemcpyHostToDevice, st));
        ctx->launches++;
        k_scatter_zone<<<(nLocal * nComp + 127) / 128, 128, 0, st>>>(nLocal, dAddr.p, dP.p, nComp, dG.p);
        CK(ctx, cudaGetLastError());
    }
    if (ctx->nranks > 1 && nZoneFaces > 0)
    {
        if (ctx->comm)
            NK(ctx, g_nccl.AllReduce(dG.p, dG.p, (size_t)nZoneFaces * nComp, ncclDouble, ncclSum, ctx->comm, st)); // reduce(gField, sumOp)
        else
        { // no NCCL communicator (B200_TRANSPORT=p2p): chunks of the zone array through the peers' exchange areas
            PeerLink& L = ctx->peer;
            const size_t total = (size_t)nZoneFaces * nComp;
            for (size_t o = 0; o < total; o += kXchgCap)
            {
                const int n = (int)std::min<size_t>(kXchgCap, total - o);
                const unsigned long long seq = ++L.xSeq;
                const int par = (int)(seq & 1ull);
                ctx->launches += 2;
                k_xchg_push<<<std::min(64, (n + 255) / 256), 256, 0, st>>>(n, dG.p + o, L.dPeerX, ctx->rank, ctx->nranks, kXchgCap, par, seq, L.counter);
                k_xchg_sum<<<std::min(64, (n + 255) / 256), 256, 0, st>>>(n, dG.p + o, L.xbuf, ctx->nranks, kXchgCap, par, seq, L.err);
                CK(ctx, cudaGetLastError());
            }
        }
    }
    if (nZoneFaces) CK(ctx, cudaMemcpyAsync(gField, dG.p, sizeof(double) * nZoneFaces * nComp, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return B200_OK;
}

extern "C" int b200_global_face_to_patch(b200_ctx* ctx, int32_t nLocal, const int32_t* faceToGlobalAddr, const double* gField,
                                         int nComp, double* pField)
{
    if (!ctx || nLocal < 0 || nComp < 1) return set_err(ctx, B200_EINVAL, "b200_global_face_to_patch: bad arguments");
    CK(ctx, cudaSetDevice(ctx->device));
    if (nLocal == 0) return B200_OK;
    int nZone = 0;
    for (int i = 0; i < nLocal; i++)
    {
        if (faceToGlobalAddr[i] < 0) return set_err(ctx, B200_EINVAL, "faceToGlobalAddr out of range");
        nZone = std::max(nZone, faceToGlobalAddr[i] + 1);
    }
    DevBuf<int> dAddr;
    DevBuf<double> dP, dG;
    cudaStream_t st = ctx->stream;
    CK(ctx, dAddr.alloc(nLocal));
    CK(ctx, dP.alloc((size_t)nLocal * nComp));
    CK(ctx, dG.alloc((size_t)nZone * nComp));
    CK(ctx, cudaMemcpyAsync(dAddr.p, faceToGlobalAddr, sizeof(int) * nLocal, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(dG.p, gField, sizeof(double) * nZone * nComp, cudaMemcpyHostToDevice, st));
    ctx->launches++;
    k_gather_zone<<<(nLocal * nComp + 127) / 128, 128, 0, st>>>(nLocal, dAddr.p, dG.p, nComp, dP.p);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(pField, dP.p, sizeof(double) * nLocal * nComp, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return B200_OK;
}

#include "direct_map.cuh"

// block-coupled (vector4) systems: include/b200_blk.h
#include "blk_system.cuh"
#include "gs_smoother.cuh"
