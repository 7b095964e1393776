"""N > 1 host logic on CPU: two processes (torch.distributed, gloo) each own one z-slab of the CHT case,
build the device tables (slot permutation, SELL, interface plan, halo layout) with the production
csrc/schedule.hpp, exchange the packed halo values over gloo in the plan's peer order and apply
Amul + interface update.  The gathered result must equal the oracle's Amul of the same decomposed
system bit for bit -- i.e. what rank A packs is what rank B's plan expects."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORLD = 2


def _emu():
    from multiregionfoam_b200.build import build_schedule_emulator
    L = C.CDLL(build_schedule_emulator())
    ip, dp, vp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p
    L.emu_sys_create.restype = vp
    L.emu_sys_create.argtypes = [C.c_int, C.c_int]
    L.emu_sys_destroy.argtypes = [vp]
    L.emu_sys_set_region.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp]
    L.emu_sys_add_iface.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp, dp]
    L.emu_sys_finalize.argtypes = [vp]
    L.emu_sys_npeers.argtypes = [vp]
    L.emu_sys_peer.argtypes = [vp, C.c_int, ip, ip, ip, ip, ip]
    L.emu_sys_pack.argtypes = [vp, dp, dp]
    L.emu_sys_amul.argtypes = [vp, dp, dp, dp]
    return L


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from multiregionfoam_b200.assembly import cht_rank_slab
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _emu()
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    I = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rs = cht_rank_slab(1, 2, rank, world)
    h = L.emu_sys_create(len(rs.regions), rank)
    keep = []
    for r, reg in enumerate(rs.regions):
        l, u = np.ascontiguousarray(reg.lowerAddr, np.int32), np.ascontiguousarray(reg.upperAddr, np.int32)
        lo = None if reg.lower is None else P(np.ascontiguousarray(reg.lower))
        L.emu_sys_set_region(h, r, reg.nCells, reg.nFaces, I(l), I(u), P(reg.diag), P(reg.upper), lo)
    for r, reg in enumerate(rs.regions):
        for itf in reg.interfaces:
            fc = np.ascontiguousarray(itf.faceCells, np.int32)
            keep.append(fc)
            L.emu_sys_add_iface(h, r, itf.kind, itf.nFaces, I(fc), itf.peerRank, itf.peerRegion, itf.peerIface, itf.nFaces,
                                None, None, None, P(np.ascontiguousarray(itf.bouCoeffs)))
    assert L.emu_sys_finalize(h) == 0
    n = sum(reg.nCells for reg in rs.regions)
    x = np.random.default_rng(10 + rank).standard_normal(n)
    peers = []
    for p in range(L.emu_sys_npeers(h)):
        v = [C.c_int() for _ in range(5)]
        L.emu_sys_peer(h, p, *[C.byref(t) for t in v])
        peers.append(tuple(t.value for t in v))
    nSend = sum(p[2] for p in peers)
    nRecv = sum(p[4] for p in peers)
    send, recv = np.zeros(max(nSend, 1)), np.zeros(max(nRecv, 1))
    L.emu_sys_pack(h, P(x), P(send))
    # the exchange the library does with ncclSend/ncclRecv inside one group, here over gloo
    reqs = []
    for (pr, so, ns, ro, nr) in peers:
        if ns:
            reqs.append(dist.isend(torch.from_numpy(send[so:so + ns].copy()), pr))
    for (pr, so, ns, ro, nr) in peers:
        if nr:
            t = torch.zeros(nr, dtype=torch.float64)
            dist.recv(t, pr)
            recv[ro:ro + nr] = t.numpy()
    for rq in reqs:
        rq.wait()
    y = np.empty(n)
    L.emu_sys_amul(h, P(x), P(recv), P(y))
    out = [None] * world
    dist.all_gather_object(out, (x, y))
    # a global reduction like gSumProd: per-rank partial, all-reduce(sum)
    t = torch.tensor([float(np.dot(x, y))], dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        q.put((out, float(t.item())))
    L.emu_sys_destroy(h)
    dist.destroy_process_group()


def test_two_rank_halo_plan_matches_oracle():
    import torch.multiprocessing as mp
    from multiregionfoam_b200.assembly import cht_rank_slab
    from multiregionfoam_b200.case import Case
    from oracle import pyoracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    out, dot = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = Case("slabs", [cht_rank_slab(1, 2, g, WORLD) for g in range(WORLD)])
    O = pyoracle.OracleSystem(case)
    x = np.concatenate([o[0] for o in out])
    y = np.concatenate([o[1] for o in out])
    yo = O.amul(x)
    assert np.array_equal(y, yo)
    assert abs(dot - O.gsumprod(x, yo)) <= 1e-12 * np.abs(x * yo).sum()
