"""Multi-rank parity WITHOUT NCCL and WITHOUT a second GPU: `nranks` processes, all on cuda:0, talk through the
library's own peer-to-peer transport (B200_TRANSPORT=p2p: mailboxes and halo buffers shared with cudaIpc; on several
GPUs the very same kernels store over NVLink).  Covers the decomposed coupled solve (halo exchange of the processor
patches + fused all-reduce) against the CPU oracle on the SAME decomposition, the early-exit cases in which every
rank has to leave the solver loop at the same iteration, and the zone all-reduce of globalPolyPatch::patchFaceToGlobal.

    python scripts/selfpeer_parity.py <workdir> <rank> <nranks> <uid hex> [r] [layers per rank]

Launched by tests/test_gpu_zz_selfpeer.py (one process per rank); rank 0 gathers the others' results from <workdir>."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200_TRANSPORT"] = "p2p"
import faulthandler

import numpy as np

faulthandler.enable()
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_rank_slab
from multiregionfoam_b200.case import Case


def wait_for(path, timeout=300.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.01)
    return path


def main():
    work, rank, world, uid = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), bytes.fromhex(sys.argv[4])
    r = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    L = int(sys.argv[6]) if len(sys.argv) > 6 else 3
    ctx = ldu.Context(0, rank, world, uid)
    rs = cht_rank_slab(r, L, rank, world)
    S = ldu.LduSystem(ctx, rs)
    x0 = np.concatenate([g.psi for g in rs.regions])
    b = np.concatenate([g.source for g in rs.regions])
    xr = np.random.default_rng(100 + rank).standard_normal(x0.size)
    out = dict(xr=xr, y=S.amul(xr), res=S.residual(xr, b))
    xs, info = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=1e-12, maxIter=300)
    out.update(xs=xs, hist=info["history"], nIter=info["nIterations"])
    xp, ip = S.solve(x0, b, ldu.SOLVER_PCG, ldu.PRECOND_DIAGONAL, tolerance=0.0, minIter=5, maxIter=5)
    out.update(xp=xp, histp=ip["history"])
    # every rank has to leave the loop at the same iteration: converged before the first iteration, after
    # one, two, three ..., repeatedly, and the systems must stay usable afterwards
    early = []
    for rep in range(6):
        _, i0 = S.solve(xs, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=1e-6, maxIter=50)      # 0 iterations
        early.append(i0["nIterations"])
        for tol in (1e-1, 1e-2, 1e-3):
            _, ik = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=tol, maxIter=50)  # a few iterations
            early.append(ik["nIterations"])
    out["early"] = np.array(early)
    # zone all-reduce: every rank scatters its share of a 1000-face zone
    nZone = 1000
    mine = np.arange(rank, nZone, world, dtype=np.int32)
    pf = np.random.default_rng(7 + rank).random((mine.size, 3))
    out.update(zaddr=mine, zpf=pf, zone=ctx.patch_face_to_global(mine, pf, nZone))
    np.savez(os.path.join(work, f"rank{rank}.tmp.npz"), **out)
    os.rename(os.path.join(work, f"rank{rank}.tmp.npz"), os.path.join(work, f"rank{rank}.npz"))
    ok = True
    if rank == 0:
        from oracle import pyoracle
        G = [np.load(wait_for(os.path.join(work, f"rank{g}.npz"))) for g in range(world)]
        case = Case("slabs", [cht_rank_slab(r, L, g, world) for g in range(world)])
        O = pyoracle.OracleSystem(case)
        cat = lambda k: np.concatenate([g[k] for g in G])
        amul_exact = bool(np.array_equal(cat("y"), O.amul(cat("xr"))))
        res_exact = bool(np.array_equal(cat("res"), O.residual(cat("xr"), case.concat("source"))))
        xo, io = O.solve(case.concat("psi"), case.concat("source"), "BiCGStab", "DILU", tolerance=1e-12, maxIter=300)
        hg, ho = G[0]["hist"], io["history"]
        k = min(21, hg.size, ho.size)
        herr = float(np.max(np.abs(hg[:k] - ho[:k]) / np.maximum(np.abs(ho[:k]), 1e-300)))
        ferr = float(np.linalg.norm(cat("xs") - xo) / np.linalg.norm(xo))
        xpo, ipo = O.solve(case.concat("psi"), case.concat("source"), "PCG", "diagonal", tolerance=0.0, minIter=5, maxIter=5)
        perr = float(np.max(np.abs(G[0]["histp"][:6] - ipo["history"][:6]) / np.abs(ipo["history"][:6])))
        same = all(np.array_equal(g["hist"], hg) and np.array_equal(g["early"], G[0]["early"]) for g in G)
        early_ok = bool(np.all(G[0]["early"][0::4] == 0))
        po = np.concatenate([[0], np.cumsum([g["zaddr"].size for g in G])]).astype(np.int32)
        zo = pyoracle.patch_face_to_global(po, cat("zaddr"), np.concatenate([g["zpf"] for g in G]), nZone, 3)
        zone_ok = all(np.array_equal(g["zone"], zo) for g in G)
        print(f"selfpeer parity: ranks={world} on one device cells={case.nCells} amul_bit_exact={amul_exact} residual_bit_exact={res_exact} "
              f"hist_max_rel_err_first{k}={herr:.2e} field_rel_l2={ferr:.2e} its gpu/oracle={int(G[0]['nIter'])}/{io['nIterations']} "
              f"pcg_hist_err={perr:.2e} identical_on_all_ranks={same} early_exit_iterations={G[0]['early'][:4].tolist()} early_ok={early_ok} "
              f"zone_allreduce_bit_exact={zone_ok}", flush=True)
        ok = amul_exact and res_exact and herr < 1e-10 and ferr < 1e-8 and perr < 1e-10 and same and early_ok and zone_ok
        open(os.path.join(work, "verdict"), "w").write("ok" if ok else "fail")
    else:
        wait_for(os.path.join(work, "verdict"))
    S.close()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
