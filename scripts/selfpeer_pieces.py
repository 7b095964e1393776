"""The CHT case decomposed with ``simple; n (2 1 2)`` per region (the decomposition the reference ships:
tutorials/conjugateHeatTransfer/flowOverHeatedPlate/system/*/decomposeParDict) on 4 processes that share cuda:0
(B200_TRANSPORT=p2p): fluid and plate are cut at different x, so pieces of the regionCouple pair face other ranks and the
interface is interpolated on the global zones (b200_sys_set_interface_pieces).  Against the oracle on the same decomposition.

    python scripts/selfpeer_pieces.py <workdir> <rank> <nranks> <uid hex> [nx ny nz]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200_TRANSPORT"] = "p2p"
import faulthandler

import numpy as np

faulthandler.enable()
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case
from multiregionfoam_b200.decompose import decompose_cht_simple


def wait_for(path, timeout=300.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.01)
    return path


def main():
    work, rank, world, uid = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), bytes.fromhex(sys.argv[4])
    n = tuple(int(a) for a in sys.argv[5:8]) if len(sys.argv) >= 8 else (2, 1, 2)
    assert int(np.prod(n)) == world
    case, fluid, solid = cht_case(1, 4)
    dec = decompose_cht_simple(case, fluid, solid, n)
    rs = dec.ranks[rank]
    ctx = ldu.Context(0, rank, world, uid)
    S = ldu.LduSystem(ctx, rs)
    x0 = np.concatenate([g.psi for g in rs.regions])
    b = np.concatenate([g.source for g in rs.regions])
    xr = np.random.default_rng(100 + rank).standard_normal(x0.size) * 10 + 300
    out = dict(xr=xr, y=S.amul(xr))
    xs, info = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=1e-12, maxIter=400)
    out.update(xs=xs, hist=info["history"], nIter=info["nIterations"])
    np.savez(os.path.join(work, f"rank{rank}.tmp.npz"), **out)
    os.rename(os.path.join(work, f"rank{rank}.tmp.npz"), os.path.join(work, f"rank{rank}.npz"))
    ok = True
    if rank == 0:
        from oracle import pyoracle
        G = [np.load(wait_for(os.path.join(work, f"rank{g}.npz"))) for g in range(world)]
        O = pyoracle.OracleSystem(dec)
        cat = lambda k: np.concatenate([g[k] for g in G])
        remote = sum(1 for rk in dec.ranks for reg in rk.regions for itf in reg.interfaces
                     for p in (getattr(itf, "pieces", None) or []) if p[0] != rk.rank)
        amul_exact = bool(np.array_equal(cat("y"), O.amul(cat("xr"))))
        xo, io = O.solve(dec.concat("psi"), dec.concat("source"), "BiCGStab", "DILU", tolerance=1e-12, maxIter=400)
        hg, ho = G[0]["hist"], io["history"]
        k = min(21, hg.size, ho.size)
        herr = float(np.max(np.abs(hg[:k] - ho[:k]) / np.maximum(np.abs(ho[:k]), 1e-300)))
        ferr = float(np.linalg.norm(cat("xs") - xo) / np.linalg.norm(xo))
        same_it = all(int(g["nIter"]) == int(G[0]["nIter"]) for g in G)
        ok = amul_exact and herr < 1e-10 and ferr < 1e-8 and same_it and remote > 0
        print(f"remote_pieces={remote} amul_bit_exact={amul_exact} iterations gpu {int(G[0]['nIter'])} oracle {io['nIterations']} "
              f"history_rel_err={herr:.2e} field_rel_err={ferr:.2e} same_iterations_on_all_ranks={same_it}")
        print(f"pieces_selfpeer_ok={ok}")
    S.close()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
