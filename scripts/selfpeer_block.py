"""Decomposed block-coupled (vector4) solve on `nranks` processes that share cuda:0 (B200_TRANSPORT=p2p, as
scripts/selfpeer_parity.py): processor patches of the block matrix (halo of vector4 values, coupleUpper product) and the
rank-ordered all-reduce inside the block reductions, against oracle/pyblk_multi.py on the same decomposition.

    python scripts/selfpeer_block.py <workdir> <rank> <nranks> <uid hex>

Launched by tests/test_gpu_zz_block_iface.py; rank 0 gathers the others' results from <workdir> and prints the verdict."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["B200_TRANSPORT"] = "p2p"
import faulthandler

import numpy as np

faulthandler.enable()
from block_helpers import box_addr, random_block_coeffs
from multiregionfoam_b200 import blockldu, ldu
from oracle.pyblk_multi import MultiBlockOracle, split_block_system


def wait_for(path, timeout=300.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.01)
    return path


def problem(world):
    n, l, u = box_addr(24, 18, 4 * world)
    diag, upper, lower = random_block_coeffs(n, l, u, 16, 16, False, seed=9)
    own = (np.arange(n) * world // n).astype(np.int64)
    own[np.random.default_rng(5).integers(0, n, n // 50)] = world - 1   # ragged subdomain boundaries, more patches
    subs, cells = split_block_system(n, l, u, diag, upper, lower, own)
    rng = np.random.default_rng(21)
    return subs, cells, rng.standard_normal((n, 4)), rng.standard_normal((n, 4)), rng.standard_normal((n, 4))


def main():
    work, rank, world, uid = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), bytes.fromhex(sys.argv[4])
    subs, cells, xr, x0, b = problem(world)
    ctx = ldu.Context(0, rank, world, uid)
    s = subs[rank]
    S = blockldu.BlockSystem(ctx, s["l"], s["u"], s["n"])
    S.set_coeffs(s["diag"], s["upper"], s["lower"])
    for I in s["ifaces"]:
        k = S.add_interface(I["faceCells"], I["peer"], I["peerIface"])
        S.set_interface_coeffs(k, I["coupleUpper"])
    c = cells[rank]
    out = dict(y=S.amul(xr[c]))
    prod, cm = S.reduce(xr[c], x0[c])
    out.update(prod=prod, cm=cm)
    for tag, solver, pre in (("bc", blockldu.SOLVER_BICGSTAB, ldu.PRECOND_CHOLESKY), ("bd", blockldu.SOLVER_BICGSTAB, ldu.PRECOND_DIAGONAL)):
        xs, info = S.solve(x0[c], b[c], solver, pre, tolerance=1e-11, maxIter=200)
        out.update({f"x_{tag}": xs, f"h_{tag}": info["history"], f"nf_{tag}": info["normFactor"], f"it_{tag}": info["nIterations"]})
    np.savez(os.path.join(work, f"rank{rank}.tmp.npz"), **out)
    os.rename(os.path.join(work, f"rank{rank}.tmp.npz"), os.path.join(work, f"rank{rank}.npz"))
    ok = True
    if rank == 0:
        G = [np.load(wait_for(os.path.join(work, f"rank{g}.npz"))) for g in range(world)]
        M = MultiBlockOracle(subs)
        yo = M.amul([xr[cc] for cc in cells])
        amul_exact = all(np.array_equal(G[g]["y"], yo[g]) for g in range(world))
        po = M.sumprod([xr[cc] for cc in cells], [x0[cc] for cc in cells])
        red_same = all(G[g]["prod"] == G[0]["prod"] and np.array_equal(G[g]["cm"], G[0]["cm"]) for g in range(world))
        red_err = abs(float(G[0]["prod"]) - po) / np.abs(xr * x0).sum()
        ok = amul_exact and red_same and red_err < 1e-13
        print(f"amul_bit_exact={amul_exact} reductions_identical_on_all_ranks={red_same} sumprod_rel_err={red_err:.2e}")
        for tag, solver, pre in (("bc", "BiCGStab", "Cholesky"), ("bd", "BiCGStab", "diagonal")):
            xo, po_ = M.solve([x0[cc] for cc in cells], [b[cc] for cc in cells], solver, pre, tolerance=1e-11, maxIter=200)
            hg, ho = G[0][f"h_{tag}"], po_["history"]
            k = min(21, hg.shape[0], ho.shape[0])
            herr = float(np.max(np.abs(hg[:k] - ho[:k]) / np.maximum(np.abs(ho[:k]), 1e-300)))
            xg = np.concatenate([G[g][f"x_{tag}"] for g in range(world)])
            xoo = np.concatenate(xo)
            ferr = float(np.linalg.norm(xg - xoo) / np.linalg.norm(xoo))
            nferr = abs(float(G[0][f"nf_{tag}"]) - po_["normFactor"]) / po_["normFactor"]
            same_it = all(int(G[g][f"it_{tag}"]) == int(G[0][f"it_{tag}"]) for g in range(world))
            print(f"{tag}: iterations gpu {int(G[0][f'it_{tag}'])} oracle {po_['nIterations']} history_rel_err={herr:.2e} field_rel_err={ferr:.2e} "
                  f"normFactor_rel_err={nferr:.2e} same_iterations_on_all_ranks={same_it}")
            ok = ok and herr < 1e-10 and ferr < 1e-8 and nferr < 1e-12 and same_it
        print(f"block_selfpeer_ok={ok}")
    S.close()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
