"""Small decomposed problems through the interface kernels added in round 2, for compute-sanitizer:
the CHT case cut with simple (2 1 2) and flattened onto one device (zone-piece interface plan, k_iface gathers over
pieces), a merged sweep schedule (B200_MERGE_MIN_GROUPS=0), and a block-coupled system with paired patches
(k_blk_iface).  Results are checked against the oracle, so a run that passes is also a parity run.

    compute-sanitizer --tool memcheck python scripts/sanitize_ifaces.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["B200_MERGE_MIN_GROUPS"] = "0"
import numpy as np

from multiregionfoam_b200 import blockldu, ldu
from multiregionfoam_b200.assembly import cht_case
from multiregionfoam_b200.decompose import decompose_cht_simple, flatten_ranks
from oracle import pyoracle
from oracle.pyblk_multi import MultiBlockOracle


def main():
    ctx = ldu.Context(0, 0, 1, None)
    case, fluid, solid = cht_case(1, 4)
    dec = decompose_cht_simple(case, fluid, solid, (2, 1, 2))
    flat = flatten_ranks(dec)
    O = pyoracle.OracleSystem(dec)
    S = ldu.LduSystem(ctx, flat.ranks[0])
    x0, b = dec.concat("psi"), dec.concat("source")
    v = np.random.default_rng(3).standard_normal(O.n) * 10 + 300
    ok = bool(np.array_equal(S.amul(v), O.amul(v)))
    xo, io = O.solve(x0, b, "BiCGStab", "DILU", tolerance=0.0, minIter=4, maxIter=4)
    xg, ig = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=0.0, minIter=4, maxIter=4)
    herr = float(np.max(np.abs(ig["history"][:5] - io["history"][:5]) / np.abs(io["history"][:5])))
    print(f"pieces: amul_bit_exact={ok} history_rel_err={herr:.2e}")
    ok = ok and herr < 1e-10
    S.close()

    from test_gpu_zz_block_iface import cut_system, device_system
    sub, (n, *_r) = cut_system((16, 16, False), seed=8)
    B, M = device_system(ctx, sub), MultiBlockOracle([sub])
    rng = np.random.default_rng(6)
    xb, bb = rng.standard_normal((n, 4)), rng.standard_normal((n, 4))
    ok2 = bool(np.array_equal(B.amul(xb), M.amul([xb])[0]))
    xg, ig = B.solve(xb, bb, blockldu.SOLVER_BICGSTAB, ldu.PRECOND_CHOLESKY, tolerance=0.0, minIter=3, maxIter=3)
    xo, io = M.solve([xb], [bb], "BiCGStab", "Cholesky", tolerance=0.0, minIter=3, maxIter=3)
    herr = float(np.max(np.abs(ig["history"][:4] - io["history"][:4]) / np.maximum(io["history"][:4], 1e-300)))
    print(f"block paired patches: amul_bit_exact={ok2} history_rel_err={herr:.2e}")
    ok = ok and ok2 and herr < 1e-10
    B.close()
    ctx.close()
    print("sanitize_ifaces_ok=" + str(ok))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
