"""Multi-GPU parity check (run under torchrun, one rank per GPU):
the z-slab decomposed CHT case solved by all ranks through libb200ldu (NCCL halo exchange + all-reduce)
against the CPU oracle run on the SAME decomposition (block-Jacobi preconditioner semantics of foam-extend).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/multirank_parity.py [r] [layers_per_rank]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_rank_slab
from multiregionfoam_b200.case import Case


def main():
    r = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t = torch.frombuffer(bytearray(ldu.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    ctx = ldu.Context(local, rank, world, bytes(t.cpu().numpy().tobytes()))
    rs = cht_rank_slab(r, L, rank, world)
    S = ldu.LduSystem(ctx, rs)
    x0 = np.concatenate([g.psi for g in rs.regions])
    b = np.concatenate([g.source for g in rs.regions])
    rng = np.random.default_rng(100 + rank)
    xr = rng.standard_normal(x0.size)
    y = S.amul(xr)
    xs, info = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=1e-12, maxIter=300)
    xp, infop = S.solve(x0, b, ldu.SOLVER_PCG, ldu.PRECOND_DIC, tolerance=1e-30, maxIter=3) if False else (None, None)
    # gather everything on rank 0 and compare with the oracle on the same decomposition
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(xr=xr, y=y, xs=xs, hist=info["history"], nIter=info["nIterations"],
                                          nf=info["normFactor"]))
    ok = True
    if rank == 0:
        from oracle import pyoracle
        case = Case("slabs", [cht_rank_slab(r, L, g, world) for g in range(world)])
        O = pyoracle.OracleSystem(case)
        xr_all = np.concatenate([g["xr"] for g in gathered])
        y_all = np.concatenate([g["y"] for g in gathered])
        yo = O.amul(xr_all)
        amul_exact = bool(np.array_equal(y_all, yo))
        xo, io = O.solve(case.concat("psi"), case.concat("source"), "BiCGStab", "DILU", tolerance=1e-12, maxIter=300)
        xs_all = np.concatenate([g["xs"] for g in gathered])
        hg, ho = gathered[0]["hist"], io["history"]
        k = min(21, hg.size, ho.size)
        herr = float(np.max(np.abs(hg[:k] - ho[:k]) / np.maximum(np.abs(ho[:k]), 1e-300)))
        ferr = float(np.linalg.norm(xs_all - xo) / np.linalg.norm(xo))
        same_hist = all(np.array_equal(g["hist"], hg) for g in gathered)
        print(f"multirank parity: ranks={world} cells={case.nCells} amul_bit_exact={amul_exact} "
              f"hist_max_rel_err_first{k}={herr:.2e} field_rel_l2={ferr:.2e} its gpu/oracle={gathered[0]['nIter']}/{io['nIterations']} "
              f"hist_identical_on_all_ranks={same_hist}", flush=True)
        ok = amul_exact and herr < 1e-10 and ferr < 1e-8 and same_hist
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    S.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
