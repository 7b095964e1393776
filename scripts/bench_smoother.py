"""Timing of the device Gauss-Seidel smoother (include/b200_smooth.h) on the fluid region of a CHT workload:
smoothSolver with nSweeps sweeps per residual check, device time from b200_perf.deviceMs.

    python scripts/bench_smoother.py [workload=C2] [rounds=10] [nSweeps=2]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from multiregionfoam_b200 import ldu, smoother
from multiregionfoam_b200.assembly import WORKLOADS, cht_rank_slab


def main():
    W = sys.argv[1] if len(sys.argv) > 1 else "C2"
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    nSweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    r, L = WORKLOADS[W]
    reg = cht_rank_slab(r, L, 0, 1).regions[0]
    ctx = ldu.Context(0)
    G = smoother.GaussSeidel(ctx, reg.lowerAddr, reg.upperAddr, reg.nCells)
    G.set_coeffs(reg.diag, reg.upper, reg.lower)
    G.solve(reg.psi, reg.source, nSweeps, tolerance=0.0, minIter=nSweeps, maxIter=nSweeps)   # warm-up
    x, info = G.solve(reg.psi, reg.source, nSweeps, tolerance=0.0, minIter=rounds * nSweeps, maxIter=rounds * nSweeps)
    sweeps = info["nIterations"]
    ms = info["deviceMs"]
    nF = int(reg.lowerAddr.size)
    # per sweep: diag, b, old, new (8 B each) + per face 2 x (8 B coefficient + 4 B index) + per round one residual pass of the same size
    bytes_sweep = 32 * reg.nCells + 24 * nF
    print(json.dumps(dict(workload=W, region="fluid", cells=reg.nCells, faces=nF, sweeps=sweeps, rounds=rounds, device_ms=ms,
                          ms_per_round=ms / rounds, cell_sweeps_per_s=reg.nCells * sweeps / (ms * 1e-3),
                          approx_gbs=(bytes_sweep * (sweeps + rounds)) / (ms * 1e-3) / 1e9, initialResidual=info["initialResidual"],
                          finalResidual=info["finalResidual"])))
    G.close()
    ctx.close()


if __name__ == "__main__":
    main()
