"""Small C1 run for compute-sanitizer (memcheck / racecheck): Amul, residual, DILU precondition (both sweeps), a few BiCGStab
iterations and a PCG solve on the as-shipped flowOverHeatedPlate mesh.
    compute-sanitizer --tool memcheck  python scripts/sanitize_c1.py
    compute-sanitizer --tool racecheck python scripts/sanitize_c1.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case

case = cht_case(1, 1)[0]
ctx = ldu.Context(0)
S = ldu.LduSystem(ctx, case.ranks[0])
x0, b = case.concat("psi"), case.concat("source")
y = S.amul(x0)
r = S.residual(x0, b)
w = S.precondition(ldu.PRECOND_DILU, r)
xs, info = S.solve(x0, b, ldu.SOLVER_BICGSTAB, ldu.PRECOND_DILU, tolerance=0.0, minIter=3, maxIter=3)
xp, ip = S.solve(x0, b, ldu.SOLVER_PBICG, ldu.PRECOND_DILU, tolerance=0.0, minIter=2, maxIter=2)
print("sanitize_c1: amul", float(np.abs(y).sum()), "precondition", float(np.abs(w).sum()), "residuals", info["finalResidual"], ip["finalResidual"])
S.close()
ctx.close()
