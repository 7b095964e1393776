"""Sweep kernel time per launch (CUDA events, no debug counters): C2 and structured boxes.
usage: python scripts/sweep_time.py [nx ny nz ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case, synthetic_coeffs, single_region_case
from multiregionfoam_b200.mesh import StructuredRegion, Block

ctx = ldu.Context(0)
args = [int(a) for a in sys.argv[1:]]
cases = [] if os.environ.get("NO_C2") else [("C2", cht_case(3, 22)[0])]
for i in range(0, len(args), 3):
    nx, ny, nz = args[i:i + 3]
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    cases.append((f"box {nx}x{ny}x{nz}", single_region_case(synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False))))
for name, case in cases:
    S = ldu.LduSystem(ctx, case.ranks[0])
    r = np.random.default_rng(0).standard_normal(S.nCells)
    for _ in range(3):
        S.precondition(ldu.PRECOND_DILU, r)
    S.set_profiling(True)
    S.kernel_times(reset=True)
    reps = 10
    for _ in range(reps):
        S.precondition(ldu.PRECOND_DILU, r)
    kt = S.kernel_times()
    S.set_profiling(False)
    print(f"{name}: cells {S.nCells} fwd {kt['sweep_fwd'][0] / reps * 1e3:.1f} us  bwd {kt['sweep_bwd'][0] / reps * 1e3:.1f} us", flush=True)
    S.close()
