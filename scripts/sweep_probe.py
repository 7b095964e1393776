"""Probe of the pipelined sweep kernels on synthetic structured boxes: time per time step.
usage: python scripts/sweep_probe.py nx ny nz [nx ny nz ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.mesh import StructuredRegion, Block
from multiregionfoam_b200.assembly import synthetic_coeffs, single_region_case

ctx = ldu.Context(0)
args = [int(a) for a in sys.argv[1:]]
for i in range(0, len(args), 3):
    nx, ny, nz = args[i:i + 3]
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    reg = synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False)
    case = single_region_case(reg)
    S = ldu.LduSystem(ctx, case.ranks[0])
    r = np.random.default_rng(0).standard_normal(m.nCells)
    S.precondition(ldu.PRECOND_DILU, r)
    S.set_profiling(True)
    S.kernel_times(reset=True)
    reps = 5
    for _ in range(reps):
        S.precondition(ldu.PRECOND_DILU, r)
    kt = S.kernel_times()
    S.set_profiling(False)
    f, b = kt["sweep_fwd"][0] / reps, kt["sweep_bwd"][0] / reps
    steps = nx + min(ny, 32) - 1
    byt = 24 * m.nCells + 16 * m.nFaces
    print(f"box {nx}x{ny}x{nz}: cells {m.nCells} fwd {f*1e3:.1f} us ({f*1e6/steps:.0f} ns/step, {byt/f/1e6:.0f} GB/s)  "
          f"bwd {b*1e3:.1f} us ({b*1e6/steps:.0f} ns/step, {byt/b/1e6:.0f} GB/s)", flush=True)
    S.close()
