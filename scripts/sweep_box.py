"""Hop-cost experiments for the sweep kernel on small structured boxes (one group per 32 j-lines and z-layer):
per-group start/end/wait from the debug counters.  usage: python scripts/sweep_box.py nx ny nz [nx ny nz ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import synthetic_coeffs, single_region_case
from multiregionfoam_b200.mesh import StructuredRegion, Block

ctx = ldu.Context(0)
args = [int(a) for a in sys.argv[1:]]
for i in range(0, len(args), 3):
    nx, ny, nz = args[i:i + 3]
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    case = single_region_case(synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False))
    S = ldu.LduSystem(ctx, case.ranks[0])
    r = np.random.default_rng(0).standard_normal(S.nCells)
    for _ in range(3):
        S.precondition(ldu.PRECOND_DILU, r)
    S.set_profiling(True)
    S.kernel_times(reset=True)
    for _ in range(10):
        S.precondition(ldu.PRECOND_DILU, r)
    kt = S.kernel_times()
    S.set_profiling(False)
    for direction in (+1,):
        S.sweep_stats(direction, True)
        S.precondition(ldu.PRECOND_DILU, r)
        st = S.sweep_stats(direction, False)
        t0 = st[:, 2].min()
        print(f"box {nx}x{ny}x{nz} stages={os.environ.get('B200_SWEEP_STAGES', '4')}: fwd {kt['sweep_fwd'][0] / 10 * 1e3:.1f} us bwd {kt['sweep_bwd'][0] / 10 * 1e3:.1f} us; "
              f"groups {len(st)}; id:end_us/ns_per_step/wait_frac/prod_stagewait/prod_spin")
        print("   " + "  ".join(f"{g}:{(st[g, 3] - t0) / 1e3:.1f}/{(st[g, 3] - st[g, 2]) / max(st[g, 5], 1):.0f}/{st[g, 1] / max(st[g, 0], 1):.2f}/"
                                f"{st[g, 9] / max(st[g, 8], 1):.2f}/{st[g, 11] / max(st[g, 8], 1):.2f}[cons cyc/blk {st[g, 0] / max(st[g, 7], 1):.0f} wait {st[g, 1] / max(st[g, 7], 1):.0f} general {st[g, 6]}/{st[g, 7]}; prod0 cyc/blk {st[g, 8] / max(st[g, 7], 1):.0f} stage {st[g, 9] / max(st[g, 7], 1):.0f} val {st[g, 10] / max(st[g, 7], 1):.0f} spin {st[g, 11] / max(st[g, 7], 1):.0f}]\n   " for g in range(min(len(st), 24))), flush=True)
    S.close()
