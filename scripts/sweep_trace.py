"""Per-block timeline of one sweep group from the debug time stamps.  usage: python scripts/sweep_trace.py nx ny nz group [first n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import synthetic_coeffs, single_region_case
from multiregionfoam_b200.mesh import StructuredRegion, Block

nx, ny, nz, grp = [int(a) for a in sys.argv[1:5]]
first = int(sys.argv[5]) if len(sys.argv) > 5 else 40
n = int(sys.argv[6]) if len(sys.argv) > 6 else 24
ctx = ldu.Context(0)
if nx == 0:  # C2
    from multiregionfoam_b200.assembly import cht_case
    case = cht_case(3, 22)[0]
else:
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    case = single_region_case(synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False))
S = ldu.LduSystem(ctx, case.ranks[0])
r = np.random.default_rng(0).standard_normal(S.nCells)
for _ in range(3):
    S.precondition(ldu.PRECOND_DILU, r)
S.sweep_stats(+1, True)
S.precondition(ldu.PRECOND_DILU, r)
st = S.sweep_stats(+1, False)
tr = st[grp, 16:].reshape(-1, 8)
t0 = tr[first, 0]
print(f"box {nx}x{ny}x{nz} group {grp}: cycles relative to block {first}'s consumer start")
print("blk  c.ready c.done(dur) | loader.issue(after done of blk-NS) | p0.start p0.stageN(+wait) p0.checked(+) p0.deliver(+) p7.deliver | deliver->c.ready")
for b in range(first, first + n):
    c0, c1, li, p3, p4, p5, p6, p7 = tr[b]
    print(f"{b:3d} {c0 - t0:7d} {c1 - t0:7d}({c1 - c0:4d}) | {li - t0:7d} | {p3 - t0:7d} {p4 - t0:7d}(+{p4 - p3:4d}) {p5 - t0:7d}(+{p5 - p4:4d}) {p6 - t0:7d}(+{p6 - p5:4d}) {p7 - t0:7d} | {c0 - max(p6, p7):5d}")
