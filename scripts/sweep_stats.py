"""Per-group timeline of one forward sweep (debug counters): who waits, who is slow."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case
from multiregionfoam_b200.mesh import StructuredRegion, Block
from multiregionfoam_b200.assembly import synthetic_coeffs, single_region_case

ctx = ldu.Context(0)
if len(sys.argv) > 3:
    nx, ny, nz = map(int, sys.argv[1:4])
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    case = single_region_case(synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False))
else:
    case = cht_case(3, 22)[0]
S = ldu.LduSystem(ctx, case.ranks[0])
r = np.random.default_rng(0).standard_normal(S.nCells)
for _ in range(3):
    S.precondition(ldu.PRECOND_DILU, r)
S.sweep_stats(+1, True)
S.precondition(ldu.PRECOND_DILU, r)
st = S.sweep_stats(+1, False)
t0 = st[:, 2].min()
print("groups", len(st), "sweep span us", (st[:, 3].max() - t0) / 1e3)
print("ticket start_us end_us dur_us consumer_cycles wait_frac polls nT ns/step tma_frac tail_frac")
sel = list(range(0, len(st), max(1, len(st) // 24)))
for i in sel:
    c, w, a, b, polls, nT, tma, tail = st[i, :8]
    print(f"{i:5d} {(a-t0)/1e3:8.1f} {(b-t0)/1e3:8.1f} {(b-a)/1e3:8.1f} {c:10d} {w/max(c,1):6.2f} {polls:6d} {nT:5d} {(b-a)/max(nT,1):6.0f} {tma/max(c,1):6.2f} {tail/max(c,1):6.2f}")
print("mean wait frac", float((st[:, 1] / np.maximum(st[:, 0], 1)).mean()), "mean ns/step", float(((st[:, 3] - st[:, 2]) / np.maximum(st[:, 5], 1)).mean()))
print("end_us by ticket (fluid nT=1040 only):")
ends = [(i, (st[i, 3] - t0) / 1e3, st[i, 1] / max(st[i, 0], 1)) for i in range(len(st)) if st[i, 5] >= 1024]
for k in range(0, len(ends), 8):
    print("  ".join(f"{i:3d}:{e:6.1f}/{w:.2f}" for i, e, w in ends[k:k + 8]))
