"""End time (us) of every group of one forward sweep on a structured box, arranged by (j-group, k): shows the
lag per group hop.  usage: python scripts/sweep_table.py nx ny nz"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.build import build_schedule_emulator
from multiregionfoam_b200.mesh import StructuredRegion, Block
from multiregionfoam_b200.assembly import synthetic_coeffs, single_region_case

nx, ny, nz = map(int, sys.argv[1:4])
m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
E = C.CDLL(build_schedule_emulator())
n = m.nCells
l = np.ascontiguousarray(m.lowerAddr, np.int32); u = np.ascontiguousarray(m.upperAddr, np.int32)
grp = np.zeros(n, np.int32); tim = np.zeros(n, np.int32); lan = np.zeros(n, np.int32); gi = np.zeros(8 * 100000, np.int32)
I = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
ng = E.emu_placement(n, m.nFaces, I(l), I(u), I(grp), I(tim), I(lan), I(gi), 100000)
cells = np.arange(n)
jj, kk = (cells // nx) % ny, cells // (nx * ny)
gj = np.full(ng, 10**9); gk = np.zeros(ng, int)
np.minimum.at(gj, grp, jj)
gk[grp] = kk
ctx = ldu.Context(0)
case = single_region_case(synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False))
S = ldu.LduSystem(ctx, case.ranks[0])
r = np.random.default_rng(0).standard_normal(S.nCells)
for _ in range(3):
    S.precondition(ldu.PRECOND_DILU, r)
S.set_profiling(True)
S.kernel_times(reset=True)
for _ in range(5):
    S.precondition(ldu.PRECOND_DILU, r)
kt = S.kernel_times()
S.set_profiling(False)
print(f"untimed-by-counters run: fwd {kt['sweep_fwd'][0] / 5 * 1e3:.1f} us, bwd {kt['sweep_bwd'][0] / 5 * 1e3:.1f} us per sweep")
S.sweep_stats(+1, True)
S.precondition(ldu.PRECOND_DILU, r)
st = S.sweep_stats(+1, False)
t0 = st[:, 2].min()
js = sorted(set(gj.tolist()))
print("rows: k, columns: j-group (first j); entries: end time us / wait frac")
for k in range(nz):
    row = []
    for j0 in js:
        g = [i for i in range(ng) if gj[i] == j0 and gk[i] == k]
        row.append(" ".join(f"{(st[i,3]-t0)/1e3:6.1f}/{st[i,1]/max(st[i,0],1):.2f}" for i in g))
    print(f"{k:3d}  " + "   ".join(row))
if os.environ.get("B200_SWEEP_DEBUG") == "2":
    print("progress of the first j-group by k: time (us) at which i/8 of the blocks were finished")
    for k in range(nz):
        g = [i for i in range(ng) if gj[i] == js[0] and gk[i] == k][0]
        print(f"  k {k:2d}: " + " ".join(f"{(st[g, 8 + i] - t0) / 1e3:7.2f}" for i in range(8)))
    print("SM of each group, rows k, columns j-group:")
    for k in range(nz):
        print(f"  k {k:2d}: " + " ".join(f"{st[[i for i in range(ng) if gj[i] == j0 and gk[i] == k][0], 6]:4d}" for j0 in js))
    print("same for k = 5, by j-group")
    for j0 in js:
        g = [i for i in range(ng) if gj[i] == j0 and gk[i] == 5][0]
        print(f"  j0 {j0:3d}: " + " ".join(f"{(st[g, 8 + i] - t0) / 1e3:7.2f}" for i in range(8)))
print("sweep span us", (st[:, 3].max() - t0) / 1e3, " kernel ms", S.kernel_times())
