"""C2: end time of every fluid group of one forward sweep arranged by (j-block, k), product kernel (B200_SWEEP_DEBUG=2:
only start / end stamps) or instrumented kernel.  Shows what a hop in j and in k costs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case

ctx = ldu.Context(0)
nj, nk = 4, 22
if len(sys.argv) > 3:  # single-region box nx ny nz [number of equal mesh blocks in x]
    from multiregionfoam_b200.assembly import synthetic_coeffs, single_region_case
    from multiregionfoam_b200.mesh import StructuredRegion, Block
    nx, ny, nz = [int(a) for a in sys.argv[1:4]]
    nb = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    m = StructuredRegion("box", [Block(nx // nb, float(i), float(i + 1), 1.0) for i in range(nb)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    case = single_region_case(synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False))
    nj, nk = (ny + 31) // 32, nz
else:
    case = cht_case(3, 22)[0]
S = ldu.LduSystem(ctx, case.ranks[0])
r = np.random.default_rng(0).standard_normal(S.nCells)
for _ in range(3):
    S.precondition(ldu.PRECOND_DILU, r)
for direction in (+1, -1):
    S.sweep_stats(direction, True)
    S.precondition(ldu.PRECOND_DILU, r)
    st = S.sweep_stats(direction, False)
    t0 = st[:, 2].min()
    end = (st[:, 3] - t0) / 1e3
    big = [g for g in range(len(st)) if st[g, 5] >= (1024 if len(sys.argv) <= 3 else 0)]
    # fluid groups are numbered by forward level jb + k; inside a level by descending jb (creation order k-major)
    T = np.full((nj, nk), np.nan)
    lev = {}
    for jb in range(nj):
        for k in range(nk):
            lev.setdefault(jb + k, []).append((jb, k))
    it = iter(big)
    for L in sorted(lev):
        for (jb, k) in sorted(lev[L], key=lambda p: -p[0]):
            T[jb, k] = end[next(it)]
    print(f"dir {direction:+d}: span {end.max():.1f} us; end times [us], rows jb, columns k")
    for jb in range(nj):
        print(f"  jb{jb}: " + " ".join(f"{T[jb, k]:6.1f}" for k in range(nk)))
    dk = np.diff(T, axis=1)
    dj = np.diff(T, axis=0)
    print(f"  mean k-hop per jb: " + " ".join(f"{np.nanmean(dk[jb]):.2f}" for jb in range(nj)) + f"; mean j-hop per step jb->jb+1: " + " ".join(f"{np.nanmean(dj[j]):.2f}" for j in range(nj - 1)))
    small = [g for g in range(len(st)) if g not in big]
    if small:
        print(f"  solid groups: first end {end[small].min():.1f}, last end {end[small].max():.1f}")
