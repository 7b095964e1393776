"""Sweep kernel under several run-time settings in ONE process (the case is assembled once): time per launch and the
per-group counters of the leader groups.  usage: python scripts/sweep_env_scan.py "K=V K2=V2" "K=V" ...  ("-" = defaults)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case

ctx = ldu.Context(0)
if os.environ.get("SCAN_BOX"):  # single-region box nx,ny,nz instead of C2
    from multiregionfoam_b200.assembly import synthetic_coeffs, single_region_case
    from multiregionfoam_b200.mesh import StructuredRegion, Block
    nx, ny, nz = [int(a) for a in os.environ["SCAN_BOX"].split(",")]
    m = StructuredRegion("box", [Block(nx, 0.0, 1.0, 1.0)], ny=ny, nz=nz, y0=0.0, y1=1.0, grady=1.0).build()
    case = single_region_case(synthetic_coeffs(m.nCells, m.lowerAddr, m.upperAddr, symmetric=False))
else:
    case = cht_case(3, 22)[0]
r = np.random.default_rng(0).standard_normal(case.nCells)
KEYS = ["B200_SWEEP_STAGES", "B200_SWEEP_SMEM_KB", "B200_SWEEP_L2AHEAD", "B200_SWEEP_DEBUG"]
for spec in sys.argv[1:]:
    for k in KEYS:
        os.environ.pop(k, None)
    if spec != "-":
        for kv in spec.split():
            k, v = kv.split("=")
            os.environ[k] = v
    S = ldu.LduSystem(ctx, case.ranks[0])
    try:
        import hashlib
        for _ in range(3):
            w = S.precondition(ldu.PRECOND_DILU, r)
        print(f"[{spec}] result sha1 {hashlib.sha1(w.tobytes()).hexdigest()[:12]} (bit-identical builds print the same)", flush=True)
        S.set_profiling(True)
        S.kernel_times(reset=True)
        reps = 10
        for _ in range(reps):
            S.precondition(ldu.PRECOND_DILU, r)
        kt = S.kernel_times()
        S.set_profiling(False)
        print(f"[{spec}] {os.environ.get('SCAN_BOX', 'C2')} fwd {kt['sweep_fwd'][0] / reps * 1e3:.1f} us  bwd {kt['sweep_bwd'][0] / reps * 1e3:.1f} us", flush=True)
        if os.environ.get("SCAN_STATS"):
            S.sweep_stats(+1, True)
            S.precondition(ldu.PRECOND_DILU, r)
            st = S.sweep_stats(+1, False)
            t0 = st[:, 2].min()
            for gi in [int(g) for g in os.environ["SCAN_STATS"].split(",")]:
                nb = max(st[gi, 7], 1)
                print(f"   g{gi}: nT {st[gi, 5]} start {(st[gi, 2] - t0) / 1e3:.1f} end {(st[gi, 3] - t0) / 1e3:.1f} us, {(st[gi, 3] - st[gi, 2]) / max(st[gi, 5], 1):.0f} ns/step; "
                      f"cons cyc/blk {st[gi, 0] / nb:.0f} wait {st[gi, 1] / nb:.0f}; prod0 cyc/blk {st[gi, 8] / nb:.0f} stage {st[gi, 9] / nb:.0f} val {st[gi, 10] / nb:.0f} spin {st[gi, 11] / nb:.0f}")
                print(f"      refill landed after {st[gi, 12] / max(st[gi, 13], 1):.0f} cycles on average (max {st[gi, 15]}, {st[gi, 13]} refills)")
                if os.environ.get("B200_SWEEP_DEBUG") == "2":
                    nc = max(st[gi, 17], 1)
                    print(f"      canonical blocks {st[gi, 17]}: body {st[gi, 16] / nc:.0f} cycles; other blocks {nb - st[gi, 17]}: body {st[gi, 18] / max(nb - st[gi, 17], 1):.0f} cycles")
                tr = st[gi, 16:].reshape(-1, 8)
                for b in (range(40, 46) if os.environ.get("B200_SWEEP_DEBUG") != "2" else ()):
                    c0, c1, li, p3, p4, p5, p6, p7 = tr[b]
                    z = tr[40, 0]
                    print(f"      blk {b}: c.ready {c0 - z} c.done {c1 - z} | loader.issue {li - z} | p0.start {p3 - z} stage {p4 - z} checked {p5 - z} deliver {p6 - z} | pLast.deliver {p7 - z}")
    finally:
        S.close()
