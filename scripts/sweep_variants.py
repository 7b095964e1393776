"""Sweep kernel diagnostics for one build of the library (select with B200_LDU_LIB=...): time per launch (CUDA
events, counters off), then the per-group timeline from the debug counters of one forward sweep.
usage: python scripts/sweep_variants.py [label]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiregionfoam_b200 import ldu
from multiregionfoam_b200.assembly import cht_case

label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("B200_LDU_LIB", "default")
ctx = ldu.Context(0)
case = cht_case(3, 22)[0]
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
print(f"[{label}] C2 fwd {kt['sweep_fwd'][0] / reps * 1e3:.1f} us  bwd {kt['sweep_bwd'][0] / reps * 1e3:.1f} us", flush=True)
if os.environ.get("NO_STATS"):
    sys.exit(0)
for direction in (+1, -1):
    S.sweep_stats(direction, True)
    S.precondition(ldu.PRECOND_DILU, r)
    st = S.sweep_stats(direction, False)
    t0 = st[:, 2].min()
    dur = (st[:, 3] - st[:, 2]).astype(float)
    nT = np.maximum(st[:, 5], 1)
    big = st[:, 5] >= 1024
    print(f"[{label}] dir {direction:+d}: groups {len(st)} span {(st[:, 3].max() - t0) / 1e3:.1f} us; ns/step all {np.mean(dur / nT):.1f} fluid {np.mean((dur / nT)[big]):.1f} "
          f"solid {np.mean((dur / nT)[~big]):.1f}; consumer wait frac {np.mean(st[:, 1] / np.maximum(st[:, 0], 1)):.2f}; "
          f"consumer cycles/step {np.mean(st[:, 0] / nT):.0f}; producer0 stage-wait {np.mean(st[:, 9] / np.maximum(st[:, 8], 1)):.2f} "
          f"spin {np.mean(st[:, 11] / np.maximum(st[:, 8], 1)):.2f}; polls/group {np.mean(st[:, 4]):.0f}")
    for gi in range(min(len(st), int(os.environ.get("FIRST_GROUPS", "0")))):
        print(f"   g{gi}: nT {st[gi, 5]} end {(st[gi, 3] - t0) / 1e3:.1f} us, {(st[gi, 3] - st[gi, 2]) / max(st[gi, 5], 1):.0f} ns/step; cons cyc/blk {st[gi, 0] / max(st[gi, 7], 1):.0f} wait {st[gi, 1] / max(st[gi, 7], 1):.0f} "
              f"general {st[gi, 6]}/{st[gi, 7]}; prod0 cyc/blk {st[gi, 8] / max(st[gi, 7], 1):.0f} stage {st[gi, 9] / max(st[gi, 7], 1):.0f} val {st[gi, 10] / max(st[gi, 7], 1):.0f} spin {st[gi, 11] / max(st[gi, 7], 1):.0f}")
    idx = [i for i in np.argsort(st[:, 2]) if big[i]]
    print("   fluid groups by start: id:start/end/wait")
    for k in range(0, len(idx), 8):
        print("   " + "  ".join(f"{i:3d}:{(st[i, 2] - t0) / 1e3:5.1f}/{(st[i, 3] - t0) / 1e3:5.1f}/{st[i, 1] / max(st[i, 0], 1):.2f}" for i in idx[k:k + 8]))
S.close()
